#!/bin/bash
# Batch sweep and the other BASELINE configurations on one GPU; one JSON line per run into gpurun_out/sweep.jsonl
out=gpurun_out/sweep.jsonl
: > $out
for b in 1 2 4 8 16 32 64; do
  python bench.py --batch $b --steps 32 --warmup 4 --skip-cpu-baseline >> $out 2>> gpurun_out/sweep.err
done
# configs[2]: batch 256, every context 2048 (tensor-core regime; per-kernel path)
python bench.py --batch 256 --context-min 2048 --context-max 2048 --steps 16 --warmup 3 --skip-cpu-baseline >> $out 2>> gpurun_out/sweep.err
# configs[3]: long prompt, 4k prefill + 1.5k decode, batch 1 and 8
for b in 1 8; do
  python bench.py --batch $b --prefill-len 4096 --target-len 5632 --context-min 4000 --context-max 5536 --steps 32 --warmup 4 --skip-cpu-baseline >> $out 2>> gpurun_out/sweep.err
done
python - <<'PY'
import json
for line in open('gpurun_out/sweep.jsonl'):
    d = json.loads(line)
    c = d['config']; r = d['roofline']
    print(f"batch {c['batch_per_gpu']:4d} ctx {c['context']:45s} {d['ms_per_step']:8.3f} ms/step {d['value']:10.0f} tok/s  step {r['whole_step']['achieved_gbs']:7.0f} GB/s ({100*r['whole_step']['frac']:.1f}% of HBM)  launches/step {d['launches_per_step']}  dominant {r['kernel']}")
PY
