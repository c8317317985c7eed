#!/bin/bash
out=gpurun_out/sweep_small.jsonl
: > $out
for b in 1 4 8 16 32 64; do
  python bench.py --batch $b --steps 32 --warmup 4 --skip-cpu-baseline >> $out 2>> gpurun_out/sweep.err
done
python - <<'PY'
import json
for line in open('gpurun_out/sweep_small.jsonl'):
    d = json.loads(line); c = d['config']; r = d['roofline']; ph = r.get('persistent_step_phases') or {}
    print(f"batch {c['batch_per_gpu']:4d} {d['ms_per_step']:8.3f} ms/step {d['value']:9.0f} tok/s ({100*r['whole_step']['frac']:.1f}% of HBM) phases", {k.replace('_per_layer',''): v for k, v in ph.items() if 'per_layer' in k})
PY
