#!/bin/bash
# A/B of the persistent-kernel experiments of round 2 (one gpurun call): every line is `bench.py` at batch 64.
#   MTX_PK_VARIANT bit 0: L2 prefetch of the next layer's K/V tiles at the end of a warp's tile loop
#   MTX_PK_VARIANT bit 1: L2 prefetch of this layer's K/V tiles at the start of the layer
#   --no-fold          : RMSNorm scales applied to the activation tiles in shared memory (round-1 behaviour)
out=gpurun_out/${1:-r2b}_variants.jsonl
: > $out
run() { # name, env, args
  echo "== $1" >&2
  env $2 python bench.py --steps 100 --warmup 5 --skip-cpu-baseline --no-verify $3 2>gpurun_out/${1}.err | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
d['variant'] = '$1'
print(json.dumps(d))" >> $out
}
run nofold_v0 "MTX_PK_VARIANT=0" "--no-fold"
run fold_v0 "MTX_PK_VARIANT=0" ""
run fold_v1 "MTX_PK_VARIANT=1" ""
run fold_v2 "MTX_PK_VARIANT=2" ""
run fold_v3 "MTX_PK_VARIANT=3" ""
python - <<'PY'
import json, sys
for line in open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/r2b_variants.jsonl"):
    d = json.loads(line)
    ph = d["roofline"]["persistent_step_phases"] or {}
    print(f"{d['variant']:12s} {d['ms_per_step']:.4f} ms  e2e {d['e2e']['ms_per_step']:.4f}  " + " ".join(f"{k.replace('_per_layer','')}={v}" for k, v in ph.items() if k not in ('ctas','unit')))
PY
