#!/bin/bash
# BASELINE configs[2]: batch 256, every context 2048 (per-kernel path; gemm_rows.cuh for the GEMMs)
tag=${1:-r2c}
python bench.py --batch 256 --context-min 2048 --context-max 2048 --steps 16 --warmup 3 --skip-cpu-baseline --no-verify > gpurun_out/${tag}_batch256.json 2> gpurun_out/${tag}_batch256.err
MTX_ROWS_KERNEL=0 python bench.py --batch 256 --context-min 2048 --context-max 2048 --steps 16 --warmup 3 --skip-cpu-baseline --no-verify > gpurun_out/${tag}_batch256_old.json 2>> gpurun_out/${tag}_batch256.err
python bench.py --batch 128 --context-min 2048 --context-max 2048 --steps 16 --warmup 3 --skip-cpu-baseline --no-verify > gpurun_out/${tag}_batch128.json 2>> gpurun_out/${tag}_batch256.err
python - <<PY
import json
for f in ["gpurun_out/${tag}_batch256.json", "gpurun_out/${tag}_batch256_old.json", "gpurun_out/${tag}_batch128.json"]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as ex:
        print(f, "failed", ex); continue
    r = d["roofline"]
    print(f, round(d["ms_per_step"], 3), "ms/step", round(d["value"]), "tok/s", r["class_ms_per_step"], r["class_launches_per_step"])
PY
