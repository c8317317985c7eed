timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > gpurun_out/r2k_tests.log
python bench.py --mode prefill --steps 5 --warmup 2 > gpurun_out/r2k_prefill.json 2> gpurun_out/r2k_prefill.err
MTX_PREFILL_ATTENTION=0 python bench.py --mode prefill --steps 3 --warmup 1 > gpurun_out/r2k_prefill_oldattn.json 2>> gpurun_out/r2k_prefill.err
MTX_PREFILL_ATTENTION=0 MTX_ROWS_KERNEL=0 python bench.py --mode prefill --steps 3 --warmup 1 > gpurun_out/r2k_prefill_r1.json 2>> gpurun_out/r2k_prefill.err
python bench.py --steps 100 --warmup 5 --skip-cpu-baseline --no-verify > gpurun_out/r2k_bench_coop.json 2> gpurun_out/r2k_bench.err
MTX_PK_COOPERATIVE=0 python bench.py --steps 100 --warmup 5 --skip-cpu-baseline --no-verify > gpurun_out/r2k_bench_pdl.json 2>> gpurun_out/r2k_bench.err
python - <<'PY'
import json
for f in ["r2k_prefill", "r2k_prefill_oldattn", "r2k_prefill_r1"]:
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["value"], 2), "ms prefill", round(d["insert_ms"], 3), "ms insert", round(d["tflops"], 1), "TFLOP/s", d["gpu_launches_per_prefill"], "launches")
    except Exception as ex:
        print(f, "failed", ex)
for f in ["r2k_bench_coop", "r2k_bench_pdl"]:
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"], 4), "ms/step e2e", round(d["e2e"]["ms_per_step"], 4))
    except Exception as ex:
        print(f, "failed", ex)
PY
