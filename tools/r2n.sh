timeout 900 python -m pytest tests/test_kv_quant_gpu.py -m gpu -q -x 2>&1 | tail -30 > gpurun_out/r2n_tests.log
python bench.py --kv-int8 --batch 256 --context-min 2048 --context-max 2048 --steps 16 --warmup 3 --skip-cpu-baseline > gpurun_out/r2n_batch256_int8.json 2> gpurun_out/r2n.err
python bench.py --kv-int8 --batch 64 --steps 32 --warmup 3 --skip-cpu-baseline > gpurun_out/r2n_batch64_int8.json 2>> gpurun_out/r2n.err
python - <<'PY'
import json
for f in ["r2n_batch256_int8", "r2n_batch64_int8"]:
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        r = d["roofline"]
        print(f, round(d["ms_per_step"], 3), "ms/step", round(d["value"]), "tok/s verify", d["verify"], r["class_ms_per_step"])
    except Exception as ex:
        print(f, "failed", ex)
PY
