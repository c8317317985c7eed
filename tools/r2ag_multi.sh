# 2-GPU sanity of everything the driver launches under torchrun, plus the NCCL tests
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
timeout 900 python -m pytest tests/test_parallel.py -m gpu -q 2>&1 | tail -3
timeout 600 $TR bench.py --gpus 2 --steps 64 --warmup 4 --skip-cpu-baseline 2>gpurun_out/r2ag_n2.err | tail -1 > gpurun_out/r2ag_bench_n2.json
timeout 900 $TR bench.py --impl reference --gpus 2 --steps 4 --warmup 1 2>gpurun_out/r2ag_ref.err | tail -1 > gpurun_out/r2ag_bench_ref_n2.json
timeout 600 $TR bench.py --gpus 2 --steps 64 --warmup 4 --skip-cpu-baseline --scaling strong 2>/dev/null | tail -1 > gpurun_out/r2ag_bench_strong_n2.json
timeout 600 $TR bench.py --gpus 2 --mode vocab-parallel --sampling topk --steps 64 --warmup 4 --skip-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r2ag_bench_vp_topk_n2.json
python - <<'PY'
import json
for f in ("r2ag_bench_n2", "r2ag_bench_ref_n2", "r2ag_bench_strong_n2", "r2ag_bench_vp_topk_n2"):
  try:
    d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
    print(f, d.get("impl"), d["n_gpus"], round(d["value"], 1), d["unit"], "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"], 1), d.get("scaling"))
  except Exception as ex:
    print(f, "FAILED", ex)
PY
