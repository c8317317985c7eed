timeout 900 python -m pytest tests/test_persistent_gpu.py tests/test_engine_gpu.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2y_tests.log
cat gpurun_out/r2y_tests.log
for bo in 256 64 1024; do
MTX_PK_POLL_BACKOFF=$bo timeout 300 python bench.py --steps 100 --warmup 5 --skip-cpu-baseline --no-verify 2>gpurun_out/r2y_bench_$bo.err | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); r = d['roofline']
print('backoff $bo', round(d['ms_per_step'], 4), 'ms/step e2e', round(d['e2e']['ms_per_step'], 4), r.get('persistent_step_phases'))"
done > gpurun_out/r2y_bench.txt 2>&1
cat gpurun_out/r2y_bench.txt
