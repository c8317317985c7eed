timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_parallel.py -m gpu -q -x -k "topk or nucleus or weighted or top_k or ties" 2>&1 | tail -20 > gpurun_out/r2x_tests.log
for s in topk nucleus; do
  python bench.py --sampling $s --steps 50 --warmup 5 --skip-cpu-baseline --no-verify 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); r = d['roofline']
print('$s', round(d['ms_per_step'], 4), 'ms/step e2e', round(d['e2e']['ms_per_step'], 4), 'launches/step', d['launches_per_step'], 'sampler class ms', r['class_ms_per_step']['finalize'], 'persistent', r['class_ms_per_step']['persistent_step'])"
done > gpurun_out/r2x_bench.txt
CMD="python bench.py --sampling topk --steps 2 --warmup 3 --no-graph --skip-cpu-baseline --no-verify"
$CMD > gpurun_out/r2x_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:par_\|finalize -s 24 -c 24 --csv --log-file gpurun_out/r2x_ncu_par.csv $CMD > gpurun_out/r2x_ncu.log 2>&1
