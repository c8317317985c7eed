timeout 900 python -m pytest tests/test_engine_gpu.py -m gpu -q -x -k "topk or nucleus or weighted" 2>&1 | tail -20 > gpurun_out/r2s_tests.log
for s in topk nucleus; do
  python bench.py --sampling $s --steps 50 --warmup 5 --skip-cpu-baseline --no-verify 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); r = d['roofline']
print('$s', round(d['ms_per_step'], 4), 'ms/step e2e', round(d['e2e']['ms_per_step'], 4), 'launches/step', d['launches_per_step'], 'sampler class ms', r['class_ms_per_step']['finalize'], 'persistent', r['class_ms_per_step']['persistent_step'])"
done
