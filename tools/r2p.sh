timeout 600 python -m pytest tests/test_offline_engine.py -m gpu -q -x 2>&1 | tail -20 > gpurun_out/r2p_tests.log
for s in greedy weighted topk nucleus; do
  python bench.py --sampling $s --steps 50 --warmup 5 --skip-cpu-baseline --no-verify 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); r = d['roofline']
print('$s', round(d['ms_per_step'], 4), 'ms/step e2e', round(d['e2e']['ms_per_step'], 4), 'launches/step', d['launches_per_step'])"
done
