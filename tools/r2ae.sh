timeout 600 python -m pytest tests/test_engine_gpu.py -m gpu -q -x -k "generate_to_host" 2>&1 | tail -3
for hg in 0 1; do
MTX_HOST_GRAPH=$hg timeout 300 python bench.py --steps 200 --warmup 5 --skip-cpu-baseline --no-verify 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('host_graph $hg', round(d['ms_per_step'], 4), 'ms/step e2e', round(d['e2e']['ms_per_step'], 4))"
done
