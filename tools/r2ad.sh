for v in 0 8 16 0 8 16; do
MTX_PK_VARIANT=$v timeout 300 python bench.py --steps 200 --warmup 5 --skip-cpu-baseline --no-verify 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); r = d['roofline']
print('variant $v', round(d['ms_per_step'], 4), 'ms/step e2e', round(d['e2e']['ms_per_step'], 4), r.get('persistent_step_phases'))"
done > gpurun_out/r2ad_barrier_pause.txt 2>&1
cat gpurun_out/r2ad_barrier_pause.txt
