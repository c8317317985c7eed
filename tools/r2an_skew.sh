for v in 0 6 12 18 0 12; do
MTX_PK_ATTN_SKEW=$v timeout 300 python bench.py --steps 100 --warmup 5 --skip-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); r = d['roofline']
print('skew $v%', round(d['ms_per_step'], 4), 'ms/step verify', d['verify']['ok'], r.get('persistent_step_phases'))"
done > gpurun_out/r2an_attn_skew.txt 2>&1
cat gpurun_out/r2an_attn_skew.txt
