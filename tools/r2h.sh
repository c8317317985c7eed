for ipsm in 1 12 24 48; do
  MTX_ATTN_ITEMS_PER_SM=$ipsm python bench.py --batch 256 --context-min 2048 --context-max 2048 --steps 16 --warmup 3 --skip-cpu-baseline --no-verify 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); r = d['roofline']
print('items/SM $ipsm', round(d['ms_per_step'], 3), 'ms/step', r['class_ms_per_step']['attention'])"
done
