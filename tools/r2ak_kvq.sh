for v in "--kv-int8" "--kv-int8 --kv-axis heads_and_dkv" "--kv-fp8" "--kv-fp8 --kv-axis heads_and_dkv"; do
timeout 600 python bench.py --batch 256 --context-min 2048 --context-max 2048 --steps 30 --warmup 3 --skip-cpu-baseline $v 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); r = d['roofline']
print('$v', round(d['ms_per_step'], 3), 'ms/step', round(d['value']), 'tok/s verify', d['verify']['ok'], 'attention ms', r['class_ms_per_step']['attention'], 'qkv', r['class_ms_per_step']['qkv_rope_append'])"
done > gpurun_out/r2ak_kv_quant_batch256.txt 2>&1
cat gpurun_out/r2ak_kv_quant_batch256.txt
