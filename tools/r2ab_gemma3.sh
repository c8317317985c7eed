# gemma3-27b geometry (E 5376, 32/16 heads x 128, MLP 21504, V 262144, window 1024), 12 of its 62 layers, batch 64 and 256
for b in 64 256; do
timeout 900 python bench.py --model gemma3-27b --layers 12 --batch $b --prefill-len 1024 --target-len 3072 --steps 20 --warmup 3 --skip-cpu-baseline > gpurun_out/r2ab_gemma3_b$b.json 2> gpurun_out/r2ab_gemma3_b$b.err
python - <<PY
import json
try:
  d = json.loads(open("gpurun_out/r2ab_gemma3_b$b.json").read().strip().splitlines()[-1]); r = d["roofline"]
  print("batch $b", round(d["ms_per_step"], 3), "ms/step", round(d["value"]), d["unit"], "frac", r.get("frac"), "verify", d.get("verify", {}).get("ok"), r.get("class_ms_per_step"))
except Exception as ex:
  print("batch $b failed", ex); print(open("gpurun_out/r2ab_gemma3_b$b.err").read()[-1500:])
PY
done
