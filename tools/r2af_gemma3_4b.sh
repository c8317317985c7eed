# the full gemma3-4b model (34 layers, E 2560, 8/4 heads x 256, MLP 10240, V 262144, window 1024), batch 64 and 256
for b in 64 256; do
timeout 900 python bench.py --model gemma3-4b --batch $b --prefill-len 1024 --target-len 3072 --steps 20 --warmup 3 --skip-cpu-baseline > gpurun_out/r2af_gemma3_4b_b$b.json 2> gpurun_out/r2af_gemma3_4b_b$b.err
python - <<PY
import json
try:
  d = json.loads(open("gpurun_out/r2af_gemma3_4b_b$b.json").read().strip().splitlines()[-1]); r = d["roofline"]
  print("batch $b", round(d["ms_per_step"], 3), "ms/step", round(d["value"]), d["unit"], "frac", r.get("frac"), "whole-step frac", r.get("whole_step", {}).get("frac"), "verify", d.get("verify", {}).get("ok"), r.get("class_ms_per_step"))
except Exception as ex:
  print("batch $b failed", ex); print(open("gpurun_out/r2af_gemma3_4b_b$b.err").read()[-1500:])
PY
done
