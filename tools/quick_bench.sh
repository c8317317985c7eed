#!/bin/bash
# usage: tools/quick_bench.sh TAG  -> one bench.py run (64 steps), prints ms/step and the per-phase microseconds
tag=${1:-q}
python bench.py --steps 64 --warmup 5 > gpurun_out/$tag.json 2> gpurun_out/$tag.err || { tail -5 gpurun_out/$tag.err; exit 1; }
python - <<PY
import json
d = json.load(open("gpurun_out/$tag.json"))
ph = d["roofline"].get("persistent_step_phases") or {}
print("$tag", round(d["ms_per_step"], 4), {k.replace("_per_layer", ""): v for k, v in ph.items() if k not in ("ctas", "unit")})
PY
