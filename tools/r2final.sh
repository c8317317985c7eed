# Round-2 final evidence (last session): GPU test suite, smoke, the judged bench line (both arms), the ncu launch list, one full
# capture of the persistent kernel, and the second lines of the rows added in this session.
set -x
T=r2final
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/${T}_pytest_gpu.txt
cat gpurun_out/${T}_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${T}_smoke.txt 2>&1; tail -2 gpurun_out/${T}_smoke.txt
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -c 600 gpurun_out/${T}_bench.json
timeout 900 python bench.py --impl reference > gpurun_out/${T}_bench_reference_arm.json 2> gpurun_out/${T}_bench_ref.err; tail -c 400 gpurun_out/${T}_bench_reference_arm.json
CMD="python bench.py --steps 2 --warmup 3 --no-graph --skip-cpu-baseline --no-verify"
$CMD > gpurun_out/${T}_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:step_persistent\|prepare_rows\|finalize_kernel -c 36 --csv --log-file gpurun_out/${T}_launches_raw.csv $CMD > gpurun_out/${T}_ncu_list.log 2>&1
$CMD > gpurun_out/${T}_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:step_persistent --launch-skip 6 -c 1 -f -o gpurun_out/${T}_full $CMD > gpurun_out/${T}_ncu_full.log 2>&1
ls -la gpurun_out/${T}_full.ncu-rep
timeout 300 python bench.py --paged 32 --skip-cpu-baseline > gpurun_out/${T}_bench_paged32.json 2>/dev/null
timeout 300 python bench.py --kv-int8 --skip-cpu-baseline > gpurun_out/${T}_bench_kv_int8.json 2>/dev/null
timeout 300 python bench.py --batch 8 --prefill-len 4096 --target-len 5632 --context-min 4000 --context-max 5500 --skip-cpu-baseline > gpurun_out/${T}_bench_long_prompt_batch8.json 2>/dev/null
timeout 300 python bench.py --batch 256 --context-min 2048 --context-max 2048 --skip-cpu-baseline --steps 32 > gpurun_out/${T}_bench_batch256_ctx2048.json 2>/dev/null
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${T}_bench*.json")):
  try:
    d=json.load(open(f)); print(f, round(d["ms_per_step"],4), round(d["value"],1), d.get("e2e",{}).get("value"), (d.get("verify") or {}).get("ok"), d.get("roofline",{}).get("frac"))
  except Exception as ex: print(f, "ERR", ex)
PY
