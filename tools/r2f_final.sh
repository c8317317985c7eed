# Round-2 final evidence: GPU test suite, the judged bench line (both arms), the ncu launch list and one full capture.
set -x
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/r2f_pytest.txt
cat gpurun_out/r2f_pytest.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2f_smoke.txt 2>&1; tail -2 gpurun_out/r2f_smoke.txt
timeout 900 python bench.py > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; tail -c 600 gpurun_out/r2f_bench.json
timeout 900 python bench.py --impl reference > gpurun_out/r2f_bench_reference_arm.json 2> gpurun_out/r2f_bench_ref.err; tail -c 400 gpurun_out/r2f_bench_reference_arm.json
CMD="python bench.py --steps 2 --warmup 3 --no-graph --skip-cpu-baseline --no-verify"
$CMD > gpurun_out/r2f_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:step_persistent\|prepare_rows\|finalize_kernel -c 36 --csv --log-file gpurun_out/r2f_launches_raw.csv $CMD > gpurun_out/r2f_ncu_list.log 2>&1
