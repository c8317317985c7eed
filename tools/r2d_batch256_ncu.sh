timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_baseline_configs_gpu.py::test_batch256_context2048_against_the_oracle tests/test_engine_gpu.py tests/test_golden.py -m gpu -q -x 2>&1 | tail -30 > gpurun_out/r2d_tests.log
bash tools/r2_batch256.sh r2d > gpurun_out/r2d_summary.txt 2>&1
CMD="python bench.py --batch 256 --context-min 2048 --context-max 2048 --steps 2 --warmup 3 --no-graph --skip-cpu-baseline --no-verify"
$CMD > gpurun_out/r2d_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:gemm_rows -s 97 -c 97 --csv --log-file gpurun_out/r2d_ncu_gemm_rows.csv $CMD > gpurun_out/r2d_ncu.log 2>&1
echo ncu_rc=$? >> gpurun_out/r2d_summary.txt
