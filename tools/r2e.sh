timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_baseline_configs_gpu.py::test_batch256_context2048_against_the_oracle tests/test_engine_gpu.py tests/test_golden.py -m gpu -q -x 2>&1 | tail -30 > gpurun_out/r2e_tests.log
bash tools/r2_batch256.sh r2e > gpurun_out/r2e_summary.txt 2>&1
