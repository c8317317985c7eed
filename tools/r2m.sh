timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_golden.py tests/test_persistent_gpu.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r2m_tests.log
python tools/prefill_timeline.py 4000 > gpurun_out/r2m_prefill_timeline.txt 2>&1
python bench.py --mode prefill --steps 5 --warmup 2 > gpurun_out/r2m_prefill.json 2> gpurun_out/r2m_prefill.err
