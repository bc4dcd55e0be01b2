mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r2_step10_tests.log 2>&1; tail -6 gpurun_out/r2_step10_tests.log
