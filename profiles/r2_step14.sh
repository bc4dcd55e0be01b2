mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r2_step14_tests.log 2>&1; tail -4 gpurun_out/r2_step14_tests.log
OGB_DBG_TIMING=1 timeout 300 python profiles/exp.py --config 2 --steps 2 --warmup 1 --device-dataset --tag c2_devds > gpurun_out/r2_step14.txt 2>&1
timeout 300 python profiles/exp.py --config 3 --steps 2 --warmup 1 --device-dataset --tag c3_devds >> gpurun_out/r2_step14.txt 2>&1
timeout 300 python profiles/exp.py --config 5 --steps 2 --warmup 1 --device-dataset --tag c5_devds >> gpurun_out/r2_step14.txt 2>&1
grep -v "^\[finalize_device\]" gpurun_out/r2_step14.txt | cut -c1-330; grep "^\[finalize_device\]" gpurun_out/r2_step14.txt | tail -14
