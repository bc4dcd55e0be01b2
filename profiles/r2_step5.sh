mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r2_step5_tests.log 2>&1; tail -5 gpurun_out/r2_step5_tests.log
E2="python profiles/exp.py --config 2 --steps 4 --warmup 2"
E3="python profiles/exp.py --config 3 --steps 3 --warmup 1"
$E2 --tag c2_mb4 > gpurun_out/r2_step5.txt 2>&1
OGB_LIB=$PWD/profiles/variants/libogb_mb5.so $E2 --tag c2_mb5 >> gpurun_out/r2_step5.txt 2>&1
OGB_LIB=$PWD/profiles/variants/libogb_mb6.so $E2 --tag c2_mb6 >> gpurun_out/r2_step5.txt 2>&1
$E3 --tag c3_mb4 >> gpurun_out/r2_step5.txt 2>&1
OGB_LIB=$PWD/profiles/variants/libogb_mb5.so $E3 --tag c3_mb5 >> gpurun_out/r2_step5.txt 2>&1
OGB_LIB=$PWD/profiles/variants/libogb_mb6.so $E3 --tag c3_mb6 >> gpurun_out/r2_step5.txt 2>&1
cat gpurun_out/r2_step5.txt
