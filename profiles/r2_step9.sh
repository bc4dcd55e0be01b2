mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r2_step9_tests.log 2>&1; tail -4 gpurun_out/r2_step9_tests.log
E3="timeout 300 python profiles/exp.py --config 3 --steps 3 --warmup 1"
$E3 --tag c3_k1queues > gpurun_out/r2_step9.txt 2>&1
OGB_K1_QUEUES=0 $E3 --tag c3_k1direct >> gpurun_out/r2_step9.txt 2>&1
OGB_ONE_STREAM=1 $E3 --tag c3_onestream >> gpurun_out/r2_step9.txt 2>&1
timeout 300 python profiles/exp.py --config 4 --scale 0.2 --steps 3 --warmup 1 --tag c4_0.2 >> gpurun_out/r2_step9.txt 2>&1
timeout 300 python profiles/exp.py --config 2 --steps 3 --warmup 1 --tag c2 >> gpurun_out/r2_step9.txt 2>&1
cat gpurun_out/r2_step9.txt
