mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r2_step17_tests.log 2>&1; tail -3 gpurun_out/r2_step17_tests.log
E3="timeout 200 python profiles/exp.py --config 3 --steps 3 --warmup 1"
$E3 --tag c3_fwdstore > gpurun_out/r2_step17.txt 2>&1
timeout 200 python profiles/exp.py --config 2 --steps 3 --warmup 1 --tag c2 >> gpurun_out/r2_step17.txt 2>&1
timeout 200 python profiles/exp.py --config 5 --steps 3 --warmup 1 --tag c5 >> gpurun_out/r2_step17.txt 2>&1
timeout 200 python profiles/exp.py --config 4 --scale 0.2 --steps 2 --warmup 1 --tag c4 >> gpurun_out/r2_step17.txt 2>&1
grep "^\[c" gpurun_out/r2_step17.txt
