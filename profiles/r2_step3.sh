mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r2_step3_tests.log 2>&1; tail -5 gpurun_out/r2_step3_tests.log
E2="python profiles/exp.py --config 2 --steps 4 --warmup 2"
E3="python profiles/exp.py --config 3 --steps 3 --warmup 1"
$E2 --tag c2_new > gpurun_out/r2_step3.txt 2>&1
OGB_LIB=$PWD/profiles/variants/libogb_mb6.so $E2 --tag c2_mb6 >> gpurun_out/r2_step3.txt 2>&1
$E3 --tag c3_new >> gpurun_out/r2_step3.txt 2>&1
OGB_LIB=$PWD/profiles/variants/libogb_mb6.so $E3 --tag c3_mb6 >> gpurun_out/r2_step3.txt 2>&1
cat gpurun_out/r2_step3.txt
CMD2="python profiles/exp.py --config 2 --steps 1 --warmup 0"
$CMD2 > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k 'regex:^(k_mark)' -c 1 -o gpurun_out/prof_r2_mark2 $CMD2 > gpurun_out/ncu_full.log 2>&1
tail -n 2 gpurun_out/ncu_full.log
