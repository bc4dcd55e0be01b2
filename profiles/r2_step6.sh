mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r2_step6_tests.log 2>&1; tail -5 gpurun_out/r2_step6_tests.log
E2="python profiles/exp.py --config 2 --steps 4 --warmup 2"
E3="python profiles/exp.py --config 3 --steps 3 --warmup 1"
$E2 --tag c2_mb5 > gpurun_out/r2_step6.txt 2>&1
OGB_LIB=$PWD/profiles/variants/libogb_mb4.so $E2 --tag c2_mb4 >> gpurun_out/r2_step6.txt 2>&1
$E3 --tag c3_mb5 >> gpurun_out/r2_step6.txt 2>&1
OGB_LIB=$PWD/profiles/variants/libogb_mb4.so $E3 --tag c3_mb4 >> gpurun_out/r2_step6.txt 2>&1
python profiles/exp.py --config 4 --scale 0.2 --steps 3 --warmup 1 --tag c4_0.2 >> gpurun_out/r2_step6.txt 2>&1
python profiles/exp.py --config 5 --steps 3 --warmup 1 --tag c5 >> gpurun_out/r2_step6.txt 2>&1
cat gpurun_out/r2_step6.txt
CMD="python profiles/exp.py --config 3 --steps 1 --warmup 1"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k 'regex:^(k_mark|k_keep|k_rows|k_emit)' -c 40 --csv --log-file gpurun_out/launches_r2_step6.csv $CMD > gpurun_out/ncu1.log 2>&1
