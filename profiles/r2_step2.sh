# Round 2: parity after the heavy-row fix; A/B of round-1 library vs new one on the same box (ncu launch lists); full capture of k_mark
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r2_step2_tests.log 2>&1; tail -5 gpurun_out/r2_step2_tests.log
CMD="python profiles/exp.py --config 2 --steps 2 --warmup 1"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r2_step2_new.csv $CMD > gpurun_out/ncu1.log 2>&1
OGB_LIB=$PWD/profiles/libogb_r1.so $CMD > gpurun_out/plain_r1.log 2>&1 && OGB_LIB=$PWD/profiles/libogb_r1.so ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r2_step2_r1.csv $CMD > gpurun_out/ncu2.log 2>&1
CMD2="python profiles/exp.py --config 2 --steps 1 --warmup 0"
$CMD2 > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k 'regex:^(k_mark|k_keep|k_rows_finish)' -c 4 -o gpurun_out/prof_r2_mark $CMD2 > gpurun_out/ncu_full.log 2>&1
tail -n 2 gpurun_out/ncu_full.log
