mkdir -p gpurun_out
CMD="python profiles/exp.py --config 3 --steps 1 --warmup 1"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2_step4_new.csv $CMD > gpurun_out/ncu1.log 2>&1
OGB_LIB=$PWD/profiles/libogb_r1.so $CMD > gpurun_out/plain_r1.log 2>&1 && OGB_LIB=$PWD/profiles/libogb_r1.so ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2_step4_r1.csv $CMD > gpurun_out/ncu2.log 2>&1
tail -2 gpurun_out/ncu2.log
