# Round 2, first GPU call: round-1 code on config 3 (the new bench workload) -- step phases, chunk-size sweep,
# ncu launch list. Run on the GPU box: bash profiles/r2_base.sh
mkdir -p gpurun_out
E="python profiles/exp.py --config 3 --steps 4 --warmup 2"
$E --tag c3_default > gpurun_out/r2_base.txt 2>&1
OGB_CHUNK_READS=524288 $E --tag c3_chunk512k >> gpurun_out/r2_base.txt 2>&1
OGB_CHUNK_READS=1048576 $E --tag c3_chunk1M >> gpurun_out/r2_base.txt 2>&1
OGB_CHUNK_READS=1048576 OGB_SUB_PARTITIONS=24 $E --tag c3_chunk1M_sub24 >> gpurun_out/r2_base.txt 2>&1
OGB_CHUNK_READS=1048576 OGB_ONE_STREAM=1 $E --tag c3_chunk1M_onestream >> gpurun_out/r2_base.txt 2>&1
OGB_SUB_PARTITIONS=24 $E --tag c3_sub24 >> gpurun_out/r2_base.txt 2>&1
cat gpurun_out/r2_base.txt
CMD="python profiles/exp.py --config 3 --steps 1 --warmup 1"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2_base_c3.csv $CMD > gpurun_out/ncu_launches.log 2>&1
tail -3 gpurun_out/ncu_launches.log
