# chunk size vs index size on one GPU: config 3 at scale 4 (34 M reads, 2.2 GB index, 2.2 GB read store) emulates what a rank of 4 sees, scale 1 what a rank of 1 sees
mkdir -p gpurun_out
E="python profiles/exp.py --config 3 --scale 4.0 --steps 2 --warmup 1"
$E --tag s4_256k > gpurun_out/r2_step8.txt 2>&1
OGB_CHUNK_READS=1048576 $E --tag s4_1M >> gpurun_out/r2_step8.txt 2>&1
OGB_CHUNK_READS=4194304 $E --tag s4_4M >> gpurun_out/r2_step8.txt 2>&1
OGB_CHUNK_READS=4194304 OGB_SUB_PARTITIONS=24 $E --tag s4_4M_sub24 >> gpurun_out/r2_step8.txt 2>&1
OGB_CHUNK_READS=1048576 OGB_SUB_PARTITIONS=24 $E --tag s4_1M_sub24 >> gpurun_out/r2_step8.txt 2>&1
cat gpurun_out/r2_step8.txt
