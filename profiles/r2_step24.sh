# with the summary persisting in L2: chunk size, sub-partitions and load factor once more (config 3)
mkdir -p gpurun_out
E="timeout 200 python profiles/exp.py --config 3 --steps 3 --warmup 2"
$E --tag base > gpurun_out/r2_step24.txt 2>&1
OGB_CHUNK_READS=262144 $E --tag chunk256k >> gpurun_out/r2_step24.txt 2>&1
OGB_CHUNK_READS=524288 $E --tag chunk512k >> gpurun_out/r2_step24.txt 2>&1
OGB_CHUNK_READS=786432 $E --tag chunk768k >> gpurun_out/r2_step24.txt 2>&1
OGB_SUB_PARTITIONS=2 $E --tag sub2 >> gpurun_out/r2_step24.txt 2>&1
OGB_SUB_PARTITIONS=8 $E --tag sub8 >> gpurun_out/r2_step24.txt 2>&1
OGB_TABLE_BUCKETS_PER_READ=0.75 $E --tag load053 >> gpurun_out/r2_step24.txt 2>&1
OGB_TABLE_BUCKETS_PER_READ=1.25 $E --tag load032 >> gpurun_out/r2_step24.txt 2>&1
grep "^\[" gpurun_out/r2_step24.txt
