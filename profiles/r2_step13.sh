mkdir -p gpurun_out
E="timeout 200 python profiles/exp.py --config 3 --steps 3 --warmup 1"
$E --tag c3_default >> gpurun_out/r2_step13.txt 2>&1
for ch in 131072 196608 393216 524288; do OGB_CHUNK_READS=$ch $E --tag c3_chunk$ch >> gpurun_out/r2_step13.txt 2>&1; done
for b in 0.75 1.25 1.5; do OGB_TABLE_BUCKETS_PER_READ=$b $E --tag c3_buckets$b >> gpurun_out/r2_step13.txt 2>&1; done
timeout 200 python profiles/exp.py --config 2 --steps 3 --warmup 1 --tag c2 >> gpurun_out/r2_step13.txt 2>&1
timeout 200 python profiles/exp.py --config 5 --steps 3 --warmup 1 --tag c5 >> gpurun_out/r2_step13.txt 2>&1
cat gpurun_out/r2_step13.txt
