mkdir -p gpurun_out
E="timeout 200 python profiles/exp.py --config 3 --steps 3 --warmup 1"
for sub in 4 6 8 12 16; do OGB_SUB_PARTITIONS=$sub $E --tag c3_sub$sub >> gpurun_out/r2_step11.txt 2>&1; done
E2="timeout 300 python profiles/exp.py --config 3 --scale 2.0 --steps 2 --warmup 1"
for sub in 8 12 16 24; do OGB_SUB_PARTITIONS=$sub $E2 --tag c3s2_sub$sub >> gpurun_out/r2_step11.txt 2>&1; done
cat gpurun_out/r2_step11.txt
