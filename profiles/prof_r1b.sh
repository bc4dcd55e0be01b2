mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k 'regex:^(k_scan|k_mark|k_twin_keep|k_compact|k_hash_insert)$' -s 5 -c 5 -o gpurun_out/prof_r1b $CMD > gpurun_out/ncu2.log 2>&1
tail -n 3 gpurun_out/ncu2.log
