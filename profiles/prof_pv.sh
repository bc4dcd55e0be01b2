mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k 'regex:^(k_probe_uniform|k_verify|k_hash_insert)$' -s 60 -c 3 -o gpurun_out/prof_pv2 $CMD > gpurun_out/ncu2.log 2>&1
tail -n 2 gpurun_out/ncu2.log
