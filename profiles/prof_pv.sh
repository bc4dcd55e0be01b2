mkdir -p gpurun_out
bash profiles/launches.sh 2>&1 | tail -16
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
OGB_CHUNK_READS=100000000 $CMD > gpurun_out/plain.log 2>&1 && OGB_CHUNK_READS=100000000 ncu --set full --clock-control none --import-source on -k 'regex:^(k_probe|k_verify|k_sort_nodes)$' -s 3 -c 3 -o gpurun_out/prof_pv $CMD > gpurun_out/ncu2.log 2>&1
tail -n 2 gpurun_out/ncu2.log
