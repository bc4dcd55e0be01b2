# Round-1 evidence: bench lines (ours + reference arm), ncu launch list of the bench command, one full
# ncu capture of the hot kernels. Run on the GPU box: bash profiles/collect_r1.sh
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r1_reference.json 2>> gpurun_out/bench_r1.err
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1.csv $CMD > gpurun_out/ncu_launches.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k 'regex:^(k_probe_uniform|k_verify|k_sort_nodes|k_mark|k_twin_keep|k_hash_insert)$' -s 120 -c 12 -o gpurun_out/prof_r1_final $CMD > gpurun_out/ncu_full.log 2>&1
tail -n 2 gpurun_out/ncu_full.log
cat gpurun_out/bench_r1.json | head -c 3000
