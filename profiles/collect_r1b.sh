# Round-1 (second half) evidence on one B200: full GPU test suite, bench lines (ours + reference arm), ncu launch
# list of the bench command, one full ncu capture of the hot kernels. Run on the GPU box: bash profiles/collect_r1b.sh
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu --durations=5 > gpurun_out/gpu_tests_r1b.log 2>&1; tail -9 gpurun_out/gpu_tests_r1b.log
python bench.py > gpurun_out/bench_r1b.json 2> gpurun_out/bench_r1b.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r1b_reference.json 2>> gpurun_out/bench_r1b.err
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1b.csv $CMD > gpurun_out/ncu_launches.log 2>&1
CMD2="python profiles/exp.py --config 2 --steps 1 --warmup 0"
$CMD2 > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k 'regex:^(k_window_part|k_probe_parts|k_verify|k_hash_insert)' -s 16 -c 7 -o gpurun_out/prof_r1b_scan $CMD2 > gpurun_out/ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:^(k_mark|k_keep|k_emit)' -s 4 -c 4 -o gpurun_out/prof_r1b_reduce $CMD2 >> gpurun_out/ncu_full.log 2>&1
tail -n 2 gpurun_out/ncu_full.log
head -c 1500 gpurun_out/bench_r1b.json
