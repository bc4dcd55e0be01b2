mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q -k "tandem" 2>&1 | grep -E "AssertionError|passed|failed" | cut -c1-600
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k 'regex:^(k_scan|k_mark)$' -s 2 -c 2 -o gpurun_out/prof_scan $CMD > gpurun_out/ncu2.log 2>&1
tail -n 2 gpurun_out/ncu2.log
