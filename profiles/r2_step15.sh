mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r2_step15_tests.log 2>&1; tail -3 gpurun_out/r2_step15_tests.log
E3="timeout 200 python profiles/exp.py --config 3 --steps 3 --warmup 1"
$E3 --tag c3_rows_xs > gpurun_out/r2_step15.txt 2>&1
OGB_LIB=$PWD/profiles/variants/libogb_t1_12.so $E3 --tag c3_t1_4096 >> gpurun_out/r2_step15.txt 2>&1
timeout 200 python profiles/exp.py --config 2 --steps 3 --warmup 1 --tag c2 >> gpurun_out/r2_step15.txt 2>&1
OGB_LIB=$PWD/profiles/variants/libogb_t1_12.so timeout 200 python profiles/exp.py --config 2 --steps 3 --warmup 1 --tag c2_t1_4096 >> gpurun_out/r2_step15.txt 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log | cut -c1-200
cat gpurun_out/r2_step15.txt
