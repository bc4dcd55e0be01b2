# Round 2: adjacency rows + candidate staging (K5/K6 rewrite), bounded K1: parity suite + timings
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r2_step1_tests.log 2>&1; tail -15 gpurun_out/r2_step1_tests.log
python profiles/exp.py --config 2 --steps 4 --warmup 2 --tag c2 > gpurun_out/r2_step1.txt 2>&1
python profiles/exp.py --config 3 --steps 4 --warmup 2 --tag c3 >> gpurun_out/r2_step1.txt 2>&1
python profiles/exp.py --config 5 --steps 4 --warmup 2 --tag c5 >> gpurun_out/r2_step1.txt 2>&1
python profiles/exp.py --config 4 --scale 0.2 --steps 3 --warmup 1 --tag c4_0.2 >> gpurun_out/r2_step1.txt 2>&1
cat gpurun_out/r2_step1.txt
