# L2 persistence window on the bucket summary: off vs on, config 3 and config 3 x 2 (1.1 GB index, 69 MB summary) and config 2
mkdir -p gpurun_out
E="timeout 200 python profiles/exp.py --config 3 --steps 4 --warmup 2"
OGB_L2_PERSIST=0 $E --tag c3_off > gpurun_out/r2_step19.txt 2>&1
OGB_L2_PERSIST=1 $E --tag c3_on >> gpurun_out/r2_step19.txt 2>&1
E2="timeout 200 python profiles/exp.py --config 3 --scale 2.0 --steps 3 --warmup 1"
OGB_L2_PERSIST=0 $E2 --tag c3x2_off >> gpurun_out/r2_step19.txt 2>&1
OGB_L2_PERSIST=1 $E2 --tag c3x2_on >> gpurun_out/r2_step19.txt 2>&1
E3="timeout 200 python profiles/exp.py --config 2 --steps 4 --warmup 2"
OGB_L2_PERSIST=0 $E3 --tag c2_off >> gpurun_out/r2_step19.txt 2>&1
OGB_L2_PERSIST=1 $E3 --tag c2_on >> gpurun_out/r2_step19.txt 2>&1
grep "^\[c" gpurun_out/r2_step19.txt
