# L2 persistence window on the bucket summary, window open for K2 and K3 only, persisting lines reset when it closes: off vs on
mkdir -p gpurun_out
E="timeout 200 python profiles/exp.py --config 3 --steps 4 --warmup 2"
OGB_L2_PERSIST=0 $E --tag c3_off > gpurun_out/r2_step22.txt 2>&1
OGB_L2_PERSIST=1 $E --tag c3_on >> gpurun_out/r2_step22.txt 2>&1
E2="timeout 200 python profiles/exp.py --config 3 --scale 2.0 --steps 3 --warmup 1"
OGB_L2_PERSIST=0 $E2 --tag c3x2_off >> gpurun_out/r2_step22.txt 2>&1
OGB_L2_PERSIST=1 $E2 --tag c3x2_on >> gpurun_out/r2_step22.txt 2>&1
E4="timeout 300 python profiles/exp.py --config 4 --scale 0.2 --steps 3 --warmup 1"
OGB_L2_PERSIST=0 $E4 --tag c4_off >> gpurun_out/r2_step22.txt 2>&1
OGB_L2_PERSIST=1 $E4 --tag c4_on >> gpurun_out/r2_step22.txt 2>&1
E5="timeout 200 python profiles/exp.py --config 5 --steps 4 --warmup 2"
OGB_L2_PERSIST=0 $E5 --tag c5_off >> gpurun_out/r2_step22.txt 2>&1
OGB_L2_PERSIST=1 $E5 --tag c5_on >> gpurun_out/r2_step22.txt 2>&1
python -c "import torch; p=torch.cuda.get_device_properties(0); print('L2', p.L2_cache_size)" >> gpurun_out/r2_step22.txt 2>&1
grep "^\[c\|^L2" gpurun_out/r2_step22.txt
