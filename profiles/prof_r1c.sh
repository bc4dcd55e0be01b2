# round 1, fused reduction pipeline: slot-cap sensitivity + full ncu capture of K5/K6
mkdir -p gpurun_out
for C in 48 64 96; do OGB_SLOT_CAP=$C python profiles/exp.py --config 2 --steps 4 --warmup 2 --tag cap$C; done > gpurun_out/exp_cap.log 2>&1
cat gpurun_out/exp_cap.log
ncu --set full --clock-control none --import-source on -k 'regex:^(k_mark|k_keep|k_emit)' -s 3 -c 3 -o gpurun_out/prof_k56 python profiles/exp.py --config 2 --steps 1 --warmup 0 > gpurun_out/ncu_k56.log 2>&1
tail -n 2 gpurun_out/ncu_k56.log
