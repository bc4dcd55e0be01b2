# round 1: probe (pending-window queue) + verify, one-stream timing and full ncu capture
mkdir -p gpurun_out
OGB_ONE_STREAM=1 python profiles/exp.py --config 2 --steps 4 --warmup 2 --tag onestream > gpurun_out/exp_d.log 2>&1
cat gpurun_out/exp_d.log
ncu --set full --clock-control none --import-source on -k 'regex:^(k_probe_uniform|k_verify)' -s 44 -c 2 -o gpurun_out/prof_pv3 python profiles/exp.py --config 2 --steps 1 --warmup 0 > gpurun_out/ncu_pv3.log 2>&1
tail -n 2 gpurun_out/ncu_pv3.log
