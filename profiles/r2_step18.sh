mkdir -p gpurun_out
E="timeout 200 python profiles/exp.py --config 3 --scale 4.0 --steps 2 --warmup 1"
$E --tag s4_wpt4 > gpurun_out/r2_step18.txt 2>&1
OGB_LIB=$PWD/profiles/variants/libogb_wpt8.so $E --tag s4_wpt8 >> gpurun_out/r2_step18.txt 2>&1
OGB_LIB=$PWD/profiles/variants/libogb_wpt16.so $E --tag s4_wpt16 >> gpurun_out/r2_step18.txt 2>&1
E1="timeout 200 python profiles/exp.py --config 3 --steps 3 --warmup 1"
OGB_LIB=$PWD/profiles/variants/libogb_wpt8.so $E1 --tag s1_wpt8 >> gpurun_out/r2_step18.txt 2>&1
grep "^\[s" gpurun_out/r2_step18.txt
