mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 || exit 1
for G in 32 64 128; do
  OGB_L2_FETCH=$G ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum,lts__t_sectors_op_read.sum,lts__t_sectors_op_read_lookup_miss.sum,lts__t_sectors_srcunit_tex_op_read.sum,dram__sectors_read.sum \
     --clock-control none -k 'regex:^(k_scan|k_mark)$' -s 2 -c 2 --csv --log-file gpurun_out/l2fetch_$G.csv $CMD > /dev/null 2>&1
  echo "== OGB_L2_FETCH=$G"; grep -E "k_scan|k_mark" gpurun_out/l2fetch_$G.csv | awk -F'","' '{print $5, $(NF-2), $(NF)}' | tr -d '"'
done
