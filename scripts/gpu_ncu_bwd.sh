set -x
cd /root/repo
CMD="python bench.py --graphs 256 --micro-graphs 256 --skip-extras --steps 1 --warmup 1"
for pf in 0 2; do
export GCRL_WS_PF=$pf
$CMD > gpurun_out/plain_ws.log 2>&1 && ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_red_lookup_miss.sum,lts__t_sectors_srcunit_tex_op_red.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum,lts__t_sectors_srcunit_tex_op_read.sum --clock-control none --profile-from-start off -k regex:k_edge_bwd_ws -s 1 -c 1 --csv --log-file gpurun_out/dram_pf$pf.csv $CMD > gpurun_out/ncu_ws.log 2>&1
done
