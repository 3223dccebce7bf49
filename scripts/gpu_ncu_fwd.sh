cd /root/repo
CMD="python bench.py --graphs 256 --micro-graphs 256 --skip-extras --steps 1 --warmup 1"
$CMD > gpurun_out/plain_fws.log 2>&1 && ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_edge_fwd_ws -s 2 -c 1 -o gpurun_out/prof_edge_fwd_ws -f $CMD > gpurun_out/ncu_fws.log 2>&1
tail -2 gpurun_out/ncu_fws.log | cut -c1-200
