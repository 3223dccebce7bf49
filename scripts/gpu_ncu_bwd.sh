set -x
cd /root/repo
CMD="python bench.py --graphs 256 --micro-graphs 256 --skip-extras --steps 1 --warmup 1"
$CMD > gpurun_out/plain_ws.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_edge_bwd_ws -s 6 -c 1 -o gpurun_out/prof_edge_bwd_ws -f $CMD > gpurun_out/ncu_ws.log 2>&1
tail -3 gpurun_out/ncu_ws.log
