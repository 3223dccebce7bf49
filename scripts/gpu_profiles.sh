# round-1 evidence: (1) per-launch times of one step, (2) DRAM bytes of every edge kernel, (3) --set full of the two
# dominant kernels.  Run under gpurun; outputs land in gpurun_out/ and are summarised by scripts/ncu_summary.py.
cd /root/repo
CMD="python bench.py --graphs 512 --micro-graphs 512 --skip-extras --steps 1 --warmup 1"
$CMD > gpurun_out/plain_prof.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 600 --csv --log-file gpurun_out/launches_r01.csv $CMD > gpurun_out/ncu_ll.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --profile-from-start off -k regex:k_edge -c 40 --csv --log-file gpurun_out/edge_dram_r01.csv $CMD > gpurun_out/ncu_dram.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_edge_bwd_ws -s 1 -c 1 -o gpurun_out/prof_r01_edge_bwd_ws -f $CMD > gpurun_out/ncu_full1.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_edge_fwd_tc -s 1 -c 1 -o gpurun_out/prof_r01_edge_fwd_tc -f $CMD > gpurun_out/ncu_full2.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
