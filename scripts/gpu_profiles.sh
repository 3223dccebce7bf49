# ncu evidence for profiles/: (1) per-launch times of one fwd+bwd step, (2) DRAM bytes of every edge kernel, (3) --set full
# of the two dominant kernels.  Run under gpurun; outputs land in gpurun_out/ and are summarised by scripts/ncu_summary.py.
#   gpurun --timeout 1500 -- 'bash scripts/gpu_profiles.sh r01b'
TAG=${1:-r01b}
cd /root/repo
CMD="python bench.py --graphs 512 --micro-graphs 512 --skip-extras --steps 1 --warmup 1"
$CMD > gpurun_out/plain_prof.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 600 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_ll.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --profile-from-start off -k regex:k_edge -c 40 --csv --log-file gpurun_out/edge_dram_$TAG.csv $CMD > gpurun_out/ncu_dram.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_edge_bwd_wg -s 1 -c 1 -o gpurun_out/prof_${TAG}_edge_bwd_wg -f $CMD > gpurun_out/ncu_full1.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_edge_fwd_wg -s 1 -c 1 -o gpurun_out/prof_${TAG}_edge_fwd_wg -f $CMD > gpurun_out/ncu_full2.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
