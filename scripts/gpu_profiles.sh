# ncu evidence for profiles/: (1) per-launch times of one fwd+bwd step, (2) DRAM bytes of every edge kernel, (3) --set full
# of the two dominant kernels.  Run under gpurun; outputs land in gpurun_out/ and are summarised by scripts/ncu_summary.py.
#   gpurun --timeout 1500 -- 'bash scripts/gpu_profiles.sh r01b'
TAG=${1:-r02}
cd ${GRAFT_REPO_ROOT:-/root/repo}; mkdir -p gpurun_out
CMD="python bench.py --graphs 512 --micro-graphs 512 --skip-extras --steps 1 --warmup 1"
$CMD > gpurun_out/plain_prof.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 600 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_ll.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --profile-from-start off -k regex:k_edge -c 40 --csv --log-file gpurun_out/edge_dram_$TAG.csv $CMD > gpurun_out/ncu_dram.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_edge_bwd_wg -s 1 -c 1 -o gpurun_out/prof_${TAG}_edge_bwd_wg -f $CMD > gpurun_out/ncu_full1.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_edge_fwd_wg -s 1 -c 1 -o gpurun_out/prof_${TAG}_edge_fwd_wg -f $CMD > gpurun_out/ncu_full2.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
# (4) launch lists of the small-batch paths: a 64-transition learn step and three rollout steps of 10 000 x 50-vertex graphs
NCU=1 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 1500 --csv --log-file gpurun_out/launches_learn64_$TAG.csv python scripts/learn_step_small_launches.py > gpurun_out/ncu_learn64.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file gpurun_out/launches_roll_$TAG.csv python scripts/rollout_launches.py > gpurun_out/ncu_roll.log 2>&1
python scripts/ncu_summary.py launches gpurun_out/launches_$TAG.csv gpurun_out/${TAG}_launches_fwd_bwd_step_512x200.json
python scripts/ncu_summary.py launches gpurun_out/launches_learn64_$TAG.csv gpurun_out/${TAG}_launches_learn_step_64x15-50.json
python scripts/ncu_summary.py launches gpurun_out/launches_roll_$TAG.csv gpurun_out/${TAG}_launches_rollout_3steps_10000x50.json
python scripts/ncu_summary.py full gpurun_out/prof_${TAG}_edge_bwd_wg.ncu-rep gpurun_out/${TAG}_ncu_full_k_edge_bwd_wg.json 20377600
python scripts/ncu_summary.py full gpurun_out/prof_${TAG}_edge_fwd_wg.ncu-rep gpurun_out/${TAG}_ncu_full_k_edge_fwd_wg.json 20377600
# (5) --set full of the fused vertex-side kernels (round 2) inside the same 512 x 200-vertex step
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_vertex_fwd_tc -s 1 -c 1 -o gpurun_out/prof_${TAG}_vertex_fwd -f $CMD > gpurun_out/ncu_full3.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_linear_bwd_tc -s 3 -c 1 -o gpurun_out/prof_${TAG}_linear_bwd -f $CMD > gpurun_out/ncu_full4.log 2>&1
python scripts/ncu_summary.py full gpurun_out/prof_${TAG}_vertex_fwd.ncu-rep gpurun_out/${TAG}_ncu_full_k_vertex_fwd_tc.json
python scripts/ncu_summary.py full gpurun_out/prof_${TAG}_linear_bwd.ncu-rep gpurun_out/${TAG}_ncu_full_k_linear_bwd_tc.json
