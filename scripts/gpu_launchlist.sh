# per-launch device times of one fwd+bwd step (cold-cache, serialised: compare SHARES, not absolutes)
cd /root/repo
CMD="python bench.py --graphs 512 --micro-graphs 512 --skip-extras --steps 1 --warmup 1"
$CMD > gpurun_out/plain_ll.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 600 --csv --log-file gpurun_out/launches_ws.csv $CMD > gpurun_out/ncu_ll.log 2>&1
tail -2 gpurun_out/ncu_ll.log | cut -c1-300
