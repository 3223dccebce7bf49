cd /root/repo
for dbg in 0 0 2; do
GCRL_FW_DBG=$dbg timeout 300 python bench.py --graphs 512 --micro-graphs 256 --skip-extras --steps 2 --warmup 1 2>gpurun_out/bench_ws_$dbg.err | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('fwdbg$dbg', round(d['value']), round(d['ms_per_step'],1), {k:(round(v['avg_ms'],3), round(v['gbps'])) for k,v in d['kernels'].items() if 'fwd' in k})"
grep -E "gcrl|rror" gpurun_out/bench_ws_$dbg.err | head -3
done
