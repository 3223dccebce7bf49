cd /root/repo
for pf in 0 1 2 3; do
GCRL_WS_PF=$pf timeout 300 python bench.py --graphs 512 --micro-graphs 256 --skip-extras --steps 3 --warmup 1 2>gpurun_out/bench_ws.err | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('pf$pf', round(d['value']), round(d['ms_per_step'],1), {k:(round(v['avg_ms'],3), round(v['gbps'])) for k,v in d['kernels'].items() if 'bwd' in k})"
done
