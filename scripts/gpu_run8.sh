cd /root/repo
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x --timeout 300 -k "forward" 2>&1 | tail -1
timeout 300 python bench.py --graphs 512 --micro-graphs 256 --skip-extras --steps 3 --warmup 1 2>gpurun_out/bench_ws.err | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('bench', round(d['value']), round(d['ms_per_step'],1), {k:(round(v['avg_ms'],3), round(v['gbps'])) for k,v in d['kernels'].items()})"
