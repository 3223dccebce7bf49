timeout 400 python -m pytest tests/test_kernels_gpu.py -q -x --timeout 200 -k "tensor_core" 2>&1 | tail -15
for m in bf16x3 bf16; do timeout 250 python bench.py --graphs 256 --micro-graphs 128 --skip-extras --steps 2 --warmup 1 --math $m 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('$m', round(d['value']), round(d['ms_per_step'],1), {k:(round(v['avg_ms'],3), round(v['gbps'])) for k,v in d['kernels'].items()})"; done
