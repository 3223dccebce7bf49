# quick A/B on the B200: parity tests selected by $1, then a short fwd+bwd bench line per kernel kind
[ -n "$1" ] && timeout 400 python -m pytest tests/test_kernels_gpu.py -q -x --timeout 120 -k "$1" 2>&1 | tail -3
run() { timeout 300 python bench.py --graphs 1024 --micro-graphs 512 --skip-extras --steps 2 --warmup 1 --math $1 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('$1 $2', round(d['value']), round(d['ms_per_step'],1), {k:(round(v['avg_ms'],3), round(v['gbps'])) for k,v in d['kernels'].items()})"; }
run bf16x3 "$QB_TAG"
