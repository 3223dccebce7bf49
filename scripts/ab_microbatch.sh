for mb in 512 256 1024 512; do
timeout 300 python bench.py --skip-extras --steps 6 --warmup 3 --micro-graphs $mb 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('mb $mb', round(d['value']), d['clocks']['sm_mhz'], {k:round(v['avg_ms'],3) for k,v in d['kernels'].items()})"
done
