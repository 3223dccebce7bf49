mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_bf16x3.json 2> gpurun_out/bench_bf16x3.err; tail -2 gpurun_out/bench_bf16x3.err
timeout 600 python bench.py --math bf16 --skip-extras > gpurun_out/bench_bf16.json 2>/dev/null
timeout 600 python bench.py --impl reference --steps 3 > gpurun_out/bench_reference.json 2>/dev/null
for f in bf16x3 bf16 reference; do python - <<PY
import json
d=json.load(open('gpurun_out/bench_$f.json'))
print('$f', d.get('value'), d.get('ms_per_step'), d.get('e2e'), d.get('roofline',{}).get('frac'), d.get('rollouts'), d.get('learn_step'), d.get('pna',{}).get('frac'), d.get('cpu_baseline'), d.get('clocks'))
PY
done
