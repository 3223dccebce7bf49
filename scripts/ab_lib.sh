#!/bin/bash
# A/B of two builds of the library on ONE box: bash scripts/ab_lib.sh <alt.so> [tag]   (default build vs the alternative, alternating)
ALT=$1; TAG=${2:-alt}
run() { GCRL_LIB=$1 timeout 300 python bench.py --skip-extras --steps 8 --warmup 3 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('$2', round(d['value']), d['clocks']['sm_mhz'], {k:round(v['avg_ms'],3) for k,v in d['kernels'].items()})"; }
run "" default; run $ALT $TAG; run "" default; run $ALT $TAG
