#!/bin/bash
# compute-sanitizer over the hot path.  Run on the B200:
#   gpurun --timeout 1500 -- bash scripts/sanitize.sh
# Writes gpurun_out/r02_sanitize_<tool>.log (full tool output) and gpurun_out/r02_sanitize_summary.txt.
# The kernels reuse one 32 KB shared-memory tile as TMA landing tile -> MMA operand -> staging tile under named barriers
# and mbarriers: racecheck looks at exactly that (it does not model async-proxy / tcgen05 accesses, so its verdict
# covers the generic-proxy reads and writes of those tiles); memcheck covers global / shared bounds, synccheck the
# barrier usage, initcheck reads of uninitialised global memory.
mkdir -p gpurun_out
SUM=gpurun_out/r02_sanitize_summary.txt
: > $SUM
run() {  # tool, modes, big-graph size, per-run timeout
  local tool=$1 modes=$2 big=$3 to=$4
  local log=gpurun_out/r02_sanitize_${tool}.log
  timeout $to compute-sanitizer --tool $tool --print-limit 20 --launch-timeout 120 --kernel-name kns=gcrl \
      --log-file $log python scripts/sanitize_target.py $modes $big > gpurun_out/r02_sanitize_${tool}.out 2>&1
  local rc=$?
  echo "== $tool (modes $modes, big graph $big vertices): exit $rc" >> $SUM
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Internal|Error:" $log | sort | uniq -c | head -20 >> $SUM
  grep "sanitize target" gpurun_out/r02_sanitize_${tool}.out >> $SUM
}
run memcheck  fp32,bf16x3,bf16 200 500
run synccheck fp32,bf16x3,bf16 200 500
run racecheck fp32,bf16x3,bf16 64  700
run initcheck bf16x3 64 400
cat $SUM
