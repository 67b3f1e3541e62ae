#!/bin/bash
mkdir -p gpurun_out
cd /root/repo
L=gpurun_out/r02e.log
: > $L
step() { echo "=== $1" >> $L; shift; timeout "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
step "fast tests" 300 python -m pytest tests/test_gpu_fast.py -x -q --timeout 120
for mb in 4 5 6; do
  step "bench Ts5 minb=$mb" 200 env XSQ_FAST_MINB=$mb tools/quick_bench.sh ts5_minb$mb --steps 3 --warmup 3 --no-extras --no-cpu
done
step "bench Ts5 nostiff" 200 tools/quick_bench.sh ts5_nostiff --steps 3 --warmup 3 --no-extras --no-cpu --stiff 0
step "bench CK5" 200 tools/quick_bench.sh ck5 --steps 3 --warmup 3 --no-extras --no-cpu --method CK5
step "exact tests" 600 python -m pytest tests/test_gpu_exact.py -q -x --timeout 240
grep -E "^===|rc=|passed|failed|steps/s|Error|error" $L | tail -40
