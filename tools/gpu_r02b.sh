#!/bin/bash
# round 2, call B: SAFE first contact of rk_fast with the GPU -- every step under
# its own short timeout, tiny sizes first.
mkdir -p gpurun_out
cd /root/repo
L=gpurun_out/r02b.log
: > $L
step() { echo "=== $1" >> $L; shift; timeout "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
step "smoke" 120 python -c "import __graft_entry__ as g; g.smoke()"
step "fast tests" 420 python -m pytest tests/test_gpu_fast.py -x -q --timeout 120
step "exact tests" 600 python -m pytest tests/test_gpu_exact.py -x -q -s --timeout 240
tail -5 $L
step "bench fast" 240 tools/quick_bench.sh fast --steps 3 --warmup 3 --no-extras
step "bench generic" 240 env XSQ_NO_FAST=1 tools/quick_bench.sh generic --steps 3 --warmup 3 --no-extras
step "bench nostiff" 240 tools/quick_bench.sh fast_nostiff --steps 3 --warmup 3 --stiff 0 --no-extras
BENCH="python bench.py --lanes 1250000 --t-end 100 --steps 1 --warmup 3 --no-cpu --no-extras"
step "ncu" 420 ncu --set full --clock-control none --import-source on -k regex:rk_fast -s 3 -c 1 -f -o gpurun_out/prof_r02b $BENCH
grep -E "^===|rc=|passed|failed|steps/s|Error|error" $L | tail -40
