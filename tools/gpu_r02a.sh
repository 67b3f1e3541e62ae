#!/bin/bash
# round 2, call A: correctness of rk_fast + first timing + ncu source profile
mkdir -p gpurun_out
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_fast.py -x -q 2>&1 | tail -15 > gpurun_out/r02a_fast_tests.log
timeout 1200 python -m pytest tests/test_gpu_rk.py -q 2>&1 | tail -25 > gpurun_out/r02a_rk_tests.log
tools/quick_bench.sh fast --steps 3 --warmup 3 > gpurun_out/r02a_bench.log 2>&1
XSQ_NO_FAST=1 tools/quick_bench.sh generic --steps 3 --warmup 3 >> gpurun_out/r02a_bench.log 2>&1
tools/quick_bench.sh fast_nostiff --steps 3 --warmup 3 --stiff 0 >> gpurun_out/r02a_bench.log 2>&1
tools/quick_bench.sh fast_ck5 --steps 3 --warmup 3 --method CK5 >> gpurun_out/r02a_bench.log 2>&1
BENCH="python bench.py --lanes 1250000 --t-end 100 --steps 1 --warmup 3 --no-cpu"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rk_fast -s 3 -c 1 \
    -f -o gpurun_out/prof_r02a $BENCH > gpurun_out/ncu_full_r02a.log 2>&1
cat gpurun_out/r02a_fast_tests.log gpurun_out/r02a_rk_tests.log gpurun_out/r02a_bench.log
