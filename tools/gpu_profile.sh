#!/bin/bash
# Runs on the GPU box (under gpurun): launch list + one full ncu capture of the
# dominant kernel for the bench command.  Outputs under gpurun_out/.
set -x
mkdir -p gpurun_out
TAG=${1:-r01}
BENCH="python bench.py --lanes ${LANES:-300000} --t-end ${TEND:-5} --steps 2 --warmup 3 --no-cpu"
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv \
    --log-file gpurun_out/launches_${TAG}.csv $BENCH > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rk_persistent -s 3 -c 1 \
    -f -o gpurun_out/prof_${TAG} $BENCH > gpurun_out/ncu_full_${TAG}.log 2>&1
ls -la gpurun_out
# the probe-queue kernel of the same command (stiffness diagnosis on by default)
ncu --set full --clock-control none --import-source on -k regex:stiff_queue -s 3 -c 1 \
    -f -o gpurun_out/prof_${TAG}_queue $BENCH > gpurun_out/ncu_full_${TAG}_queue.log 2>&1
