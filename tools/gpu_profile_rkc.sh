#!/bin/bash
# ncu capture of the SSV2stab stage kernel (run under gpurun)
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_stage -s 5 -c 1 -f \
    -o gpurun_out/prof_rkc_${1:-r01} python tools/rkc_bw.py > gpurun_out/ncu_rkc.log 2>&1
ls -la gpurun_out | tail -3
