#!/bin/bash
mkdir -p gpurun_out
cd /root/repo
L=gpurun_out/r02v.log
: > $L
timeout 900 python -m pytest tests/test_gpu_rkc.py tests/test_gpu_sens.py tests/test_gpu_swag.py tests/test_gpu_tma.py tests/test_gpu_wide.py tests/test_gpu_rkn.py -q --timeout 300 >> $L 2>&1
echo "rc=$?" >> $L
grep -E "passed|failed|^FAILED|^E  |rc=" $L | head -40
