#!/bin/bash
mkdir -p gpurun_out
cd /root/repo
L=gpurun_out/r02w.log
: > $L
timeout 600 python -m pytest tests/test_gpu_rkc.py -q --timeout 300 -s -k "combustion or notebook" >> $L 2>&1
echo "rc=$?" >> $L
grep -E "passed|failed|^FAILED|^E  |combustion tol|rc=" $L | head -30
