#!/bin/bash
mkdir -p gpurun_out
cd /root/repo
L=gpurun_out/r02q.log
: > $L
timeout 900 python -m pytest tests/test_gpu_rkn.py -q --timeout 300 -s >> $L 2>&1
echo "rc=$?" >> $L
grep -E "passed|failed|^FAILED|^E  |identical|rc=" $L | head -60
