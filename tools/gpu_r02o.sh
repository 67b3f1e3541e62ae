#!/bin/bash
mkdir -p gpurun_out
cd /root/repo
L=gpurun_out/r02o.log
: > $L
timeout 900 python -m pytest tests/test_gpu_wide.py -q -x --timeout 300 >> $L 2>&1
echo "rc=$?" >> $L
tail -40 $L
