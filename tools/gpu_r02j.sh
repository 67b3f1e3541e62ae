#!/bin/bash
# full GPU test-suite + the default bench line + ncu captures for profiles/
mkdir -p gpurun_out
cd /root/repo
L=gpurun_out/r02j.log
: > $L
step() { echo "=== $1" >> $L; shift; timeout "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
( time timeout 1500 python -m pytest tests -m gpu -q --timeout 300 -x ) >> $L 2>&1
echo "=== full bench" >> $L
( time timeout 600 python bench.py ) > gpurun_out/r02j_bench_full.json 2>> $L
grep -E "^===|rc=|passed|failed|real|Error|error" $L | tail -30
tail -c 600 gpurun_out/r02j_bench_full.json
