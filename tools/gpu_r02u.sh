#!/bin/bash
# full GPU suite + default bench (what the driver runs at round end)
mkdir -p gpurun_out
cd /root/repo
L=gpurun_out/r02u.log
: > $L
timeout 1500 python -m pytest tests/ -x -q -m gpu --timeout 600 >> $L 2>&1
echo "rc=$?" >> $L
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" >> $L 2>&1
echo "smoke rc=$?" >> $L
timeout 600 python bench.py > gpurun_out/r02u_bench.json 2>> $L
echo "bench rc=$?" >> $L
grep -E "passed|failed|^FAILED|rc=|Error" $L | tail -20
