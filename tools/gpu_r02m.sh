#!/bin/bash
mkdir -p gpurun_out
cd /root/repo
L=gpurun_out/r02m.log
: > $L
step() { echo "=== $1" >> $L; shift; timeout "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
step "swag tests" 400 python -m pytest tests/test_gpu_swag.py tests/test_gpu_exact.py -q -x -s --timeout 200 -k "swag or SWAG"
echo "=== bench c4a" >> $L
timeout 300 python bench.py --no-cpu --only c4a --steps 1 --warmup 3 > gpurun_out/r02m_c4a.json 2>> $L
python -c "
import json
for f in ('gpurun_out/r02m_c4a.json',):
    d = json.load(open(f)); c = d['configs']
    print(f, {k: (round(v['value']/1e9, 3), round(v['ms'], 1), v.get('ok_frac', v.get('ok'))) for k, v in c.items()})
" >> $L 2>&1
grep -E "^===|rc=|passed|failed|SWAG|gpurun_out|Error|error" $L | tail -30
