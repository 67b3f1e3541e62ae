#!/bin/bash
mkdir -p gpurun_out
cd /root/repo
L=gpurun_out/r02x.log
: > $L
for occ in 1 2 3; do
  echo "=== occ $occ" >> $L
  XSQ_LIB=/root/repo/extensisq_b200/libxsq_swagsweep.so XSQ_SWAG_OCC=$occ timeout 300 python bench.py --no-cpu --only c4a --steps 1 --warmup 2 > gpurun_out/r02x_$occ.json 2>> $L
  python -c "
import json
d = json.load(open('gpurun_out/r02x_$occ.json')); c = d['configs']
print('occ $occ', {k: (round(v['value']/1e9, 3), round(v['ms'], 1)) for k, v in c.items() if 'SWAG' in k})
" >> $L 2>&1
done
cat $L | tail -12
