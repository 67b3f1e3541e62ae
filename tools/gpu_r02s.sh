#!/bin/bash
mkdir -p gpurun_out
cd /root/repo
L=gpurun_out/r02s.log
: > $L
cat > /tmp/t.py <<'PY'
import ctypes as C
from extensisq_b200 import _lib
lib = _lib.load()
ms, d = C.c_double(), C.c_double()
print(lib.xsq_rkc_stage_bench_tma(512, 64, 1, C.byref(ms), C.byref(d), None), ms.value, d.value)
PY
PYTHONPATH=/root/repo timeout 300 compute-sanitizer --tool memcheck python /tmp/t.py >> $L 2>&1
echo "rc=$?" >> $L
grep -v "^$" $L | head -60
