#!/bin/bash
mkdir -p gpurun_out
cd /root/repo
L=gpurun_out/r02r.log
: > $L
timeout 300 python -m pytest tests/test_gpu_tma.py -q -x --timeout 120 >> $L 2>&1
echo "rc=$?" >> $L
timeout 300 python - >> $L 2>&1 <<'PY'
import ctypes as C
from extensisq_b200 import _lib
lib = _lib.load()
for nx, rows in ((16384, 2048), (16384, 8192), (16384, 16384)):
    a, b, d = C.c_double(), C.c_double(), C.c_double()
    for rep in range(2):
        r1 = lib.xsq_rkc_stage_bench(nx, rows, 40, C.byref(a), None)
        r2 = lib.xsq_rkc_stage_bench_tma(nx, rows, 40, C.byref(b), C.byref(d), None)
        gb = nx * rows * 40 / 1e9
        print(f"{rows}x{nx}: k_stage {a.value:.4f} ms = {gb/a.value*1e3:.0f} GB/s   k_stage_tma {b.value:.4f} ms = {gb/b.value*1e3:.0f} GB/s   diff {d.value}  rc {r1} {r2}")
PY
echo "rc=$?" >> $L
cat $L | tail -20
