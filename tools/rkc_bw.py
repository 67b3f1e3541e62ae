"""GPU box: bandwidth of the fused RKC stage kernel on a slab (roofline)."""
import ctypes as C, sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from extensisq_b200 import _lib
lib = _lib.load()
torch.cuda.init(); torch.zeros(1, device="cuda")
for nx, rows in ((16384, 2048), (16384, 4096), (8192, 8192), (4096, 4096)):
    ms = C.c_double()
    rc = lib.xsq_rkc_stage_bench(nx, rows, 50, C.byref(ms), None)
    gb = nx * rows * 40 / 1e9
    print(json.dumps(dict(nx=nx, rows=rows, rc=rc, ms_per_stage=ms.value, algorithmic_GBps=gb / (ms.value * 1e-3))))
