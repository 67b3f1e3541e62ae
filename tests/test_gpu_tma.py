"""GPU: the TMA-staged variant of the SSV2stab stage kernel (xsq_rkc_tma.cuh,
cp.async.bulk.tensor.2d) computes bit for bit what the shipped kernel computes,
also on slabs whose size is not a multiple of the tile."""
import ctypes as C

import pytest

from extensisq_b200 import _lib

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("nx,rows", [(512, 64), (1000, 37), (4096, 2048), (132, 9)])
def test_tma_stage_is_bit_identical(nx, rows):
    lib = _lib.load()
    ms, diff = C.c_double(), C.c_double(-1.0)
    rc = lib.xsq_rkc_stage_bench_tma(nx, rows, 3, C.byref(ms), C.byref(diff), None)
    assert rc == 0, lib.xsq_last_error_detail()
    assert diff.value == 0.0
    assert ms.value > 0.0
