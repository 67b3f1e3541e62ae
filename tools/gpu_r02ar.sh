#!/bin/bash
# round 2, second session: the Nystrom methods bit-identical to the C oracle in device arithmetic
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_rkn.py -m gpu -q -x -k "bit_identical" 2>&1 | tail -15 > gpurun_out/r02ar_tests.log
cat gpurun_out/r02ar_tests.log
