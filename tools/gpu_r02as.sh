#!/bin/bash
# round 2, second session: SWAG events bit-identical to the C oracle
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_exact.py -m gpu -q -x -k "swag_events" 2>&1 | tail -12 > gpurun_out/r02as_tests.log
cat gpurun_out/r02as_tests.log
