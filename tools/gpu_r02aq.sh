#!/bin/bash
# round 2, second session: events on the device bit-identical to the C oracle with events
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_exact.py -m gpu -q -x -k "events_bit_identical" 2>&1 | tail -12 > gpurun_out/r02aq_tests.log
cat gpurun_out/r02aq_tests.log
