#!/bin/bash
# round 2, second session: the whole GPU suite as the driver runs it, final sources
mkdir -p gpurun_out
timeout 900 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -8 > gpurun_out/r02au_tests.log
cat gpurun_out/r02au_tests.log
