#!/bin/bash
# round 2, second session: dense_output tests again, compute-sanitizer on the event paths
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_reference_suite.py tests/test_gpu_events.py -m gpu -q 2>&1 | tail -8 > gpurun_out/r02ak_tests.log
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_smoke.py > gpurun_out/r02ak_memcheck.log 2>&1
timeout 600 compute-sanitizer --tool racecheck python tools/sanitize_smoke.py > gpurun_out/r02ak_racecheck.log 2>&1
cat gpurun_out/r02ak_tests.log; tail -6 gpurun_out/r02ak_memcheck.log; tail -4 gpurun_out/r02ak_racecheck.log
