#!/bin/bash
# round 2, second session: dense_output / event tests again, compute-sanitizer on the event paths
mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_gpu_reference_suite.py tests/test_gpu_events.py -m gpu -q -k "dense_output or queue or ensemble" 2>&1 | tail -8 > gpurun_out/r02ak_tests.log
SANITIZE_QUICK=1 timeout 240 compute-sanitizer --tool memcheck python tools/sanitize_smoke.py > gpurun_out/r02ak_memcheck.log 2>&1
cat gpurun_out/r02ak_tests.log; tail -8 gpurun_out/r02ak_memcheck.log
