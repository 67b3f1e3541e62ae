#!/bin/bash
# round 2, second session: no-terminal build of the fast event kernel
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_events.py -m gpu -q -x -k "queue or ensemble or early or reference_event" 2>&1 | tail -6 > gpurun_out/r02an_tests.log
timeout 300 python tools/bench_events.py > gpurun_out/r02an_bench_events.json 2> gpurun_out/r02an_bench_events.err
cat gpurun_out/r02an_tests.log; cut -c1-420 gpurun_out/r02an_bench_events.json; tail -3 gpurun_out/r02an_bench_events.err
