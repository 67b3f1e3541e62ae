#!/bin/bash
# round 2, second session: 32-bit CTA counter, pool threshold, events in the fast kernel (timing)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_events.py -q -x -k "queue or ensemble" 2>&1 | tail -8 > gpurun_out/r02ah_events.log
timeout 600 python tools/bench_events.py > gpurun_out/r02ah_bench_events.json 2> gpurun_out/r02ah_bench_events.err
XSQ_NO_FAST=1 timeout 300 python tools/bench_events.py 2>&1 | grep "event queue\"" > gpurun_out/r02ah_bench_events_nofast.json
cat gpurun_out/r02ah_events.log gpurun_out/r02ah_bench_events.json gpurun_out/r02ah_bench_events_nofast.json; tail -5 gpurun_out/r02ah_bench_events.err
