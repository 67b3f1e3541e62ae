#!/bin/bash
# round 2, second session: event queue with shared-memory counter per CTA, AoS records
mkdir -p gpurun_out
python -m pytest tests/test_gpu_events.py -q -x -k "queue or ensemble" 2>&1 | tail -8 > gpurun_out/r02af_events.log
MINB_SWEEP=3,2 timeout 600 python tools/bench_events.py > gpurun_out/r02af_bench_events.json 2> gpurun_out/r02af_bench_events.err
cat gpurun_out/r02af_events.log gpurun_out/r02af_bench_events.json; tail -5 gpurun_out/r02af_bench_events.err
