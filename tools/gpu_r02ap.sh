#!/bin/bash
# round 2, second session: occupancy of the event queue kernel
mkdir -p gpurun_out
EVQ_MINB_SWEEP=5,6,8 timeout 400 python tools/bench_events.py > gpurun_out/r02ap_bench_events.json 2> gpurun_out/r02ap_bench_events.err
cut -c1-100 gpurun_out/r02ap_bench_events.json; grep -o '"kernels": {[^}]*}' gpurun_out/r02ap_bench_events.json; tail -3 gpurun_out/r02ap_bench_events.err
