#!/bin/bash
# round 2, second session: CKdisc exact parity, C4 collision lanes, event queue
mkdir -p gpurun_out
python -m pytest tests/test_gpu_exact.py -q -x -k "ckdisc or c4" -s 2>&1 | tail -15 > gpurun_out/r02ab_exact.log
python -m pytest tests/test_gpu_events.py tests/test_gpu_ckdisc.py -q -x 2>&1 | tail -15 > gpurun_out/r02ab_events.log
timeout 600 python tools/bench_events.py > gpurun_out/r02ab_bench_events.json 2> gpurun_out/r02ab_bench_events.err
cat gpurun_out/r02ab_exact.log gpurun_out/r02ab_events.log gpurun_out/r02ab_bench_events.json; tail -5 gpurun_out/r02ab_bench_events.err
