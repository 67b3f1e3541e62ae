#!/bin/bash
# round 2, second session: the whole GPU suite and the event timings with the final sources
mkdir -p gpurun_out
timeout 300 python tools/bench_events.py > gpurun_out/r02aj_bench_events.json 2> gpurun_out/r02aj_bench_events.err
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r02aj_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02aj_smoke.log 2>&1
cat gpurun_out/r02aj_bench_events.json | cut -c1-420; tail -3 gpurun_out/r02aj_bench_events.err; cat gpurun_out/r02aj_tests.log; tail -2 gpurun_out/r02aj_smoke.log
