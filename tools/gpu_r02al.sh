#!/bin/bash
# round 2, second session: final build -- event tests, smoke, default bench
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_events.py tests/test_gpu_exact.py -m gpu -q -x -k "queue or ensemble or golden or ckdisc or c4" 2>&1 | tail -6 > gpurun_out/r02al_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02al_smoke.log 2>&1
timeout 500 python bench.py > gpurun_out/r02al_bench.json 2> gpurun_out/r02al_bench.err
cat gpurun_out/r02al_tests.log; tail -2 gpurun_out/r02al_smoke.log; head -c 600 gpurun_out/r02al_bench.json; tail -3 gpurun_out/r02al_bench.err
