#!/bin/bash
# round 2, second session, final sources: headline capture for profiles/traffic.json, default bench
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:rk_fast -s 3 -c 1 \
    -f -o gpurun_out/prof_r02ao_rk_fast python bench.py --steps 1 --warmup 3 --no-cpu --no-extras > gpurun_out/r02ao_ncu.log 2>&1
timeout 500 python bench.py > gpurun_out/r02ao_bench.json 2> gpurun_out/r02ao_bench.err
tail -2 gpurun_out/r02ao_ncu.log; head -c 400 gpurun_out/r02ao_bench.json; tail -2 gpurun_out/r02ao_bench.err
