#!/bin/bash
# round 2, second session: kept log of the CKdisc / events / sens_forward throughput numbers,
# and bench.py's C2_CKdisc configuration
mkdir -p gpurun_out
timeout 50 python bench.py --only c2ckdisc --no-cpu --steps 1 --warmup 3 > gpurun_out/r02av_bench_ckdisc.json 2> gpurun_out/r02av_bench_ckdisc.err
timeout 90 python tools/bench_extras.py > gpurun_out/r02av_bench_extras.json 2> gpurun_out/r02av_bench_extras.err
grep -o '"configs": {.*' gpurun_out/r02av_bench_ckdisc.json | cut -c1-400; tail -2 gpurun_out/r02av_bench_ckdisc.err
cut -c1-260 gpurun_out/r02av_bench_extras.json; tail -2 gpurun_out/r02av_bench_extras.err
