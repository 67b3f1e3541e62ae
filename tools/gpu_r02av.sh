#!/bin/bash
# round 2, second session: kept log of the CKdisc / events / sens_forward throughput numbers
mkdir -p gpurun_out
timeout 140 python tools/bench_extras.py > gpurun_out/r02av_bench_extras.json 2> gpurun_out/r02av_bench_extras.err
cut -c1-260 gpurun_out/r02av_bench_extras.json; tail -2 gpurun_out/r02av_bench_extras.err
