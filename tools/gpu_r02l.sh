#!/bin/bash
mkdir -p gpurun_out
cd /root/repo
timeout 600 ncu --set full --clock-control none --import-source on -k regex:swag_fast -s 1 -c 1 -f -o gpurun_out/prof_r02l_swag python tools/prof_swag.py > gpurun_out/r02l.log 2>&1
tail -3 gpurun_out/r02l.log
