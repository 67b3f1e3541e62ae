#!/bin/bash
mkdir -p gpurun_out
cd /root/repo
L=gpurun_out/r02t.log
: > $L
for cfg in "0 132 10 -2" "0 132 10 254" "0 132 10 126" "0 132 10 2" "0 132 10 4" "0 132 10 -4" "0 132 10 384" "0 136 10 -4" "0 136 10 124"; do
  timeout 60 tools/devtest/tma_probe $cfg >> $L 2>&1
done
cat $L
