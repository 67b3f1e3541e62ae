#!/bin/bash
mkdir -p gpurun_out
cd /root/repo
for mode in 0 1; do
  echo "=== sync mode $mode"
  XSQ_RKC_SYNC_MODE=$mode timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$mode tools/bench_rkc_mp.py 2>&1 | grep -E "^\{|Error|error" | tail -2
done
echo "=== one rank"
timeout 200 python tools/bench_rkc_mp.py 2>&1 | grep -E "^\{" | tail -1
nvidia-smi topo -m 2>&1 | head -8
