#!/bin/bash
cd /root/repo
echo "=== weak 2 ranks"
XSQ_RKC_DEBUG=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/bench_rkc_mp.py 2>&1 | grep -E "rkc r0|^\{" | tail -8
echo "=== strong 2 ranks"
XSQ_RKC_DEBUG=1 ROWS=8192 T=4.883e-5 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 tools/bench_rkc_mp.py 2>&1 | grep -E "rkc r0|^\{" | tail -8
echo "=== 8192 rows 1 rank"
XSQ_RKC_DEBUG=1 ROWS=8192 T=4.883e-5 timeout 300 python tools/bench_rkc_mp.py 2>&1 | grep -E "rkc r0|^\{" | tail -5
