#!/bin/bash
mkdir -p gpurun_out
cd /root/repo
timeout 300 python -m pytest tests/test_gpu_rkc.py -q -x --timeout 200 2>&1 | tail -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/bench_rkc_mp.py 2>&1 | grep -E "^\{|Error|error" | tail -2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/r02h_bench_n2.json 2> gpurun_out/r02h_bench_n2.err
echo "rc=$?"
python -c "
import json
d = json.loads(open('gpurun_out/r02h_bench_n2.json').read().strip().splitlines()[-1])
print('value %.4g n_gpus %d frac %.4f' % (d['value'], d['n_gpus'], d['roofline']['frac']))
s = d.get('ssv2stab', {})
for k in ('strong','weak','one_gpu_same_grid','parity_vs_1rank','checksum_rel_err_vs_1rank','strong_speedup_vs_1gpu','strong_efficiency','error'):
    print(k, s.get(k))
"
