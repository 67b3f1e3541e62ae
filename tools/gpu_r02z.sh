#!/bin/bash
# 4 GPUs: the two-neighbour (middle rank) path of the SSV2stab halo, under torchrun
mkdir -p gpurun_out
cd /root/repo
nvidia-smi -L | head -8
timeout 240 python -m pytest tests/test_gpu_rkc.py -q -x --timeout 200 -k "rank" 2>&1 | tail -1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 2 --warmup 3 --no-cpu --only c5 > gpurun_out/r02z_bench_n8.json 2> gpurun_out/r02z_bench_n8.err
echo "rc=$?"
tail -5 gpurun_out/r02z_bench_n8.err
python -c "
import json
d = json.loads(open('gpurun_out/r02z_bench_n8.json').read().strip().splitlines()[-1])
print('value %.4g n_gpus %d frac %.4f e2e %.4g' % (d['value'], d['n_gpus'], d['roofline']['frac'], d['e2e']['value']))
s = d.get('ssv2stab', {})
for k in ('strong','weak','one_gpu_same_grid','parity_vs_1rank','checksum_rel_err_vs_1rank','strong_speedup_vs_1gpu','strong_efficiency','error'):
    print(k, s.get(k))
"
