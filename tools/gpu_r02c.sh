#!/bin/bash
# round 2, call C: geometry of rk_fast, remaining exactness tests, the full bench line
mkdir -p gpurun_out
cd /root/repo
L=gpurun_out/r02c.log
: > $L
step() { echo "=== $1" >> $L; shift; timeout "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
for mb in 4 5 6; do
  step "bench Ts5 minb=$mb" 200 env XSQ_FAST_MINB=$mb tools/quick_bench.sh ts5_minb$mb --steps 3 --warmup 3 --no-extras --no-cpu
done
for mb in 4 5 6; do
  step "bench CK5 minb=$mb" 200 env XSQ_FAST_MINB=$mb tools/quick_bench.sh ck5_minb$mb --steps 3 --warmup 3 --no-extras --no-cpu --method CK5
done
step "bench Ts5 nostiff default" 200 tools/quick_bench.sh ts5_nostiff --steps 3 --warmup 3 --no-extras --no-cpu --stiff 0
step "fast tests" 420 python -m pytest tests/test_gpu_fast.py -x -q --timeout 120
step "exact tests" 600 python -m pytest tests/test_gpu_exact.py -q -s --timeout 240
echo "=== full bench" >> $L
timeout 600 python bench.py > gpurun_out/r02c_bench_full.json 2>> $L; echo "rc=$?" >> $L
grep -E "^===|rc=|passed|failed|steps/s|identical|accepted/lane|Error|error" $L | tail -60
python -c "
import json
d = json.load(open('gpurun_out/r02c_bench_full.json'))
print('value %.4g frac %.4f e2e %.4g' % (d['value'], d['roofline']['frac'], d['e2e']['value']))
for k, v in d.get('configs', {}).items(): print(k, {kk: (round(vv, 4) if isinstance(vv, float) else vv) for kk, vv in v.items() if kk in ('value','ms','ok','ok_frac','t_eval_overhead_ms')}, 'frac', v.get('roofline', {}).get('frac'))
s = d.get('ssv2stab', {}); print({k: s.get(k) for k in ('strong','weak','stage_kernel','error')})
print({k: d.get(k) for k in ('cpu_baseline','cpu_baseline_serial','cpu_baseline_c')})
"
