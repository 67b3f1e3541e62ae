#!/bin/bash
# resume test + the ncu captures the round-1 verdict asked for (t_eval variant of
# rk_persistent<Pr8, VanDerPol>, the events kernel)
mkdir -p gpurun_out
cd /root/repo
L=gpurun_out/r02aa.log
: > $L
step() { echo "=== $1" >> $L; shift; timeout "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
step "resume test" 300 python -m pytest tests/test_gpu_rk.py -q -x -k "resume" --timeout 200
step "full c3 t_eval" 500 ncu --set full --clock-control none --import-source on -k regex:rk_persistent -c 1 \
    -f -o gpurun_out/prof_r02aa_c3 python bench.py --steps 1 --warmup 3 --no-cpu --only c3
cat > /tmp/ev.py <<'PY'
import numpy as np, torch, sys
sys.path.insert(0, "/root/repo")
import extensisq_b200 as xb
import bench
N = 300000
y0, prm = bench.make_lanes(N, 0)
src = """
__device__ double event(int k, double t, const double* y, const double* p) {
    if (k == 0) return y[2] - (p[1] - 1.0);     // Poincare section z = rho - 1
    if (k == 1) return y[0];
    return y[1] - y[0];
}"""
ev = xb.DeviceEvents.from_source(src, "event", 3, terminal=[0, 0, 0], direction=[-1, 0, 1])
for it in range(2):
    r = xb.solve_ivp_batched("lorenz63", (0.0, 20.0), y0, xb.Ts5, params=prm, rtol=1e-8, atol=1e-10,
                             events=ev, max_event_records=4)
    torch.cuda.synchronize()
print("events", int(r.event_counts.sum()), "steps", int(r.n_accepted.sum()))
PY
step "full events" 500 ncu --set full --clock-control none --import-source on -k regex:xsq_user_kernel -s 1 -c 1 \
    -f -o gpurun_out/prof_r02aa_events python /tmp/ev.py
grep -E "^===|rc=|passed|failed|Error|events " $L | tail -20
ls -la gpurun_out/*.ncu-rep | tail -4
