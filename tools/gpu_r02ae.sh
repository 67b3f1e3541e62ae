#!/bin/bash
# round 2, second session: ncu capture of the events kernel with the event queue
mkdir -p gpurun_out
cat > /tmp/ev.py <<'PY'
import numpy as np, torch, sys
sys.path.insert(0, "/root/repo")
import extensisq_b200 as xb
import bench
N = 300000
y0, prm = bench.make_lanes(N, 0)
src = """
__device__ double event(int k, double t, const double* y, const double* p) {
    if (k == 0) return y[2] - 27.0;
    if (k == 1) return y[0];
    return y[0] * y[1] - 30.0;
}"""
ev = xb.DeviceEvents.from_source(src, "event", 3, terminal=[0, 0, 0], direction=[1, 0, 0])
for it in range(2):
    r = xb.solve_ivp_batched("lorenz63", (0.0, 20.0), y0, xb.Ts5, params=prm, rtol=1e-8, atol=1e-10,
                             events=ev, max_event_records=64)
    torch.cuda.synchronize()
print("events", int(r.event_counts.sum()), "steps", int(r.n_accepted.sum()))
PY
XSQ_USER_MINB=3 timeout 500 ncu --set full --clock-control none --import-source on -k regex:xsq_user_kernel -s 1 -c 1 \
    -f -o gpurun_out/prof_r02ae_events python /tmp/ev.py > gpurun_out/r02ae.log 2>&1
XSQ_USER_MINB=3 timeout 300 ncu --set full --clock-control none -k regex:xsq_user_evq -s 1 -c 1 \
    -f -o gpurun_out/prof_r02ae_evq python /tmp/ev.py >> gpurun_out/r02ae.log 2>&1
tail -5 gpurun_out/r02ae.log; ls -la gpurun_out/*r02ae*
