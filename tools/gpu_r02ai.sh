#!/bin/bash
# round 2, second session: events timing, the whole GPU suite, the default bench,
# launch list and captures (headline kernel for profiles/traffic.json, fast event kernel)
mkdir -p gpurun_out
L=gpurun_out/r02ai.log
: > $L
step() { echo "=== $1" >> $L; shift; timeout "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
timeout 600 python tools/bench_events.py > gpurun_out/r02ai_bench_events.json 2> gpurun_out/r02ai_bench_events.err
XSQ_NO_FAST=1 timeout 300 python tools/bench_events.py 2>&1 | grep "event queue\"" > gpurun_out/r02ai_bench_events_nofast.json
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r02ai_tests.log
echo "=== bench default" >> $L
timeout 600 python bench.py > gpurun_out/r02ai_bench.json 2>> $L
echo "rc=$?" >> $L
step "launch list" 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/r02ai_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-extras
step "full rk_fast" 500 ncu --set full --clock-control none --import-source on -k regex:rk_fast -s 3 -c 1 \
    -f -o gpurun_out/prof_r02ai_rk_fast python bench.py --steps 1 --warmup 3 --no-cpu --no-extras
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
                             events=ev, max_event_records=48)
    torch.cuda.synchronize()
print("events", int(r.event_counts.sum()), "steps", int(r.n_accepted.sum()))
PY
step "full fast events" 400 ncu --set full --clock-control none --import-source on -k regex:xsq_user_kernel -s 1 -c 1 \
    -f -o gpurun_out/prof_r02ai_events_fast python /tmp/ev.py
step "full evq" 300 ncu --set full --clock-control none -k regex:xsq_user_evq -s 1 -c 1 \
    -f -o gpurun_out/prof_r02ai_evq python /tmp/ev.py
cat gpurun_out/r02ai_bench_events.json gpurun_out/r02ai_bench_events_nofast.json gpurun_out/r02ai_tests.log
grep -E "^===|rc=|Error" $L | tail -20
head -c 1500 gpurun_out/r02ai_bench.json
