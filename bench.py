#!/usr/bin/env python
"""bench.py -- accepted RK steps/sec of the Lorenz-63 ensemble (BASELINE.json
configs[1]: Ts5/CK5, 10 M lanes over 8 GPUs = 1.25 M lanes per GPU, randomised
initial conditions and parameters, rtol 1e-8, atol 1e-10, t in [0, 100]).

A "step" of the benchmark = one pass of the hot path over this rank's shard of
the ensemble (one xsq_rk_solve call = one persistent-kernel launch).

  python bench.py [--gpus N] [--steps K] [--warmup W]          # our arm
  python bench.py --impl reference ...                         # CPU arm

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for how every
field is obtained.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LANES_PER_GPU = 1_250_000           # 10 M lanes / 8 GPUs (BASELINE.json configs[1])
T_END = 100.0
RTOL, ATOL = 1e-8, 1e-10
SEED = 12345


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--method", default="Ts5")
    ap.add_argument("--lanes", type=int, default=LANES_PER_GPU,
                    help="lanes per GPU (weak scaling)")
    ap.add_argument("--t-end", type=float, default=T_END)
    ap.add_argument("--cpu-lanes", type=int, default=0,
                    help="lanes of the CPU sample (0 = auto)")
    ap.add_argument("--stiff", type=int, default=5000,
                    help="nfev_stiff_detect (reference default 5000; 0 = off)")
    ap.add_argument("--no-cpu", action="store_true",
                    help="skip the cpu_baseline leg (profiling runs)")
    return ap.parse_args()


def make_lanes(n, rank):
    """y0 ~ U([-15,15]x[-20,20]x[5,40]),
    sigma ~ U(9,11), rho ~ U(24,32), beta ~ U(2.4,2.9); one stream per rank
    (seed 12345 + rank) so shards are independent of the rank count."""
    rng = np.random.default_rng(SEED + rank)
    # one row of six uniforms per lane: lane i does not depend on n, so any
    # prefix of a shard is a true subset of it (tests/test_gpu_exact.py)
    u = rng.random((n, 6))
    lo = np.array([-15.0, -20.0, 5.0, 9.0, 24.0, 2.4])
    hi = np.array([15.0, 20.0, 40.0, 11.0, 32.0, 2.9])
    v = lo + (hi - lo) * u
    return np.ascontiguousarray(v[:, :3]), np.ascontiguousarray(v[:, 3:])


# ---- algorithmic flops -------------------------------
def flops_per_attempt_and_accept(method_cls, n, F):
    """FMA = 2 flops.  Per ATTEMPTED step: stage sums 2n*nnzA, y+h*dy 2n(s-1),
    t+c*h 2(s-1), (s-1+FSAL) RHS evaluations, y_new 2n*nnzB+2n, scale 2n,
    error 2n*nnzE+5n+2, controller 10.  Non-FSAL: one more RHS evaluation per
    ACCEPTED step."""
    s = method_cls.n_stages
    fsal = 1 if method_cls.E[s] != 0 else 0
    nnzA = int(np.count_nonzero(method_cls.A))
    nnzB = int(np.count_nonzero(method_cls.B))
    nnzE = int(np.count_nonzero(method_cls.E[:s + fsal]))
    att = (2 * n * nnzA + 2 * n * (s - 1) + 2 * (s - 1) + (s - 1 + fsal) * F +
           2 * n * nnzB + 2 * n + 2 * n + 2 * n * nnzE + 5 * n + 2 + 10)
    acc = 0 if fsal else F
    return att, acc


# ---- clocks during the timed region ----------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,"
         "clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "100", "-i",
                 str(self.index)], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def mark(self):
        """Samples before this moment (nvidia-smi starting up during the last
        warm-up step) are not part of the report."""
        self.t_mark = time.perf_counter()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                 "sw_power_cap"]
        for ts, r in self.rows:
            if ts < getattr(self, "t_mark", 0.0):
                continue
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                pw.append(float(r[3]))
                for nme, v in zip(names, r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ---- CPU legs ----------------------------------------------------------------
def _np_worker(args):
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    from oracle import rk_oracle as O
    method, y0, prm, t_end = args
    tab = O.load_tableaux()[method]
    acc = 0
    for i in range(len(y0)):
        r = O.rk_solve(tab, O.lorenz63(*prm[i]), (0.0, t_end), y0[i],
                       rtol=RTOL, atol=ATOL)
        acc += r["n_accepted"]
    return acc


def cpu_numpy_port(method, lanes, t_end, cores, target_s=15.0):
    """The NumPy restatement (bit-identical to the reference's own Python,
    same per-step cost model), one process per core.  lanes == 0: calibrate
    on one lane per core and size the sample for ~target_s seconds."""
    import multiprocessing as mp
    with mp.get_context("spawn").Pool(cores) as pool:
        y0, prm = make_lanes(max(lanes, 4 * cores), 0)
        pool.map(_np_worker, [(method, y0[:1], prm[:1], 0.05)] * cores)  # warm
        if lanes <= 0:
            t0 = time.perf_counter()
            pool.map(_np_worker, [(method, y0[i:i + 1], prm[i:i + 1], t_end)
                                  for i in range(cores)])
            per_lane = time.perf_counter() - t0
            lanes = max(cores, int(cores * target_s / max(per_lane, 1e-3)))
            y0, prm = make_lanes(lanes, 0)
        chunks = [(method, y0[i:lanes:cores], prm[i:lanes:cores], t_end)
                  for i in range(cores)]
        t0 = time.perf_counter()
        accs = pool.map(_np_worker, chunks)
        dt = time.perf_counter() - t0
    return sum(accs) / dt, sum(accs), dt, lanes


def cpu_c_port(method, lanes, t_end, threads, target_s=10.0):
    from oracle import c_oracle as CO
    from oracle import rk_oracle as O
    tab = O.load_tableaux()[method]
    y0, prm = make_lanes(64 * threads, 0)
    t0 = time.perf_counter()
    CO.rk_batch(tab, "lorenz63", (0.0, t_end), y0, params=prm,
                rtol=RTOL, atol=ATOL, n_threads=threads)
    per_lane = (time.perf_counter() - t0) / (64 * threads)
    if lanes <= 0:
        lanes = max(64 * threads, int(target_s / max(per_lane, 1e-9)))
    y0, prm = make_lanes(lanes, 0)
    t0 = time.perf_counter()
    r = CO.rk_batch(tab, "lorenz63", (0.0, t_end), y0, params=prm, rtol=RTOL,
                    atol=ATOL, n_threads=threads)
    dt = time.perf_counter() - t0
    acc = int(r["n_accepted"].sum())
    return acc / dt, acc, dt, lanes


def config_dict(args, world):
    return {"workload": f"{args.method} on a Lorenz-63 ensemble, "
                        f"{args.lanes} lanes per GPU x {world} GPU(s) "
                        f"(BASELINE.json configs[1] shard: 10M lanes / 8 GPUs), "
                        f"t in [0,{args.t_end:g}], rtol {RTOL:g}, atol {ATOL:g}, "
                        "randomised y0 and (sigma, rho, beta), seed 12345+rank",
            "method": args.method, "lanes_per_gpu": args.lanes,
            "t_end": args.t_end, "rtol": RTOL, "atol": ATOL,
            "sharding": f"lanes/{world}", "l2": "flushed between timed iterations (512 MiB write)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    cores = len(os.sched_getaffinity(0))
    # bounded sample: ~15 s of CPU work per step, calibrated on this host
    lanes = args.cpu_lanes
    vals, t_all = [], []
    for i in range(args.warmup + args.steps):
        v, acc, dt, lanes = cpu_numpy_port(args.method, lanes, args.t_end,
                                           cores)
        if i >= args.warmup:
            vals.append(v)
            t_all.append(dt)
    value = float(np.mean(vals))
    sample = (f"{lanes} lanes of the same ensemble (seed 12345), full t span, "
              f"NumPy restatement of the reference (oracle/rk_oracle.py), "
              f"one process per core")
    line = {"impl": "reference", "metric": "accepted RK steps/sec (ensemble)",
            "value": value, "unit": "steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": float(np.mean(t_all) * 1e3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": config_dict(args, world),
            "cpu_baseline": {"value": value, "unit": "steps/s", "cores": cores,
                             "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "steps/s",
                    "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import ctypes as C
    import extensisq_b200 as xb
    from extensisq_b200 import _lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    method = getattr(xb, args.method)
    N = args.lanes

    y0_np, prm_np = make_lanes(N, rank)
    y0_pin = torch.from_numpy(y0_np).pin_memory()
    prm_pin = torch.from_numpy(prm_np).pin_memory()
    y0_d, prm_d = y0_pin.to(dev), prm_pin.to(dev)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream(dev)

    def solve_resident():
        return xb.solve_ivp_batched("lorenz63", (0.0, args.t_end), y0_d, method,
                                    params=prm_d, rtol=RTOL, atol=ATOL,
                                    nfev_stiff_detect=args.stiff)

    out_pin = {"y": torch.empty((N, 3), dtype=torch.float64).pin_memory(),
               "acc": torch.empty(N, dtype=torch.int32).pin_memory(),
               "rej": torch.empty(N, dtype=torch.int32).pin_memory(),
               "st": torch.empty(N, dtype=torch.int32).pin_memory()}

    def solve_e2e():
        a = y0_pin.to(dev, non_blocking=True)
        b = prm_pin.to(dev, non_blocking=True)
        r = xb.solve_ivp_batched("lorenz63", (0.0, args.t_end), a, method,
                                 params=b, rtol=RTOL, atol=ATOL,
                                 nfev_stiff_detect=args.stiff)
        out_pin["y"].copy_(r.y_final, non_blocking=True)
        out_pin["acc"].copy_(r.n_accepted, non_blocking=True)
        out_pin["rej"].copy_(r.n_rejected, non_blocking=True)
        out_pin["st"].copy_(r.status, non_blocking=True)
        return r
    h2d = y0_pin.numel() * 8 + prm_pin.numel() * 8
    d2h = sum(t.numel() * t.element_size() for t in out_pin.values())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # nvidia-smi is started BEFORE the last warm-up step: its start-up stalls
    # the GPU for tens of ms, which must not land in the first timed step; only
    # the samples taken inside the timed region are reported (sampler.mark()).
    sampler = ClockSampler(local)
    for i in range(args.warmup):
        if rank == 0 and i == args.warmup - 1:
            sampler.start()
        r = solve_resident()
    barrier()
    acc_total = int(r.n_accepted.sum().item())
    rej_total = int(r.n_rejected.sum().item())
    assert bool((r.status == 0).all()), "lanes failed"

    # ---- timed region 1: inputs resident in HBM ---------------------------
    if rank == 0 and sampler.proc is None:
        sampler.start()
    sampler.mark()
    lib.xsq_launch_count(1)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(args.steps)]
    barrier()
    t_wall0 = time.perf_counter()
    for s0, s1 in ev:
        flush.fill_(1)                  # evict L2 between timed iterations
        s0.record(stream)
        solve_resident()
        s1.record(stream)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = int(lib.xsq_launch_count(0))
    kern_ms = [a.elapsed_time(b) for a, b in ev]
    my_ms = float(sum(kern_ms))
    clocks = sampler.stop() if rank == 0 else None

    # ---- timed region 2: end to end through the public API, host buffers ---
    for _ in range(1):
        solve_e2e()
    barrier()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        solve_e2e()
    e1.record(stream)
    barrier()
    my_e2e_ms = e0.elapsed_time(e1)
    assert int(out_pin["acc"].sum().item()) == acc_total

    stats = torch.tensor([my_ms, my_e2e_ms], dtype=torch.float64, device=dev)
    tot = torch.tensor([acc_total, rej_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    max_ms, max_e2e_ms = stats.tolist()
    acc_all, rej_all = tot.tolist()

    if rank == 0:
        value = acc_all * args.steps / (max_ms * 1e-3)
        e2e_value = acc_all * args.steps / (max_e2e_ms * 1e-3)
        # roofline of the dominant (only) kernel, this rank
        att_f, acc_f = flops_per_attempt_and_accept(method, 3, 8)
        flops_launch = (acc_total + rej_total) * att_f + acc_total * acc_f
        ms_launch = float(np.mean(kern_ms))
        achieved = flops_launch / (ms_launch * 1e-3) / 1e12
        peak = C.c_double()
        _lib.check(lib.xsq_fp64_peak(local, 2000, C.byref(peak)))
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        hbm = None
        if os.path.exists(peaks_path):
            hbm = json.load(open(peaks_path)).get("hbm_gbs")
        line = {
            "metric": "accepted RK steps/sec (ensemble)",
            "value": value, "unit": "steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": max_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": config_dict(args, world),
            "accepted_steps_per_pass": acc_all, "rejected_steps_per_pass": rej_all,
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "steps/s",
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": max_e2e_ms / args.steps},
            "gpu_launches": launches,
            "roofline": {
                "bound": "fp64", "achieved": achieved, "peak": peak.value,
                "unit": "TFLOP/s", "frac": achieved / peak.value,
                # dram__bytes_read.sum + dram__bytes_write.sum of one launch at
                # this size from the committed capture profiles/r01_ncu_full_
                # rk_persistent_Ts5_lorenz_1250k_T100.txt: 0.23 GB + 2.54 GB with
                # the stiffness diagnosis on (the 2.4 GB are the probe-queue
                # records, 17.5 M x 160 B); 69.9 MB + 32.4 MB with it off
                "traffic": (None if not (args.method == "Ts5" and N == LANES_PER_GPU
                                         and args.t_end == T_END)
                            else 2.765e9 if args.stiff > 0 else 102.2e6),
                "traffic_unit": "bytes of DRAM traffic per launch (ncu)",
                "kernel": f"rk_persistent<{args.method}, Lorenz63>",
                "flops_per_attempted_step": att_f,
                "peak_source": "xsq_fp64_peak: dependent-chain DFMA microbenchmark "
                               "measured live on this GPU (MEASURED_PEAKS.json has no fp64 entry); "
                               "the kernel keeps state in registers, HBM traffic is ~100 B per "
                               "trajectory plus 160 B per queued stiffness probe (<1% of HBM "
                               "bandwidth), so the bound is the fp64 pipe, not HBM",
                "hbm_peak_gbs_measured": hbm},
            "wall_s_timed_region": t_wall,
            "kernel_ms_each_step": kern_ms,
        }
        # second hot path of the north star: SSV2stab's fused stage kernel
        # (HBM-bound, 40 B algorithmic per grid point and stage) on this GPU's
        # slab of the 16384^2 grid (configs[4]: 2048 rows x 16384 at 8 GPUs)
        ms_stage = C.c_double()
        if lib.xsq_rkc_stage_bench(16384, 2048, 40, C.byref(ms_stage), None) == 0:
            gbs = 16384 * 2048 * 40 / 1e9 / (ms_stage.value * 1e-3)
            line["ssv2stab"] = {
                "kernel": "k_stage (stencil RHS + three-term recurrence)",
                "slab": "2048 x 16384 (1/8 of the 16384^2 grid)",
                "ms_per_stage": ms_stage.value,
                "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm,
                             "unit": "GB/s",
                             "frac": gbs / hbm if hbm else None,
                             # ncu, profiles/r01_ncu_full_rkc_k_stage_
                             # 2048x16384.txt: 1.074 GB read + 0.2525 GB
                             # written per launch (algorithmic 1.342 GB)
                             "traffic": 1.3266e9,
                             "bytes_per_point_stage": 40}}
            try:    # end-to-end SSV2stab solve on the same slab (host loop,
                    # error-norm reductions, first/final stages included)
                import time as _t
                nxs, rws = 16384, 2048
                xg = torch.arange(1, nxs + 1, dtype=torch.float64, device=dev) / (nxs + 1)
                yg = torch.arange(1, rws + 1, dtype=torch.float64, device=dev) / (rws + 1)
                u0 = torch.outer(torch.sin(math.pi * yg), torch.sin(math.pi * xg))
                rho = 8.0 * (nxs + 1.0) ** 2 + 2.0
                for _ in range(2):
                    torch.cuda.synchronize()
                    t0 = _t.perf_counter()
                    rr = xb.solve_pde_rkc("heat2d_reaction", (0.0, 3e-4), u0,
                                          rho_jac=rho, rtol=1e-4, atol=1e-4,
                                          max_steps=100)
                    torch.cuda.synchronize()
                    dt = _t.perf_counter() - t0
                line["ssv2stab"]["solve"] = {
                    "t_span": [0.0, 3e-4], "accepted": rr.n_accepted,
                    "rejected": rr.n_rejected, "nfev": rr.nfev, "s_max": rr.maxm,
                    "seconds": dt, "kernel_launches": rr.kernel_launches,
                    "algorithmic_GBps": nxs * rws * 40 * rr.nfev / dt / 1e9}
                del u0, rr
            except Exception as exc:     # never lose the headline line
                line["ssv2stab"]["solve"] = {"error": repr(exc)}
        if not args.no_cpu:
            cores = len(os.sched_getaffinity(0))
            v_np, a_np, dt_np, lanes_np = cpu_numpy_port(
                args.method, args.cpu_lanes, args.t_end, cores)
            v_c, a_c, dt_c, lanes_c = cpu_c_port(args.method, 0, args.t_end,
                                                 cores)
            line["cpu_baseline"] = {
                "value": v_np, "unit": "steps/s", "cores": cores, "kind": "port",
                "sample": f"{lanes_np} lanes of the same ensemble, full t span, "
                          f"{a_np} accepted steps in {dt_np:.1f} s; NumPy restatement "
                          "of the reference (bit-identical to it, same Python/NumPy "
                          "cost model), one process per core"}
            line["cpu_baseline_c"] = {
                "value": v_c, "unit": "steps/s", "cores": cores, "kind": "port",
                "sample": f"{lanes_c} lanes, full t span, {a_c} accepted steps in "
                          f"{dt_c:.1f} s; plain-C restatement (oracle/xsq_oracle.c), "
                          "OpenMP over lanes -- a much stronger baseline than the "
                          "reference's own Python"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
