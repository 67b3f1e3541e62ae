"""TEST INFRASTRUCTURE (dev-time checker: compares the device path with oracle/;
not part of the product, not used by bench.py).  Diagnostic (GPU box): per-method count parity GPU vs C oracle."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import extensisq_b200 as xb
from oracle import c_oracle as CO, rk_oracle as O
from test_gpu_rk import lorenz_lanes, vdp_lanes, arenstorf_lanes, to_np, METHODS
TABS = O.load_tableaux()
for prob, lanes, span in (("lorenz63", lorenz_lanes, (0., 5.)), ("vanderpol", vdp_lanes, (0., 20.)), ("arenstorf", arenstorf_lanes, (0., 3.))):
    for m in METHODS:
        N = 256
        y0, prm = lanes(N)
        res = to_np(xb.solve_ivp_batched(prob, span, y0, m, params=prm, rtol=1e-8, atol=1e-10))
        ref = CO.rk_batch(TABS[m.__name__], prob, span, y0, params=prm, rtol=1e-8, atol=1e-10, n_threads=8)
        same = (res["n_accepted"] == ref["n_accepted"]) & (res["n_rejected"] == ref["n_rejected"])
        bad = np.where(~same)[0]
        scale = np.abs(ref["y_final"]).max(axis=1) + 1e-300
        err = np.abs(res["y_final"] - ref["y_final"]).max(axis=1) / scale
        print(prob, m.__name__, "same %.4f" % same.mean(), "maxerr_same %.2e" % (err[same].max() if same.any() else -1), "median err %.2e" % np.median(err))
        for i in bad[:6]:
            print("   lane", i, "p", prm[i], "gpu", res["n_accepted"][i], res["n_rejected"][i], "ref", ref["n_accepted"][i], ref["n_rejected"][i], "err %.2e" % err[i])
