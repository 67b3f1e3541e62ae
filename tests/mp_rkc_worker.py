"""torchrun worker: SSV2stab on a grid split over WORLD_SIZE ranks vs the
single-rank solve of the same grid (run on rank 0)."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import extensisq_b200 as xb  # noqa: E402
from oracle.problems import heat2d_reaction  # noqa: E402


def main():
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    comm = xb.SlabComm()
    report = dict(ok=True, cases=[])
    for nx, use_rho, tol in ((64, True, 1e-4), (256, True, 1e-4),
                             (64, False, 1e-4)):
        _, y0, rho = heat2d_reaction(nx)
        u0 = y0.reshape(nx, nx)
        lo, hi = xb.slab_of(nx, rank, world)
        te = np.linspace(0, 0.05, 5)
        r = xb.solve_pde_rkc("heat2d_reaction", (0.0, 0.05), u0[lo:hi],
                             rows_global=nx, row0=lo, t_eval=te,
                             rho_jac=float(rho) if use_rho else None,
                             rtol=tol, atol=tol, comm=comm, max_steps=10000)
        parts = [torch.empty((xb.slab_of(nx, q, world)[1] -
                              xb.slab_of(nx, q, world)[0], nx),
                             dtype=torch.float64, device="cuda")
                 for q in range(world)]
        dist.all_gather(parts, r.y_final)
        full = torch.cat(parts, 0)
        eparts = [torch.empty((te.size, p.shape[0], nx), dtype=torch.float64,
                              device="cuda") for p in parts]
        dist.all_gather(eparts, r.y.contiguous())
        efull = torch.cat(eparts, 1)
        if rank == 0:
            one = xb.solve_pde_rkc("heat2d_reaction", (0.0, 0.05), u0,
                                   t_eval=te,
                                   rho_jac=float(rho) if use_rho else None,
                                   rtol=tol, atol=tol, max_steps=10000)
            same = (one.n_accepted, one.n_rejected, one.nfev, one.nfesig,
                    one.maxm) == (r.n_accepted, r.n_rejected, r.nfev,
                                  r.nfesig, r.maxm)
            err = float((full - one.y_final).abs().max())
            eerr = float((efull - one.y).abs().max())
            ok = same and err <= 1e-12 and eerr <= 1e-12 and r.status == 0
            report["cases"].append(dict(nx=nx, use_rho=use_rho, same=same,
                                        err=err, eerr=eerr, nfev=r.nfev,
                                        launches=r.kernel_launches))
            report["ok"] = report["ok"] and ok
    comm.close()
    dist.barrier()
    if rank == 0:
        with open(sys.argv[1], "w") as fh:
            json.dump(report, fh)
        print(json.dumps(report))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
