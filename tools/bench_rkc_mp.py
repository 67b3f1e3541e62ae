"""torchrun worker: SSV2stab weak scaling -- every rank owns a 2048 x 16384 slab
(the 8-GPU share of BASELINE.json configs[4]); reports time per stage."""
import json, os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import extensisq_b200 as xb

rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
comm = None
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    comm = xb.SlabComm()
nx = int(os.environ.get("NX", 16384)); rows = int(os.environ.get("ROWS", 2048))
rows_global = rows * world
h = 1.0 / (nx + 1)
x = torch.arange(1, nx + 1, dtype=torch.float64, device="cuda") * h
yrow = torch.arange(rank * rows + 1, (rank + 1) * rows + 1, dtype=torch.float64, device="cuda") / (rows_global + 1)
u0 = torch.outer(torch.sin(np.pi * yrow), torch.sin(np.pi * x))
rho = 8.0 * (nx + 1.0) ** 2 + 2.0
T = float(os.environ.get("T", 2e-5))
for rep in range(2):
    if world > 1: dist.barrier()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r = xb.solve_pde_rkc("heat2d_reaction", (0.0, T), u0, rows_global=rows_global, row0=rank * rows,
                         rho_jac=float(rho), rtol=1e-4, atol=1e-4, comm=comm, max_steps=50)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
if rank == 0:
    print(json.dumps(dict(world=world, nx=nx, rows_per_gpu=rows, s=dt, nfev=r.nfev, accepted=r.n_accepted,
                          maxm=r.maxm, ms_per_stage=dt / r.nfev * 1e3,
                          algorithmic_GBps_per_gpu=nx * rows * 40 * r.nfev / dt / 1e9, status=r.status)))
if world > 1:
    comm.close(); dist.destroy_process_group()
