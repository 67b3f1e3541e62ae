"""extensisq_b200 -- B200 (sm_100a) implementation of extensisq's explicit
Runge-Kutta / SWAG / SSV2stab stepping path for ensembles and large PDEs.

Same export names as the reference for the hot-path classes
(``extensisq/__init__.py:4-12``); everything else of the reference (ESDIRK,
Nystrom, sensitivity) is out of scope.
"""
from .tableaux import (RungeKutta, Ts5, BS5, CK5, CKdisc, Me4, Pr7, Pr8, Pr9,
                       CFMR7osc, SWAG, BUILTIN, REFERENCE_VERSION,
                       RungeKuttaNystrom, Fi4N, Fi5N, Mu5Nmb, MR6NN, BUILTIN_RKN)
from .batched import (DeviceRHS, DeviceEvents, BatchedOdeResult, BatchedOdeSolution, solve_ivp_batched, NFS, trim_memory,
                      update_nfs)
from .sharding import shard_bounds, gather_result
from .sensitivity import sens_forward, SensitivityOutput
from .pde import (SSV2stab, SlabComm, PdeResult, PdeRHS, solve_pde_rkc, slab_of,
                  nfesig, maxm)

__version__ = "0.1.0"
__all__ = ["RungeKutta", "Ts5", "BS5", "CK5", "CKdisc", "Me4", "Pr7", "Pr8", "Pr9",
           "CFMR7osc", "SWAG", "RungeKuttaNystrom", "Fi4N", "Fi5N", "Mu5Nmb", "MR6NN", "DeviceRHS", "DeviceEvents", "BatchedOdeResult", "BatchedOdeSolution", "solve_ivp_batched",
           "NFS", "update_nfs", "trim_memory", "shard_bounds", "gather_result", "sens_forward", "SensitivityOutput", "SSV2stab",
           "SlabComm", "PdeResult", "PdeRHS", "solve_pde_rkc", "slab_of", "nfesig",
           "maxm"]
