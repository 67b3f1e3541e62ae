"""``SSV2stab`` for one large semi-discretised parabolic PDE on the device(s).

Counterpart of ``solve_ivp(fun, t_span, y0, method=SSV2stab, rtol=, atol=,
first_step=, max_step=, const_jac=, rho_jac=, t_eval=)`` of the reference
(``extensisq/sommeijer.py:17-406``).  The state is a 2-D grid; with more than
one rank it is split into contiguous row slabs and the
kernels exchange one halo row per stage over NCCL.  Host logic only; the
stepping loop and every kernel live in ``libxsq.so`` (``xsq_rkc_solve``).
"""
import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from .sharding import shard_bounds

__all__ = ["SSV2stab", "SlabComm", "PdeResult", "PdeRHS", "solve_pde_rkc",
           "nfesig", "maxm"]

# module-level counters of the reference (sommeijer.py:12-14), last solve
nfesig = np.array(0)
maxm = np.array(0)

PDES = {"heat2d_reaction": 0}


class PdeRHS:
    """Right-hand side of a 2-D parabolic PDE on the unit square (homogeneous
    Dirichlet boundaries, 5-point neighbourhood) as CUDA source defining

        __device__ double <entry>(double t, double x, double y, double inv_h2,
                                  double uc, double un, double us, double uw,
                                  double ue, const double* p);

    compiled with NVRTC into the fused SSV2stab stage kernels.  Replaces the
    Python callable `fun` of the reference (sommeijer.py:93)."""

    def __init__(self, handle, n_param, name, n_state=0):
        self.handle, self.n_param, self.name = handle, n_param, name
        self.n_state = n_state          # > 0: a general system (from_vector_source)

    @classmethod
    def from_source(cls, cuda_src, entry, n_param=0):
        lib = _lib.load()
        h = C.c_int32()
        _lib.check(lib.xsq_pde_register_source(cuda_src.encode(),
                                               entry.encode(), int(n_param),
                                               C.byref(h)))
        return cls(h.value, int(n_param), f"user:{entry}")

    @classmethod
    def from_vector_source(cls, cuda_src, entry, n_state, n_param=0):
        """A GENERAL system ``y' = f(t, y)`` of ``n_state`` equations (the
        reference's SSV2stab takes any ``fun``, sommeijer.py:93-145; its published
        examples are 3-D and multi-component problems): CUDA source defining

            __device__ double <entry>(int i, double t, const double* y, const double* p);

        the i-th component of f with the whole state in view.  ``solve_pde_rkc``
        then takes a 1-D ``u0`` of length ``n_state`` (single GPU)."""
        lib = _lib.load()
        h = C.c_int32()
        _lib.check(lib.xsq_pde_register_vector_source(cuda_src.encode(), entry.encode(),
                                                      int(n_state), int(n_param), C.byref(h)))
        return cls(h.value, int(n_param), f"user:{entry}", n_state=int(n_state))


class SSV2stab:
    """Method description (reference: sommeijer.py:17-145).  Options of
    ``solve_pde_rkc``: ``const_jac`` (bool), ``rho_jac`` (None | float |
    callable ``rho_jac(t) -> float``)."""
    _xsq_method = 300


class SlabComm:
    """NCCL communicator owned by libxsq for the halo exchange.  The unique id
    is created on rank 0 and broadcast through ``torch.distributed`` (any
    backend)."""

    def __init__(self, group=None):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        lib = _lib.load()
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        buf = C.create_string_buffer(128)
        if self.rank == 0:
            _lib.check(lib.xsq_comm_unique_id(buf))
        backend = dist.get_backend(group)
        dev = (torch.device("cuda", torch.cuda.current_device())
               if backend == "nccl" else torch.device("cpu"))
        t = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).to(dev)
        dist.broadcast(t, src=0, group=group)
        self._id = bytes(t.cpu().numpy().tobytes())
        self._handle = C.c_void_p()
        _lib.check(lib.xsq_comm_create(self.rank, self.world, self._id,
                                       C.byref(self._handle)))

    def close(self):
        if self._handle:
            _lib.load().xsq_comm_destroy(self._handle)
            self._handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


@dataclass
class PdeResult:
    t: object                 # t_eval (numpy) or None
    y: object                 # [n_eval, rows_local, nx] device tensor or None
    y_final: torch.Tensor     # [rows_local, nx]
    t_final: float
    n_accepted: int
    n_rejected: int
    nfev: int
    nfesig: int
    maxm: int
    status: int
    kernel_launches: int
    row0: int = 0
    rows_global: int = 0

    @property
    def success(self):
        return self.status >= 0

    @property
    def message(self):
        return _lib.LANE_MESSAGES[self.status]


def validate_rkc_options(rtol, atol, first_step, max_step, const_jac, rho_jac,
                         t0, tf):
    """Argument checks of SSV2stab.__init__ (sommeijer.py:99-118) with the
    reference's exception types; returns (rho_const, rho_callable)."""
    if first_step is not None:
        if first_step <= 0:
            raise ValueError("`first_step` must be positive.")
        if first_step > abs(tf - t0):
            raise ValueError("`first_step` exceeds bounds.")
    if not isinstance(const_jac, bool):
        raise TypeError('`const_jac` should be True or False')
    rho_const, rho_fn = 0.0, None
    if rho_jac is not None:
        if callable(rho_jac):
            v = rho_jac(t0)
            if not isinstance(v, float):
                raise TypeError('`rho_jac` should return a float')
            if v <= 0:
                raise ValueError('`rho_jac` should return a positive float')
            rho_fn = rho_jac
        elif isinstance(rho_jac, float):
            if rho_jac <= 0:
                raise ValueError('`rho_jac` should return a positive float')
            rho_const = rho_jac
        else:
            raise TypeError('`rho_jac` should be None or a function: '
                            '`sprad = rho_jac(t, y)`')
    if max_step <= 0:
        raise ValueError("`max_step` must be positive.")
    if not isinstance(rtol, float):
        raise ValueError("`rtol` must be a float.")
    if rtol < 0:
        raise ValueError("`rtol` must be positive.")
    if np.ndim(atol) != 0:
        raise ValueError("`atol` must be a scalar for the PDE path.")
    if atol < 0:
        raise ValueError("`atol` must be positive.")
    return rho_const, rho_fn


def solve_pde_rkc(pde, t_span, u0, rows_global=None, row0=0, t_eval=None,
                  rtol=1e-3, atol=1e-6, first_step=None, max_step=np.inf,
                  const_jac=False, rho_jac=None, max_steps=None, comm=None,
                  stream=None, method=SSV2stab, pde_params=None):
    """Integrate the PDE `pde` ("heat2d_reaction") with SSV2stab.

    u0 : [rows_local, nx] float64 tensor -- this rank's row slab of the grid
    rows_global, row0 : size of the whole grid and first row of this slab
        (defaults: the slab is the whole grid)
    comm : SlabComm when the grid is split over ranks
    Other arguments as in the reference; `rho_jac` may be a positive float, a
    callable of t, or None (nonlinear power iteration, sommeijer.py:331-398).
    """
    global nfesig, maxm
    lib = _lib.load()
    if not torch.cuda.is_available():
        raise RuntimeError("extensisq_b200 needs a CUDA device; there is no "
                           "CPU fallback")
    if method is not SSV2stab:
        raise ValueError("solve_pde_rkc only implements SSV2stab")
    if isinstance(pde, PdeRHS):
        pde_id, n_prm = pde.handle, pde.n_param
    elif pde in PDES:
        pde_id, n_prm = PDES[pde], 0
    else:
        raise ValueError(f"unknown pde {pde!r}; available: {sorted(PDES)} or "
                         "a PdeRHS")
    prm_np = np.ascontiguousarray(np.asarray(
        pde_params if pde_params is not None else [], dtype=float))
    if prm_np.size != n_prm:
        raise ValueError(f"`pde_params` must have {n_prm} entries")
    t0, tf = map(float, t_span)
    rho_const, rho_fn = validate_rkc_options(rtol, atol, first_step, max_step,
                                             const_jac, rho_jac, t0, tf)
    if not isinstance(u0, torch.Tensor):
        u0 = torch.as_tensor(np.asarray(u0), dtype=torch.float64)
    dev = u0.device if u0.is_cuda else torch.device(
        "cuda", torch.cuda.current_device())
    u0 = u0.to(device=dev, dtype=torch.float64).contiguous()
    n_vec = pde.n_state if isinstance(pde, PdeRHS) else 0
    if n_vec:
        # a general system: one slab row, padded with zeros to a multiple of 4
        if u0.ndim != 1 or u0.numel() != n_vec:
            raise ValueError(f"`u0` must be 1-dimensional with {n_vec} entries")
        if comm is not None:
            raise ValueError("general systems run on one GPU")
        nx_pad = (n_vec + 3) // 4 * 4
        padded = torch.zeros((1, nx_pad), dtype=torch.float64, device=dev)
        padded[0, :n_vec] = u0
        u0 = padded
    if u0.ndim != 2:
        raise ValueError("`u0` must be [rows_local, nx]")
    rows_local, nx = u0.shape
    if nx % 4 or nx < 4:
        raise ValueError("nx must be a positive multiple of 4")
    rows_global = rows_local if rows_global is None else int(rows_global)
    world = comm.world if comm is not None else 1
    rank = comm.rank if comm is not None else 0
    if world == 1 and rows_global != rows_local:
        raise ValueError("a slab of a larger grid needs a SlabComm")
    te = None
    n_eval = 0
    if t_eval is not None:
        te = np.ascontiguousarray(np.asarray(t_eval, dtype=float))
        if te.ndim != 1:
            raise ValueError("`t_eval` must be 1-dimensional.")
        if te.size and (te.min() < min(t0, tf) or te.max() > max(t0, tf)):
            raise ValueError("Values in `t_eval` are not within `t_span`.")
        d = np.diff(te)
        if (tf > t0 and np.any(d <= 0)) or (tf < t0 and np.any(d >= 0)):
            raise ValueError("Values in `t_eval` are not properly sorted.")
        n_eval = te.size
    with torch.cuda.device(dev):
        u_final = torch.empty_like(u0)
        u_eval = (torch.empty((n_eval, rows_local, nx), dtype=torch.float64,
                              device=dev) if n_eval else None)
        res = _lib.XsqRkcResult()
        a = _lib.XsqRkcArgs()
        a.struct_size = C.sizeof(_lib.XsqRkcArgs)
        a.pde = pde_id
        a.pde_params = (prm_np.ctypes.data_as(C.POINTER(C.c_double))
                        if n_prm else None)
        a.n_pde_params = n_prm
        a.nx, a.rows_global, a.rows_local, a.row0 = nx, rows_global, \
            rows_local, int(row0)
        a.rank, a.world = rank, world
        a.u0 = u0.data_ptr()
        a.t0, a.t_bound = t0, tf
        a.rtol, a.atol = rtol, float(atol)
        a.first_step = float(first_step) if first_step is not None else 0.0
        a.max_step = float(max_step)
        a.const_jac = 1 if const_jac else 0
        a.max_steps = int(max_steps) if max_steps else 0
        a.rho_const = rho_const
        cb = _lib.RHO_FN(lambda t, _u: float(rho_fn(t))) if rho_fn else \
            _lib.RHO_FN()
        a.rho_cb = cb
        a.t_eval = te.ctypes.data_as(C.POINTER(C.c_double)) if n_eval else None
        a.n_eval = n_eval
        a.u_eval = u_eval.data_ptr() if n_eval else None
        a.u_final = u_final.data_ptr()
        a.result = C.pointer(res)
        st = stream if stream is not None else torch.cuda.current_stream(dev)
        _lib.check(lib.xsq_rkc_solve(C.byref(a),
                                     comm._handle if comm is not None else None,
                                     C.c_void_p(st.cuda_stream)))
    nfesig[()] = res.nfesig
    maxm[()] = res.maxm
    if n_vec:
        u_final = u_final[0, :n_vec]
        u_eval = u_eval[:, 0, :n_vec] if u_eval is not None else None
    return PdeResult(t=te, y=u_eval, y_final=u_final, t_final=res.t_final,
                     n_accepted=res.n_accepted, n_rejected=res.n_rejected,
                     nfev=res.nfev, nfesig=res.nfesig, maxm=res.maxm,
                     status=res.status, kernel_launches=res.kernel_launches,
                     row0=int(row0), rows_global=rows_global)


def slab_of(rows_global, rank, world):
    """[row0, row1) of this rank's contiguous row slab."""
    return shard_bounds(rows_global, rank, world)
