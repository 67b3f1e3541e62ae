/*
 * xsq.h -- C ABI of libxsq.so: the B200 (sm_100a) implementation of
 * extensisq's explicit adaptive Runge-Kutta / SWAG / SSV2stab stepping path
 * for ensembles of small ODE systems and for one large parabolic PDE.
 *
 * The reference (WRKampi/extensisq v0.6.0) is pure Python and has NO FFI; each
 * entry point below names the reference interface it replaces (file:line in
 * the reference tree).  INTEGRATION.md shows the ctypes stub a maintainer of
 * the reference would add.
 *
 * Conventions
 *   - plain C types only; no C++ exceptions cross the boundary;
 *   - every function returns XSQ_OK (0) or a negative xsq_err;
 *   - the CALLER owns every buffer; the library borrows device pointers for
 *     the duration of the call and allocates only its own scratch on the
 *     given stream (freed before returning);
 *   - calls are asynchronous on `stream` (a cudaStream_t passed as void*)
 *     unless the name ends in _host;
 *   - per-lane outcome is reported in status[] with the xsq_lane_status codes,
 *     mirroring the reference's in-band (False, message) -> status=-1.
 */
#ifndef XSQ_H
#define XSQ_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XSQ_ABI_VERSION 2
#define XSQ_MAX_STAGES 18      /* Pr9: n_stages=17, +1 row for f(t+h, y_new) */
#define XSQ_MAX_POLY 8         /* Pr9 interpolant has 8 columns             */
#define XSQ_MAX_LANE_STATE 16  /* lane-per-system kernels: n_state <= 16    */
#define XSQ_MAX_WARP_STATE 1024 /* warp-per-system user right-hand sides    */

typedef enum xsq_err {
    XSQ_OK = 0,
    XSQ_ERR_ARG = -1,         /* bad argument (Python layer raises ValueError) */
    XSQ_ERR_CUDA = -2,        /* CUDA runtime/driver error                   */
    XSQ_ERR_NVRTC = -3,       /* user RHS failed to compile/link             */
    XSQ_ERR_UNSUPPORTED = -4, /* combination not compiled into the library   */
    XSQ_ERR_NOMEM = -5
} xsq_err;

/* per-lane status[] codes */
typedef enum xsq_lane_status {
    XSQ_LANE_FINISHED = 0,
    XSQ_LANE_EVENT = 1,           /* "A termination event occurred." (scipy
                                     solve_ivp status 1)                       */
    XSQ_LANE_STEP_TOO_SMALL = -1, /* OdeSolver.TOO_SMALL_STEP, common.py:234   */
    XSQ_LANE_OVERFLOW = -2,       /* "Overflow or underflow", common.py:286    */
    XSQ_LANE_TOL_TOO_TIGHT = -3,  /* SWAG, shampine.py:235-238                 */
    XSQ_LANE_SPRAD_FAILED = -4,   /* SSV2stab, sommeijer.py:179-182            */
    XSQ_LANE_STEP_BUDGET = -5,    /* max_steps exhausted (no reference analogue;
                                     guards the GPU against runaway lanes)     */
    XSQ_LANE_EVENT_QUEUE = -6     /* the event queue ran out of records (defensive:
                                     the queue is sized for every possible record
                                     whenever the kernel without an in-lane root
                                     finder is selected)                        */
} xsq_lane_status;

/* Built-in tableau methods: the reference's classes
 * Ts5 (tsitouras.py:83), BS5 (bogacki.py:103), CK5 (cash.py:82),
 * Me4 (merson.py:82), Pr7/Pr8/Pr9 (prince.py:79,205,449),
 * CFMR7osc (calvo.py:89), CKdisc (cash.py:115).  XSQ_METHOD_USER selects the tableau uploaded with
 * xsq_tableau_load (user subclasses of common.RungeKutta, common.py:88-121). */
typedef enum xsq_method {
    XSQ_TS5 = 0, XSQ_BS5 = 1, XSQ_CK5 = 2, XSQ_ME4 = 3,
    XSQ_PR7 = 4, XSQ_PR8 = 5, XSQ_PR9 = 6, XSQ_CFMR7OSC = 7,
    XSQ_CKDISC = 8,         /* cash.py:115-416: variable order (5,3,2); ignores
                               sc_params, never runs the stiffness diagnosis,
                               no forced step sequences */
    /* Runge-Kutta-Nystrom methods for second order problems in first order form
     * [v, a] = f(t, [x, v]) (common.py:1207-1320): Fi4N, Fi5N (fine.py:6,115),
     * Mu5Nmb (murua.py:6, scale_embedded=True), MR6NN (mikkawy.py:5, velocity
     * independent problems only).  n_state must be even, the first half of the
     * state are positions, the second half their velocities.  Final state and
     * counters only: no t_eval, no events, no stiffness diagnosis. */
    XSQ_FI4N = 9, XSQ_FI5N = 10, XSQ_MU5NMB = 11, XSQ_MR6NN = 12,
    XSQ_METHOD_USER = 100,
    XSQ_METHOD_SWAG = 200   /* internal tag used by xsq_swag_solve */
} xsq_method;

/* Built-in right-hand sides (the reference takes a Python callable `fun`,
 * common.py:187; on the device the RHS must be device code). Handles returned
 * by xsq_rhs_register_source are >= XSQ_RHS_USER_BASE. */
typedef enum xsq_rhs_id {
    XSQ_RHS_LORENZ63 = 0,   /* n=3, p=(sigma, rho, beta)                     */
    XSQ_RHS_VANDERPOL = 1,  /* n=2, p=(mu)                                   */
    XSQ_RHS_ARENSTORF = 2,  /* n=4, p=(mu)                                   */
    XSQ_RHS_NBODY32 = 3,    /* n=192 (32 bodies, 3-D), p=(eps2, m[32]); warp per system */
    XSQ_RHS_USER_BASE = 1000
} xsq_rhs_id;

/* what the reference reports as warnings (common.py:459-516) */
typedef enum xsq_stiff_flag {
    XSQ_STIFF_REAL = 1,        /* real dominant root, diagnosed as stiff        */
    XSQ_STIFF_COMPLEX = 2,     /* complex dominant pair, diagnosed as stiff     */
    XSQ_STIFF_OSCILLATORY = 4  /* complex pair near the imaginary axis and many
                                  recently failed steps                         */
} xsq_stiff_flag;

typedef enum xsq_interpolant {      /* BS5 only, bogacki.py:217-236 */
    XSQ_INTERP_DEFAULT = 0, XSQ_INTERP_FREE = 1, XSQ_INTERP_LOW = 2,
    XSQ_INTERP_BEST = 3
} xsq_interpolant;

/* POD image of a user tableau: the class attributes of a common.RungeKutta
 * subclass (common.py:88-121).  P may be absent (n_poly = 0): the cubic
 * Hermite fallback of common.py:793-821 is used. */
typedef struct xsq_tableau {
    int32_t n_stages, order, order_secondary, n_poly;
    double A[XSQ_MAX_STAGES][XSQ_MAX_STAGES];
    double B[XSQ_MAX_STAGES];
    double C[XSQ_MAX_STAGES];
    double E[XSQ_MAX_STAGES + 1];
    double P[XSQ_MAX_STAGES + 1][XSQ_MAX_POLY];
    double sc_params[4];            /* (kb1, kb2, a, g), common.py:166-185   */
    double stbrad, tanang;          /* stiffness detection (common.py:113-115);
                                       <= 0: not implemented for this method */
} xsq_tableau_t;

/* Arguments of one batched explicit-RK solve: replaces
 *   solve_ivp(fun, t_span, y0, method=Cls, t_eval=..., **options)
 * (scipy ivp.py:161-760 driving RungeKutta.__init__/_step_impl/
 * _dense_output_impl, common.py:187-368; BS5 bogacki.py:238-393; CFMR7osc
 * calvo.py:152-261) for n_lanes independent trajectories.
 * All array pointers are DEVICE pointers unless stated otherwise. */
typedef struct xsq_rk_args {
    int32_t struct_size;      /* sizeof(xsq_rk_args_t), ABI guard            */
    int32_t method;           /* xsq_method                                  */
    int32_t rhs;              /* xsq_rhs_id or registered handle             */
    int32_t n_state;          /* n                                           */
    int32_t n_param;          /* parameters per lane                         */
    int32_t interpolant;      /* xsq_interpolant (BS5)                       */
    int64_t n_lanes;          /* N trajectories                              */
    const double* y0;         /* SoA [n_state][n_lanes]                      */
    const double* params;     /* SoA [n_param][n_lanes], NULL if n_param==0  */
    double t0, t_bound;       /* t_span                                      */
    double rtol;              /* clipped like validate_tol, common.py:30-54  */
    const double* atol;       /* HOST pointer, n_atol in {1, n_state}        */
    int32_t n_atol;
    int32_t use_sc_params;    /* 0: the method's default controller          */
    double sc_params[4];      /* (kb1, kb2, a, g) if use_sc_params           */
    double first_step;        /* <= 0: Watts' h_start, common.py:519-763     */
    double max_step;          /* +inf for none                               */
    const double* t_eval;     /* [n_eval] sorted along the direction, or NULL */
    int32_t n_eval;
    int32_t max_steps;        /* attempted-step budget per lane, <=0: 2^31-1 */
    double* y_eval;           /* [n_lanes][n_state][pitch], pitch = n_eval
                                 rounded up to a multiple of 4 (32-byte rows) */
    const double* h_forced;   /* forced |h| sequence [n_forced] or NULL      */
    int32_t n_forced;
    int32_t reserved0;
    double* t_final;          /* [n_lanes]                                   */
    double* y_final;          /* SoA [n_state][n_lanes]                      */
    double* h_next;           /* [n_lanes] next |h| proposal, may be NULL    */
    int32_t* n_accepted;      /* [n_lanes]                                   */
    int32_t* n_rejected;      /* [n_lanes]  (the reference's NFS counter)    */
    int32_t* nfev;            /* [n_lanes]                                   */
    int32_t* status;          /* [n_lanes]  xsq_lane_status                  */
    int32_t* n_eval_done;     /* [n_lanes] t_eval points written, may be NULL */
    int32_t nfev_stiff_detect;/* stiffness diagnosis every this many RHS
                                 evaluations (common.py:150-164, 370-516);
                                 0 = off; the reference's default is 5000   */
    int32_t reserved1;
    int32_t* stiff_flags;     /* [n_lanes] OR of xsq_stiff_flag, may be NULL  */
    /* events: scipy solve_ivp(events=...) on the device -- find_active_events,
     * handle_events, solve_event_equation (scipy/integrate/_ivp/ivp.py) and
     * brentq (scipy/optimize/Zeros/brentq.c), evaluated on the method's own
     * dense output (Horner / cubic / SWAG's interpolant).  Not for forced
     * steps or XSQ_RHS_NBODY32. */
    int32_t events;           /* handle of xsq_events_register_source, 0 = none */
    int32_t n_event_fns;      /* must equal the handle's n_events               */
    const int32_t* ev_terminal;  /* HOST [n_event_fns]: 0 never terminal, k > 0
                                    stop at the k-th occurrence (event.terminal) */
    const int32_t* ev_direction; /* HOST [n_event_fns]: -1, 0, +1 (event.direction) */
    int32_t ev_capacity;      /* records kept per event function and lane       */
    int32_t reserved2;
    double* t_events;         /* [n_lanes][n_event_fns][ev_capacity]; records
                                 beyond min(ev_count, ev_capacity) are left as
                                 the caller filled them (xsq_rk_solve_host: NaN) */
    double* y_events;         /* [n_lanes][n_event_fns][ev_capacity][n_state]   */
    int32_t* ev_count;        /* [n_lanes][n_event_fns] occurrences found (can
                                 exceed ev_capacity: later ones are not kept)   */
    /* resume: a first |h| PER LANE, e.g. the h_next of the solve that ended at
     * this call's t0 (the reference's manual stepping, tests/test_ivp.py:839-868,
     * continues with the step size the controller proposed).  [n_lanes] or NULL;
     * overrides first_step; each value is clipped to |t_bound - t0|; no h_start
     * evaluations are made.  Not with forced steps. */
    const double* first_step_lanes;
} xsq_rk_args_t;

int xsq_abi_version(void);
const char* xsq_strerror(int err);
/* text of the last NVRTC / CUDA failure on this thread ("" if none) */
const char* xsq_last_error_detail(void);

/* Number of SMs / device name probe; returns XSQ_ERR_CUDA if there is no
 * usable device (the product has no CPU fallback). */
int xsq_device_info(int device, int32_t* n_sm, int32_t* cc_major,
                    int32_t* cc_minor);

/* Upload a user tableau into the device's constant-memory slot.
 * Replaces: subclassing common.RungeKutta with own A/B/C/E/P
 * (common.py:88-121, docs/Demo_own_RK.ipynb). */
int xsq_tableau_load(const xsq_tableau_t* tab);
/* Read back the constant-memory image of a built-in method (for the
 * order-condition checks of the reference's tests/test_rk.py:14-72). */
int xsq_tableau_get(int32_t method, xsq_tableau_t* out);

/* Look up a built-in RHS by name: "lorenz63", "vanderpol", "arenstorf",
 * "nbody32".  Replaces the `fun` argument of common.py:187. */
int xsq_rhs_builtin(const char* name, int32_t* rhs_out, int32_t* n_state,
                    int32_t* n_param);
/* Register CUDA source defining
 *   __device__ void <entry>(double t, const double* y, const double* p,
 *                           double* dy);
 * It is compiled with NVRTC together with the solver template so the RHS
 * inlines into the persistent kernel.  Systems up to XSQ_MAX_LANE_STATE states
 * run one per thread.  Larger ones (up to XSQ_MAX_WARP_STATE; the reference
 * takes any n, common.py:187-217) run one per WARP, and the entry returns one
 * component of the derivative with the whole stage vector in view:
 *   __device__ double <entry>(int i, double t, const double* y, const double* p);
 * (n_param <= 16 there; no events). */
int xsq_rhs_register_source(const char* cuda_src, const char* entry,
                            int32_t n_state, int32_t n_param,
                            int32_t* rhs_out);

/* Compile-only probe: does NVRTC accept the specialised kernel for
 * (method, rhs)?  Needs libnvrtc but no device. */
int xsq_user_compile_check(int32_t method, int32_t rhs);

/* Event functions (the `events=` argument of scipy's solve_ivp, ivp.py) as
 * CUDA source defining
 *     __device__ double <entry>(int k, double t, const double* y, const double* p)
 * for k = 0 .. n_events-1; terminal / direction attributes travel in
 * xsq_rk_args_t.  The kernel for (method, rhs, events) is compiled with NVRTC
 * on first use. */
#define XSQ_MAX_EVENTS 8
int xsq_events_register_source(const char* cuda_src, const char* entry, int32_t n_events,
                               int32_t* handle_out);
int xsq_events_compile_check(int32_t method, int32_t rhs, int32_t events);

/* Batched adaptive explicit RK solve, device buffers, asynchronous. */
int xsq_rk_solve(const xsq_rk_args_t* args, void* stream);
/* Same, but EVERY array pointer in args is a HOST pointer; H2D/D2H copies and
 * the solve happen inside the call, which returns when the results are in the
 * host buffers. */
int xsq_rk_solve_host(const xsq_rk_args_t* args, int device);

/* Batched SWAG solve (Shampine-Gordon-Watts variable-order Adams PECE):
 * replaces solve_ivp(fun, t_span, y0, method=SWAG, t_eval=..., k_max=...)
 * i.e. SWAG.__init__/_step_impl (shampine.py:99-480) and SwagDenseOutput
 * (shampine.py:498-587).  Same argument block as the RK solve; the fields
 * method, interpolant, use_sc_params, sc_params, h_forced and n_forced are
 * ignored.  n_rejected receives the failed-step count (the reference's NFS).
 * 1 <= k_max <= 12 (shampine.py:102-103). */
typedef xsq_rk_args_t xsq_swag_args_t;
int xsq_swag_solve(const xsq_swag_args_t* args, int32_t k_max, void* stream);

/* ---- SSV2stab: one large parabolic PDE, row-slab decomposed -------------- */
typedef enum xsq_pde_id {
    /* u_t = Lap(u) + u - u^3 on (0,1)^2, Dirichlet 0, nx x rows_global interior
     * points, 5-point stencil, h = 1/(nx+1)   */
    XSQ_PDE_HEAT2D_REACTION = 0,
    XSQ_PDE_USER_BASE = 1000      /* handles from xsq_pde_register_source */
} xsq_pde_id;

typedef double (*xsq_rho_fn)(double t, void* user);   /* rho_jac(t, .) */

typedef struct xsq_rkc_result {     /* HOST struct filled by xsq_rkc_solve */
    double t_final;
    int32_t n_accepted, n_rejected; /* rejected = the reference's NFS/nrejct  */
    int32_t nfev, nfesig, maxm;     /* sommeijer.py:12-14 counters            */
    int32_t status;                 /* xsq_lane_status                        */
    int32_t n_eval_done, reserved;
    int64_t kernel_launches;
} xsq_rkc_result_t;

/* Replaces solve_ivp(fun, t_span, y0, method=SSV2stab, rtol, atol, first_step,
 * max_step, const_jac, rho_jac, t_eval) -- SSV2stab.__init__/_step_impl/
 * _stages/_rho/_dense_output_impl, sommeijer.py:93-406 -- for the built-in
 * PDE right-hand sides.  Each rank owns rows [row0, row0 + rows_local) of the
 * rows_global x nx grid. */
typedef struct xsq_rkc_args {
    int32_t struct_size;
    int32_t pde;                  /* xsq_pde_id                               */
    int32_t nx;                   /* points per row (multiple of 4)           */
    int32_t rows_global, rows_local, row0;
    int32_t rank, world;          /* position in the row-slab decomposition   */
    const double* u0;             /* DEVICE [rows_local][nx]                  */
    double t0, t_bound;
    double rtol, atol;            /* scalars; clipped like validate_tol       */
    double first_step;            /* <= 0: _init_step_size                    */
    double max_step;
    int32_t const_jac;            /* sommeijer.py:95                          */
    int32_t max_steps;            /* attempted-step budget, <= 0: unlimited   */
    double rho_const;             /* > 0: rho_jac(t, y) = rho_const           */
    xsq_rho_fn rho_cb;            /* else if non-NULL: host callback          */
    void* rho_user;               /* else: nonlinear power iteration (_rho)   */
    const double* t_eval;         /* HOST [n_eval], sorted along direction    */
    int32_t n_eval, reserved;
    double* u_eval;               /* DEVICE [n_eval][rows_local][nx]          */
    double* u_final;              /* DEVICE [rows_local][nx]                  */
    xsq_rkc_result_t* result;     /* HOST                                     */
    const double* pde_params;     /* HOST [n_pde_params], user PDE parameters */
    int32_t n_pde_params, reserved2;
} xsq_rkc_args_t;

/* Register the right-hand side of a 2-D parabolic PDE on the unit square with
 * homogeneous Dirichlet boundaries and a 5-point neighbourhood, as CUDA source
 * defining
 *   __device__ double <entry>(double t, double x, double y, double inv_h2,
 *                             double uc, double un, double us, double uw,
 *                             double ue, const double* p);
 * (n = row above = smaller y).  It is compiled with NVRTC into the fused stage
 * kernels.  Replaces the Python callable `fun` handed to SSV2stab
 * (sommeijer.py:93). */
int xsq_pde_register_source(const char* cuda_src, const char* entry, int32_t n_param,
                            int32_t* pde_out);

/* A GENERAL system y' = f(t, y) for SSV2stab (the reference takes any `fun`,
 * sommeijer.py:93-145; its published tables are a 3-D heat problem and a
 * two-species combustion problem, docs/Demo_SSV2stab.ipynb): CUDA source
 * defining one component of the right-hand side with the whole state in view,
 *   __device__ double <entry>(int i, double t, const double* y, const double* p);
 * Solve it with nx = n_state rounded up to a multiple of 4 (the padding must
 * be zero in u0), rows_local = rows_global = 1, world = 1. */
int xsq_pde_register_vector_source(const char* cuda_src, const char* entry, int32_t n_state,
                                   int32_t n_param, int32_t* pde_out);

/* NCCL communicator for the slab decomposition (new; the reference has no
 * communication layer).  Rank 0 calls xsq_comm_unique_id and distributes the
 * 128 bytes out of band (e.g. torch.distributed.broadcast); every rank then
 * calls xsq_comm_create with its CUDA device current. */
int xsq_comm_unique_id(char id[128]);
int xsq_comm_create(int32_t rank, int32_t world, const char id[128], void** comm);
int xsq_comm_destroy(void* comm);

/* comm may be NULL when world == 1.  Synchronous with respect to the host
 * (the step-size decisions need the error norm); work runs on `stream`. */
int xsq_rkc_solve(const xsq_rkc_args_t* args, void* comm, void* stream);

/* Stage-kernel microbenchmark (roofline of the HBM-streaming stage):
 * average device milliseconds per fused stage on a rows x nx slab. */
int xsq_rkc_stage_bench(int32_t nx, int32_t rows, int32_t reps, double* ms_per_stage,
                        void* stream);
/* The same stage with its stencil operand staged through shared memory by the
 * TMA unit (cp.async.bulk.tensor.2d), an experiment kept for comparison:
 * average milliseconds per stage, and the largest |difference| of one stage
 * against the shipped kernel on the same random slab (must be 0). */
int xsq_rkc_stage_bench_tma(int32_t nx, int32_t rows, int32_t reps, double* ms_per_stage,
                            double* max_abs_diff, void* stream);

/* The library caches its scratch (work queue, init pass, stiffness probe queue)
 * in the device's stream-ordered memory pool between calls.  This returns all
 * of it to the driver (synchronises the device). */
int xsq_trim_memory(int device);

/* Device timing of the kernels of the LAST xsq_rk_solve / xsq_swag_solve issued
 * after xsq_profile_enable(1): CUDA events on the launching stream around the
 * init pass, the persistent kernel and the probe-queue kernel (milliseconds;
 * synchronises with the end of that solve).  For benchmarks: the roofline of
 * the dominant kernel is quoted on ITS duration. */
int xsq_profile_enable(int on);
int xsq_profile_last(double* ms_init, double* ms_main, double* ms_probe);
/* ... and of the solve `back` calls before the last one (0 = the last; the
 * library keeps 8), so that a benchmark can read the kernel times of a timed
 * loop afterwards instead of synchronising inside it. */
int xsq_profile_get(int back, double* ms_init, double* ms_main, double* ms_probe);

/* Kernel-launch bookkeeping for benchmarks: number of kernels this library
 * launched since the last reset. */
int64_t xsq_launch_count(int reset);

/* fp64 FMA peak microbenchmark (roofline denominator for the ensemble
 * kernels): runs `iters` dependent-chain DFMA blocks on all SMs and returns
 * the achieved TFLOP/s in *tflops. */
int xsq_fp64_peak(int device, int32_t iters, double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* XSQ_H */
