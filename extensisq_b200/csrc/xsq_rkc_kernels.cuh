// xsq_rkc_kernels.cuh -- the three SSV2stab kernels that contain the PDE's
// right-hand side, templated on a Pde policy so that a user-supplied RHS
// (xsq_pde_register_source, compiled with NVRTC) inlines into the fused stage
// kernel exactly like the built-in one.  Reference: the user `fun(t, y)` of
// sommeijer.py:93 evaluated inside _stages (:311), at t+h (:214) and in
// __init__/_rho (:139, :372).
//
// This header is also compiled by NVRTC: no host includes.
#pragma once

namespace xsq {
namespace rkc {

constexpr int TX = 32, TY = 8, PX = 4;   // CTA = 32 x 8 threads, 4 points/thread

// One rank's slab of the grid; passed by value to every kernel.
struct Slab {
    int nx, rows;          // points per row, local interior rows
    int row0;              // global index of the first local row
    int pad;
    double inv_h2;         // (nx+1)^2
    double hgrid;          // 1/(nx+1)
    const double* prm;     // device array of PDE parameters (may be null)
    __host__ __device__ size_t n() const { return (size_t)nx * rows; }
    __host__ __device__ size_t n_alloc() const { return (size_t)nx * (rows + 2); }
};

// ---- PDE policies ------------------------------------------------------------
// rhs(t, x, y, inv_h2, uc, un, us, uw, ue, p) -> du/dt at one grid point with
// its 5-point neighbourhood (n = row above, s = row below); Dirichlet-0 values
// outside the unit square.
namespace pde {
// u_t = Lap(u) + u - u^3  
struct Heat2dReaction {
    static constexpr bool VECTOR = false;
    __device__ __forceinline__ static double rhs(double, double, double, double inv_h2,
                                                 double c, double n, double s, double w,
                                                 double e, const double*) {
        const double lap = (((n + s) + (w + e)) - 4.0 * c) * inv_h2;
        return lap + (c - c * c * c);
    }
};
}  // namespace pde

__device__ __forceinline__ void load4(const double* __restrict__ p, double (&v)[PX]) {
    const double2 a = __ldg(reinterpret_cast<const double2*>(p));
    const double2 b = __ldg(reinterpret_cast<const double2*>(p) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
__device__ __forceinline__ void store4(double* __restrict__ p, const double (&v)[PX]) {
    reinterpret_cast<double2*>(p)[0] = make_double2(v[0], v[1]);
    reinterpret_cast<double2*>(p)[1] = make_double2(v[2], v[3]);
}
// Rows above the slab's first row and below its last row are read through
// `up_row` / `dn_row`: the slab's own (zero) ghost row at the domain edge, or --
// multi-GPU -- the neighbour rank's boundary row in ITS memory, mapped with
// CUDA IPC and loaded over NVLink by the threads that need it.  The halo is
// therefore part of the stage kernel; there is no separate exchange step.
__device__ __forceinline__ void load4_peer(const double* p, double (&v)[PX]) {
    const double2 a = __ldcv(reinterpret_cast<const double2*>(p));
    const double2 b = __ldcv(reinterpret_cast<const double2*>(p) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}

// Neighbour barrier, fused into every kernel that reads a halo row.  A stencil
// kernel with sequence number `seq` starts only after everything enqueued before
// it on this rank's stream has finished, so at its start (a) the vector the
// neighbours are about to read is complete and (b) this rank no longer reads
// the buffers they are about to overwrite.  One thread says so -- a release
// store of `seq` into the neighbours' flag words over NVLink -- and only the
// CTAs that hold the slab's first / last row wait (acquire loads of this rank's
// own flag words) until the neighbour has said the same.  Interior CTAs never
// wait: the handshake latency hides behind their work, and there is no separate
// synchronisation kernel between two stages.  The spin is bounded; on timeout
// an error word is set and the host aborts the solve.
struct PeerSync {
    long long seq;             // 0: single rank, nothing to do
    long long* up_remote;      // upper rank's "from below" flag (null at the domain edge)
    long long* dn_remote;      // lower rank's "from above" flag
    long long* mine;           // [0] written by the upper rank, [1] by the lower, [2] error
};
__device__ __forceinline__ void peer_handshake(const PeerSync& ps) {
    if (ps.seq == 0) return;
    const bool first = blockIdx.y == 0, last = blockIdx.y == gridDim.y - 1;
    const bool t0 = threadIdx.x == 0 && threadIdx.y == 0;
    if (t0 && first && blockIdx.x == 0) {
        if (ps.up_remote)
            asm volatile("st.release.sys.global.s64 [%0], %1;" ::"l"(ps.up_remote), "l"(ps.seq) : "memory");
        if (ps.dn_remote)
            asm volatile("st.release.sys.global.s64 [%0], %1;" ::"l"(ps.dn_remote), "l"(ps.seq) : "memory");
    }
    const bool wait_up = first && ps.up_remote != nullptr;
    const bool wait_dn = last && ps.dn_remote != nullptr;
    if (!(wait_up || wait_dn)) return;            // uniform over the CTA
    if (t0) {
        const long long start = clock64(), budget = 4000000000LL;      // ~2 s
        bool ok = true;
        for (int i = 0; i < 2 && ok; ++i) {
            if (!(i == 0 ? wait_up : wait_dn)) continue;
            for (;;) {
                long long v;
                asm volatile("ld.acquire.sys.global.s64 %0, [%1];" : "=l"(v) : "l"(ps.mine + i) : "memory");
                if (v >= ps.seq) break;
                if (clock64() - start > budget) { ok = false; break; }
            }
        }
        if (!ok) ps.mine[2] = ps.seq;
    }
    __syncthreads();
}

// f(t, u) at 4 consecutive points of one row
template <class Pde>
__device__ __forceinline__ void rhs4(const Slab& S, const double* __restrict__ u,
                                     const double* up_row, const double* dn_row, double t,
                                     int row, size_t idx, int col, double (&f)[PX]) {
    if constexpr (Pde::VECTOR) {
        // A general system y' = f(t, y) (the reference takes any `fun`,
        // sommeijer.py:93-145): the state is ONE row of the slab, the policy returns
        // component i with the whole vector in view.  Entries beyond Pde::N pad
        // the row to a multiple of 4 and stay zero.
        const double* y = u + S.nx;                    // skip the ghost row
#pragma unroll
        for (int k = 0; k < PX; ++k)
            f[k] = col + k < Pde::N ? Pde::at(col + k, t, y, S.prm) : 0.0;
        return;
    }
    double c[PX], up[PX], dn[PX];
    load4(u + idx, c);
    if (row == 0) load4_peer(up_row + col, up);
    else load4(u + idx - S.nx, up);
    if (row == S.rows - 1) load4_peer(dn_row + col, dn);
    else load4(u + idx + S.nx, dn);
    const double left = col > 0 ? __ldg(u + idx - 1) : 0.0;
    const double right = col + PX < S.nx ? __ldg(u + idx + PX) : 0.0;
    const double y = (double)(S.row0 + row + 1) * S.hgrid;
#pragma unroll
    for (int k = 0; k < PX; ++k) {
        const double w = k == 0 ? left : c[k - 1];
        const double e = k == PX - 1 ? right : c[k + 1];
        const double x = (double)(col + k + 1) * S.hgrid;
        f[k] = Pde::rhs(t, x, y, S.inv_h2, c[k], up[k], dn[k], w, e, S.prm);
    }
}

#define XSQ_RKC_INDEX                                                     \
    const int col = (blockIdx.x * TX + threadIdx.x) * PX;                 \
    const int row = blockIdx.y * TY + threadIdx.y;                        \
    const bool active = col < S.nx && row < S.rows;                       \
    const size_t idx = (size_t)(row + 1) * S.nx + col;

__device__ __forceinline__ void block_sum_to(double s, double* __restrict__ partial) {
    __shared__ double sm[TX * TY / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const int tid = threadIdx.y * TX + threadIdx.x;
    if ((tid & 31) == 0) sm[tid >> 5] = s;
    __syncthreads();
    if (tid == 0) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < TX * TY / 32; ++w) t += sm[w];
        partial[blockIdx.y * gridDim.x + blockIdx.x] = t;
    }
}

// dy = f(t, u)
template <class Pde>
__device__ __forceinline__ void eval_body(const Slab& S, const PeerSync& ps,
                                          const double* __restrict__ u,
                                          const double* up_row, const double* dn_row, double t,
                                          double* __restrict__ dy) {
    peer_handshake(ps);
    XSQ_RKC_INDEX
    if (!active) return;
    double f[PX];
    rhs4<Pde>(S, u, up_row, dn_row, t, row, idx, col, f);
    store4(dy + idx, f);
}

// Stage j >= 2 (sommeijer.py:311-313), fused with the RHS evaluation:
//   Y_j = mu*Y_{j-1} + nu*Y_{j-2} + (1-mu-nu)*y_n + h*mus*(f(t_j, Y_{j-1}) - a_{j-1}*f_n)
template <class Pde>
__device__ __forceinline__ void stage_body(const Slab& S, const PeerSync& ps,
                                           const double* __restrict__ yjm1,
                                           const double* up_row, const double* dn_row,
                                           const double* __restrict__ yjm2,
                                           const double* __restrict__ yn,
                                           const double* __restrict__ fn, double* __restrict__ yj,
                                           double t, double mu, double nu, double c3,
                                           double hmus, double ajm1) {
    peer_handshake(ps);
    XSQ_RKC_INDEX
    if (!active) return;
    double f[PX], a[PX], b[PX], c[PX], d[PX], o[PX];
    rhs4<Pde>(S, yjm1, up_row, dn_row, t, row, idx, col, f);
    load4(yjm1 + idx, a);
    load4(yjm2 + idx, b);
    load4(yn + idx, c);
    load4(fn + idx, d);
#pragma unroll
    for (int k = 0; k < PX; ++k)
        o[k] = ((mu * a[k] + nu * b[k]) + c3 * c[k]) + hmus * (f[k] - ajm1 * d[k]);
    store4(yj + idx, o);
}

// Final evaluation fused with the error estimate (sommeijer.py:214-220):
//   f1 = f(t+h, y);  est = 0.8*(yn - y) + 0.4*h*(fn + f1);
//   wt = atol + rtol*max(|y|,|yn|);  partial[block] = sum (est/wt)^2
template <class Pde>
__device__ __forceinline__ void final_body(const Slab& S, const PeerSync& ps,
                                           const double* __restrict__ y,
                                           const double* up_row, const double* dn_row,
                                           const double* __restrict__ yn,
                                           const double* __restrict__ fn, double* __restrict__ f1,
                                           double t, double h, double rtol, double atol,
                                           double* __restrict__ partial) {
    peer_handshake(ps);
    XSQ_RKC_INDEX
    double s = 0.0;
    if (active) {
        double f[PX], a[PX], b[PX], c[PX];
        rhs4<Pde>(S, y, up_row, dn_row, t, row, idx, col, f);
        load4(y + idx, a);
        load4(yn + idx, b);
        load4(fn + idx, c);
        store4(f1 + idx, f);
        const double h04 = 0.4 * h;
#pragma unroll
        for (int k = 0; k < PX; ++k) {
            const double est = 0.8 * (b[k] - a[k]) + h04 * (c[k] + f[k]);
            const double wt = atol + rtol * fmax(fabs(a[k]), fabs(b[k]));
            const double q = est / wt;
            s = fma(q, q, s);
        }
    }
    block_sum_to(s, partial);
}

template <class Pde>
__global__ void __launch_bounds__(TX* TY)
    k_eval(Slab S, PeerSync ps, const double* u, const double* up_row, const double* dn_row,
           double t, double* dy) {
    eval_body<Pde>(S, ps, u, up_row, dn_row, t, dy);
}
template <class Pde>
__global__ void __launch_bounds__(TX* TY)
    k_stage(Slab S, PeerSync ps, const double* yjm1, const double* up_row, const double* dn_row,
            const double* yjm2, const double* yn, const double* fn, double* yj, double t,
            double mu, double nu, double c3, double hmus, double ajm1) {
    stage_body<Pde>(S, ps, yjm1, up_row, dn_row, yjm2, yn, fn, yj, t, mu, nu, c3, hmus, ajm1);
}
template <class Pde>
__global__ void __launch_bounds__(TX* TY)
    k_final(Slab S, PeerSync ps, const double* y, const double* up_row, const double* dn_row,
            const double* yn, const double* fn, double* f1, double t, double h, double rtol,
            double atol, double* partial) {
    final_body<Pde>(S, ps, y, up_row, dn_row, yn, fn, f1, t, h, rtol, atol, partial);
}

}  // namespace rkc
}  // namespace xsq
