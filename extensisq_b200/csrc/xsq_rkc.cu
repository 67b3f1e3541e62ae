// xsq_rkc.cu -- SSV2stab (Sommeijer-Shampine-Verwer Runge-Kutta-Chebyshev) for
// ONE large semi-discretised parabolic PDE, row-slab decomposed over the GPUs
// of a box.
//
// Reference (file:line in /root/reference/extensisq/sommeijer.py):
//   SSV2stab.__init__      :93-145   host scalars in rkc_solve()
//   _init_step_size        :147-160
//   _step_impl (RKCLOW)    :162-271  host loop in rkc_solve(): every decision is
//                                    scalar and is taken redundantly and
//                                    identically on every rank
//   _stages (STEP)         :273-329  one fused kernel per stage: 5-point
//                                    stencil RHS + three-term recurrence
//   _rho (RKCRHO)          :331-398  power iteration with device norms
//   _dense_output_impl     :400-406  cubic Hermite (common.py:793-821)
//
// Data layout in HBM: each vector is a slab of (rows + 2) x nx doubles; row 0
// and row rows+1 are ghost rows (the neighbour rank's boundary row, or the
// Dirichlet zero at the domain edge).  Per stage and grid point the kernel
// reads Y_{j-1} (stencil: centre from HBM, neighbours from L1/L2), Y_{j-2},
// y_n, f_n and writes Y_j: 40 B of algorithmic HBM traffic.  The reference's
// two full-vector copies per stage (sommeijer.py:318-319) become pointer
// rotation.  Multi-GPU: before each stage the two boundary rows of Y_{j-1}
// travel to the neighbours (NCCL send/recv over NVLink, 2 x nx doubles each
// way); per step attempt one all-gather of a scalar per rank gives the
// error norm, summed in rank order so every rank takes the same decision.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <ctime>
#include <string>
#include <vector>

#include "xsq.h"
#include "xsq_comm.h"
#include "xsq_user.h"   // set_detail, count_launch
#include "xsq_rkc_kernels.cuh"
#include "xsq_rkc_tma.cuh"

namespace xsq {

void keep_pool_memory();   // xsq_api.cu

namespace {

constexpr double kURound = 0x1.0000000000001p-53;
constexpr double kSqrtTiny = 0x1.0p-511;
using namespace rkc;

// Neighbour handshake (one thread): tell both neighbours that every kernel
// enqueued before this one has finished (so the buffer about to be read is
// complete and the buffer about to be overwritten is no longer being read),
// then wait for the same message from them.  Flags live in the RECEIVER's
// memory (remote store, local spin).  The spin is bounded: on timeout an
// error word is set and the host aborts the solve instead of hanging the GPU.
// the neighbour barrier on its own (end of a solve): a one-CTA grid is both the
// first and the last block row, so it posts and waits for both neighbours
__global__ void k_peer_barrier(PeerSync ps) { peer_handshake(ps); }

// out = a + s * b      (first stage, sommeijer.py:289; step-size probe :152)
__global__ void __launch_bounds__(TX* TY) k_axpy(Slab S, const double* __restrict__ a,
                                                 const double* __restrict__ b, double s,
                                                 double* __restrict__ out) {
    XSQ_RKC_INDEX
    if (!active) return;
    double va[PX], vb[PX], o[PX];
    load4(a + idx, va);
    load4(b + idx, vb);
#pragma unroll
    for (int k = 0; k < PX; ++k) o[k] = va[k] + s * vb[k];
    store4(out + idx, o);
}

// partial[block] = sum over the block of g(a, b)^2 with
//   MODE 0: a                      (2-norm of a)
//   MODE 1: a - b                  (2-norm of a difference)
//   MODE 2: (a - b) / (atol + rtol*|c|)   (step-size probe, sommeijer.py:154-155)
template <int MODE>
__global__ void __launch_bounds__(TX* TY)
    k_sumsq(Slab S, const double* __restrict__ a, const double* __restrict__ b,
            const double* __restrict__ c, double rtol, double atol,
            double* __restrict__ partial) {
    XSQ_RKC_INDEX
    double s = 0.0;
    if (active) {
        double va[PX], vb[PX], vc[PX];
        load4(a + idx, va);
        if (MODE >= 1) load4(b + idx, vb);
        if (MODE == 2) load4(c + idx, vc);
#pragma unroll
        for (int k = 0; k < PX; ++k) {
            double q = va[k];
            if (MODE >= 1) q -= vb[k];
            if (MODE == 2) q /= atol + rtol * fabs(vc[k]);
            s = fma(q, q, s);
        }
    }
    block_sum_to(s, partial);
}

// deterministic second pass: one CTA adds the block partials in a fixed order
__global__ void __launch_bounds__(256) k_reduce(const double* __restrict__ partial, int n,
                                                double* __restrict__ out) {
    __shared__ double sm[256];
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) s += partial[i];
    sm[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = sm[0];
}

// out = a + (b - c) * s   (power iteration update, sommeijer.py:389) and
// out = a + b * s (c == nullptr, :354);  MODE 3: out = b * s (:357, :360)
__global__ void __launch_bounds__(TX* TY)
    k_combine(Slab S, const double* __restrict__ a, const double* __restrict__ b,
              const double* __restrict__ c, double s, int mode, double* __restrict__ out) {
    XSQ_RKC_INDEX
    if (!active) return;
    double va[PX], vb[PX], vc[PX], o[PX];
    if (mode != 4) load4(b + idx, vb);
    if (mode != 3 && mode != 4) load4(a + idx, va);
    if (mode == 1) load4(c + idx, vc);
#pragma unroll
    for (int k = 0; k < PX; ++k) {
        if (mode == 1) o[k] = va[k] + (vb[k] - vc[k]) * s;
        else if (mode == 0) o[k] = va[k] + vb[k] * s;
        else if (mode == 2) o[k] = va[k] - vb[k];          // V = v - yn (a=v, b=yn)
        else if (mode == 3) o[k] = vb[k] * s;
        else o[k] = s;                                     // constant fill
    }
    store4(out + idx, o);
}

// v[index] = -v[index]  (sommeijer.py:391-395, degenerate power iteration)
__global__ void k_flip(double* __restrict__ v, size_t padded_index) {
    v[padded_index] = -v[padded_index];
}

// cubic Hermite dense output at one point (common.py:806-816)
__global__ void __launch_bounds__(TX* TY)
    k_hermite(Slab S, const double* __restrict__ y_old, const double* __restrict__ f_old,
              const double* __restrict__ y, const double* __restrict__ f, double h00,
              double h10, double h01, double h11, double* __restrict__ out_compact) {
    XSQ_RKC_INDEX
    if (!active) return;
    double a[PX], b[PX], c[PX], d[PX], o[PX];
    load4(y_old + idx, a);
    load4(f_old + idx, b);
    load4(y + idx, c);
    load4(f + idx, d);
#pragma unroll
    for (int k = 0; k < PX; ++k) o[k] = ((h00 * a[k] + h10 * b[k]) + h01 * c[k]) + h11 * d[k];
    store4(out_compact + (size_t)row * S.nx + col, o);
}

// compact [rows][nx]  <->  padded [(rows+2)][nx]
__global__ void __launch_bounds__(TX* TY)
    k_copy(Slab S, const double* __restrict__ src, double* __restrict__ dst, int to_padded) {
    XSQ_RKC_INDEX
    if (!active) return;
    double v[PX];
    const size_t cidx = (size_t)row * S.nx + col;
    load4(src + (to_padded ? cidx : idx), v);
    store4(dst + (to_padded ? idx : cidx), v);
}

// The three kernels that contain the PDE right-hand side: the built-in PDE is
// launched through the runtime API, a user PDE (NVRTC module) through the
// driver API (xsq_user.cpp).
struct PdeLaunch {
    void* user_fn[3] = {nullptr, nullptr, nullptr};   // CUfunction eval/stage/final
};

struct Ctx {
    Slab S;
    PdeLaunch pl;
    dim3 grid, block;
    int nblocks;
    cudaStream_t st;
    Comm* comm;
    int rank, world;
    double* partial;      // [nblocks]
    double* scalar_dev;   // [world] (all-gather target)
    double* scalar_host;  // pinned [world]
    long long n_total;
    int64_t launches = 0;
    int rc = XSQ_OK;

    void fail(const char* what) {
        cudaError_t e = cudaGetLastError();
        if (rc == XSQ_OK) {
            rc = XSQ_ERR_CUDA;
            set_detail(std::string(what) + ": " + cudaGetErrorString(e));
        }
    }
    void launched() { count_launch(); ++launches; }

    // ---- peer (NVLink) halo -------------------------------------------------
    double* base = nullptr;            // start of this rank's vector storage
    const double* peer_up = nullptr;   // IPC mapping of the upper rank's storage
    const double* peer_dn = nullptr;
    int rows_up = 0;                   // interior rows of the upper rank's slab
    PeerSync ps = {0, nullptr, nullptr, nullptr};   // barrier carried by the next stencil kernel
    long long* flags = nullptr;        // [0] from upper rank, [1] from lower, [2] error
    long long* up_flag_remote = nullptr;   // upper rank's flags[1]
    long long* dn_flag_remote = nullptr;   // lower rank's flags[0]
    long long seq = 0;
    const double* up_row = nullptr;    // set by halo() for the vector about to be read
    const double* dn_row = nullptr;

    // Prepare the stencil read of the padded vector u: at a domain edge the
    // neighbour row is u's own zero ghost row; between ranks it is the
    // neighbour's boundary row of the SAME vector (identical layout and
    // rotation state on every rank), read in place over NVLink after a
    // neighbour handshake.  No data is copied.
    void halo(const double* u) {
        const size_t nx = S.nx;
        up_row = u;                                   // top ghost row
        dn_row = u + nx * (S.rows + 1);               // bottom ghost row
        if (world == 1) return;
        const size_t off = (size_t)(u - base);
        if (peer_up) up_row = peer_up + off + nx * rows_up;   // its last interior row
        if (peer_dn) dn_row = peer_dn + off + nx;             // its first interior row
        ++seq;                 // the next stencil kernel carries the barrier (peer_handshake)
        ps = PeerSync{seq, up_flag_remote, dn_flag_remote, flags};
    }

    // global sum of this rank's block partials: deterministic on every rank
    double global_sum() {
        k_reduce<<<1, 256, 0, st>>>(partial, nblocks, scalar_dev + rank);
        launched();
        if (world > 1 && comm) {
            if (comm_allgather1(comm, scalar_dev + rank, scalar_dev, st) != 0 && rc == XSQ_OK) {
                rc = XSQ_ERR_CUDA;
                set_detail("NCCL all-gather failed");
            }
        }
        if (cudaMemcpyAsync(scalar_host, scalar_dev, sizeof(double) * world,
                            cudaMemcpyDeviceToHost, st) != cudaSuccess ||
            cudaStreamSynchronize(st) != cudaSuccess) {
            fail("global_sum");
            return NAN;
        }
        if (world > 1 && flags) {
            long long e = 0;
            cudaMemcpy(&e, flags + 2, sizeof e, cudaMemcpyDeviceToHost);
            if (e != 0 && rc == XSQ_OK) {
                rc = XSQ_ERR_CUDA;
                set_detail("rkc: neighbour handshake timed out (peer rank stalled or failed)");
            }
        }
        double s = 0.0;
        for (int r = 0; r < world; ++r) s += scalar_host[r];
        return s;
    }

    double t_eval_arg = 0.0;                     // time argument of the next eval()
    void launch_eval(const double* u, double t, double* dy) {
        if (pl.user_fn[0]) {
            void* args[] = {&S, &ps, &u, &up_row, &dn_row, &t, &dy};
            if (user_launch(pl.user_fn[0], grid.x, grid.y, block.x, block.y, args, st) != 0) fail("user pde eval");
        } else {
            k_eval<pde::Heat2dReaction><<<grid, block, 0, st>>>(S, ps, u, up_row, dn_row, t, dy);
        }
        launched();
    }
    void launch_stage(const double* yjm1, const double* yjm2, const double* yn, const double* fn,
                      double* yj, double t, double mu, double nu, double c3, double hmus,
                      double ajm1) {
        if (pl.user_fn[1]) {
            void* args[] = {&S, &ps, &yjm1, &up_row, &dn_row, &yjm2, &yn, &fn, &yj, &t, &mu, &nu, &c3, &hmus, &ajm1};
            if (user_launch(pl.user_fn[1], grid.x, grid.y, block.x, block.y, args, st) != 0) fail("user pde stage");
        } else {
            k_stage<pde::Heat2dReaction><<<grid, block, 0, st>>>(S, ps, yjm1, up_row, dn_row, yjm2, yn, fn,
                                                                yj, t, mu, nu, c3, hmus, ajm1);
        }
        launched();
    }
    void launch_final(const double* y, const double* yn, const double* fn, double* f1, double t,
                      double h, double rtol, double atol) {
        double* part = partial;
        if (pl.user_fn[2]) {
            void* args[] = {&S, &ps, &y, &up_row, &dn_row, &yn, &fn, &f1, &t, &h, &rtol, &atol, &part};
            if (user_launch(pl.user_fn[2], grid.x, grid.y, block.x, block.y, args, st) != 0) fail("user pde final");
        } else {
            k_final<pde::Heat2dReaction><<<grid, block, 0, st>>>(S, ps, y, up_row, dn_row, yn, fn, f1, t, h,
                                                                rtol, atol, part);
        }
        launched();
    }
    void eval(double* u, double* dy) {           // dy = f(t_eval_arg, u), with halo
        halo(u);
        launch_eval(u, t_eval_arg, dy);
    }
    double norm2(const double* a) {
        k_sumsq<0><<<grid, block, 0, st>>>(S, a, nullptr, nullptr, 0, 0, partial);
        launched();
        return std::sqrt(global_sum());
    }
    double norm2_diff(const double* a, const double* b) {
        k_sumsq<1><<<grid, block, 0, st>>>(S, a, b, nullptr, 0, 0, partial);
        launched();
        return std::sqrt(global_sum());
    }
};

}  // namespace

int rkc_solve(const xsq_rkc_args_t* A, Comm* comm, cudaStream_t st) {
    if (!A || A->struct_size != (int32_t)sizeof(xsq_rkc_args_t)) {
        set_detail("xsq_rkc_args_t.struct_size mismatch");
        return XSQ_ERR_ARG;
    }
    PdeLaunch pl;
    int n_pde_param_expected = 0, n_vector = 0;
    if (A->pde >= XSQ_PDE_USER_BASE) {
        int rc = user_pde_kernels(A->pde, pl.user_fn, &n_pde_param_expected);
        if (rc != XSQ_OK) return rc;
        n_vector = user_pde_vector_size(A->pde);
        if (n_vector > 0 && (A->rows_local != 1 || A->rows_global != 1 || A->world != 1 ||
                             A->nx < n_vector)) {
            set_detail("rkc: a general system is one row of nx >= n entries on one GPU");
            return XSQ_ERR_ARG;
        }
    } else if (A->pde != XSQ_PDE_HEAT2D_REACTION) {
        set_detail("unknown pde");
        return XSQ_ERR_UNSUPPORTED;
    }
    if (A->n_pde_params != n_pde_param_expected || (A->n_pde_params > 0 && !A->pde_params)) {
        set_detail("rkc: pde_params does not match the registered PDE");
        return XSQ_ERR_ARG;
    }
    if (A->nx < 4 || A->nx % 4 != 0 || A->rows_local < 1 || A->rows_global < A->rows_local ||
        !A->u0 || !A->u_final || !A->result) {
        set_detail("rkc: bad grid (nx must be a positive multiple of 4) or NULL pointer");
        return XSQ_ERR_ARG;
    }
    if (!(A->rtol >= 0) || !(A->atol >= 0)) { set_detail("`rtol`/`atol` must be positive."); return XSQ_ERR_ARG; }
    if (!(A->max_step > 0)) { set_detail("`max_step` must be positive."); return XSQ_ERR_ARG; }
    if (A->world > 1 && !comm) { set_detail("rkc: world > 1 needs a communicator"); return XSQ_ERR_ARG; }
    const double t0 = A->t0, tf = A->t_bound;
    if (A->first_step > 0 && A->first_step > std::fabs(tf - t0)) {
        set_detail("`first_step` exceeds bounds.");
        return XSQ_ERR_ARG;
    }
    Ctx C;
    C.pl = pl;
    C.S.nx = A->nx;
    C.S.rows = A->rows_local;
    C.S.row0 = A->row0;
    C.S.pad = 0;
    C.S.inv_h2 = ((double)A->nx + 1.0) * ((double)A->nx + 1.0);
    C.S.hgrid = 1.0 / ((double)A->nx + 1.0);
    C.S.prm = nullptr;
    C.block = dim3(TX, TY);
    C.grid = dim3((A->nx / PX + TX - 1) / TX, (A->rows_local + TY - 1) / TY);
    C.nblocks = C.grid.x * C.grid.y;
    C.st = st;
    C.comm = comm;
    C.rank = A->world > 1 ? A->rank : 0;
    C.world = A->world > 1 ? A->world : 1;
    C.n_total = (long long)A->nx * A->rows_global;
    if (n_vector > 0) C.n_total = n_vector;      // padding is not part of the norm
    const size_t na = C.S.n_alloc();
    // device scratch: yn, fn, w0, w1, w2, V + partials + scalars
    keep_pool_memory();
    const bool dbg0 = getenv("XSQ_RKC_DEBUG") != nullptr;
    auto wall0 = []() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; };
    const double t_enter = wall0();
    double* buf = nullptr;
    const size_t total = na * 6 + C.nblocks + C.world + 8 + 64 + (size_t)A->n_pde_params;
    const bool multi = C.world > 1;
    const size_t flag_off = na * 6 + C.nblocks + C.world + 8;
    CommWorkspace* ws = multi ? comm_workspace(comm) : nullptr;
    if (multi && !ws) { set_detail("rkc: world > 1 needs a communicator"); return XSQ_ERR_ARG; }
    auto release = [&]() {        // single rank: back to the pool; multi: kept in the workspace
        if (!multi && buf) cudaFreeAsync(buf, st);
    };
    if (!multi) {
        if (cudaMallocAsync((void**)&buf, total * sizeof(double), st) != cudaSuccess) {
            set_detail("rkc: out of device memory");
            return XSQ_ERR_NOMEM;
        }
        cudaMemsetAsync(buf, 0, total * sizeof(double), st);      // ghost rows = Dirichlet 0
    } else {
        // Storage exported with CUDA IPC (plain cudaMalloc), mapped by the two
        // neighbours, and KEPT in the communicator for the next solve of the same
        // shape.  Whether to rebuild is decided collectively: every rank
        // contributes "my shape changed", any one of them makes all rebuild (a
        // stale mapping of a neighbour's freed storage must never be used).
        const long long key[6] = {A->nx, A->rows_global, A->rows_local, C.world, C.rank,
                                  (long long)A->n_pde_params};
        const bool mine_changed = !ws->buf || ws->doubles != total ||
                                  std::memcmp(ws->key, key, sizeof key) != 0;
        struct Card { cudaIpcMemHandle_t h; long long rows; long long changed; };
        static_assert(sizeof(Card) == 80, "card");
        bool ok = true;
        if (!ws->xchg) ok = cudaMalloc((void**)&ws->xchg, sizeof(Card) * (C.world + 1)) == cudaSuccess;
        std::vector<Card> cards(C.world);
        auto exchange = [&](const Card& mine) {
            bool k = cudaMemcpyAsync(ws->xchg, &mine, sizeof(Card), cudaMemcpyHostToDevice, st) == cudaSuccess;
            k = k && comm_allgather_bytes(comm, ws->xchg, ws->xchg + sizeof(Card), sizeof(Card), st) == 0;
            k = k && cudaMemcpyAsync(cards.data(), ws->xchg + sizeof(Card), sizeof(Card) * C.world,
                                     cudaMemcpyDeviceToHost, st) == cudaSuccess;
            return k && cudaStreamSynchronize(st) == cudaSuccess;
        };
        Card vote;
        std::memset(&vote, 0, sizeof vote);
        vote.changed = mine_changed ? 1 : 0;
        ok = ok && exchange(vote);
        bool rebuild = false;
        for (int r = 0; ok && r < C.world; ++r) rebuild = rebuild || cards[r].changed != 0;
        if (ok && rebuild) {
            comm_workspace_release(ws);
            ok = cudaMalloc((void**)&ws->xchg, sizeof(Card) * (C.world + 1)) == cudaSuccess;
            if (ok && cudaMalloc((void**)&ws->buf, total * sizeof(double)) != cudaSuccess) {
                set_detail("rkc: out of device memory");
                return XSQ_ERR_NOMEM;
            }
            ws->doubles = total;
            std::memcpy(ws->key, key, sizeof key);
            ok = ok && cudaMemsetAsync(ws->buf, 0, total * sizeof(double), st) == cudaSuccess;
            Card mine;
            std::memset(&mine, 0, sizeof mine);
            ok = ok && cudaIpcGetMemHandle(&mine.h, ws->buf) == cudaSuccess;
            mine.rows = A->rows_local;
            ok = ok && exchange(mine);
            void* p = nullptr;
            if (ok && C.rank > 0) {
                ok = cudaIpcOpenMemHandle(&p, cards[C.rank - 1].h, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess;
                ws->peer_up = (const double*)p;
                ws->rows_up = (int)cards[C.rank - 1].rows;
            }
            if (ok && C.rank + 1 < C.world) {
                ok = cudaIpcOpenMemHandle(&p, cards[C.rank + 1].h, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess;
                ws->peer_dn = (const double*)p;
            }
        }
        if (!ok) {
            set_detail(std::string("rkc: peer storage set-up failed: ") + cudaGetErrorString(cudaGetLastError()));
            comm_workspace_release(ws);
            return XSQ_ERR_CUDA;
        }
        buf = ws->buf;
        C.peer_up = ws->peer_up;
        C.peer_dn = ws->peer_dn;
        C.rows_up = ws->rows_up;
        C.seq = ws->seq;           // flags hold the last sequence number: keep counting
        if (C.peer_up) C.up_flag_remote = (long long*)(C.peer_up + flag_off) + 1;
        if (C.peer_dn) C.dn_flag_remote = (long long*)(C.peer_dn + flag_off) + 0;
    }
    C.base = buf;
    C.flags = reinterpret_cast<long long*>(buf + flag_off);
    if (multi) cudaMemsetAsync(C.flags + 2, 0, sizeof(long long), st);   // the time-out word
    double *yn = buf, *fn = buf + na, *w0 = buf + 2 * na, *w1 = buf + 3 * na,
           *w2 = buf + 4 * na, *V = buf + 5 * na;
    C.partial = buf + 6 * na;
    C.scalar_dev = C.partial + C.nblocks;
    if (cudaMallocHost((void**)&C.scalar_host, sizeof(double) * C.world) != cudaSuccess) {
        release();
        return XSQ_ERR_NOMEM;
    }
    xsq_rkc_result_t* R = A->result;
    std::memset(R, 0, sizeof(*R));

    // ---- SSV2stab.__init__, sommeijer.py:118-145 ---------------------------
    const double rtol = std::fmin(std::fmax(A->rtol, 0x1.4p-50), 0.1);
    const double atol = std::fmax(A->atol, kSqrtTiny);
    const double direction = (tf != t0) ? (tf > t0 ? 1.0 : -1.0) : 1.0;
    const double sqrtu = std::sqrt(kURound);
    int mmax = (int)std::nearbyint(std::sqrt(rtol / (10.0 * kURound)));
    if (mmax < 2) mmax = 2;
    bool newspc = true, jacatt = false, have_V = false, have_absh = false, have_hold = false;
    double max_step = std::fmin(std::fmin(A->max_step, std::fabs(tf - t0)), 0x1.fffffffffffffp+511);
    double hmin0 = std::fabs(t0);
    if (tf != INFINITY) hmin0 = std::fmax(hmin0, std::fabs(max_step));
    hmin0 = std::fmax(kSqrtTiny, 10.0 * kURound * hmin0);
    double t = t0, absh = 0.0, hold = 0.0, errold = 0.0, sprad = 0.0;
    if (A->first_step > 0) { absh = A->first_step; have_absh = true; }
    int nstsig = 0, n_acc = 0, n_rej = 0, nfev = 0, nfesig = 0, maxm = 0, status = 1;
    const int max_steps = A->max_steps > 0 ? A->max_steps : 2147483647;
    const double one3rd = 1.0 / 3.0, two3rd = 2.0 / 3.0;

    k_copy<<<C.grid, C.block, 0, st>>>(C.S, A->u0, yn, 1);
    C.launched();
    if (A->n_pde_params > 0) {       // PDE parameters live behind the flag words
        double* dprm = buf + na * 6 + C.nblocks + C.world + 8 + 64;
        cudaMemcpyAsync(dprm, A->pde_params, sizeof(double) * A->n_pde_params,
                        cudaMemcpyHostToDevice, st);
        C.S.prm = dprm;
    }
    C.t_eval_arg = t0;
    C.eval(yn, fn);
    nfev = 1;
    int ieval = 0;
    auto emit_point = [&](double t_old, double t_new, const double* y_old, const double* f_old,
                          double te, double* out) {
        const double hh = t_new - t_old, x = (te - t_old) / hh, omx = 1.0 - x;
        const double h00 = (1.0 + 2.0 * x) * (omx * omx), h10 = x * (omx * omx) * hh;
        const double h01 = (x * x) * (3.0 - 2.0 * x), h11 = (x * x) * (x - 1.0) * hh;
        k_hermite<<<C.grid, C.block, 0, st>>>(C.S, y_old, f_old, yn, fn, h00, h10, h01, h11, out);
        C.launched();
    };
    if (t0 == tf) {                                   // scipy base.py:195-200
        for (; ieval < A->n_eval; ++ieval) {
            k_copy<<<C.grid, C.block, 0, st>>>(C.S, yn, A->u_eval + (size_t)ieval * C.S.n(), 0);
            C.launched();
        }
        status = 0;
    }
    double* W[3] = {w0, w1, w2};                     // rotating work vectors
    const bool dbg = getenv("XSQ_RKC_DEBUG") != nullptr;
    auto wall_ms = []() {
        timespec ts;
        clock_gettime(CLOCK_MONOTONIC, &ts);
        return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
    };
    double t_dbg = wall_ms();
    if (dbg) { cudaStreamSynchronize(st); std::fprintf(stderr, "[rkc r%d] setup %.3f ms\n", C.rank, wall_ms() - t_enter); t_dbg = wall_ms(); }

    while (status == 1 && C.rc == XSQ_OK) {
        // ---------------- one step (sommeijer.py:162-271) -------------------
        double h = 0.0, hmin = 0.0, err = 0.0;
        const double t_old = t;
        for (;;) {
            if (n_acc + n_rej >= max_steps) { status = XSQ_LANE_STEP_BUDGET; break; }
            if (newspc) {
                if (A->rho_const > 0.0) {
                    sprad = A->rho_const;
                } else if (A->rho_cb) {
                    sprad = A->rho_cb(t, A->rho_user);
                } else {
                    // ---- _rho, sommeijer.py:331-398 ------------------------
                    double* v = W[0];
                    double* fv = W[1];
                    const double small = 1.0 / max_step;
                    if (!have_V) {
                        cudaMemcpyAsync(V, fn, na * sizeof(double), cudaMemcpyDeviceToDevice, st);
                        have_V = true;
                    }
                    const double ynrm = C.norm2(yn), vnrm = C.norm2(V);
                    double dynrm;
                    if (ynrm != 0.0 && vnrm != 0.0) {
                        dynrm = ynrm * sqrtu;
                        k_combine<<<C.grid, C.block, 0, st>>>(C.S, yn, V, nullptr, dynrm / vnrm, 0, v);
                    } else if (ynrm != 0.0) {
                        dynrm = ynrm * sqrtu;
                        k_combine<<<C.grid, C.block, 0, st>>>(C.S, nullptr, V, nullptr, 1.0 + sqrtu, 3, v);
                    } else if (vnrm != 0.0) {
                        dynrm = kURound;
                        k_combine<<<C.grid, C.block, 0, st>>>(C.S, nullptr, V, nullptr, dynrm / vnrm, 3, v);
                    } else {
                        dynrm = kURound;
                        k_combine<<<C.grid, C.block, 0, st>>>(C.S, nullptr, nullptr, nullptr, dynrm, 4, v);
                    }
                    C.launched();
                    double sigma = 0.0;
                    bool converged = false;
                    for (int iter = 0; iter < 50; ++iter) {
                        C.t_eval_arg = t;
                        C.eval(v, fv);
                        ++nfesig;
                        const double dfnrm = C.norm2_diff(fv, fn);
                        const double sigmal = sigma;
                        sigma = dfnrm / dynrm;
                        sprad = 1.2 * sigma;
                        if (C.rc != XSQ_OK) break;
                        if (iter && std::fabs(sigma - sigmal) <= std::fmax(sigma, small) * 0.01) {
                            k_combine<<<C.grid, C.block, 0, st>>>(C.S, v, yn, nullptr, 0.0, 2, V);
                            C.launched();
                            converged = true;
                            break;
                        }
                        if (dfnrm != 0.0) {
                            k_combine<<<C.grid, C.block, 0, st>>>(C.S, yn, fv, fn, dynrm / dfnrm, 1, v);
                            C.launched();
                        } else {
                            // flip the sign of global component iter % n
                            const long long gi = iter % C.n_total;
                            const long long grow = gi / C.S.nx, gcol = gi % C.S.nx;
                            const long long lrow = grow - A->row0;
                            if (lrow >= 0 && lrow < C.S.rows) {
                                k_flip<<<1, 1, 0, st>>>(v, (size_t)(lrow + 1) * C.S.nx + gcol);
                                C.launched();
                            }
                        }
                    }
                    if (!converged) { status = XSQ_LANE_SPRAD_FAILED; break; }
                }
                jacatt = true;
            }
            if (!have_absh) {                         // _init_step_size, :147-160
                absh = max_step;
                if (sprad * absh > 1.0) absh = 1.0 / sprad;
                absh = std::fmax(absh, hmin0);
                k_axpy<<<C.grid, C.block, 0, st>>>(C.S, yn, fn, absh, W[0]);
                C.launched();
                C.t_eval_arg = t + absh;
                C.eval(W[0], W[1]);
                ++nfev;
                k_sumsq<2><<<C.grid, C.block, 0, st>>>(C.S, W[1], fn, yn, rtol, atol, C.partial);
                C.launched();
                const double est = absh * std::sqrt(C.global_sum() / (double)C.n_total);
                if (0.1 * absh < max_step * std::sqrt(est))
                    absh = std::fmax(0.1 * absh / std::sqrt(est), hmin0);
                else
                    absh = max_step;
                have_absh = true;
            }
            if (1.1 * absh >= std::fabs(tf - t)) absh = std::fabs(tf - t);
            int m = 1 + (int)std::sqrt(1.54 * absh * sprad + 1.0);
            if (m > mmax) {
                m = mmax;
                absh = ((double)m * m - 1) / (1.54 * sprad);
            }
            if (m > maxm) maxm = m;
            h = direction * absh;
            hmin = std::fmax(kSqrtTiny, 13.3 * kURound * (std::fabs(t) + absh) * ((double)m * m - 1));

            // ---- _stages, sommeijer.py:273-329 -----------------------------
            const double w0c = 1.0 + 2.0 / (13.0 * ((double)m * m));
            const double temp1 = w0c * w0c - 1.0, temp2 = std::sqrt(temp1);
            const double arg = m * std::log(w0c + temp2);
            const double w1c = std::sinh(arg) * temp1 /
                               (std::cosh(arg) * m * temp2 - w0c * std::sinh(arg));
            double bjm1 = 1.0 / ((2.0 * w0c) * (2.0 * w0c)), bjm2 = bjm1;
            double mus = w1c * bjm1;
            // Y_0 = y_n is read in place (the reference copies it, :287);
            // Y_1 = y_n + h*mus*f_n.  i1: slot of Y_{j-1}, i0: slot Y_j is
            // written to, i2: slot of Y_{j-2} (-1: it is y_n itself).
            int i1 = 0, i0 = 1, i2 = -1;
            k_axpy<<<C.grid, C.block, 0, st>>>(C.S, yn, fn, h * mus, W[i1]);
            C.launched();
            double thjm2 = 0.0, thjm1 = mus, zjm1 = w0c, zjm2 = 1.0, dzjm1 = 1.0, dzjm2 = 0.0,
                   d2zjm1 = 0.0, d2zjm2 = 0.0;
            for (int j = 2; j <= m; ++j) {
                const double zj = 2.0 * w0c * zjm1 - zjm2;
                const double dzj = 2.0 * w0c * dzjm1 - dzjm2 + 2.0 * zjm1;
                const double d2zj = 2.0 * w0c * d2zjm1 - d2zjm2 + 4.0 * dzjm1;
                const double bj = d2zj / (dzj * dzj);
                const double ajm1 = 1.0 - zjm1 * bjm1;
                const double mu = 2.0 * w0c * bj / bjm1;
                const double nu = -bj / bjm2;
                mus = mu * w1c / w0c;
                C.halo(W[i1]);
                C.launch_stage(W[i1], i2 < 0 ? yn : W[i2], yn, fn, W[i0], t + h * thjm1, mu, nu,
                               1.0 - mu - nu, h * mus, ajm1);
                ++nfev;
                const double thj = mu * thjm1 + nu * thjm2 + mus * (1.0 - ajm1);
                if (j < m) {       // rotate slots instead of the copies of :318-319
                    const int freed = i2 < 0 ? 3 - i1 - i0 : i2;
                    i2 = i1; i1 = i0; i0 = freed;
                    thjm2 = thjm1; thjm1 = thj; bjm2 = bjm1; bjm1 = bj;
                    zjm2 = zjm1; zjm1 = zj; dzjm2 = dzjm1; dzjm1 = dzj;
                    d2zjm2 = d2zjm1; d2zjm1 = d2zj;
                }
            }
            // result in W[i0]; Y_{j-1} (slot i1) is dead and receives f(t+h, y)
            double* y = W[i0];
            double* f1 = W[i1];
            double* third = W[3 - i0 - i1];
            // ---- final evaluation + error estimate (:214-220) ---------------
            C.halo(y);
            C.launch_final(y, yn, fn, f1, t + h, h, rtol, atol);
            ++nfev;
            err = std::sqrt(C.global_sum() / (double)C.n_total);
            if (C.rc != XSQ_OK) break;
            if (dbg) {
                const double now = wall_ms();
                std::fprintf(stderr, "[rkc r%d] t=%.3e h=%.3e m=%d err=%.3e  %.3f ms (%.4f ms/stage)\n",
                             C.rank, t, h, m, err, now - t_dbg, (now - t_dbg) / m);
                t_dbg = now;
            }
            if (err < 1.0) {
                // accepted: (yn, fn) <- (y, f1); the old (yn, fn) are the
                // interpolation data (:246-251), then become work vectors
                double* y_old = yn;
                double* f_old = fn;
                yn = y;
                fn = f1;
                t += h;
                ++n_acc;
                while (ieval < A->n_eval && direction * (A->t_eval[ieval] - t) <= 0.0) {
                    emit_point(t_old, t, y_old, f_old, A->t_eval[ieval],
                               A->u_eval + (size_t)ieval * C.S.n());
                    ++ieval;
                }
                W[0] = y_old; W[1] = f_old; W[2] = third;
                break;
            }
            if (std::isnan(err) || std::isinf(err)) { status = XSQ_LANE_OVERFLOW; break; }
            ++n_rej;
            absh = 0.8 * absh / std::pow(err, one3rd);
            if (absh < hmin) { status = XSQ_LANE_STEP_TOO_SMALL; break; }
            newspc = !jacatt;
        }
        if (status != 1 || C.rc != XSQ_OK) break;
        // ---- accepted-step bookkeeping (:239-266) ---------------------------
        jacatt = A->const_jac != 0;
        nstsig = (nstsig + 1) % 25;
        newspc = false;
        if (A->rho_const > 0.0 || A->rho_cb || nstsig == 0) newspc = !jacatt;
        double fac = 10.0;
        if (!have_hold) {
            const double t2 = std::pow(err, one3rd);
            if (0.8 < fac * t2) fac = 0.8 / t2;
        } else {
            const double t1 = 0.8 * absh * std::pow(errold, one3rd);
            const double t2 = std::fabs(hold) * std::pow(err, two3rd);
            if (t1 < fac * t2) fac = t1 / t2;
        }
        absh = std::fmax(0.1, fac) * absh;
        absh = std::fmax(hmin, std::fmin(max_step, absh));
        errold = err;
        hold = h;
        have_hold = true;
        if (direction * (t - tf) >= 0.0) status = 0;
    }
    k_copy<<<C.grid, C.block, 0, st>>>(C.S, yn, A->u_final, 0);
    C.launched();
    cudaError_t e = cudaStreamSynchronize(st);
    R->t_final = t;
    R->n_accepted = n_acc;
    R->n_rejected = n_rej;
    R->nfev = nfev;
    R->nfesig = nfesig;
    R->maxm = maxm;
    R->status = status == 1 ? XSQ_LANE_STEP_BUDGET : status;
    R->n_eval_done = ieval;
    R->kernel_launches = C.launches;
    if (multi) {              // nobody frees storage a neighbour may still read
        C.halo(yn);
        k_peer_barrier<<<dim3(1, 1), dim3(1, 1), 0, st>>>(C.ps);
        C.launched();
        cudaStreamSynchronize(st);
        ws->seq = C.seq;
    }
    release();
    cudaFreeHost(C.scalar_host);
    if (dbg0) std::fprintf(stderr, "[rkc r%d] total %.3f ms\n", C.rank, wall0() - t_enter);
    if (C.rc != XSQ_OK) return C.rc;
    if (e != cudaSuccess) { set_detail(std::string("rkc: ") + cudaGetErrorString(e)); return XSQ_ERR_CUDA; }
    return XSQ_OK;
}

// Stage-kernel micro-benchmark for the roofline: runs `reps` stage launches on
// a rows x nx slab and returns the average device time per launch.
int rkc_stage_bench(int nx, int rows, int reps, double* ms_per_stage, cudaStream_t st) {
    Slab S{nx, rows, 0, 0, ((double)nx + 1.0) * ((double)nx + 1.0), 1.0 / ((double)nx + 1.0), nullptr};
    const size_t na = S.n_alloc();
    double* buf = nullptr;
    if (cudaMalloc((void**)&buf, na * 5 * sizeof(double)) != cudaSuccess) return XSQ_ERR_NOMEM;
    cudaMemsetAsync(buf, 0, na * 5 * sizeof(double), st);
    dim3 block(TX, TY), grid((nx / PX + TX - 1) / TX, (rows + TY - 1) / TY);
    double* v[3] = {buf, buf + na, buf + 2 * na};
    double *yn = buf + 3 * na, *fn = buf + 4 * na;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int w = 0; w < 3; ++w) {
        k_stage<pde::Heat2dReaction><<<grid, block, 0, st>>>(
            S, PeerSync{0, nullptr, nullptr, nullptr}, v[1], v[1], v[1] + (size_t)nx * (rows + 1), v[2],
            yn, fn, v[0], 0.0, 1.9, -0.95, 0.05, 1e-9, 0.3);
        count_launch();
    }
    cudaEventRecord(e0, st);
    for (int r = 0; r < reps; ++r) {
        const double* in = v[(r + 1) % 3];
        k_stage<pde::Heat2dReaction><<<grid, block, 0, st>>>(
            S, PeerSync{0, nullptr, nullptr, nullptr}, in, in, in + (size_t)nx * (rows + 1),
            v[(r + 2) % 3], yn, fn, v[r % 3], 0.0, 1.9, -0.95, 0.05, 1e-9, 0.3);
        count_launch();
    }
    cudaEventRecord(e1, st);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaError_t e = cudaGetLastError();
    cudaFree(buf);
    if (e != cudaSuccess) { set_detail(cudaGetErrorString(e)); return XSQ_ERR_CUDA; }
    *ms_per_stage = ms / reps;
    return XSQ_OK;
}

// k_stage against its TMA-staged variant (xsq_rkc_tma.cuh) on the same random
// slab: the largest |difference| of one stage (must be 0) and the average time
// per stage of the TMA variant.
__global__ void k_fill_pattern(double* p, size_t n, unsigned long long seed) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x) {
        unsigned long long z = (i + seed) * 0x9E3779B97F4A7C15ULL;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
        z ^= z >> 31;
        p[i] = (double)(z >> 11) * 0x1.0p-53 - 0.5;
    }
}
__global__ void k_max_abs_diff(const double* a, const double* b, size_t n, unsigned long long* out) {
    double m = 0.0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x) {
        const double d = fabs(a[i] - b[i]);
        m = (d > m || d != d) ? (d != d ? __longlong_as_double(0x7ff0000000000000LL) : d) : m;
    }
    atomicMax(out, (unsigned long long)__double_as_longlong(m));   // non-negative doubles order as integers
}

int rkc_stage_bench_tma(int nx, int rows, int reps, double* ms_per_stage, double* max_abs_diff,
                        cudaStream_t st) {
    Slab S{nx, rows, 0, 0, ((double)nx + 1.0) * ((double)nx + 1.0), 1.0 / ((double)nx + 1.0), nullptr};
    const size_t na = S.n_alloc();
    double* buf = nullptr;
    unsigned long long* dmax = nullptr;
    if (cudaMalloc((void**)&buf, na * 6 * sizeof(double)) != cudaSuccess) return XSQ_ERR_NOMEM;
    if (cudaMalloc((void**)&dmax, 8) != cudaSuccess) { cudaFree(buf); return XSQ_ERR_NOMEM; }
    cudaMemsetAsync(dmax, 0, 8, st);
    k_fill_pattern<<<1184, 256, 0, st>>>(buf, na * 6, 12345ULL);
    double* v[3] = {buf, buf + na, buf + 2 * na};
    double *yn = buf + 3 * na, *fn = buf + 4 * na, *ref = buf + 5 * na;
    // ghost rows hold the Dirichlet zero
    for (int k = 0; k < 3; ++k) {
        cudaMemsetAsync(v[k], 0, (size_t)nx * sizeof(double), st);
        cudaMemsetAsync(v[k] + (size_t)nx * (rows + 1), 0, (size_t)nx * sizeof(double), st);
    }
    CUtensorMap maps[3];
    for (int k = 0; k < 3; ++k)
        if (make_stage_map(&maps[k], v[k], nx, rows) != 0) {
            cudaFree(buf);
            cudaFree(dmax);
            set_detail("cuTensorMapEncodeTiled failed");
            return XSQ_ERR_CUDA;
        }
    dim3 block(TX, TY), grid((nx / PX + TX - 1) / TX, (rows + TY - 1) / TY);
    const double mu = 1.9, nu = -0.95, c3 = 0.05, hmus = 1e-9, ajm1 = 0.3;
    // one stage each way from the same inputs
    k_stage<pde::Heat2dReaction><<<grid, block, 0, st>>>(
        S, PeerSync{0, nullptr, nullptr, nullptr}, v[1], v[1], v[1] + (size_t)nx * (rows + 1), v[2], yn,
        fn, ref, 0.0, mu, nu, c3, hmus, ajm1);
    k_stage_tma<pde::Heat2dReaction><<<grid, block, 0, st>>>(maps[1], S, v[2], yn, fn, v[0], 0.0, mu,
                                                             nu, c3, hmus, ajm1);
    k_max_abs_diff<<<1184, 256, 0, st>>>(ref + nx, v[0] + nx, (size_t)nx * rows, dmax);
    count_launch();
    count_launch();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int w = 0; w < 3; ++w)
        k_stage_tma<pde::Heat2dReaction><<<grid, block, 0, st>>>(maps[1], S, v[2], yn, fn, v[0], 0.0,
                                                                 mu, nu, c3, hmus, ajm1);
    cudaEventRecord(e0, st);
    for (int r = 0; r < reps; ++r) {
        k_stage_tma<pde::Heat2dReaction><<<grid, block, 0, st>>>(
            maps[(r + 1) % 3], S, v[(r + 2) % 3], yn, fn, v[r % 3], 0.0, mu, nu, c3, hmus, ajm1);
        count_launch();
    }
    cudaEventRecord(e1, st);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    unsigned long long bits = 0;
    cudaMemcpy(&bits, dmax, 8, cudaMemcpyDeviceToHost);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaError_t e = cudaGetLastError();
    cudaFree(buf);
    cudaFree(dmax);
    if (e != cudaSuccess) { set_detail(cudaGetErrorString(e)); return XSQ_ERR_CUDA; }
    *ms_per_stage = ms / reps;
    if (max_abs_diff) std::memcpy(max_abs_diff, &bits, 8);
    return XSQ_OK;
}

}  // namespace xsq
