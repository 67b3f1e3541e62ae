// One object file per tableau: compile with -DXSQ_INST_TAB=<name>.
// Instantiates rk_persistent<Tab, Rhs> for every built-in right-hand side and
// provides launch_<Tab>(), the dispatcher xsq_api.cu calls.
#include "xsq_launch.h"
#include "xsq_rk_fast.cuh"
#include "xsq_rhs.cuh"
#include "xsq.h"
#include <cstdlib>
#include <string>
#include <cstdio>
#include <cstring>
#include <cmath>
#include <limits>

#ifndef XSQ_INST_TAB
#error "compile with -DXSQ_INST_TAB=<tableau>"
#endif
#define XSQ_CAT2(a, b) a##b
#define XSQ_CAT(a, b) XSQ_CAT2(a, b)

namespace xsq {

// Launch geometry.  Lane-per-system kernels are register-bound (all stage
// vectors live in registers), so the CTA is small (128 threads) and the
// register cap follows from MINB resident CTAs per SM.
template <class Tab, class R>
struct Geometry {
    static constexpr int BLOCK = 128;
    static constexpr int KDOUBLES = (Tab::S + 1) * R::NL;
    // ~2 regs per live double + ~60 of scalars/addresses
    static constexpr int MINB =
        KDOUBLES <= 21 ? 4 : (KDOUBLES <= 36 ? 3 : 2);
};

static unsigned hi_word(double x) {
    unsigned long long b;
    std::memcpy(&b, &x, 8);
    return (unsigned)(b >> 32);
}

// The ensemble hot path (xsq_rk_fast.cuh): adaptive, final state only, default
// step budget, controller without the alpha term.  XSQ_NO_FAST=1 forces the
// generic kernel (tests compare the two bit for bit).
template <class Tab, class R>
static bool fast_eligible(const RkDev& P) {
    if constexpr (Tab::VARIANT != tab::GENERIC || R::WARP) {
        return false;
    } else {
        if (P.n_forced != 0 || P.n_eval != 0 || P.n_events != 0) return false;
        if (P.minalpha != 0.0 || P.max_steps != std::numeric_limits<int>::max()) return false;
        if (P.n_lanes >= (1LL << 31)) return false;
        const char* e = getenv("XSQ_NO_FAST");
        return !(e && e[0] == '1');
    }
}

template <class Tab, class R>
static int launch_fast(const RkDev& P0, cudaStream_t st, LaunchInfo* info) {
    if constexpr (Tab::VARIANT != tab::GENERIC || R::WARP) {
        return XSQ_ERR_UNSUPPORTED;
    } else {
        constexpr int BLOCK = Geometry<Tab, R>::BLOCK, MINB = Geometry<Tab, R>::MINB;
        RkDev P = P0;
        // min_step = max(H_MIN_A (|t| + h0), sqrt(tiny)) <= M for every step that
        // starts with 2 h0 < |t_bound - t| (common.py:123-148, 310-331)
        const double tmax = std::fmax(std::fabs(P.t0), std::fabs(P.t_bound));
        const double span = std::fabs(P.t_bound - P.t0);
        const double M = std::fmax(Tab::H_MIN_A * (tmax + 0.5 * span), 0x1.0p-511) * (1.0 + 0x1.0p-30);
        const long long lo = hi_word(M), hi = hi_word(P.max_step);
        P.fast_hi_min = (int)lo;
        P.fast_hi_span = (int)(hi - lo - 1 > 0 ? hi - lo - 1 : 0);
        auto kern = rk_fast<Tab, R, BLOCK, MINB>;
        int dev = 0, n_sm = 0, occ = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return XSQ_ERR_CUDA;
        if (cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
            return XSQ_ERR_CUDA;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, BLOCK, 0) != cudaSuccess ||
            occ < 1)
            return XSQ_ERR_CUDA;
        long long want_blocks = (P.n_lanes + BLOCK - 1) / BLOCK;
        long long grid = (long long)n_sm * occ;
        if (want_blocks < grid) grid = want_blocks;
        if (grid < 1) grid = 1;
        kern<<<(unsigned)grid, BLOCK, 0, st>>>(P);
        count_launch();
        if (info) {
            cudaFuncAttributes fa;
            cudaFuncGetAttributes(&fa, kern);
            info->grid = (int)grid;
            info->block = BLOCK;
            info->blocks_per_sm = occ;
            info->regs = fa.numRegs;
        }
        return cudaGetLastError() == cudaSuccess ? XSQ_OK : XSQ_ERR_CUDA;
    }
}

template <class Tab, class R, int BLOCK = Geometry<Tab, R>::BLOCK,
          int MINB = Geometry<Tab, R>::MINB>
static int launch_one(const RkDev& P, cudaStream_t st, LaunchInfo* info) {
    using G = Geometry<Tab, R>;
    if (BLOCK == G::BLOCK && MINB == G::MINB && fast_eligible<Tab, R>(P))
        return launch_fast<Tab, R>(P, st, info);
#ifdef XSQ_TUNE
    // occupancy sweep for profiling builds: XSQ_GEOM=<block>x<CTAs per SM>
    if (BLOCK == G::BLOCK && MINB == G::MINB) {
        const char* e = getenv("XSQ_GEOM");
        const std::string want = e ? e : "";
#define XSQ_TRY(B, M) if (want == #B "x" #M) return launch_one<Tab, R, B, M>(P, st, info);
        XSQ_TRY(128, 3) XSQ_TRY(128, 5) XSQ_TRY(32, 16) XSQ_TRY(32, 17)
        XSQ_TRY(32, 18) XSQ_TRY(32, 19) XSQ_TRY(32, 20) XSQ_TRY(64, 9)
        XSQ_TRY(64, 10) XSQ_TRY(256, 2)
#undef XSQ_TRY
        const char* mr = getenv("XSQ_MAXREG");
        if (mr) {
            const int want_mr = atoi(mr);
            int dev = 0, n_sm = 0, occ = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
#define XSQ_MR(N)                                                              \
    if (want_mr == N) {                                                        \
        auto k = rk_persistent_mr<Tab, R, 32, N>;                              \
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, 32, 0);         \
        k<<<n_sm * occ, 32, 0, st>>>(P);                                       \
        count_launch();                                                        \
        if (getenv("XSQ_VERBOSE")) fprintf(stderr, "maxreg %d occ %d\n", N, occ); \
        return cudaGetLastError() == cudaSuccess ? XSQ_OK : XSQ_ERR_CUDA;      \
    }
            XSQ_MR(120) XSQ_MR(112) XSQ_MR(104) XSQ_MR(96)
#undef XSQ_MR
        }
    }
#endif
    auto kern = rk_persistent<Tab, R, BLOCK, MINB>;
    int dev = 0, n_sm = 0, occ = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return XSQ_ERR_CUDA;
    if (cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev) !=
        cudaSuccess)
        return XSQ_ERR_CUDA;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(
            &occ, kern, BLOCK,
            P.n_eval > 0 ? sizeof(double) * 4 * R::NL * BLOCK : 0) != cudaSuccess ||
        occ < 1)
        return XSQ_ERR_CUDA;
    const long long per_block = R::WARP ? BLOCK / 32 : BLOCK;
    long long want_blocks = (P.n_lanes + per_block - 1) / per_block;
    long long grid = (long long)n_sm * occ;          // persistent: fill the GPU
    if (want_blocks < grid) grid = want_blocks;
    if (grid < 1) grid = 1;
    // dense-output staging: 4 points x NL components per thread
    const size_t smem = P.n_eval > 0 ? sizeof(double) * 4 * R::NL * BLOCK : 0;
    kern<<<(unsigned)grid, BLOCK, smem, st>>>(P);
    count_launch();
    if (info) {
        cudaFuncAttributes fa;
        cudaFuncGetAttributes(&fa, kern);
        info->grid = (int)grid;
        info->block = BLOCK;
        info->blocks_per_sm = occ;
        info->regs = fa.numRegs;
    }
    return cudaGetLastError() == cudaSuccess ? XSQ_OK : XSQ_ERR_CUDA;
}

int XSQ_CAT(launch_, XSQ_INST_TAB)(int rhs, const RkDev& P, cudaStream_t st,
                                   LaunchInfo* info) {
    using T = tab::XSQ_INST_TAB;
    switch (rhs) {
        case XSQ_RHS_LORENZ63: return launch_one<T, rhs::Lorenz63>(P, st, info);
        case XSQ_RHS_VANDERPOL: return launch_one<T, rhs::VanDerPol>(P, st, info);
        case XSQ_RHS_ARENSTORF: return launch_one<T, rhs::Arenstorf>(P, st, info);
        case XSQ_RHS_NBODY32: return launch_one<T, rhs::NBody32>(P, st, info);
        default: return XSQ_ERR_UNSUPPORTED;
    }
}

}  // namespace xsq
