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

// CTAs per SM of the fast kernel: the same as the generic kernel's.  One or two
// more CTAs per SM (96 / 80 registers) were measured on B200 and are slower
// (Ts5/Lorenz: 0.543 / 0.536 / 0.514 of the fp64 peak at 4 / 5 / 6 CTAs): the
// kernel is bound by instruction issue, not by latency, so more warps do not
// help and the tighter register budget costs moves.  XSQ_FAST_MINB=<n>
// selects the neighbours for profiling.
template <class Tab, class R>
struct FastGeometry {
    static constexpr int BASE = Geometry<Tab, R>::MINB;
    static constexpr int MINB = BASE;
};

template <class Tab, class R, int MINB = FastGeometry<Tab, R>::MINB>
static int launch_fast(const RkDev& P0, cudaStream_t st, LaunchInfo* info) {
    if constexpr (Tab::VARIANT != tab::GENERIC || R::WARP) {
        return XSQ_ERR_UNSUPPORTED;
    } else {
        constexpr int BLOCK = Geometry<Tab, R>::BLOCK;
        if constexpr (MINB == FastGeometry<Tab, R>::MINB) {
            if (const char* e = getenv("XSQ_FAST_MINB")) {
                const int want = atoi(e);
                constexpr int B = FastGeometry<Tab, R>::BASE;
                if (want == B + 1) return launch_fast<Tab, R, B + 1>(P0, st, info);
                if (want == B + 2) return launch_fast<Tab, R, B + 2>(P0, st, info);
            }
        }
        RkDev P = P0;
        fast_prepare<Tab>(P);
        auto kern = P.nfev_stiff_detect > 0 ? rk_fast<Tab, R, BLOCK, MINB, true>
                                            : rk_fast<Tab, R, BLOCK, MINB, false>;
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

// One entry per tableau: the right-hand sides it is instantiated for.
template <class T>
static int launch_tab(int rhs, const RkDev& P, cudaStream_t st, LaunchInfo* info) {
    if constexpr (T::VARIANT == tab::NYSTROMV) {
        // second order problems only: the built-in Van der Pol and Arenstorf
        // systems depend on the velocity (not for MR6NN, mikkawy.py), the
        // N-body problem does not
        switch (rhs) {
            case XSQ_RHS_VANDERPOL:
                if constexpr (T::VELOCITY_DEPENDENT) return launch_one<T, rhs::VanDerPol>(P, st, info);
                else return XSQ_ERR_UNSUPPORTED;
            case XSQ_RHS_ARENSTORF:
                if constexpr (T::VELOCITY_DEPENDENT) return launch_one<T, rhs::Arenstorf>(P, st, info);
                else return XSQ_ERR_UNSUPPORTED;
            case XSQ_RHS_NBODY32: return launch_one<T, rhs::NBody32>(P, st, info);
            default: return XSQ_ERR_UNSUPPORTED;
        }
    } else {
        switch (rhs) {
            case XSQ_RHS_LORENZ63: return launch_one<T, rhs::Lorenz63>(P, st, info);
            case XSQ_RHS_VANDERPOL: return launch_one<T, rhs::VanDerPol>(P, st, info);
            case XSQ_RHS_ARENSTORF: return launch_one<T, rhs::Arenstorf>(P, st, info);
            case XSQ_RHS_NBODY32: return launch_one<T, rhs::NBody32>(P, st, info);
            default: return XSQ_ERR_UNSUPPORTED;
        }
    }
}

int XSQ_CAT(launch_, XSQ_INST_TAB)(int rhs, const RkDev& P, cudaStream_t st,
                                   LaunchInfo* info) {
    return launch_tab<tab::XSQ_INST_TAB>(rhs, P, st, info);
}

}  // namespace xsq
