// One object file per tableau: compile with -DXSQ_INST_TAB=<name>.
// Instantiates rk_persistent<Tab, Rhs> for every built-in right-hand side and
// provides launch_<Tab>(), the dispatcher xsq_api.cu calls.
#include "xsq_launch.h"
#include "xsq_rhs.cuh"
#include "xsq.h"

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

template <class Tab, class R>
static int launch_one(const RkDev& P, cudaStream_t st, LaunchInfo* info) {
    using G = Geometry<Tab, R>;
    auto kern = rk_persistent<Tab, R, G::BLOCK, G::MINB>;
    int dev = 0, n_sm = 0, occ = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return XSQ_ERR_CUDA;
    if (cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev) !=
        cudaSuccess)
        return XSQ_ERR_CUDA;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, G::BLOCK,
                                                      0) != cudaSuccess ||
        occ < 1)
        return XSQ_ERR_CUDA;
    const long long per_block = R::WARP ? G::BLOCK / 32 : G::BLOCK;
    long long want = (P.n_lanes + per_block - 1) / per_block;
    long long grid = (long long)n_sm * occ;          // persistent: fill the GPU
    if (want < grid) grid = want;
    if (grid < 1) grid = 1;
    kern<<<(unsigned)grid, G::BLOCK, 0, st>>>(P);
    count_launch();
    if (info) {
        cudaFuncAttributes fa;
        cudaFuncGetAttributes(&fa, kern);
        info->grid = (int)grid;
        info->block = G::BLOCK;
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
