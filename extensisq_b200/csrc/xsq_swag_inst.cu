// Instantiates swag_persistent<Rhs> for the built-in right-hand sides and
// provides launch_swag(), called by xsq_swag_solve (xsq_api.cu).
#include "xsq_launch.h"
#include "xsq_swag_core.cuh"
#include "xsq_swag_fast.cuh"
#include "xsq_rhs.cuh"
#include "xsq.h"

namespace xsq {

// Small systems, final state only: registers for phi, shared memory for the
// coefficient arrays (xsq_swag_fast.cuh).  64-thread CTAs: 47.6 KB of
// coefficient storage each, four per SM.
template <class R, int BLOCK, int MINB, int MAXREG = 0>
static int launch_swag_fast_geom(const RkDev& P, cudaStream_t st) {
#ifdef XSQ_SWAG_GEOMETRY_SWEEP
    auto kern = MAXREG ? swag_fast_maxreg<R, BLOCK, MAXREG ? MAXREG : 255> : swag_fast<R, BLOCK, MINB>;
#else
    auto kern = swag_fast<R, BLOCK, MINB>;
#endif
    int dev = 0, n_sm = 0, occ = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return XSQ_ERR_CUDA;
    if (cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
        return XSQ_ERR_CUDA;
    constexpr size_t smem = sizeof(SwagCoefs<BLOCK>);
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
        cudaSuccess)
        return XSQ_ERR_CUDA;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, BLOCK, smem) != cudaSuccess ||
        occ < 1)
        return XSQ_ERR_CUDA;
#ifdef XSQ_SWAG_GEOMETRY_SWEEP
    if (const char* e = getenv("XSQ_SWAG_OCC")) {      // fewer resident CTAs than fit
        const int cap = atoi(e);
        if (cap >= 1 && cap < occ) occ = cap;
    }
#endif
    long long want = (P.n_lanes + BLOCK - 1) / BLOCK;
    long long grid = (long long)n_sm * occ;
    if (want < grid) grid = want;
    if (grid < 1) grid = 1;
    kern<<<(unsigned)grid, BLOCK, smem, st>>>(P);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? XSQ_OK : XSQ_ERR_CUDA;
}

template <class R>
static int launch_swag_fast(const RkDev& P, cudaStream_t st) {
    if constexpr (R::WARP || R::NL > 4) {
        return XSQ_ERR_UNSUPPORTED;
    } else {
#ifdef XSQ_SWAG_GEOMETRY_SWEEP
        const char* e = getenv("XSQ_SWAG_GEOM");
        if (e && e[0] == '5') return launch_swag_fast_geom<R, 64, 4, 200>(P, st);
        if (e && e[0] == 'b') return launch_swag_fast_geom<R, 32, 4, 184>(P, st);
        if (e && e[0] == 'c') return launch_swag_fast_geom<R, 32, 4, 216>(P, st);
#endif
        return launch_swag_fast_geom<R, 64, 4>(P, st);
    }
}

template <class R>
static int launch_swag_one(const RkDev& P, cudaStream_t st) {
    if (swag_fast_eligible<R>(P)) return launch_swag_fast<R>(P, st);
    constexpr int BLOCK = 128, MINB = 2;
    auto kern = swag_persistent<R, BLOCK, MINB>;
    int dev = 0, n_sm = 0, occ = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return XSQ_ERR_CUDA;
    if (cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev) !=
        cudaSuccess)
        return XSQ_ERR_CUDA;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(
            &occ, kern, BLOCK, P.n_eval > 0 ? sizeof(double) * 4 * R::NL * BLOCK : 0) !=
            cudaSuccess || occ < 1)
        return XSQ_ERR_CUDA;
    const long long per_block = R::WARP ? BLOCK / 32 : BLOCK;
    long long want = (P.n_lanes + per_block - 1) / per_block;
    long long grid = (long long)n_sm * occ;
    if (want < grid) grid = want;
    if (grid < 1) grid = 1;
    // dense-output staging: 4 points x NL components per thread
    const size_t smem = P.n_eval > 0 ? sizeof(double) * 4 * R::NL * BLOCK : 0;
    kern<<<(unsigned)grid, BLOCK, smem, st>>>(P);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? XSQ_OK : XSQ_ERR_CUDA;
}

template <class R>
static int launch_init_one(const RkDev& P, cudaStream_t st) {
    constexpr int BLOCK = 128;
    const long long threads = R::WARP ? P.n_lanes * 32 : P.n_lanes;
    const long long grid = (threads + BLOCK - 1) / BLOCK;
    if (grid < 1) return XSQ_OK;
    ens_init<R, BLOCK><<<(unsigned)grid, BLOCK, 0, st>>>(P);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? XSQ_OK : XSQ_ERR_CUDA;
}

int launch_ens_init(int rhs, const RkDev& P, cudaStream_t st) {
    switch (rhs) {
        case XSQ_RHS_LORENZ63: return launch_init_one<rhs::Lorenz63>(P, st);
        case XSQ_RHS_VANDERPOL: return launch_init_one<rhs::VanDerPol>(P, st);
        case XSQ_RHS_ARENSTORF: return launch_init_one<rhs::Arenstorf>(P, st);
        case XSQ_RHS_NBODY32: return launch_init_one<rhs::NBody32>(P, st);
        default: return XSQ_ERR_UNSUPPORTED;
    }
}

template <class R>
static int launch_queue_one(const RkDev& P, int cost, double stbrad, double tanang,
                            cudaStream_t st) {
    int dev = 0, n_sm = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
        return XSQ_ERR_CUDA;
    stiff_queue<R><<<(unsigned)(n_sm * 12), 128, 0, st>>>(P, cost, stbrad, tanang);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? XSQ_OK : XSQ_ERR_CUDA;
}

// The stiffness probes the persistent kernel queued (xsq_rk_core.cuh, Lane::diagnose).
int launch_stiff_queue(int rhs, const RkDev& P, int cost, double stbrad, double tanang,
                       cudaStream_t st) {
    switch (rhs) {
        case XSQ_RHS_LORENZ63: return launch_queue_one<rhs::Lorenz63>(P, cost, stbrad, tanang, st);
        case XSQ_RHS_VANDERPOL: return launch_queue_one<rhs::VanDerPol>(P, cost, stbrad, tanang, st);
        case XSQ_RHS_ARENSTORF: return launch_queue_one<rhs::Arenstorf>(P, cost, stbrad, tanang, st);
        default: return XSQ_OK;          // warp-per-system policies do not queue
    }
}

int launch_swag(int rhs, const RkDev& P, cudaStream_t st) {
    switch (rhs) {
        case XSQ_RHS_LORENZ63: return launch_swag_one<rhs::Lorenz63>(P, st);
        case XSQ_RHS_VANDERPOL: return launch_swag_one<rhs::VanDerPol>(P, st);
        case XSQ_RHS_ARENSTORF: return launch_swag_one<rhs::Arenstorf>(P, st);
        case XSQ_RHS_NBODY32: return launch_swag_one<rhs::NBody32>(P, st);
        default: return XSQ_ERR_UNSUPPORTED;
    }
}

}  // namespace xsq
