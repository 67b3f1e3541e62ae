// xsq_rkc_tma.cuh -- the SSV2stab stage kernel with its stencil operand staged
// through shared memory by the TMA unit (cp.async.bulk.tensor.2d, SASS UTMALDG):
// an EXPERIMENT measured against k_stage (xsq_rkc_kernels.cuh), which reads the
// same operand with 128-bit __ldg loads and takes the vertical neighbours from
// L1/L2.  Same arithmetic, operation by operation, as stage_body (the result is
// compared bit for bit by xsq_rkc_stage_bench_tma).  Single GPU only: the halo
// rows of a multi-GPU slab live in the neighbour's memory, outside any tensor map
// of this rank's buffer.
//
// One CTA = one tile of TY rows x (TX * PX) columns.  The (TY + 2) x (TX*PX + 4)
// box of Y_{j-1} around it arrives with ONE bulk tensor copy; columns outside the
// grid are filled with zeros by the TMA unit, which is the Dirichlet value.  Thread
// (tx, ty) computes columns tx, tx + 32, tx + 64, tx + 96 of row ty: conflict-free
// 64-bit shared-memory reads and fully coalesced 64-bit global accesses.
#pragma once
#include <cuda.h>
#include "xsq_rkc_kernels.cuh"

namespace xsq {
namespace rkc {

constexpr int TMA_COLS = TX * PX;                 // 128
// the box starts TWO columns left of the tile: the innermost start coordinate of a
// tensor copy must be a multiple of 16 bytes (an odd column index of doubles is an
// illegal instruction, tools/devtest/tma_probe.cu)
constexpr int TMA_PAD = 2;
constexpr int TMA_BOX_X = TMA_COLS + 2 * TMA_PAD, TMA_BOX_Y = TY + 2;

__device__ __forceinline__ unsigned smem_u32(const void* p) {
    return (unsigned)__cvta_generic_to_shared(p);
}

template <class Pde>
__global__ void __launch_bounds__(TX* TY)
    k_stage_tma(const __grid_constant__ CUtensorMap map, Slab S, const double* __restrict__ yjm2,
                const double* __restrict__ yn, const double* __restrict__ fn,
                double* __restrict__ yj, double t, double mu, double nu, double c3, double hmus,
                double ajm1) {
    __shared__ __align__(128) double tile[TMA_BOX_Y][TMA_BOX_X];
    __shared__ __align__(8) unsigned long long bar;
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int col0 = blockIdx.x * TMA_COLS, row0 = blockIdx.y * TY;    // interior coordinates
    if (tx == 0 && ty == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)),
                     "r"((unsigned)sizeof(tile))
                     : "memory");
        // storage row of interior row r is r + 1; the box starts one row above and
        // two columns to the left of the tile
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
            " [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(&tile[0][0])),
            "l"(&map), "r"(col0 - TMA_PAD), "r"(row0), "r"(smem_u32(&bar))
            : "memory");
    }
    __syncthreads();
    unsigned done = 0;
    while (!done) {
        asm volatile(
            "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; "
            "selp.u32 %0, 1, 0, p; }"
            : "=r"(done)
            : "r"(smem_u32(&bar))
            : "memory");
    }
    const int row = row0 + ty;
    if (row >= S.rows) return;
    const double y = (double)(S.row0 + row + 1) * S.hgrid;
#pragma unroll
    for (int k = 0; k < PX; ++k) {
        const int lc = tx + 32 * k, col = col0 + lc;
        if (col >= S.nx) continue;
        const double c = tile[ty + 1][lc + TMA_PAD];
        const double n = tile[ty][lc + TMA_PAD], s = tile[ty + 2][lc + TMA_PAD];
        const double w = tile[ty + 1][lc + TMA_PAD - 1], e = tile[ty + 1][lc + TMA_PAD + 1];
        const double x = (double)(col + 1) * S.hgrid;
        const double f = Pde::rhs(t, x, y, S.inv_h2, c, n, s, w, e, S.prm);
        const size_t idx = (size_t)(row + 1) * S.nx + col;
        yj[idx] = ((mu * c + nu * __ldg(yjm2 + idx)) + c3 * __ldg(yn + idx)) +
                  hmus * (f - ajm1 * __ldg(fn + idx));
    }
}

// A tensor map of one slab vector: (rows + 2) x nx doubles, box as above.
inline int make_stage_map(CUtensorMap* map, const double* base, int nx, int rows) {
    typedef CUresult (*encode_t)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                 const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                 const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess || !fn)
        return -1;
    const cuuint64_t dims[2] = {(cuuint64_t)nx, (cuuint64_t)(rows + 2)};
    const cuuint64_t strides[1] = {(cuuint64_t)nx * sizeof(double)};
    const cuuint32_t box[2] = {(cuuint32_t)TMA_BOX_X, (cuuint32_t)TMA_BOX_Y};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = ((encode_t)fn)(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)base, dims,
                                      strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : -1;
}

}  // namespace rkc
}  // namespace xsq
