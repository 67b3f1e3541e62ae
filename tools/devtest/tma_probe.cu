// Which 2-D tensor-map configurations does UTMALDG accept for a (rows x nx)
// array of doubles?  usage: tma_probe <type: 0 f64, 1 u64, 2 u32> <box_x elements of 8 B> <box_y> <x0>
// One configuration per process (an illegal instruction poisons the context).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

__device__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__global__ void probe(const __grid_constant__ CUtensorMap map, int bytes, int x0, int y0, double* out, int n) {
    extern __shared__ __align__(128) unsigned char tile[];
    __shared__ __align__(8) unsigned long long bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
            ::"r"(smem_u32(tile)), "l"(&map), "r"(x0), "r"(y0), "r"(smem_u32(&bar)) : "memory");
    }
    __syncthreads();
    unsigned done = 0;
    while (!done)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
    for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = ((const double*)tile)[i];
}

int main(int argc, char** argv) {
    const int type = atoi(argv[1]), bx = atoi(argv[2]), by = atoi(argv[3]), x0 = atoi(argv[4]);
    const int nx = 512, rows = 66;
    double* d; double* out;
    cudaMalloc(&d, sizeof(double) * nx * rows);
    cudaMalloc(&out, sizeof(double) * bx * by);
    double* h = (double*)malloc(sizeof(double) * nx * rows);
    for (int i = 0; i < nx * rows; ++i) h[i] = i;
    cudaMemcpy(d, h, sizeof(double) * nx * rows, cudaMemcpyHostToDevice);
    typedef CUresult (*encode_t)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    const int mul = type == 2 ? 2 : 1;
    const CUtensorMapDataType dt = type == 0 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : type == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT64 : CU_TENSOR_MAP_DATA_TYPE_UINT32;
    const cuuint64_t dims[2] = {(cuuint64_t)nx * mul, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)nx * 8};
    const cuuint32_t box[2] = {(cuuint32_t)bx * mul, (cuuint32_t)by};
    const cuuint32_t es[2] = {1, 1};
    CUtensorMap map;
    CUresult r = ((encode_t)fn)(&map, dt, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("type %d box %dx%d x0 %d: encode %d ", type, bx, by, x0, (int)r);
    if (r != CUDA_SUCCESS) { printf("\n"); return 1; }
    probe<<<1, 128, bx * by * 8>>>(map, bx * by * 8, x0 * mul, 1, out, bx * by);
    cudaError_t e = cudaDeviceSynchronize();
    double* ho = (double*)malloc(sizeof(double) * bx * by);
    cudaMemcpy(ho, out, sizeof(double) * bx * by, cudaMemcpyDeviceToHost);
    printf("run: %s  first %g %g second-row %g (want %d %d %d)\n", cudaGetErrorString(e), ho[0], ho[1], ho[bx],
           x0 < 0 ? 0 : nx + x0, nx + x0 + 1, 2 * nx + x0);
    return 0;
}
