// GPU box: dumps what MUFU.RCP64H (rcp.approx.ftz.f64) returns, so that the C
// oracle can restate the kernel's rcp_scale / rcp_fast / log2_fast bit for bit.
// The instruction reads only the HIGH 32 bits of its operand and writes only
// the high 32 bits of the result (low word zero), so the whole function is a
// table over sign/exponent/20 mantissa bits.
//   out: gpurun_out/rcp64h_e0.bin     2^20 uint32: high word of rcp(1.m) for every m
//        gpurun_out/rcp64h_check.txt  exponent independence + low-word checks
//        gpurun_out/devmath_vectors.bin  random (x, log2_fast, exp2_fast(z), rcp_scale, rcp_fast)
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include "xsq_rk_core.cuh"
using namespace xsq;

__global__ void k_table(uint32_t* out, uint32_t* lo_or, int ebias) {
    const uint32_t m = blockIdx.x * blockDim.x + threadIdx.x;       // 20 bits
    const uint32_t hi = ((uint32_t)ebias << 20) | m;
    uint32_t acc = 0;
    // low word must not matter: try three different ones
    const uint32_t lows[3] = {0u, 0xffffffffu, 0x9e3779b9u};
    uint32_t r_hi = 0;
    for (int i = 0; i < 3; ++i) {
        const double x = __hiloint2double((int)hi, (int)lows[i]);
        double r;
        asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
        const uint32_t h = (uint32_t)__double2hiint(r);
        if (i == 0) r_hi = h; else acc |= (h ^ r_hi);
        acc |= (uint32_t)__double2loint(r);
    }
    out[m] = r_hi;
    if (acc) atomicOr(lo_or, acc);
}

__global__ void k_vectors(const double* x, const double* z, double* out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[4 * i + 0] = log2_fast(x[i]);
    out[4 * i + 1] = exp2_fast(z[i]);
    out[4 * i + 2] = rcp_scale(x[i]);
    out[4 * i + 3] = rcp_fast(x[i]);
}

// dependent-chain DFMA latency and throughput per SM sub-partition
__global__ void k_lat(double* out, long long* cyc, int iters) {
    double a = threadIdx.x * 1e-9 + 1.0, b = 1.0000001, c = 1e-9;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) a = fma(a, b, c);
    }
    long long t1 = clock64();
    out[threadIdx.x + blockIdx.x * blockDim.x] = a;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
    system("mkdir -p gpurun_out");
    const int M = 1 << 20;
    uint32_t *d_t, *d_lo;
    cudaMalloc(&d_t, M * 4); cudaMalloc(&d_lo, 4);
    uint32_t* h0 = (uint32_t*)malloc(M * 4);
    uint32_t* h1 = (uint32_t*)malloc(M * 4);
    FILE* chk = fopen("gpurun_out/rcp64h_check.txt", "w");
    cudaMemset(d_lo, 0, 4);
    k_table<<<M / 256, 256>>>(d_t, d_lo, 1023);
    cudaMemcpy(h0, d_t, M * 4, cudaMemcpyDeviceToHost);
    uint32_t lo = 0; cudaMemcpy(&lo, d_lo, 4, cudaMemcpyDeviceToHost);
    fprintf(chk, "e=1023 low-word/low-input dependence mask: %08x (0 = none)\n", lo);
    FILE* f = fopen("gpurun_out/rcp64h_e0.bin", "wb"); fwrite(h0, 4, M, f); fclose(f);
    // exponent independence: result(e, m) == result(1023, m) with exponent 2046 - e - (m ? 1 : 0) ... just compare mantissa+relative exponent
    const int es[] = {1, 2, 500, 900, 989, 1000, 1022, 1024, 1025, 1060, 1500, 2000, 2045};
    for (int e : es) {
        cudaMemset(d_lo, 0, 4);
        k_table<<<M / 256, 256>>>(d_t, d_lo, e);
        cudaMemcpy(h1, d_t, M * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(&lo, d_lo, 4, cudaMemcpyDeviceToHost);
        long long mism = 0; int first = -1;
        for (int m = 0; m < M; ++m) {
            // expected: shift exponent by (1023 - e)
            const int64_t expect = (int64_t)h0[m] + ((int64_t)(1023 - e) << 20);
            if ((int64_t)h1[m] != expect) { ++mism; if (first < 0) first = m; }
        }
        fprintf(chk, "e=%4d lowmask %08x mismatches vs shifted e=1023 table: %lld (first m=%d h1=%08x h0=%08x)\n",
                e, lo, mism, first, first >= 0 ? h1[first] : 0, first >= 0 ? h0[first] : 0);
    }
    // negative inputs
    fprintf(chk, "table[0]=%08x table[1]=%08x table[M/2]=%08x table[M-1]=%08x\n", h0[0], h0[1], h0[M / 2], h0[M - 1]);

    // random vectors for the C restatement of log2_fast / exp2_fast / rcp_scale / rcp_fast
    const int N = 1 << 18;
    double* hx = (double*)malloc(N * 8), *hz = (double*)malloc(N * 8), *ho = (double*)malloc(N * 32);
    srand48(12345);
    for (int i = 0; i < N; ++i) {
        const int kind = i & 3;
        if (kind == 0) hx[i] = exp2(-60.0 + 120.0 * drand48());           // wide range
        else if (kind == 1) hx[i] = exp2(-8.0 + 10.0 * drand48());        // ss around n
        else if (kind == 2) hx[i] = 3.0 * (1.0 + (drand48() - 0.5) * 1e-6);
        else hx[i] = 1e-10 * (1.0 + 1e3 * drand48());                      // scales
        hz[i] = -40.0 + 80.0 * drand48();
        if ((i & 7) == 5) hz[i] = -2.0 + 4.0 * drand48();
    }
    double *dx, *dz, *dout;
    cudaMalloc(&dx, N * 8); cudaMalloc(&dz, N * 8); cudaMalloc(&dout, N * 32);
    cudaMemcpy(dx, hx, N * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(dz, hz, N * 8, cudaMemcpyHostToDevice);
    k_vectors<<<N / 256, 256>>>(dx, dz, dout, N);
    cudaMemcpy(ho, dout, N * 32, cudaMemcpyDeviceToHost);
    f = fopen("gpurun_out/devmath_vectors.bin", "wb");
    fwrite(hx, 8, N, f); fwrite(hz, 8, N, f); fwrite(ho, 8, 4 * N, f); fclose(f);

    // DFMA latency / throughput
    double* dl; long long* dc; cudaMalloc(&dl, 1024 * 148 * 8); cudaMalloc(&dc, 148 * 8);
    for (int threads : {32, 64, 128, 256, 512, 1024}) {
        k_lat<<<1, threads>>>(dl, dc, 1000);
        k_lat<<<1, threads>>>(dl, dc, 4000);
        long long c; cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
        fprintf(chk, "dfma chain: %4d threads/SM: %.2f cycles per dependent DFMA per warp (=> %.3f warp-DFMA/cycle/SM)\n",
                threads, c / (4000.0 * 16), (threads / 32) * 4000.0 * 16 / c);
    }
    fprintf(chk, "cuda status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    fclose(chk);
    printf("done\n");
    return 0;
}
