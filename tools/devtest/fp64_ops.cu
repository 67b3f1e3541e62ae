// GPU box: throughput of the fp64-pipe instructions one by one (DFMA, DMUL, DADD,
// DSETP, DFMA with a constant-bank / immediate operand), 16 warps per SM, 8
// independent chains per thread: cycles per warp instruction per scheduler.
#include <cstdio>
#include <cstdlib>
__constant__ double c_k[4] = {1.0000001, 1e-9, 0.5, 2.0};
template <int OP>
__global__ void __launch_bounds__(512) k(double* out, long long* cyc, int iters, double b, double c) {
    double a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-9 + 1.0 + i;
    int cnt = 0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (OP == 0) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(a[i]) : "d"(b), "d"(c));
                if (OP == 1) asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(a[i]) : "d"(b));
                if (OP == 2) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(a[i]) : "d"(c));
                if (OP == 3) { int p; asm volatile("{ .reg .pred q; setp.lt.f64 q, %1, %2; selp.s32 %0, 1, 0, q; }" : "=r"(p) : "d"(a[i]), "d"(b)); cnt += p; }
                if (OP == 4) a[i] = fma(a[i], c_k[0], c_k[1]);            // constant-bank operands
                if (OP == 5) asm volatile("fma.rn.f64 %0, %0, %1, 0d0000000000000000;" : "+d"(a[i]) : "d"(b));
                if (OP == 6) asm volatile("fma.rn.f64 %0, %0, 0d3FF0000000000000, %1;" : "+d"(a[i]) : "d"(c));
                if (OP == 7) asm volatile("abs.f64 %0, %0;" : "+d"(a[i]));
                if (OP == 8) asm volatile("mul.rn.f64 %0, %0, 0d3FE0000000000000;" : "+d"(a[i]));
            }
        }
    }
    const long long t1 = clock64();
    double s = cnt;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int OP>
void run(FILE* fp, const char* name, double* d, long long* dc) {
    const int iters = 2000;
    k<OP><<<148, 512>>>(d, dc, iters, 1.0000001, 1e-9);
    k<OP><<<148, 512>>>(d, dc, iters, 1.0000001, 1e-9);
    long long c; cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
    fprintf(fp, "%-28s %.2f cycles per warp instruction per scheduler (4 warps resident)\n", name,
            (double)c / iters / 32.0 / 4.0);
}
int main() {
    system("mkdir -p gpurun_out");
    FILE* fp = fopen("gpurun_out/fp64_ops.txt", "w");
    double* d; long long* dc;
    cudaMalloc(&d, 148 * 512 * 8); cudaMalloc(&dc, 148 * 8);
    run<0>(fp, "DFMA r,r,r", d, dc); run<1>(fp, "DMUL r,r", d, dc); run<2>(fp, "DADD r,r", d, dc);
    run<3>(fp, "DSETP + SEL", d, dc); run<4>(fp, "DFMA r,c[],c[]", d, dc); run<5>(fp, "DFMA r,r,0 (mul as fma)", d, dc);
    run<6>(fp, "DFMA r,1.0,r (add as fma)", d, dc); run<7>(fp, "abs.f64", d, dc); run<8>(fp, "DMUL r,imm", d, dc);
    fprintf(fp, "status %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    fclose(fp);
    return 0;
}
