// GPU box: how do fp64 and non-fp64 instructions share an SM sub-partition's
// issue port?  Each kernel runs a fixed mix per loop iteration, 16 warps per SM
// (4 per scheduler), all chains independent (no latency limit), and reports
// cycles per iteration per scheduler.  If an fp64 instruction blocks issue for
// two cycles the cost is 2 F + O; with perfect overlap it is max(2 F, F + O).
#include <cstdio>
#include <cstdlib>

template <int F, int O, int KIND>
__global__ void __launch_bounds__(512) mix(double* out, int* iout, long long* cyc, int iters) {
    double a0 = threadIdx.x * 1e-9 + 1.0, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5,
           a6 = a0 + 6, a7 = a0 + 7;
    const double b = 1.0000001, c = 1e-9;
    int i0 = threadIdx.x, i1 = i0 + 1, i2 = i0 + 2, i3 = i0 + 3, i4 = i0 + 4, i5 = i0 + 5, i6 = i0 + 6, i7 = i0 + 7;
    float f0 = threadIdx.x, f1 = f0 + 1, f2 = f0 + 2, f3 = f0 + 3;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int k = 0; k < F; ++k) {
                switch ((u * F + k) & 7) {
#define DF(v) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(v) : "d"(b), "d"(c))
                    case 0: DF(a0); break; case 1: DF(a1); break; case 2: DF(a2); break; case 3: DF(a3); break;
                    case 4: DF(a4); break; case 5: DF(a5); break; case 6: DF(a6); break; default: DF(a7); break;
                }
            }
#pragma unroll
            for (int k = 0; k < O; ++k) {
                if (KIND == 0) {            // integer ALU: one LOP3 each
#define LO(v) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v) : "r"(it), "r"(i7))
                    switch ((u * O + k) % 7) {
                        case 0: LO(i0); break; case 1: LO(i1); break; case 2: LO(i2); break; case 3: LO(i3); break;
                        case 4: LO(i4); break; case 5: LO(i5); break; default: LO(i6); break;
                    }
                } else {                    // fp32 FMA pipe
#define FF(v) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(v) : "f"(1.0001f), "f"(0.5f))
                    switch ((u * O + k) & 3) {
                        case 0: FF(f0); break; case 1: FF(f1); break; case 2: FF(f2); break; default: FF(f3); break;
                    }
                }
            }
        }
    }
    const long long t1 = clock64();
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    out[g] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7)) + f0 + f1 + f2 + f3;
    iout[g] = i0 ^ i1 ^ i2 ^ i3 ^ i4 ^ i5 ^ i6 ^ i7;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int F, int O, int KIND>
void run(FILE* fp, double* d, int* di, long long* dc) {
    const int iters = 2000;
    mix<F, O, KIND><<<148, 512>>>(d, di, dc, iters);
    mix<F, O, KIND><<<148, 512>>>(d, di, dc, iters);
    long long c;
    cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
    // per scheduler: 4 warps, each executes 8*(F+O) instructions per iteration
    const double per_round = (double)c / iters / 8.0;      // cycles for 4 warps x (F + O)
    fprintf(fp, "F=%d O=%d kind=%s: %.2f cycles per (4 warps x %d instr);  2F+O model %.0f, max(2F,F+O) model %.0f\n",
            F, O, KIND ? "ffma" : "ialu", per_round, F + O, 4.0 * (2 * F + O),
            4.0 * ((2 * F > F + O) ? 2 * F : F + O));
}

int main() {
    system("mkdir -p gpurun_out");
    FILE* fp = fopen("gpurun_out/issue_model.txt", "w");
    double* d; int* di; long long* dc;
    cudaMalloc(&d, 148 * 512 * 8); cudaMalloc(&di, 148 * 512 * 4); cudaMalloc(&dc, 148 * 8);
    run<1, 0, 0>(fp, d, di, dc); run<0, 1, 0>(fp, d, di, dc); run<0, 1, 1>(fp, d, di, dc);
    run<1, 1, 0>(fp, d, di, dc); run<1, 2, 0>(fp, d, di, dc); run<1, 3, 0>(fp, d, di, dc);
    run<2, 1, 0>(fp, d, di, dc); run<2, 2, 0>(fp, d, di, dc); run<3, 1, 0>(fp, d, di, dc);
    run<1, 1, 1>(fp, d, di, dc); run<1, 2, 1>(fp, d, di, dc); run<2, 2, 1>(fp, d, di, dc);
    fprintf(fp, "status %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    fclose(fp);
    return 0;
}
