// GPU box: accuracy of the controller's math helpers against libdevice.
#include <cstdio>
#include <cmath>
#include "xsq_rk_core.cuh"
using namespace xsq;
__global__ void k(double* out, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // x sweeps 1e-30 .. 1e30 logarithmically
    double x = exp2(-100.0 + 200.0 * i / n);
    double a = log2_fast(x), b = log2(x);
    double e1 = fabs(a - b) / fmax(fabs(b), 1.0);
    double z = -60.0 + 120.0 * i / n;
    double c = exp2_fast(z), d = exp2(z);
    double e2 = fabs(c - d) / d;
    double r = rcp_scale(x), e3 = fabs(r * x - 1.0);
    double r2 = rcp_fast(x), e4 = fabs(r2 * x - 1.0);
    out[4 * i] = e1; out[4 * i + 1] = e2; out[4 * i + 2] = e3; out[4 * i + 3] = e4;
}
int main() {
    const int n = 1 << 20;
    double* d; cudaMalloc(&d, 4 * n * sizeof(double));
    k<<<n / 256, 256>>>(d, n);
    double* h = new double[4 * n];
    cudaMemcpy(h, d, 4 * n * sizeof(double), cudaMemcpyDeviceToHost);
    double m[4] = {0, 0, 0, 0}; int at[4] = {0,0,0,0};
    for (int i = 0; i < n; ++i) for (int j = 0; j < 4; ++j) if (h[4 * i + j] > m[j]) { m[j] = h[4 * i + j]; at[j] = i; }
    printf("max rel err: log2_fast %.3e (i=%d)  exp2_fast %.3e (i=%d)  rcp_scale %.3e  rcp_fast %.3e\n", m[0], at[0], m[1], at[1], m[2], m[3]);
    return 0;
}
