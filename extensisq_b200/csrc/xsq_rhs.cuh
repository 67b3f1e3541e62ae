// Built-in right-hand sides, as device policies consumed by the solver
// templates.  The reference takes a Python callable `fun(t, y)`
// (extensisq/common.py:187, called once per stage at common.py:356); on the
// device the RHS is a type so that it inlines into the persistent kernel.
//
// Policy concept
//   N      total state size of one system
//   NL     components held by one thread (== N for lane-per-system,
//          N/32 for warp-per-system)
//   NPAR   parameters per system in global memory (SoA [NPAR][n_lanes])
//   NPL    parameters held by one thread
//   WARP   true: one warp integrates one system
//   PADDED true: N is not 32 * NL, slots with comp(k, lane) >= N are padding
//   comp(k, lane)        global component index of local slot k
//   load_params(...)     fill the thread's parameter registers
//   f(t, y, p, dy)       the derivative (may use warp shuffles if WARP)
//   FLOPS  nominal flop count of one evaluation (FMA = 2), for the roofline
//
// The library is compiled with -fmad=false and every fused operation is
// written as an explicit fma(): the expression trees below are then exactly
// the ones of oracle/xsq_oracle.c, which makes forced-step runs bit-comparable.
#pragma once

namespace xsq {
namespace rhs {

struct Lorenz63 {
    static constexpr int N = 3, NL = 3, NPAR = 3, NPL = 3;
    static constexpr bool WARP = false, PADDED = false;
    static constexpr int FLOPS = 8;
    __device__ __forceinline__ static int comp(int k, int) { return k; }
    __device__ __forceinline__ static void load_params(
        const double* __restrict__ params, long long sys, long long n_lanes,
        int, double (&p)[NPL]) {
#pragma unroll
        for (int k = 0; k < NPL; ++k) p[k] = params[k * n_lanes + sys];
    }
    __device__ __forceinline__ static void f(double, const double (&y)[NL],
                                             const double (&p)[NPL],
                                             double (&dy)[NL]) {
        dy[0] = p[0] * (y[1] - y[0]);
        dy[1] = fma(y[0], p[1] - y[2], -y[1]);
        dy[2] = fma(y[0], y[1], -(p[2] * y[2]));
    }
};

struct VanDerPol {
    static constexpr int N = 2, NL = 2, NPAR = 1, NPL = 1;
    static constexpr bool WARP = false, PADDED = false;
    static constexpr int FLOPS = 6;
    __device__ __forceinline__ static int comp(int k, int) { return k; }
    __device__ __forceinline__ static void load_params(
        const double* __restrict__ params, long long sys, long long n_lanes,
        int, double (&p)[NPL]) {
        p[0] = params[sys];
    }
    __device__ __forceinline__ static void f(double, const double (&y)[NL],
                                             const double (&p)[NPL],
                                             double (&dy)[NL]) {
        dy[0] = y[1];
        dy[1] = fma(p[0] * fma(-y[0], y[0], 1.0), y[1], -y[0]);
    }
};

// Planar restricted three-body problem in the rotating frame (Arenstorf
// orbit; Hairer-Norsett-Wanner I, eq. II.0.1).  y = (x, y, x', y').
struct Arenstorf {
    static constexpr int N = 4, NL = 4, NPAR = 1, NPL = 1;
    static constexpr bool WARP = false, PADDED = false;
    static constexpr int FLOPS = 60;
    __device__ __forceinline__ static int comp(int k, int) { return k; }
    __device__ __forceinline__ static void load_params(
        const double* __restrict__ params, long long sys, long long n_lanes,
        int, double (&p)[NPL]) {
        p[0] = params[sys];
    }
    __device__ __forceinline__ static void f(double, const double (&y)[NL],
                                             const double (&p)[NPL],
                                             double (&dy)[NL]) {
        const double mu = p[0], mup = 1.0 - mu;
        const double xa = y[0] + mu, xb = y[0] - mup;
        double d1 = fma(xa, xa, y[1] * y[1]);
        d1 = d1 * sqrt(d1);
        double d2 = fma(xb, xb, y[1] * y[1]);
        d2 = d2 * sqrt(d2);
        dy[0] = y[2];
        dy[1] = y[3];
        dy[2] = fma(2.0, y[3], y[0]) - mup * xa / d1 - mu * xb / d2;
        dy[3] = fma(-2.0, y[2], y[1]) - mup * y[1] / d1 - mu * y[1] / d2;
    }
};

// 32-body softened gravity, G = 1, warp per system: lane b owns body b.
// State order y = [pos(3*32), vel(3*32)], body-major (SURVEY.md 8d, C4 ii).
// params = (eps2, m[0..31]).
struct NBody32 {
    static constexpr int NB = 32;
    static constexpr int N = 6 * NB, NL = 6, NPAR = 1 + NB, NPL = 2;
    static constexpr bool WARP = true, PADDED = false;
    static constexpr int FLOPS = 31 * 20 * 32;
    __device__ __forceinline__ static int comp(int k, int lane) {
        return k < 3 ? 3 * lane + k : 3 * NB + 3 * lane + (k - 3);
    }
    __device__ __forceinline__ static void load_params(
        const double* __restrict__ params, long long sys, long long n_lanes,
        int lane, double (&p)[NPL]) {
        p[0] = params[sys];
        p[1] = params[(1 + lane) * n_lanes + sys];
    }
    __device__ __forceinline__ static void f(double, const double (&y)[NL],
                                             const double (&p)[NPL],
                                             double (&dy)[NL]) {
        const unsigned full = 0xffffffffu;
        const int lane = threadIdx.x & 31;
        double ax = 0.0, ay = 0.0, az = 0.0;
#pragma unroll 4
        for (int j = 0; j < NB; ++j) {
            const double xj = __shfl_sync(full, y[0], j);
            const double yj = __shfl_sync(full, y[1], j);
            const double zj = __shfl_sync(full, y[2], j);
            const double mj = __shfl_sync(full, p[1], j);
            const double dx = xj - y[0], dyy = yj - y[1], dz = zj - y[2];
            const double r2 = fma(dz, dz, fma(dyy, dyy, fma(dx, dx, p[0])));
            const double inv = 1.0 / (r2 * sqrt(r2));
            const double w = (j == lane) ? 0.0 : mj * inv;
            ax = fma(w, dx, ax);
            ay = fma(w, dyy, ay);
            az = fma(w, dz, az);
        }
        dy[0] = y[3];
        dy[1] = y[4];
        dy[2] = y[5];
        dy[3] = ax;
        dy[4] = ay;
        dy[5] = az;
    }
};

// A user system too large for one thread (n_state > 16; the reference takes any
// n, common.py:187-217): one warp per system, component i in slot i / 32 of
// lane i % 32, so that every access of the warp to the state is one contiguous
// run.  The user's device function returns ONE component,
//     double f_i(int i, double t, const double* y, const double* p)
// and sees the whole stage vector y[0..N) through shared memory (one buffer
// per warp of the 128-thread CTA).  Slots with i >= N are padding (zero).
template <int N_, int NPAR_, class Fn>
struct WideSystem {
    static constexpr int N = N_, NL = (N_ + 31) / 32, NPAR = NPAR_, NPL = NPAR_ > 0 ? NPAR_ : 1;
    static constexpr bool WARP = true, PADDED = (N_ % 32) != 0;
    static constexpr int FLOPS = 0;
    __device__ __forceinline__ static int comp(int k, int lane) { return k * 32 + lane; }
    __device__ __forceinline__ static void load_params(
        const double* __restrict__ params, long long sys, long long n_lanes,
        int, double (&p)[NPL]) {
#pragma unroll
        for (int k = 0; k < NPAR; ++k) p[k] = params[k * n_lanes + sys];
    }
    __device__ __forceinline__ static void f(double t, const double (&y)[NL],
                                             const double (&p)[NPL], double (&dy)[NL]) {
        __shared__ double stage[4][NL * 32];
        double* s = stage[(threadIdx.x >> 5) & 3];
        const int lane = threadIdx.x & 31;
        __syncwarp();                       // the previous evaluation has been read
#pragma unroll
        for (int k = 0; k < NL; ++k) s[k * 32 + lane] = y[k];
        __syncwarp();
#pragma unroll
        for (int k = 0; k < NL; ++k) {
            const int i = k * 32 + lane;
            dy[k] = i < N ? Fn::at(i, t, s, p) : 0.0;
        }
    }
};

}  // namespace rhs
}  // namespace xsq
