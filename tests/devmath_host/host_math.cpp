// Compiles the DEVICE arithmetic (extensisq_b200/csrc/xsq_math.cuh) for the
// host: the three bit-cast intrinsics are shimmed, everything else is the very
// source the kernels use.  tests/test_devmath_oracle.py compares it bit for
// bit with the C oracle's restatement -- a check of the device code that needs
// no GPU.
#include <cmath>
#include <cstdint>
#include <cstring>
#define XSQ_MATH_FN static inline
static inline int __double2hiint(double x) { uint64_t b; std::memcpy(&b, &x, 8); return (int)(uint32_t)(b >> 32); }
static inline int __double2loint(double x) { uint64_t b; std::memcpy(&b, &x, 8); return (int)(uint32_t)b; }
static inline double __hiloint2double(int hi, int lo) {
    const uint64_t b = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo;
    double x; std::memcpy(&x, &b, 8); return x;
}
using std::fma;
#include "xsq_math.cuh"
#include "xsq_devmath_tables.h"     // same generator as xsq_math_tables_gen.cuh

namespace {
double h_log2(double x) {
    const double* T = c_xsq_lg_tab + xsq::log2_tab_offset(x);
    return xsq::log2_arith(x, T[0], T[1], T[2], c_xsq_lg_pol);
}
struct HostExp2 {
    double operator()(double z) const {
        const xsq::Exp2Split s = xsq::exp2_split(z);
        const double* T = c_xsq_e2_tab + ((s.N & 63) << 1);
        return xsq::exp2_arith(s, T[0], T[1], c_xsq_e2_pol);
    }
};
}  // namespace

extern "C" {
void host_log2(const double* x, double* o, long n) { for (long i = 0; i < n; ++i) o[i] = h_log2(x[i]); }
void host_exp2(const double* x, double* o, long n) { for (long i = 0; i < n; ++i) o[i] = HostExp2()(x[i]); }
// c = {a1s, a0s, a1c, a2c, a0c}; flags bit0 accept, 1 second, 2 rej, 3 tiny, 4 extra
void host_ctl(const double* c, const double* l2, const double* l2_old, const double* zx,
              const int* flags, const double* mf, double* o, long n) {
    xsq::CtlConst C{c[0], c[1], c[2], c[3], c[4]};
    for (long i = 0; i < n; ++i) {
        const int f = flags[i];
        if (f & 16)
            o[i] = xsq::ctl_factor_arith<true>(C, l2[i], l2_old[i], zx[i], f & 1, f & 2, f & 4, f & 8, mf[i], HostExp2());
        else
            o[i] = xsq::ctl_factor_arith<false>(C, l2[i], l2_old[i], 0.0, f & 1, f & 2, f & 4, f & 8, mf[i], HostExp2());
    }
}
}
