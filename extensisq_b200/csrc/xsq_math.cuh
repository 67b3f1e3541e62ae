// xsq_math.cuh -- the arithmetic of the controller's log2 / exp2 on table
// entries that are already loaded (tables: xsq_math_tables_gen.cuh).  Kept free
// of everything device specific except three bit-cast intrinsics, so that
// tests/test_devmath_oracle.py can compile THIS file for the host (g++, with
// shims for the intrinsics) and compare it bit for bit with the C oracle's
// restatement (oracle/xsq_devmath.h) without a GPU.
#pragma once
#ifndef XSQ_MATH_FN
#define XSQ_MATH_FN __device__ __forceinline__
#endif

namespace xsq {

XSQ_MATH_FN int log2_tab_offset(double x) {      // in doubles
    return (__double2hiint(x) >> 11) & (127 << 2);
}
XSQ_MATH_FN double log2_arith(double x, double inv, double l_hi, double l_lo,
                              const double* pol) {
    const int hi = __double2hiint(x);
    const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(x));
    const double r = fma(m, inv, -1.0);
    double q = fma(r, pol[5], pol[4]);
    q = fma(r, q, pol[3]);
    q = fma(r, q, pol[2]);
    q = fma(r, q, pol[1]);
    q = fma(r, q, pol[0]);
    const double t = fma(r, q, l_lo);
    // the unbiased exponent as a double without a conversion instruction: the
    // biased one (0..2047) in the low mantissa bits of 2^52, minus 2^52 + 1023
    const double ed = __hiloint2double(0x43300000, (hi >> 20) & 0x7ff) - 0x1.00000000003ffp52;
    return (ed + l_hi) + t;
}

// |z| < 1000, so 2^n never leaves the exponent range
struct Exp2Split { double r; int N; };
XSQ_MATH_FN Exp2Split exp2_split(double z) {
    const double t = z + 0x1.8p46;                     // rounds z to a multiple of 1/64
    Exp2Split s;
    s.N = __double2loint(t);                           // 64 n + j
    s.r = z - (t - 0x1.8p46);                          // |r| <= 2^-7, exact
    return s;
}
XSQ_MATH_FN double exp2_arith(const Exp2Split& s, double t_hi, double t_lo, const double* pol) {
    double p = fma(s.r, pol[4], pol[3]);
    p = fma(s.r, p, pol[2]);
    p = fma(s.r, p, pol[1]);
    p = fma(s.r, p, pol[0]);
    p = s.r * p;                                        // 2^r - 1
    const double v = fma(t_hi, p, t_lo) + t_hi;
    return __hiloint2double(__double2hiint(v) + ((s.N >> 6) << 20), __double2loint(v));
}

// ---- the step-size controller, common.py:249-287, in the log2 domain --------
// With l2 = log2(sum((err/scale)^2)) = 2 log2(error_norm) + log2 n:
//   safety    * error_norm^err_exp                     = 2^(a1s l2 + a0s)
//   safety_sc * error_norm^minbeta1 * err_old^minbeta2 = 2^(a1c l2 + a2c l2_old + a0c)
// (constants formed on the host, xsq_api.cu build_params).  Returns the factor
// by which h_abs is multiplied after an attempt; the caller has decided accept /
// tiny and handles a NaN / Inf error norm (the reference's max(0.2, nan) is 0.2).
// Accepting and rejecting lanes share ONE exp2.
//   z_extra: minalpha * log2(h / h_prev) when EXTRA (user sc_params), else unused
XSQ_MATH_FN unsigned long long dbits(double x) {
    return ((unsigned long long)(unsigned)__double2hiint(x) << 32) | (unsigned)__double2loint(x);
}
struct CtlConst {
    double a1s, a0s, a1c, a2c, a0c;
};
template <bool EXTRA, class E2>
XSQ_MATH_FN double ctl_factor_arith(const CtlConst& C, double l2, double l2_old, double z_extra,
                                    bool accept, bool second, bool rej, bool tiny,
                                    double max_factor, E2 e2) {
    const double z_std = fma(C.a1s, l2, C.a0s);
    double z_sc = fma(C.a1c, l2, fma(C.a2c, l2_old, C.a0c));
    if (EXTRA) z_sc += z_extra;
    const double raw = e2(second ? z_sc : z_std);
    // max(min_factor, .) on rejection and in the second order branch only
    double factor = raw;
    if ((!accept || second) && !(raw > 0.2)) factor = 0.2;
    // min(max_factor, .) in the second order branch; min(1, .) after a rejection
    const double hi = (accept && rej) ? 1.0
                    : (second ? max_factor : __hiloint2double(0x7ff00000, 0));
    if (!(factor < hi)) factor = hi;
    if (accept && tiny) factor = rej ? 1.0 : max_factor;
    return factor;
}

}  // namespace xsq
