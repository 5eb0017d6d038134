// hnm_detmath.h -- deterministic f64 elementary functions (sin, cos, exp, pow, acos).
//
// The reference (Rust f64::sin/cos/powf/exp/acos -> platform libm) is only
// defined up to the libm in use; CUDA's libdevice differs from glibc by 1-2
// ulp on a sizeable fraction of inputs.  These routines use nothing but IEEE
// + - * / sqrt and explicit fma, so the SAME source gives the SAME bits under
// gcc (-ffp-contract=off) and nvcc (-fmad=false).  The device code calls them
// everywhere the reference calls libm; the oracle has a build flavour that
// calls them too (bit-exact GPU-vs-oracle parity) next to its default flavour
// that calls glibc (tests/test_detmath.py bounds the difference: <= 1 ulp).
//
// Accuracy target: < 0.52 ulp for exp/pow, < 0.6 ulp for sin/cos on the ranges
// the renderer uses (|x| <= 2*pi for sin/cos, base in [0,1] for pow).
#ifndef HNM_DETMATH_H
#define HNM_DETMATH_H

#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define HNM_HD __host__ __device__ __forceinline__
#define HNM_HD_NOINLINE __host__ __device__ __noinline__
#else
#define HNM_HD inline
#define HNM_HD_NOINLINE inline
#endif

namespace hnm {
namespace dm {

HNM_HD double fma_(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
    return __fma_rn(a, b, c);
#else
    return __builtin_fma(a, b, c);
#endif
}
HNM_HD double rint_(double x) {
#if defined(__CUDA_ARCH__)
    return rint(x);
#else
    return __builtin_rint(x);
#endif
}
HNM_HD double sqrt_(double x) {
#if defined(__CUDA_ARCH__)
    return __dsqrt_rn(x);
#else
    return __builtin_sqrt(x);
#endif
}
HNM_HD uint64_t bits_(double x) {
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(x);
#else
    uint64_t u; memcpy(&u, &x, 8); return u;
#endif
}
HNM_HD double from_bits_(uint64_t u) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    double x; memcpy(&x, &u, 8); return x;
#endif
}

struct dd { double hi, lo; };

HNM_HD dd two_sum(double a, double b) {
    double s = a + b;
    double bb = s - a;
    double e = (a - (s - bb)) + (b - bb);
    return dd{s, e};
}
HNM_HD dd fast_two_sum(double a, double b) {  // |a| >= |b|
    double s = a + b;
    double e = b - (s - a);
    return dd{s, e};
}
HNM_HD dd two_prod(double a, double b) {
    double p = a * b;
    double e = fma_(a, b, -p);
    return dd{p, e};
}
HNM_HD dd dd_add(dd a, dd b) {
    dd s = two_sum(a.hi, b.hi);
    dd t = two_sum(a.lo, b.lo);
    double c = s.lo + t.hi;
    dd v = fast_two_sum(s.hi, c);
    double w = t.lo + v.lo;
    return fast_two_sum(v.hi, w);
}
HNM_HD dd dd_add_d(dd a, double b) {
    dd s = two_sum(a.hi, b);
    double w = s.lo + a.lo;
    return fast_two_sum(s.hi, w);
}
HNM_HD dd dd_mul(dd a, dd b) {
    dd p = two_prod(a.hi, b.hi);
    double t = fma_(a.hi, b.lo, a.lo * b.hi);
    return fast_two_sum(p.hi, p.lo + t);
}
HNM_HD dd dd_mul_d(dd a, double b) {
    dd p = two_prod(a.hi, b);
    return fast_two_sum(p.hi, fma_(a.lo, b, p.lo));
}
HNM_HD dd dd_div(dd a, dd b) {
    double q1 = a.hi / b.hi;
    // r = a - q1*b
    dd p = dd_mul_d(b, q1);
    dd r = dd_add(a, dd{-p.hi, -p.lo});
    double q2 = r.hi / b.hi;
    return fast_two_sum(q1, q2);
}

// ---- constants (tools/gen_detmath_constants.py, mpmath at 400 bits) --------
#define HNM_LN2_HI 0x1.62e42fee00000p-1   /* top 32 bits: k*LN2_HI exact for |k| < 2^20 */
#define HNM_LN2_LO 0x1.a39ef35793c76p-33
#define HNM_LN2_LO2 0x1.cc01f97b57a08p-87
#define HNM_INV_LN2 0x1.71547652b82fep+0
#define HNM_PIO2_1 0x1.921fb54442d18p+0
#define HNM_PIO2_2 0x1.1a62633145c07p-54
#define HNM_PIO2_3 -0x1.f1976b7ed8fbcp-110
#define HNM_TWO_OVER_PI 0x1.45f306dc9c883p-1
#define HNM_PI_HI 0x1.921fb54442d18p+1
#define HNM_PI_LO 0x1.1a62633145c07p-53
#define HNM_THIRD_HI 0x1.5555555555555p-2
#define HNM_THIRD_LO 0x1.5555555555555p-56
#define HNM_SQRT2 0x1.6a09e667f3bcdp+0

// 2^k for -1022 <= k <= 1023
HNM_HD double pow2i_(int k) { return from_bits_((uint64_t)(k + 1023) << 52); }

// exp(hi + lo), |lo| << |hi|.  Core of exp() and pow().
HNM_HD double exp_dd_(double hi, double lo) {
    if (hi != hi) return hi;
    if (hi > 709.782712893384) return from_bits_(0x7ff0000000000000ull);
    if (hi < -745.1332191019412) return 0.0;
    double k = rint_(hi * HNM_INV_LN2);
    // r = hi - k*ln2 as a double-double (t exact: k*LN2_HI has <= 43 bits)
    double t = fma_(-k, HNM_LN2_HI, hi);
    dd c = two_prod(k, HNM_LN2_LO);
    dd s = two_sum(t, -c.hi);
    double rl = ((s.lo - c.lo) - k * HNM_LN2_LO2) + lo;
    dd r = fast_two_sum(s.hi, rl);
    double x = r.hi;
    // exp(x) - 1 - x = x^2 * q(x), Taylor to x^13 (|x| <= 0.3466 -> rel. err < 2^-57)
    double q = 1.0 / 6227020800.0;            // 1/13!
    q = fma_(q, x, 1.0 / 479001600.0);        // 1/12!
    q = fma_(q, x, 1.0 / 39916800.0);
    q = fma_(q, x, 1.0 / 3628800.0);
    q = fma_(q, x, 1.0 / 362880.0);
    q = fma_(q, x, 1.0 / 40320.0);
    q = fma_(q, x, 1.0 / 5040.0);
    q = fma_(q, x, 1.0 / 720.0);
    q = fma_(q, x, 1.0 / 120.0);
    q = fma_(q, x, 1.0 / 24.0);
    q = fma_(q, x, 1.0 / 6.0);
    q = fma_(q, x, 0.5);
    // tail = x^2 q + r.lo * (1 + x)   (d/dx exp = exp ~ 1 + x)
    double tail = fma_(x * x, q, fma_(r.lo, x, r.lo));
    dd one_x = fast_two_sum(1.0, x);
    double y = one_x.hi + (one_x.lo + tail);
    int ki = (int)k;
    // scale in two steps so that subnormal results round once, correctly enough
    if (ki > 1000) { y *= pow2i_(1000); ki -= 1000; }
    else if (ki < -1000) { y *= pow2i_(-1000); ki += 1000; }
    return y * pow2i_(ki);
}

HNM_HD double exp(double x) { return exp_dd_(x, 0.0); }

// log(x) as a double-double, x > 0 finite.
HNM_HD dd log_dd_(double x) {
    uint64_t u = bits_(x);
    int e = 0;
    if ((u >> 52) == 0) {  // subnormal
        x *= 0x1p54;
        u = bits_(x);
        e = -54;
    }
    e += (int)(u >> 52) - 1023;
    double m = from_bits_((u & 0x000fffffffffffffull) | 0x3ff0000000000000ull);  // [1,2)
    if (m > HNM_SQRT2) { m *= 0.5; e += 1; }                                      // [0.7071, 1.4142]
    double f = m - 1.0;  // exact
    dd den = two_sum(2.0, f);
    dd s = dd_div(dd{f, 0.0}, den);  // s = f/(2+f), |s| <= 0.1716
    dd s2 = dd_mul(s, s);
    double z = s2.hi;
    // sum_{k>=2} z^k/(2k+1), k up to 13
    double q = 1.0 / 27.0;
    q = fma_(q, z, 1.0 / 25.0);
    q = fma_(q, z, 1.0 / 23.0);
    q = fma_(q, z, 1.0 / 21.0);
    q = fma_(q, z, 1.0 / 19.0);
    q = fma_(q, z, 1.0 / 17.0);
    q = fma_(q, z, 1.0 / 15.0);
    q = fma_(q, z, 1.0 / 13.0);
    q = fma_(q, z, 1.0 / 11.0);
    q = fma_(q, z, 1.0 / 9.0);
    q = fma_(q, z, 1.0 / 7.0);
    q = fma_(q, z, 1.0 / 5.0);
    q = q * (z * z);
    dd u3 = dd_mul(s2, dd{HNM_THIRD_HI, HNM_THIRD_LO});
    dd uu = dd_add_d(u3, q);      // s2/3 + s2^2/5 + ...
    dd su = dd_mul(s, uu);
    dd lm = dd_add(s, su);        // atanh(s)
    lm.hi *= 2.0; lm.lo *= 2.0;   // log(m)
    double ed = (double)e;
    dd el = fast_two_sum(ed * HNM_LN2_HI, ed * HNM_LN2_LO);
    return dd_add(el, lm);
}

// pow(x, y) with the libm special cases the renderer can reach
// (src/color.rs:26-48: base in [0,1] after saturate / texel/255).
HNM_HD double pow(double x, double y) {
    if (y == 0.0) return 1.0;
    if (x == 1.0) return 1.0;
    if (x != x || y != y) return x + y;
    const double inf = from_bits_(0x7ff0000000000000ull);
    bool y_int = (rint_(y) == y);
    bool y_odd = y_int && (y > -0x1p53 && y < 0x1p53) && (((long long)y) & 1);
    if (x == 0.0) {
        bool neg = (bits_(x) >> 63) && y_odd;
        if (y > 0.0) return neg ? -0.0 : 0.0;
        return neg ? -inf : inf;
    }
    double ax = x < 0.0 ? -x : x;
    double sign = 1.0;
    if (x < 0.0) {
        if (!y_int) return from_bits_(0x7ff8000000000000ull);
        if (y_odd) sign = -1.0;
    }
    if (y == inf) return ax > 1.0 ? inf : (ax < 1.0 ? 0.0 : 1.0);
    if (y == -inf) return ax > 1.0 ? 0.0 : (ax < 1.0 ? inf : 1.0);
    if (ax == inf) return y > 0.0 ? sign * inf : sign * 0.0;
    dd l = log_dd_(ax);
    dd p = dd_mul_d(l, y);
    return sign * exp_dd_(p.hi, p.lo);
}

// ---- sin / cos ---------------------------------------------------------------
// r = x - k*pi/2 as a double-double, k = rint(x*2/pi).  Good for |x| < ~1e5
// (the renderer only uses x = 2*pi*u, u in [0,1): src/material.rs:238,263,
// src/scene.rs:93).
HNM_HD int rem_pio2_(double x, double& rh, double& rl) {
    double k = rint_(x * HNM_TWO_OVER_PI);
    double t = fma_(-k, HNM_PIO2_1, x);
    dd c = two_prod(k, HNM_PIO2_2);
    dd s = two_sum(t, -c.hi);
    double lo = (s.lo - c.lo) - k * HNM_PIO2_3;
    dd r = fast_two_sum(s.hi, lo);
    rh = r.hi; rl = r.lo;
    return (int)((long long)k & 3);
}
// sin(x + y), |x| <= pi/4 + eps, |y| << |x|; Taylor to x^17
HNM_HD double ksin_(double x, double y) {
    double z = x * x;
    double p = 1.0 / 355687428096000.0;              //  1/17!
    p = fma_(p, z, -1.0 / 1307674368000.0);          // -1/15!
    p = fma_(p, z, 1.0 / 6227020800.0);              //  1/13!
    p = fma_(p, z, -1.0 / 39916800.0);               // -1/11!
    p = fma_(p, z, 1.0 / 362880.0);                  //  1/9!
    p = fma_(p, z, -1.0 / 5040.0);                   // -1/7!
    p = fma_(p, z, 1.0 / 120.0);                     //  1/5!
    p = fma_(p, z, -1.0 / 6.0);                      // -1/3!
    // x + [x^3 p + y (1 - z/2)]
    double corr = fma_(y, fma_(-0.5, z, 1.0), (x * z) * p);
    return x + corr;
}
// cos(x + y); Taylor to x^18
HNM_HD double kcos_(double x, double y) {
    double z = x * x;
    double p = -1.0 / 6402373705728000.0;            // -1/18!
    p = fma_(p, z, 1.0 / 20922789888000.0);          //  1/16!
    p = fma_(p, z, -1.0 / 87178291200.0);            // -1/14!
    p = fma_(p, z, 1.0 / 479001600.0);               //  1/12!
    p = fma_(p, z, -1.0 / 3628800.0);                // -1/10!
    p = fma_(p, z, 1.0 / 40320.0);                   //  1/8!
    p = fma_(p, z, -1.0 / 720.0);                    // -1/6!
    p = fma_(p, z, 1.0 / 24.0);                      //  1/4!
    // 1 - z/2 + [z^2 p - x y]; h = 1 - z/2 with its rounding error recovered
    double hz = 0.5 * z;
    double w = 1.0 - hz;
    double werr = (1.0 - w) - hz;
    return w + (werr + fma_(z * z, p, -(x * y)));
}
HNM_HD void sincos(double x, double& s, double& c) {
    double rh, rl;
    int q = rem_pio2_(x, rh, rl);
    double sn = ksin_(rh, rl);
    double cs = kcos_(rh, rl);
    switch (q) {
        case 0: s = sn; c = cs; break;
        case 1: s = cs; c = -sn; break;
        case 2: s = -sn; c = -cs; break;
        default: s = -cs; c = sn; break;
    }
}
HNM_HD double sin(double x) { double s, c; sincos(x, s, c); return s; }
HNM_HD double cos(double x) { double s, c; sincos(x, s, c); return c; }

// ---- acos (sphere uv, src/scene.rs:69-73) -------------------------------------
// asin(x)/x - 1 = z*R(z), z = x^2 in [0, 0.25]: Chebyshev fit (coefficients
// generated by tools/gen_detmath_constants.py).
HNM_HD double asin_r_(double z) {
    double p = 0x1.e529c6fce9bb4p-6;
    p = fma_(p, z, -0x1.3b416bb7d9257p-6);
    p = fma_(p, z, 0x1.406192d124629p-6);
    p = fma_(p, z, 0x1.8f193743418ffp-9);
    p = fma_(p, z, 0x1.31622469ce5adp-7);
    p = fma_(p, z, 0x1.3b49de7121487p-7);
    p = fma_(p, z, 0x1.7b027ee1dd585p-7);
    p = fma_(p, z, 0x1.c990ad3d8fdcap-7);
    p = fma_(p, z, 0x1.1c4efce23019fp-6);
    p = fma_(p, z, 0x1.6e8ba123e494cp-6);
    p = fma_(p, z, 0x1.f1c71c7a52ba3p-6);
    p = fma_(p, z, 0x1.6db6db6dac1e0p-5);
    p = fma_(p, z, 0x1.3333333333388p-4);
    p = fma_(p, z, 0x1.5555555555555p-3);
    return p;
}
HNM_HD double acos(double x) {
    double ax = x < 0.0 ? -x : x;
    if (!(ax <= 1.0)) return from_bits_(0x7ff8000000000000ull);
    if (ax <= 0.5) {
        // pi/2 - (x + x*z*R(z))
        double z = x * x;
        double w = x * (z * asin_r_(z));
        return (0.5 * HNM_PI_HI) - (x - ((0.5 * HNM_PI_LO) - w));
    }
    // acos(|x|) = 2 asin(sqrt((1-|x|)/2))
    double z = (1.0 - ax) * 0.5;
    double s = sqrt_(z);
    if (s == 0.0) return x > 0.0 ? 0.0 : HNM_PI_HI;  // acos(+-1)
    double c = fma_(-s, s, z) / (2.0 * s);  // sqrt residual
    double w = fma_(s, z * asin_r_(z), c);
    if (x > 0.0) return 2.0 * (s + w);
    return HNM_PI_HI - 2.0 * (s + (w - 0.5 * HNM_PI_LO));
}

}  // namespace dm
}  // namespace hnm
#endif
