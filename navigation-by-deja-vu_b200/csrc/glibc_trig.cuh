// glibc_trig.cuh -- sin / cos with the bits of the GNU C Library's double-precision
// sin() / cos(), for the device-resident stepping loop.
//
// The reference takes its trigonometry from the host libm: the glimpse rotation through libc
// cos/sin (navsim/util.pyx:7,144-145) and the move through np.cos/np.sin
// (navsim/NavBySceneFamiliarity.py:319-320; NumPy's float64 sin/cos are glibc's bit for bit,
// SURVEY.md H2).  glibc's functions are accurate to 0.55 ulp but NOT correctly rounded: a
// correctly rounded sin/cos differs from them in about 0.13 % of the arguments, CUDA's in
// more.  Bit-identical positions therefore need glibc's own
// algorithm: this file restates it -- third-party dependency of the reference, not vendored
// there: GNU C Library 2.39, sysdeps/ieee754/dbl-64/s_sin.c (IBM Accurate Mathematical
// Library, LGPL-2.1-or-later), table in glibc_sincos_table.h -- including the placement of
// the fused multiply-adds of the x86-64 FMA build of that file (the variant `sin` / `cos`
// resolve to on every AVX2+FMA host), taken from the instruction sequence of libm.so.6:
// with other contractions the last bit changes for a few arguments in a million.
//
// Domain: |x| < 105414350 (the stepping loop only ever passes angles in (-pi, 2 pi]);
// larger arguments, which glibc hands to a Payne-Hanek reduction, are the caller's business
// (nvb_glibc_sincos falls back to CUDA's sincos there, on the device).
// tests/test_glibc_trig.py compares the host build of this file with the host libm on
// 10^7 arguments; tests/test_gpu_glibc_trig.py does the same for the device build.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "glibc_sincos_table.h"

#if defined(__CUDACC__)
// device build: every function below is compiled for the device only (the host side of
// the engine takes its trigonometry from libm itself)
__device__ const double nvb_sincos_tab_dev[4 * NVB_SINCOS_ROWS] = NVB_SINCOS_TABLE_INIT;
#define NVB_TRIG_HD __device__ __forceinline__
#define NVB_FMA(a, b, c) __fma_rn((a), (b), (c))
#define NVB_MUL(a, b) __dmul_rn((a), (b))
#define NVB_ADD(a, b) __dadd_rn((a), (b))
#define NVB_SUB(a, b) __dsub_rn((a), (b))
#define NVB_TRIG_TAB nvb_sincos_tab_dev
#else
#define NVB_TRIG_HD static inline
// host build (tests): compile with -ffp-contract=off so that only the explicit fma() fuse
#define NVB_FMA(a, b, c) fma((a), (b), (c))
#define NVB_MUL(a, b) ((a) * (b))
#define NVB_ADD(a, b) ((a) + (b))
#define NVB_SUB(a, b) ((a) - (b))
#define NVB_TRIG_TAB nvb_sincos_tab_host
static const double nvb_sincos_tab_host[4 * NVB_SINCOS_ROWS] = NVB_SINCOS_TABLE_INIT;
#endif

namespace nvb_trig {

// s_sin.c / usncs.h constants (bit patterns read from libm.so.6)
#define NVB_T_BIG 0x1.8000000000000p+45      /* 1.5 * 2^45: adding it rounds to a multiple of 2^-7 */
#define NVB_T_SN3 (-0x1.5555555555515p-3)
#define NVB_T_SN5 0x1.11110e829872fp-7
#define NVB_T_CS2 0x1.0000000000000p-1
#define NVB_T_CS4 (-0x1.5555555555535p-5)
#define NVB_T_CS6 0x1.6c16bedd9e239p-10
#define NVB_T_S1 (-0x1.5555555555555p-3)
#define NVB_T_S2 0x1.1111111110ecep-7
#define NVB_T_S3 (-0x1.a01a019db08b8p-13)
#define NVB_T_S4 0x1.71de27b9a7ed9p-19
#define NVB_T_S5 (-0x1.addffc2fcdf59p-26)
#define NVB_T_TOINT 0x1.8000000000000p+52
#define NVB_T_HPINV 0x1.45f306dc9c883p-1
#define NVB_T_MP1 0x1.921fb58000000p+0
#define NVB_T_MP2 (-0x1.dde973c000000p-27)
#define NVB_T_PP3 (-0x1.cb3b398000000p-55)
#define NVB_T_PP4 (-0x1.d747f23e32ed7p-83)
#define NVB_T_HP0 0x1.921fb54442d18p+0
#define NVB_T_HP1 0x1.1a62633145c07p-54

NVB_TRIG_HD uint64_t bits(double x)
{
#if defined(__CUDACC__)
    return (uint64_t)__double_as_longlong(x);
#else
    uint64_t u;
    memcpy(&u, &x, 8);
    return u;
#endif
}

// TAYLOR_SIN(xx, x, dx), |x| < 0.126
NVB_TRIG_HD double taylor_sin(double x, double dx)
{
    const double xx = NVB_MUL(x, x);
    double p = NVB_FMA(xx, NVB_T_S5, NVB_T_S4);
    p = NVB_FMA(xx, p, NVB_T_S3);
    p = NVB_FMA(xx, p, NVB_T_S2);
    p = NVB_FMA(xx, p, NVB_T_S1);
    const double t = NVB_FMA(p, x, -NVB_MUL(0.5, dx));   // p * x - 0.5 * dx, one rounding
    return NVB_ADD(x, NVB_FMA(xx, t, dx));
}

// do_sin(x, dx): sin(x + dx), |x| < 0.8555, dx a small correction
NVB_TRIG_HD double do_sin(double x, double dx)
{
    const double ax = fabs(x);
    if (ax < 0.126) return taylor_sin(x, dx);
    if (x <= 0.0) dx = -dx;
    const double u = NVB_ADD(ax, NVB_T_BIG);
    const int k = (int)(uint32_t)bits(u) * 4;
    const double xr = NVB_SUB(ax, NVB_SUB(u, NVB_T_BIG));
    const double xx = NVB_MUL(xr, xr);
    const double sn = NVB_TRIG_TAB[k], ssn = NVB_TRIG_TAB[k + 1], cs = NVB_TRIG_TAB[k + 2], ccs = NVB_TRIG_TAB[k + 3];
    const double s = NVB_ADD(xr, NVB_FMA(NVB_MUL(xr, xx), NVB_FMA(xx, NVB_T_SN5, NVB_T_SN3), dx));
    const double q = NVB_FMA(xx, NVB_FMA(xx, NVB_T_CS6, NVB_T_CS4), NVB_T_CS2);
    const double c = NVB_FMA(xr, dx, NVB_MUL(xx, q));
    const double cor = NVB_FMA(s, cs, NVB_FMA(-c, sn, NVB_FMA(s, ccs, ssn)));
    return copysign(NVB_ADD(sn, cor), x);
}

// do_cos(x, dx): cos(x + dx)
NVB_TRIG_HD double do_cos(double x, double dx)
{
    if (x < 0.0) dx = -dx;
    const double ax = fabs(x);
    const double u = NVB_ADD(ax, NVB_T_BIG);
    const int k = (int)(uint32_t)bits(u) * 4;
    const double xr = NVB_ADD(NVB_SUB(ax, NVB_SUB(u, NVB_T_BIG)), dx);
    const double xx = NVB_MUL(xr, xr);
    const double sn = NVB_TRIG_TAB[k], ssn = NVB_TRIG_TAB[k + 1], cs = NVB_TRIG_TAB[k + 2], ccs = NVB_TRIG_TAB[k + 3];
    const double s = NVB_FMA(NVB_MUL(xr, xx), NVB_FMA(xx, NVB_T_SN5, NVB_T_SN3), xr);
    const double q = NVB_FMA(xx, NVB_FMA(xx, NVB_T_CS6, NVB_T_CS4), NVB_T_CS2);
    const double c = NVB_MUL(xx, q);
    const double cor = NVB_FMA(-s, sn, NVB_FMA(-c, cs, NVB_FMA(-s, ssn, ccs)));
    return NVB_ADD(cs, cor);
}

// reduce_sincos(x): x = n * pi/2 + (a + da), |a| <= pi/4; returns n & 3
NVB_TRIG_HD int reduce(double x, double *a, double *da)
{
    const double t = NVB_FMA(x, NVB_T_HPINV, NVB_T_TOINT);
    const double xn = NVB_SUB(t, NVB_T_TOINT);
    const int n = (int)(uint32_t)bits(t) & 3;
    const double y = NVB_FMA(-xn, NVB_T_MP2, NVB_FMA(-xn, NVB_T_MP1, x));
    const double t2 = NVB_FMA(-xn, NVB_T_PP3, y);
    double db = NVB_FMA(-xn, NVB_T_PP3, NVB_SUB(y, t2));
    const double b = NVB_FMA(-xn, NVB_T_PP4, t2);
    db = NVB_ADD(db, NVB_FMA(-xn, NVB_T_PP4, NVB_SUB(t2, b)));
    *a = b;
    *da = db;
    return n;
}

NVB_TRIG_HD double do_sincos(double a, double da, int n)
{
    const double r = (n & 1) ? do_cos(a, da) : do_sin(a, da);
    return (n & 2) ? -r : r;
}

// high word of |x| as glibc compares it
NVB_TRIG_HD int32_t hi_abs(double x) { return (int32_t)((bits(x) >> 32) & 0x7fffffffu); }

#define NVB_TRIG_MAX_K 0x419921FB   /* |x| < 105414350: the reduction above is valid */

NVB_TRIG_HD double sin_(double x)
{
    const int32_t k = hi_abs(x);
    if (k < 0x3e500000) return x;                                  // |x| < 2^-26
    if (k < 0x3feb6000) return do_sin(x, 0.0);                     // |x| < 0.855469
    if (k < 0x400368fd) {                                          // |x| < 2.426265
        const double t = NVB_SUB(NVB_T_HP0, fabs(x));
        return copysign(do_cos(t, NVB_T_HP1), x);
    }
    double a, da;
    const int n = reduce(x, &a, &da);
    return do_sincos(a, da, n);
}

NVB_TRIG_HD double cos_(double x)
{
    const int32_t k = hi_abs(x);
    if (k < 0x3e400000) return 1.0;                                // |x| < 2^-27
    if (k < 0x3feb6000) return do_cos(x, 0.0);
    if (k < 0x400368fd) {
        const double y = NVB_SUB(NVB_T_HP0, fabs(x));
        const double a = NVB_ADD(y, NVB_T_HP1);
        const double da = NVB_ADD(NVB_SUB(y, a), NVB_T_HP1);
        return do_sin(a, da);
    }
    double a, da;
    const int n = reduce(x, &a, &da);
    return do_sincos(a, da, n + 1);
}

}  // namespace nvb_trig

// sin(x) and cos(x) with glibc's bits for |x| < 105414350
NVB_TRIG_HD void nvb_glibc_sincos(double x, double *s, double *c)
{
#if defined(__CUDACC__)
    if (nvb_trig::hi_abs(x) >= NVB_TRIG_MAX_K) {   // never reached by the stepping loop (angles are reduced mod 2 pi)
        sincos(x, s, c);
        return;
    }
#endif
    *s = nvb_trig::sin_(x);
    *c = nvb_trig::cos_(x);
}
