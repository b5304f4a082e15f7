/*
 * glibc_sincos.h -- sin(x) and cos(x) with the results of glibc's libm, bit for bit, in device code.
 *
 * Why: the reference evaluates the rotation of every Gauss-Newton iterate with libm sin / cos (viso.cpp:1410-1411,
 * :122-127) and everything downstream -- residuals, pivots, the signed convergence test, per-hypothesis support counts,
 * which hypothesis wins -- is a discontinuous function of those twelve numbers.  CUDA's sin / cos are accurate to 1-2 ulp
 * but not identical to glibc's (correct to ~0.55 ulp, not correctly rounded either), and one ulp was enough to change 6
 * of 4096 support counts at BASELINE configs[3] (round-1 verdict).  North_star asks for bit-exact inlier sets, so the
 * device follows glibc's algorithm operation for operation.
 *
 * What: glibc 2.39, sysdeps/ieee754/dbl-64/s_sin.c (IBM Accurate Mathematical Library; __sin / __cos, do_sin, do_cos,
 * TAYLOR_SIN, reduce_sincos), in the form the x86-64 build selected at run time on FMA-capable CPUs actually executes
 * (sysdeps/x86_64/fpu/multiarch/s_sin-fma.c = the same source compiled with -mfma -mavx2): which multiply-adds are
 * fused and which are not was read off the machine code of libm.so.6 (__sin_fma, __cos_fma) and is written out
 * explicitly below with fma() / separate roundings -- the translation unit is compiled with -fmad=false, so nothing
 * else gets contracted.  The interpolation table is glibc's own (glibc_sincostab.inc, tools/gen_sincostab.py).
 *
 * Range: high word of |x| below 0x419921FB (|x| < 0x1.921fbp+26 = 105414336) follows glibc exactly: tiny, Taylor (|x| < 0.126), table (|x| < 0.855469),
 * pi/2 - |x| (|x| < 2.426265) and the three-constant Cody-Waite reduction (reduce_sincos).  Larger arguments
 * (glibc: __branred) and non-finite ones take CUDA's sin / cos; the estimation never gets there with finite data
 * (the largest angle seen over 4096 hypotheses x 10000 outlier-ridden correspondences is 1.6e6).
 *
 * The same text compiles as host C++ (tests/host/sincos_replica.cpp, -ffp-contract=off) where it is compared with the
 * local libm over millions of arguments; tests/test_gpu_parity.py does the same with the device code.
 */
#ifndef VISO_GLIBC_SINCOS_H_
#define VISO_GLIBC_SINCOS_H_

#include <math.h>
#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define VSC_FN __device__ __forceinline__
#define VSC_FMA(a, b, c) __fma_rn((a), (b), (c))
#define VSC_ADD(a, b) __dadd_rn((a), (b))
#define VSC_SUB(a, b) __dsub_rn((a), (b))
#define VSC_MUL(a, b) __dmul_rn((a), (b))
#define VSC_TAB_DECL __device__ const double
#define VSC_TAB(i) __ldg(&viso_sincostab[i])
#define VSC_HI(x) __double2hiint(x)
#define VSC_LO(x) __double2loint(x)
#else
#define VSC_FN static inline
#define VSC_FMA(a, b, c) fma((a), (b), (c))
#define VSC_ADD(a, b) ((a) + (b))
#define VSC_SUB(a, b) ((a) - (b))
#define VSC_MUL(a, b) ((a) * (b))
#define VSC_TAB_DECL static const double
#define VSC_TAB(i) viso_sincostab[i]
static inline int vsc_hi(double x) { int64_t b; memcpy(&b, &x, 8); return (int)(b >> 32); }
static inline int vsc_lo(double x) { int64_t b; memcpy(&b, &x, 8); return (int)(b & 0xffffffff); }
#define VSC_HI(x) vsc_hi(x)
#define VSC_LO(x) vsc_lo(x)
#endif

VSC_TAB_DECL viso_sincostab[440] = {
#include "glibc_sincostab.inc"
};

namespace viso_sc {

/* s_sin.c / usncs.h / trigo.h constants (hexadecimal: exactly the doubles of the binary) */
#define VSC_S1 (-0x1.5555555555555p-3)
#define VSC_S2 (0x1.1111111110ecep-7)
#define VSC_S3 (-0x1.a01a019db08b8p-13)
#define VSC_S4 (0x1.71de27b9a7ed9p-19)
#define VSC_S5 (-0x1.addffc2fcdf59p-26)
#define VSC_SN3 (-0x1.5555555555515p-3)
#define VSC_SN5 (0x1.11110e829872fp-7)
#define VSC_CS2 (0.5)
#define VSC_CS4 (-0x1.5555555555535p-5)
#define VSC_CS6 (0x1.6c16bedd9e239p-10)
#define VSC_BIG (0x1.8p+45)
#define VSC_HP0 (0x1.921fb54442d18p+0)
#define VSC_HP1 (0x1.1a62633145c07p-54)
#define VSC_TOINT (0x1.8p+52)
#define VSC_HPINV (0x1.45f306dc9c883p-1)
#define VSC_MP1 (0x1.921fb58000000p+0)
#define VSC_MP2 (-0x1.dde973c000000p-27)
#define VSC_PP3 (-0x1.cb3b398000000p-55)
#define VSC_PP4 (-0x1.d747f23e32ed7p-83)

/* TAYLOR_SIN (xx, x, dx): x + ((POLYNOMIAL(xx) * x - 0.5 * dx) * xx + dx) */
VSC_FN double taylor_sin(double x, double dx)
{
    const double xx = VSC_MUL(x, x);
    double p = VSC_S5;
    p = VSC_FMA(p, xx, VSC_S4);
    p = VSC_FMA(p, xx, VSC_S3);
    p = VSC_FMA(p, xx, VSC_S2);
    p = VSC_FMA(p, xx, VSC_S1);
    const double t = VSC_FMA(p, x, -VSC_MUL(dx, 0.5));
    return VSC_ADD(x, VSC_FMA(xx, t, dx));
}

/* do_sin (x, dx) */
VSC_FN double do_sin(double x, double dx)
{
    const double ax = fabs(x);
    if (ax < 0.126) return taylor_sin(x, dx);
    if (!(x > 0.0)) dx = -dx; /* if (x <= 0) dx = -dx */
    const double u = VSC_ADD(VSC_BIG, ax);
    const double x1 = VSC_SUB(ax, VSC_SUB(u, VSC_BIG));
    const int k = VSC_LO(u) << 2;
    const double xx = VSC_MUL(x1, x1);
    const double p = VSC_FMA(VSC_SN5, xx, VSC_SN3);
    const double s = VSC_ADD(x1, VSC_FMA(VSC_MUL(x1, xx), p, dx));
    double q = VSC_FMA(VSC_CS6, xx, VSC_CS4);
    q = VSC_FMA(q, xx, VSC_CS2);
    const double c = VSC_FMA(x1, dx, VSC_MUL(xx, q));
    const double sn = VSC_TAB(k), ssn = VSC_TAB(k + 1), cs = VSC_TAB(k + 2), ccs = VSC_TAB(k + 3);
    double e = VSC_FMA(s, ccs, ssn);
    e = VSC_FMA(-c, sn, e);
    const double cor = VSC_FMA(s, cs, e);
    return copysign(VSC_ADD(sn, cor), x);
}

/* do_cos (x, dx) */
VSC_FN double do_cos(double x, double dx)
{
    if (x < 0.0) dx = -dx;
    const double ax = fabs(x);
    const double u = VSC_ADD(VSC_BIG, ax);
    const double x1 = VSC_ADD(VSC_SUB(ax, VSC_SUB(u, VSC_BIG)), dx);
    const int k = VSC_LO(u) << 2;
    const double xx = VSC_MUL(x1, x1);
    const double p = VSC_FMA(VSC_SN5, xx, VSC_SN3);
    const double s = VSC_FMA(VSC_MUL(x1, xx), p, x1);
    double q = VSC_FMA(VSC_CS6, xx, VSC_CS4);
    q = VSC_FMA(q, xx, VSC_CS2);
    const double c = VSC_MUL(xx, q);
    const double sn = VSC_TAB(k), ssn = VSC_TAB(k + 1), cs = VSC_TAB(k + 2), ccs = VSC_TAB(k + 3);
    double e = VSC_FMA(-s, ssn, ccs);
    e = VSC_FMA(-c, cs, e);
    const double cor = VSC_FMA(-s, sn, e);
    return VSC_ADD(cs, cor);
}

/* reduce_sincos (x, &a, &da): x = n * pi/2 + (a + da), returns n (low bits of the rounded quotient) */
VSC_FN int reduce_sincos(double x, double* a, double* da)
{
    const double t = VSC_FMA(x, VSC_HPINV, VSC_TOINT);
    const double xn = VSC_SUB(t, VSC_TOINT);
    const int n = VSC_LO(t);
    double y = VSC_FMA(-xn, VSC_MP1, x);
    y = VSC_FMA(-xn, VSC_MP2, y);
    const double t2 = VSC_FMA(-xn, VSC_PP3, y);
    const double db1 = VSC_FMA(-xn, VSC_PP3, VSC_SUB(y, t2));
    const double b = VSC_FMA(-xn, VSC_PP4, t2);
    const double db2 = VSC_FMA(-xn, VSC_PP4, VSC_SUB(t2, b));
    *a = b;
    *da = VSC_ADD(db1, db2);
    return n;
}

VSC_FN double sin_glibc(double x)
{
    const int k = VSC_HI(x) & 0x7fffffff;
    if (k < 0x3e500000) return x; /* |x| < 2^-26 */
    if (k < 0x3feb6000) return do_sin(x, 0.0);
    if (k < 0x400368fd) {
        const double t = VSC_SUB(VSC_HP0, fabs(x));
        return copysign(do_cos(t, VSC_HP1), x);
    }
    if (k < 0x419921fb) {
        double a, da;
        const int n = reduce_sincos(x, &a, &da);
        const double r = (n & 1) ? do_cos(a, da) : do_sin(a, da);
        return (n & 2) ? -r : r;
    }
    return sin(x); /* __branred territory and inf / nan */
}

VSC_FN double cos_glibc(double x)
{
    const int k = VSC_HI(x) & 0x7fffffff;
    if (k < 0x3e400000) return 1.0; /* |x| < 2^-27 */
    if (k < 0x3feb6000) return do_cos(x, 0.0);
    if (k < 0x400368fd) {
        const double y = VSC_SUB(VSC_HP0, fabs(x));
        const double a = VSC_ADD(y, VSC_HP1);
        const double da = VSC_ADD(VSC_SUB(y, a), VSC_HP1);
        return do_sin(a, da);
    }
    if (k < 0x419921fb) {
        double a, da;
        const int n = reduce_sincos(x, &a, &da) + 1;
        const double r = (n & 1) ? do_cos(a, da) : do_sin(a, da);
        return (n & 2) ? -r : r;
    }
    return cos(x);
}

} // namespace viso_sc

#endif
