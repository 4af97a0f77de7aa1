// bsyn_math.cuh -- branch-free FP64 reciprocal / division / rsqrt / exp for the hot path.
//
// CUDA's IEEE-exact double division, rsqrt() and exp() carry special-case branches (slow-path calls, BSSY /
// BSYNC reconvergence) and integer fix-up code that cost 20-35 instructions per call site; the model needs
// ~60 such calls per evaluation per lane.  Here: hardware seed (MUFU.RCP64H / RSQ64H, ~20 bits) + Newton
// steps in DFMA, and a magic-number range reduction + degree-11 near-minimax polynomial for exp.
// Accuracy ~1-2 ulp on normal arguments (parity budget: 1e-10 relative); arguments are well inside the
// normal range everywhere they are used (guards below keep inf / 0 / NaN semantics where the model relies
// on them).
#pragma once
#include <cuda_runtime.h>

namespace bsyn {

// Polynomial coefficients and range-reduction constants live in constant memory so that they are DFMA
// operands (c[bank][off]) instead of 64-bit immediates, each of which costs two UMOVs at every use.
static __constant__ double c_exp[16] = {
    2.51100376175620502e-08, 2.76326396541237579e-07, 2.75572409185476252e-06, 2.48014854822877313e-05,
    1.98412698900471429e-04, 1.38888889523148119e-03, 8.33333333331960115e-03, 4.16666666664880919e-02,
    1.66666666666666796e-01, 5.00000000000001887e-01,
    1.4426950408889634074,          // [10] log2(e)
    6755399441055744.0,             // [11] 1.5 * 2^52
    -6.93147180369123816490e-01,    // [12] -ln2 hi
    -1.90821492927058770002e-10,    // [13] -ln2 lo
    708.0, -708.0};

// Library log / exp / log10 for rarely executed paths (LPTN tails, > 4 replicates per cell, generated
// quantities): out of line, so that their ~100-instruction bodies are not replicated at every inlined site.
static __device__ __noinline__ double cold_log(double x) { return log(x); }
static __device__ __noinline__ double cold_exp(double x) { return exp(x); }

// log(x) for normal positive x (the algorithm of fdlibm's e_log.c: x = 2^k (1+f), s = f/(2+f),
// log(1+f) = f - f^2/2 + s (f^2/2 + R(s^2)), R = degree-7 Remez polynomial; Lg1..Lg7 are its published
// coefficients).  No denormal / zero / negative handling: callers pass products of (f + lambda) >= lambda^n
// or 1 + positive.  Error < 1 ulp.
static __constant__ double c_log[10] = {
    6.666666666666735130e-01, 3.999999999940941908e-01, 2.857142874366239149e-01, 2.222219843214978396e-01,
    1.818357216161805012e-01, 1.531383769920937332e-01, 1.479819860511658591e-01,
    6.93147180369123816490e-01,   // [7] ln2 hi
    1.90821492927058770002e-10,   // [8] ln2 lo
    1.41421356237309504880};      // [9] sqrt(2)

__device__ __forceinline__ double fast_rcp(double d);   // below

__device__ __forceinline__ double fast_log(double x)
{
    int hi = __double2hiint(x);
    int k = (hi >> 20) - 1023;
    double mnt = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(x));   // [1, 2)
    if (mnt > c_log[9]) { mnt *= 0.5; ++k; }
    const double f = mnt - 1.0;
    const double s = f * fast_rcp(2.0 + f);
    const double z = s * s, w = z * z;
    const double t1 = w * fma(w, fma(w, c_log[5], c_log[3]), c_log[1]);
    const double t2 = z * fma(w, fma(w, fma(w, c_log[6], c_log[4]), c_log[2]), c_log[0]);
    const double R = t1 + t2;
    const double hfsq = 0.5 * f * f;
    const double dk = (double)k;
    return fma(dk, c_log[7], f - (hfsq - fma(s, hfsq + R, dk * c_log[8])));
}

__device__ __forceinline__ double fast_rcp(double d)
{
    double x;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d));
    // one third-order step x (1 + e + e^2): the seed is good to ~2^-20, so the result to ~2^-60; the dependent
    // chain is 3 DFMA after the MUFU instead of the 4 of two Newton steps (the kernels are latency-bound)
    const double e = fma(-d, x, 1.0);
    x = fma(x, fma(e, e, e), x);
    return x;   // d = 0 -> inf, d = inf -> 0 (NaN from the correction is mapped back by the seed's value)
}
// a / b with one residual correction
__device__ __forceinline__ double fast_div(double a, double b)
{
    const double x = fast_rcp(b);
    const double q = a * x;
    const double r = fma(-b, q, a);
    return fma(r, x, q);
}
__device__ __forceinline__ double fast_rsqrt(double d)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    const double e = fma(-d * y, y, 1.0);    // 1 - d y^2
    y = fma(y * fma(0.375, e, 0.5), e, y);   // y (1 + e/2 + 3 e^2 / 8): third order, ~2^-58 from a 2^-20 seed
    return y;
}
// exp(x): x is clamped to [-708, 708] (results stay normal: exp(-708) = 3e-308 stands in for 0, exp(708) =
// 3e307 for inf -- both only feed 1/(1+E)-type expressions here); NaN propagates.
__device__ __forceinline__ double fast_exp(double x)
{
    // clamp on the sign-magnitude form: ONE compare of |x| against an immediate and two integer selects (fmin / fmax on
    // doubles cost ~16 instructions per call in SASS: DSETP.MIN/MAX + NaN-quieting LOP3s + register-pair shuffling)
    const int hx = __double2hiint(x);
    const bool big = fabs(x) > 708.0;          // false for NaN, which flows on and is restored at the end
    const double xc = __hiloint2double(big ? ((hx & 0x80000000) | 0x40862000) : hx, big ? 0 : __double2loint(x));
    const double magic = c_exp[11];            // (t + magic) - magic rounds t to an integer
    const double tm = fma(xc, c_exp[10], magic);
    const double t = tm - magic;
    double r = fma(t, c_exp[12], xc);
    r = fma(t, c_exp[13], r);
    // 1 + r + a2 r^2 + ... + a11 r^11 (a_k = c_exp[11 - k]) as even and odd Horner chains in r^2: every DFMA has exactly one
    // coefficient, which is then a constant-bank operand (two-coefficient DFMAs, as in an Estrin scheme, make the compiler
    // hold the coefficients in registers); 12 FP64 instructions, dependent depth 7
    const double r2 = r * r, q0 = 1.0 + r;
    double ev = fma(c_exp[1], r2, c_exp[3]);   // a10 r2 + a8
    double od = fma(c_exp[0], r2, c_exp[2]);   // a11 r2 + a9
    ev = fma(ev, r2, c_exp[5]); od = fma(od, r2, c_exp[4]);
    ev = fma(ev, r2, c_exp[7]); od = fma(od, r2, c_exp[6]);
    ev = fma(ev, r2, c_exp[9]); od = fma(od, r2, c_exp[8]);
    const double p = fma(fma(od, r, ev), r2, q0);
    const int k = __double2loint(tm);           // the integer t (low word of t + magic)
    const double res = __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
    return (x != x) ? x : res;
}

}  // namespace bsyn
