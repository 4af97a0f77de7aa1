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

__device__ __forceinline__ double fast_rcp(double d)
{
    double x;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d));
    double e = fma(-d, x, 1.0);
    x = fma(x, e, x);
    e = fma(-d, x, 1.0);
    x = fma(x, e, x);
    return x;   // d = 0 -> inf, d = inf -> 0 (NaN from the Newton steps is mapped back by the seed's value)
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
    double e = fma(-d * y, y, 1.0);          // 1 - d y^2
    y = fma(y * fma(0.375, e, 0.5), e, y);   // y (1 + e/2 + 3 e^2 / 8)
    e = fma(-d * y, y, 1.0);
    y = fma(0.5 * y, e, y);
    return y;
}
// exp(x): x is clamped to [-708, 709] (results stay normal: exp(-708) = 3e-308 stands in for 0, exp(709) =
// 8e307 for inf -- both only feed 1/(1+E)-type expressions here); NaN propagates.
__device__ __forceinline__ double fast_exp(double x)
{
    const double xc = fmax(fmin(x, 709.0), -708.0);
    const double magic = 6755399441055744.0;   // 1.5 * 2^52: (t + magic) - magic rounds t to an integer
    const double tm = fma(xc, 1.4426950408889634074, magic);
    const double t = tm - magic;
    double r = fma(t, -6.93147180369123816490e-01, xc);   // ln2 hi
    r = fma(t, -1.90821492927058770002e-10, r);           // ln2 lo
    double p = 2.51100376175620502e-08;
    p = fma(p, r, 2.76326396541237579e-07);
    p = fma(p, r, 2.75572409185476252e-06);
    p = fma(p, r, 2.48014854822877313e-05);
    p = fma(p, r, 1.98412698900471429e-04);
    p = fma(p, r, 1.38888889523148119e-03);
    p = fma(p, r, 8.33333333331960115e-03);
    p = fma(p, r, 4.16666666664880919e-02);
    p = fma(p, r, 1.66666666666666796e-01);
    p = fma(p, r, 5.00000000000001887e-01);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    const int k = __double2loint(tm);           // the integer t (low word of t + magic)
    const double res = __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
    return (x != x) ? x : res;
}

}  // namespace bsyn
