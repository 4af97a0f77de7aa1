// bsyn_model.cuh -- warp-cooperative log-density, hand-derived gradient and write_array of the
// gp_grid / nointeraction models.  One warp evaluates one (combination, chain).
//
// What it computes = reference model_gp_grid::log_prob<false,jacobian,var> + reverse sweep
// (src/stanExports_gp_grid.h:1076-1577; inst/stan/gp_grid.stan:64-233) and model_nointeraction::log_prob
// (src/stanExports_nointeraction.h:660-955); write_array = h:1705-2235.  It is NOT a translation of the
// generated C++: the Kronecker product uses the triangular structure, the Normal likelihoods use per-cell
// sufficient statistics, the Delta transform is evaluated without the logit, and all adjoints are
// hand-derived.
#pragma once
#include "bsyn_common.cuh"
#include "bsyn_math.cuh"

namespace bsyn {

#define BSYN_LOG_2PI_HALF 0.91893853320467274178
#define BSYN_LN10 2.30258509299404568402
#define BSYN_LOG_PI 1.14472988584940017414
#define BSYN_LOG_2 0.69314718055994530942
// constant-bank operands in device code (a 64-bit literal costs two moves at every use)
static __constant__ double c_kern[3] = {1.73205080756887729353, 2.23606797749978969641, 2.30258509299404568402};
#define BSYN_SQRT3 c_kern[0]
#define BSYN_SQRT5 c_kern[1]
// lgamma(2.25) - lgamma(1) - lgamma(1.25)  (beta(1,1.25) normaliser) = log(1.25)
#define BSYN_BETA_C 0.22314355131420927
#define BSYN_IG55_C 4.869135731822556          /* 5 log 5 - lgamma(5) */
#define BSYN_IG32_C 1.38629436111989061883     /* 3 log 2 - lgamma(3) */
#define BSYN_LOG_0_1 (-2.30258509299404568402)
// The heteroscedastic -n/2 log(f + lambda) terms of a lane are folded into ONE logarithm of a product of up to 9 factors
// v^n, n <= 4; a factor below this bound (only possible with a user lambda << the default 0.005) takes its own logarithm
// instead, so the product stays a normal double (1e-8^36 = 1e-288).
#define BSYN_VPROD_MIN 1e-8
// fixed even grids: 2 x 2 register-blocked products (block_product) instead of the 4-row strips
#ifndef BSYN_USE_BLOCK_PRODUCT
#define BSYN_USE_BLOCK_PRODUCT 1
#endif

template <int CPL> struct WVec {
    double z[CPL];
    double s;
};

// cell descriptor per lane slot t: i | j<<8 | valid<<16 | (valid && z is a parameter)<<17
template <int CPL, class SP = SpecRT>
__device__ __forceinline__ void make_cells(const Flags &F, const ProbS &P, int (&cell)[CPL])
{
    const OptV<SP> O(F);
    const int lane = threadIdx.x & 31;
    const int n2 = SP::NF ? SP::NF : P.n2, G = SP::NF ? SP::NF * SP::NF : P.G;
#pragma unroll
    for (int t = 0; t < CPL; ++t) {
        int c = lane + 32 * t;
        int valid = c < G;
        int cc = valid ? c : 0;
        int j = cc / n2, i = cc - j * n2;
        cell[t] = i | (j << 8) | (valid << 16) | ((valid && O.model() == 0) << 17);
    }
}

// kernel value from the stored exponential (and, for the backward pass, its partials)
template <class OV> __device__ __forceinline__ double kern_arg(const OV &F, double d, double iell, double alpha)
{
    // argument x of ev = fast_exp(x) (RQ: ev = fast_exp(-alpha*log1p(w)) handled by the caller)
    if (F.kernel() == 1) return -0.5 * d * d * iell * iell;
    if (F.nu_matern() == 1) return -d * iell;
    if (F.nu_matern() == 2) return -BSYN_SQRT3 * d * iell;
    return -BSYN_SQRT5 * d * iell;
}
template <class OV> __device__ __forceinline__ double kern_val(const OV &F, double d, double iell, double ev)
{
    if (F.kernel() == 2 && F.nu_matern() == 2) { double r = BSYN_SQRT3 * d * iell; return (1.0 + r) * ev; }
    if (F.kernel() == 2 && F.nu_matern() == 3) { double r = BSYN_SQRT5 * d * iell; return (1.0 + r + r * r * (1.0 / 3.0)) * ev; }
    return ev;
}
template <class OV>
__device__ __forceinline__ void kern_partials(const OV &F, double d, double iell, double alpha, double ev,
                                              double &kv, double &dk, double &da)
{
    da = 0.0;
    if (F.kernel() == 1) { kv = ev; dk = ev * d * d * iell * iell * iell; }
    else if (F.kernel() == 2) {
        if (F.nu_matern() == 1) { kv = ev; dk = ev * d * iell * iell; }
        else if (F.nu_matern() == 2) { double r = BSYN_SQRT3 * d * iell; kv = (1.0 + r) * ev; dk = r * r * ev * iell; }
        else { double r = BSYN_SQRT5 * d * iell; kv = (1.0 + r + r * r * (1.0 / 3.0)) * ev; dk = r * r * (1.0 / 3.0) * (1.0 + r) * ev * iell; }
    } else {
        double w = d * d * iell * iell / (2.0 * alpha);
        double iw = 1.0 / (1.0 + w);
        kv = ev;
        dk = 2.0 * alpha * w * ev * iw * iell;
        da = ev * (-log1p(w) + w * iw);
    }
}

// ---------------------------------------------------------------------------------------------------
// Strip product: the one matrix-product shape every Kronecker / adjoint product is reduced to.
//   Out[r0..r0+3, c] = sum_{m < nm} A[r0..r0+3 + m*NM] * B[c*SBC + m*SBM]        (c < C)
// A lane owns a 4-row strip of one output column: per m it issues 2 LDS.128 (A) + 1 LDS.64 (B, shared by the
// strip) for 4 DFMA; rows r0..r0+3 beyond the valid size only ever produce values in padding rows (NM a multiple of
// 4), or are dropped by the st4 / st4t stores of the epilogue (fixed grids: NM = the even grid size).
// CF / KF > 0: compile-time column and term counts (fixed grids): the whole product is one predicated, fully unrolled
// pass, every load address an immediate offset.
template <int NM, int SBC, int SBM, int CF, int KF, class Epi>
__device__ __forceinline__ void strip_product(const double *A, const double *B, const int C, const int nm, Epi &&epi)
{
    constexpr int NS = (NM + 3) / 4;
    const int lane = threadIdx.x & 31;
    if constexpr (CF > 0 && KF > 0 && NS * CF <= 32) {
        const int s = lane < NS * CF ? lane : 0;      // idle lanes recompute item 0 and do not store
        const int c = s / NS, r0 = 4 * (s - c * NS);
        const double2 *Ap = reinterpret_cast<const double2 *>(A + r0);
        const double *Bp = B + c * SBC;
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
        for (int m = 0; m < KF; ++m) {
            const double2 x = Ap[m * (NM / 2)], y = Ap[m * (NM / 2) + 1];
            const double bv = Bp[m * SBM];
            a0 = fma(x.x, bv, a0); a1 = fma(x.y, bv, a1); a2 = fma(y.x, bv, a2); a3 = fma(y.y, bv, a3);
        }
        if (lane < NS * CF) epi(r0, c, a0, a1, a2, a3);
    } else {
        const int Cn = CF > 0 ? CF : C, nmn = KF > 0 ? KF : nm;
        for (int s = lane; s < NS * Cn; s += 32) {
            const int c = s / NS, r0 = 4 * (s - c * NS);
            const double2 *Ap = reinterpret_cast<const double2 *>(A + r0);
            const double *Bp = B + c * SBC;
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll 2
            for (int m = 0; m < nmn; ++m) {
                const double2 x = Ap[m * (NM / 2)], y = Ap[m * (NM / 2) + 1];
                const double bv = Bp[m * SBM];
                a0 = fma(x.x, bv, a0); a1 = fma(x.y, bv, a1); a2 = fma(y.x, bv, a2); a3 = fma(y.y, bv, a3);
            }
            epi(r0, c, a0, a1, a2, a3);
        }
    }
}
__device__ __forceinline__ void st2(double *p, double x, double y) { *reinterpret_cast<double2 *>(p) = make_double2(x, y); }
// store the 4 strip values at p[0..3] (rows r0..r0+3 of a column of length NM); the upper pair is dropped when the
// column ends at r0 + 2 (NM not a multiple of 4)
template <int NM> __device__ __forceinline__ void st4(double *p, int r0, double a0, double a1, double a2, double a3)
{
    st2(p, a0, a1);
    if ((NM & 3) == 0 || r0 + 2 < NM) st2(p + 2, a2, a3);
}
// transposed store: element (r0 + e) goes to p[e * NM]
template <int NM> __device__ __forceinline__ void st4t(double *p, int r0, double a0, double a1, double a2, double a3)
{
    p[0] = a0; p[NM] = a1;
    if ((NM & 3) == 0 || r0 + 2 < NM) { p[2 * NM] = a2; p[3 * NM] = a3; }
}

// Fixed-grid variant of the product (CF columns, KF terms, both even): a lane owns a 2 x 2 block of the output,
//   Out[r0..r0+1, c0..c0+1] = sum_m A[r0..r0+1 + m*NM] * B[(c0..c0+1)*SBC + m*SBM],
// so per term it issues ONE LDS.128 for its two rows of A and one for its two columns of B (SBC == 1), or per PAIR of
// terms two LDS.128 along m (SBM == 1): 8 shared-memory wavefronts per term instead of the strip mapping's 10, and the
// transposed copy of a result is two 16-byte stores instead of four conflicting 8-byte ones.  (NM/2) x (CF/2) <= 32 lanes.
template <int NM, int SBC, int SBM, int CF, int KF, class Epi>
__device__ __forceinline__ void block_product(const double *A, const double *B, Epi &&epi)
{
    static_assert((NM & 1) == 0 && (CF & 1) == 0 && (KF & 1) == 0 && (NM / 2) * (CF / 2) <= 32, "block_product: even sizes, <= 32 blocks");
    static_assert(SBC == 1 || SBM == 1, "block_product: B must be contiguous along c or along m");
    constexpr int RB = NM / 2, NB = RB * (CF / 2);
    const int lane = threadIdx.x & 31;
    const int s = lane < NB ? lane : 0;      // idle lanes recompute block 0 and do not store
    const int cp = s / RB, r0 = 2 * (s - cp * RB), c0 = 2 * cp;
    const double2 *Ap = reinterpret_cast<const double2 *>(A + r0);
    double a00 = 0.0, a10 = 0.0, a01 = 0.0, a11 = 0.0;   // a<row><col>
    if constexpr (SBC == 1) {
        const double2 *Bp = reinterpret_cast<const double2 *>(B + c0);
#pragma unroll
        for (int m = 0; m < KF; ++m) {
            const double2 x = Ap[m * (NM / 2)], b = Bp[m * (SBM / 2)];
            a00 = fma(x.x, b.x, a00); a10 = fma(x.y, b.x, a10); a01 = fma(x.x, b.y, a01); a11 = fma(x.y, b.y, a11);
        }
    } else {
        const double2 *B0 = reinterpret_cast<const double2 *>(B + c0 * SBC), *B1 = reinterpret_cast<const double2 *>(B + (c0 + 1) * SBC);
#pragma unroll
        for (int m = 0; m < KF; m += 2) {
            const double2 x = Ap[m * (NM / 2)], y = Ap[(m + 1) * (NM / 2)], b0 = B0[m / 2], b1 = B1[m / 2];
            a00 = fma(x.x, b0.x, a00); a10 = fma(x.y, b0.x, a10); a01 = fma(x.x, b1.x, a01); a11 = fma(x.y, b1.x, a11);
            a00 = fma(y.x, b0.y, a00); a10 = fma(y.y, b0.y, a10); a01 = fma(y.x, b1.y, a01); a11 = fma(y.y, b1.y, a11);
        }
    }
    if (lane < NB) epi(r0, c0, a00, a10, a01, a11);
}
// stores of a 2 x 2 block: column-major Out[r + c*NM], and its transpose Out[c + r*NM]
template <int NM> __device__ __forceinline__ void stb(double *Out, int r0, int c0, double a00, double a10, double a01, double a11)
{
    st2(Out + r0 + c0 * NM, a00, a10);
    st2(Out + r0 + (c0 + 1) * NM, a01, a11);
}
template <int NM> __device__ __forceinline__ void stbt(double *Out, int r0, int c0, double a00, double a10, double a01, double a11)
{
    st2(Out + c0 + r0 * NM, a00, a01);
    st2(Out + c0 + (r0 + 1) * NM, a10, a11);
}

template <int CPL, class SP = SpecRT> __device__ __forceinline__ int zoff(const int cell)   // cell -> offset in a grid-shaped buffer
{
    return (cell & 255) + ((cell >> 8) & 255) * Cfg<CPL, SP>::NM;
}

// ---------------------------------------------------------------------------------------------------
// GP forward: kernel matrices -> two Cholesky factors -> GP = L2 * Z * L1^T.
// Cholesky: half-warp h factorises matrix h, lane r of the half keeps row r in registers; the pivot travels by
// shuffle, column k is published to the column-major factor in shared memory and read back (broadcast LDS.128)
// for the trailing update.  Leaves in scratch: L1r, L2r (row-major), L2c (column-major), invd (1/diag), exv
// (exp table), Zs = Z, Ts = T = L2*Z, Tt = T^T, Gs = GP.  Returns GP of this lane's cells in Gv.
template <int CPL, class SP>
__device__ __forceinline__ bool gp_forward(const OptV<SP> &F, const ProbS &P, const int ws, const int (&cell)[CPL],
                                           const double (&z)[CPL], double (&Gv)[CPL])
{
    using C = Cfg<CPL, SP>;
    constexpr int NM = C::NM, NF = SP::NF;
    const int lane = threadIdx.x & 31;
    const int n1 = NF ? NF : P.n1, n2 = NF ? NF : P.n2;
    double *const th = smem + ws + C::OFF_TH;
    double *const invd = smem + ws + C::OFF_INVD, *const L1r = smem + ws + C::OFF_L1R, *const L2r = smem + ws + C::OFF_L2R,
                 *const L2c = smem + ws + C::OFF_L2C, *const L1c = smem + ws + C::OFF_WS;
    double *const Zs = smem + ws + C::OFF_ZS, *const Ts = smem + ws + C::OFF_TS, *const Tt = smem + ws + C::OFF_TT,
                 *const Gs = smem + ws + C::OFF_GS, *const exv = smem + ws + C::OFF_EX;
    const double sigf = th[SL_SIGF], alpha = th[SL_ALPHA];
    const double iell = fast_rcp(th[SL_ELL]);
    const int P1 = NF ? NF * (NF - 1) / 2 : P.P1, Pn = NF ? NF * (NF - 1) : P.P1 + P.P2;
#pragma unroll 3
    for (int e0 = lane; e0 < Pn; e0 += 32) {
        const double d = smem[P.dist + e0];
        const int ab = SMI(P.pr + e0), a = ab >> 8, b = ab & 255;
        double ev;
        if (F.kernel() == 3) {
            const double w = d * d * iell * iell / (2.0 * alpha);
            ev = fast_exp(-alpha * log1p(w));
        } else {
            ev = fast_exp(kern_arg(F, d, iell, alpha));
        }
        exv[e0] = ev;
        const double kv = kern_val(F, d, iell, ev);
        if (e0 < P1) L1c[b * NM + a] = sigf * kv; else L2c[b * NM + a] = kv;   // lower triangle, column-major
    }
    __syncwarp();
    const int h = lane >> 4, r = lane & 15;
    const int n = h ? n2 : n1;
    bool ok = true;
    {
        double *const Lc = L1c + h * (C::OFF_L2C - C::OFF_WS);   // (arithmetic, not a select: the compiler re-derives a select per column)
        const double dg = h ? (1.0 + 1e-10) : (sigf + 1e-10);
        double a[NM];
        // (measured dead ends, profiles/r02_ab_nuts_v11_split.txt: writing the diagonal to shared memory and loading row r
        // unconditionally -- no selects, 60 instructions fewer -- is 4 % SLOWER, the store -> barrier -> load round trip sits
        // on the critical path; packing i + j * NM into the cell word, 50 instructions fewer, is neutral)
        static_for<0, NM>([&](auto kc) {
            constexpr int k = decltype(kc)::value;
            const double lo = (r < n && k < r) ? Lc[k * NM + r] : 0.0;
            a[k] = (k == r) ? ((r < n) ? dg : 1.0) : lo;
        });
        __syncwarp();
        static_for<0, NM>([&](auto kc) {
            constexpr int k = decltype(kc)::value;
            const double dkk = __shfl_sync(FULL, a[k], (lane & 16) + k);
            ok = ok && (dkk > 0.0);
            const double inv = fast_rsqrt(dkk);
            a[k] *= inv;
            if (r == k) invd[h * 16 + k] = inv;
            if (r < NM) Lc[k * NM + r] = (r >= k) ? a[k] : 0.0;   // publish column k (zeros above the diagonal)
            __syncwarp();
            // trailing update a[j] -= L[r][k] * L[j][k], j > k.  Entries j > r (upper triangle) are updated too:
            // they are never read (masked to zero when the row is stored).
            constexpr int j0 = k + 1;
            if constexpr (j0 < NM && (j0 & 1)) a[j0] = fma(-a[k], Lc[k * NM + j0], a[j0]);
            constexpr int je = j0 + (j0 & 1);
            static_for<je / 2, NM / 2>([&](auto pc) {
                constexpr int j = 2 * decltype(pc)::value;
                const double2 l = *reinterpret_cast<const double2 *>(Lc + k * NM + j);
                a[j] = fma(-a[k], l.x, a[j]);
                a[j + 1] = fma(-a[k], l.y, a[j + 1]);
            });
        });
        double *const Lr = L1r + h * (C::OFF_L2R - C::OFF_L1R);
        if (r < NM) {
            static_for<0, NM / 2>([&](auto pc) {
                constexpr int k = 2 * decltype(pc)::value;
                st2(Lr + r * NM + k, (k <= r) ? a[k] : 0.0, (k + 1 <= r) ? a[k + 1] : 0.0);
            });
        }
    }
    // Z -> Zs ; T = L2*Z -> Ts, Tt ; G = T*L1^T -> Gs
#pragma unroll
    for (int t = 0; t < CPL; ++t)
        if (cell[t] >> 16) Zs[zoff<CPL, SP>(cell[t])] = z[t];
    __syncwarp();
    if constexpr (BSYN_USE_BLOCK_PRODUCT && NF > 0 && (NF & 1) == 0) {
        block_product<NM, NM, 1, NF, NF>(L2c, Zs, [&](int r0, int c0, double a00, double a10, double a01, double a11) {
            stb<NM>(Ts, r0, c0, a00, a10, a01, a11);
            stbt<NM>(Tt, r0, c0, a00, a10, a01, a11);
        });
        __syncwarp();
        block_product<NM, NM, 1, NF, NF>(Ts, L1r, [&](int r0, int c0, double a00, double a10, double a01, double a11) {
            stb<NM>(Gs, r0, c0, a00, a10, a01, a11);
        });
    } else {
        strip_product<NM, NM, 1, NF, NF>(L2c, Zs, n1, n2, [&](int r0, int c, double a0, double a1, double a2, double a3) {
            st4<NM>(Ts + r0 + c * NM, r0, a0, a1, a2, a3);
            st4t<NM>(Tt + c + r0 * NM, r0, a0, a1, a2, a3);
        });
        __syncwarp();
        strip_product<NM, NM, 1, NF, NF>(Ts, L1r, n1, n1, [&](int r0, int c, double a0, double a1, double a2, double a3) {
            st4<NM>(Gs + r0 + c * NM, r0, a0, a1, a2, a3);
        });
    }
    __syncwarp();
#pragma unroll
    for (int t = 0; t < CPL; ++t) Gv[t] = (cell[t] >> 16) ? Gs[zoff<CPL, SP>(cell[t])] : 0.0;
    return ok;
}

// Solve L^T y = s for one right-hand side held in registers (s is overwritten, y returned in s).
// Right-looking form: once y[k] is known every remaining entry is updated independently, so the dependent
// chain is NM steps long, not NM^2/2.  L row-major in shared memory, rows >= n are identity rows.
template <int NM> __device__ __forceinline__ void solve_LT(const double *Lr, const double *idg, double (&s)[NM])
{
    static_for<0, NM>([&](auto kc) {
        constexpr int k = NM - 1 - decltype(kc)::value;
        s[k] *= idg[k];
        static_for<0, (k + 1) / 2>([&](auto pc) {          // rr = 0 .. k-1 in pairs (rr = k itself is harmless: L[k][k]
            constexpr int rr = 2 * decltype(pc)::value;     //  is only touched when rr+1 == k, handled below)
            const double2 l = *reinterpret_cast<const double2 *>(Lr + k * NM + rr);
            s[rr] = fma(-l.x, s[k], s[rr]);
            if constexpr (rr + 1 < k) s[rr + 1] = fma(-l.y, s[k], s[rr + 1]);
        });
    });
}

// ---------------------------------------------------------------------------------------------------
// Constrain this lane's scalar slot (h:1086-1215, Stan Math lub_constrain/lb_constrain semantics) and
// publish th[16] / ur[16] to the warp scratch.  Returns the constrained value; e = fast_exp(x) where
// x = -|u| (la slots) or u (lb slots).
template <class OV>
__device__ __forceinline__ double constrain_slot(const OV &F, double u, const int th, const int ur, double &e)
{
    const int lane = threadIdx.x & 31;
    const bool act = (F.active_mask() >> lane) & 1u;
    const bool is_la = lane < 2, is_lb = (lane >= SL_SL1) && (lane < SL_COUNT);
    e = exp(is_la ? -fabs(u) : u);   // exact exp here: exp(u) = inf must propagate (the reference rejects it)
    double thv;
    if (is_la) {
        double rr = fast_rcp(1.0 + e);
        if (u > 0) { thv = rr; if (thv == 1.0) thv = 1.0 - 1e-15; }
        else { thv = 1.0 - rr; if (thv == 0.0) thv = 1e-15; }
    } else if (is_lb) {
        thv = e;
    } else {
        thv = u;
    }
    if (!act) thv = 0.0;
    if (lane < 16) { smem[th + lane] = thv; smem[ur + lane] = u; }
    __syncwarp();
    return thv;
}

// Likelihood of one cell of f (gp_grid.stan:212-232) from its sufficient statistics (Normal variants) or
// its sorted observations (LPTN variants).  Adds to lpl / slogbar, returns d lp / d f.
// The per-observation constants  -log(2pi)/2 - log(s)  are added once per evaluation by the caller, and the
// heteroscedastic  -n/2 log(f + lambda)  terms of all cells of a lane are folded into ONE logarithm:
// vprod accumulates prod (f+lambda)^n (n <= 4 replicates; otherwise the log is taken directly).
template <class OV>
__device__ __forceinline__ double lik_cell(const OV &O, const ProbS &P, int fc, double f, double s, double is2,
                                           double &lpl, double &slogbar, double &vprod, bool &ok)
{
    const Flags &F = O.F;
    const double n = smem[P.cnt + fc];
    if (n == 0.0) return 0.0;
    double fbar;
    double iv = 0.0;
    if (O.het()) {
        const double v = f + F.lambda;
        if (!(v > 0.0)) { ok = false; return 0.0; }
        iv = fast_rcp(v);
        if (n <= 4.0 && v >= BSYN_VPROD_MIN) {
            const double v2 = v * v;
            vprod *= (n == 1.0) ? v : (n == 2.0) ? v2 : (n == 3.0) ? v2 * v : v2 * v2;
        } else {
            lpl += -0.5 * n * cold_log(v);
        }
    }
    if (!O.robust()) {
        const double d = smem[P.ybar + fc] - f;
        const double Q = fma(n * d, d, smem[P.ss + fc]);
        if (O.het()) {
            lpl += -0.5 * Q * is2 * iv;
            fbar = iv * (-0.5 * n + is2 * (n * d + 0.5 * Q * iv));
            slogbar += -n + Q * is2 * iv;
        } else {
            lpl += -0.5 * Q * is2;
            fbar = n * d * is2;
            slogbar += -n + Q * is2;
        }
    } else {
        double isd = fast_rcp(s);
        if (O.het()) isd *= sqrt(iv);
        fbar = -0.5 * n * iv;
        const int o1 = SMI(P.cstart + fc + 1);
        for (int o = SMI(P.cstart + fc); o < o1; ++o) {
            const double r = (smem[P.ysort + o] - f) * isd;
            const double ar = fabs(r);
            double dl;
            if (ar <= F.tau) {
                lpl += -0.5 * r * r;
                dl = -r;
            } else {
                const double az = fmax(ar, 1.0001);
                const double la = cold_log(az);
                lpl += F.lptn_c0 - la - (F.lamL + 1.0) * cold_log(la);
                dl = (ar > 1.0001) ? -copysign(1.0, r) * fast_rcp(az) * (1.0 + (F.lamL + 1.0) * fast_rcp(la)) : 0.0;
            }
            fbar += dl * (-isd - 0.5 * r * iv);
            slogbar += -dl * r - 1.0;
        }
    }
    return fbar;
}

// Branch-free version for the Normal likelihoods: an unobserved (or masked: live = false) cell has n = 0 and
// every term vanishes, so the four cells of a lane are straight-line code the scheduler can interleave.
template <class OV>
__device__ __forceinline__ double lik_cell_normal(const OV &O, const ProbS &P, int fc, double f, double is2, bool live,
                                                  double &lpl, double &slogbar, double &vprod, bool &ok)
{
    const Flags &F = O.F;
    const double n = live ? smem[P.cnt + fc] : 0.0;
    const double d = smem[P.ybar + fc] - f;
    const double Q = fma(n * d, d, (n > 0.0) ? smem[P.ss + fc] : 0.0);
    if (O.het()) {
        const double v = f + F.lambda;
        ok = ok && ((v > 0.0) || (n == 0.0));
        const double iv = fast_rcp(v);
        // v^n for n <= 4 by selects over unconditionally computed powers (the nested conditional products compile to
        // a chain of branches otherwise)
        const double v2 = v * v, v3 = v2 * v, v4 = v2 * v2;
        const bool fold = (v >= BSYN_VPROD_MIN) && (n <= 4.0);
        double pw = (n == 1.0) ? v : 1.0;
        pw = (n == 2.0) ? v2 : pw;
        pw = (n == 3.0) ? v3 : pw;
        pw = (n == 4.0) ? v4 : pw;
        vprod *= fold ? pw : 1.0;
        if (!fold && n > 0.0) lpl += -0.5 * n * cold_log(v);
        const double qi = Q * is2 * iv;
        lpl = fma(-0.5, qi, lpl);
        slogbar += qi - n;
        return iv * (-0.5 * n + is2 * (n * d + 0.5 * Q * iv));
    }
    const double qi = Q * is2;
    lpl = fma(-0.5, qi, lpl);
    slogbar += qi - n;
    return n * d * is2;
}

// Optional generated-quantities sink of an evaluation = write_array (h:1705-2235; gp_grid.stan:234-299).
// When on, the evaluation also emits one output row
//   [constrained params | p0 | p01 | p02 | Delta | GP | ec50_1 ec50_2 | CPO[N] | dss_1 dss_2 |
//    rVUS_f rVUS_p0 VUS_Delta VUS_syn VUS_ant]            (nointeraction: without Delta, GP, last three)
// and/or the generated scalars (qout) and the per-chain CPO accumulator.  Keeping this inside the ONE
// evaluation function (instead of a second forward pass) keeps a single copy of the model code in the
// sampler kernel -- its instruction footprint is what bounds the sampler (profiles/).
struct GQ {
    bool on;
    int so;   // offset (doubles) of the sink block in shared memory, see GqSlot; read only where gq.on
    bool fwd; // value only: return after the forward pass with lp (ADVI's ELBO samples, log_prob<false, true>)
};
// The sinks live in a small shared-memory block written by the caller (gq_publish) instead of in the struct: as
// struct members they would occupy ~25 registers through every evaluation, 99 % of which have gq.on == false.
enum GqSlot { GQS_ROW = 0, GQS_QOUT, GQS_CPO, GQS_Y, GQS_II, GQS_UZ = 5 /* 4 doubles */, GQS_D = 9, GQS_FACC = 10, GQ_DOUBLES = 12 };
__device__ __forceinline__ void gq_publish(const int so, const int D, const double *yglob, const int *iiglob,
                                           const double (&uz)[4], double *row, double *qout, double *cpo_acc,
                                           double *f_acc = nullptr)
{
    smem[so + GQS_FACC] = __longlong_as_double((long long)f_acc);
    // every lane stores the same values to the same addresses
    smem[so + GQS_ROW] = __longlong_as_double((long long)row);
    smem[so + GQS_QOUT] = __longlong_as_double((long long)qout);
    smem[so + GQS_CPO] = __longlong_as_double((long long)cpo_acc);
    smem[so + GQS_Y] = __longlong_as_double((long long)yglob);
    smem[so + GQS_II] = __longlong_as_double((long long)iiglob);
    smem[so + GQS_UZ] = uz[0]; smem[so + GQS_UZ + 1] = uz[1]; smem[so + GQS_UZ + 2] = uz[2]; smem[so + GQS_UZ + 3] = uz[3];
    SMI(2 * (so + GQS_D)) = D;
    __syncwarp();
}
template <class T> __device__ __forceinline__ T *gq_ptr(const GQ &gq, const int k)
{
    return reinterpret_cast<T *>(__double_as_longlong(smem[gq.so + k]));
}
__device__ __forceinline__ int gq_D(const GQ &gq) { return SMI(2 * (gq.so + GQS_D)); }

// ---------------------------------------------------------------------------------------------------
// lp(u) and d lp / d u for one (problem, point), cooperatively on one warp.
// q / g are in lane layout.  Returns false where the reference would throw (reject): lp/g undefined.
// dbg (DBG only): global scratch for intermediate dumps (tests).
template <int CPL, bool DBG, class SP = SpecRT>
__device__ __forceinline__ bool warp_logp_grad(const Flags &F, const ProbS &P, const int ws, const int (&cell)[CPL],
                                               const WVec<CPL> &q, const int jacobian, double &lp_out, WVec<CPL> &g,
                                               double *dbg, const GQ &gq)
{
    using C = Cfg<CPL, SP>;
    constexpr int NM = C::NM, NF = SP::NF;
    const OptV<SP> O(F);
    const int lane = threadIdx.x & 31;
    const int n1 = NF ? NF : P.n1, n2 = NF ? NF : P.n2, nm = n1 + n2;
    double *const th = smem + ws + C::OFF_TH, *const ur = smem + ws + C::OFF_UR, *const pm = smem + ws + C::OFF_PM;
    double *const invd = smem + ws + C::OFF_INVD, *const L1r = smem + ws + C::OFF_L1R, *const L2r = smem + ws + C::OFF_L2R;
    double *const Zs = smem + ws + C::OFF_ZS, *const Ts = smem + ws + C::OFF_TS, *const Tt = smem + ws + C::OFF_TT,
                 *const Gs = smem + ws + C::OFF_GS, *const Ws = smem + ws + C::OFF_WS;
    double *const Lb1 = smem + ws + C::OFF_LB1, *const Lb2 = smem + ws + C::OFF_LB2, *const exv = smem + ws + C::OFF_EX;
    const bool gp = (O.model() == 0);
    const double J = jacobian ? 1.0 : 0.0;
    const bool act = (O.active_mask() >> lane) & 1u;
    const bool is_la = lane < 2, is_lb = (lane >= SL_SL1) && (lane < SL_COUNT);

    __syncwarp();   // previous users of the scratch are done
    // ---- 1. scalars: constrain, Jacobian, priors (gp_grid.stan:177-209) -----------------------------
    const double u = q.s;
    double e;
    const double thv = constrain_slot(O, u, ws + C::OFF_TH, ws + C::OFF_UR, e);
    const int npre = 2 * O.est_la() + 6 + (gp ? 4 + O.est_alpha() : 0);
    if (gq.on && act) {   // constrained parameters in declaration order
        double *const row = gq_ptr<double>(gq, GQS_ROW);
        const int pos = (lane < SL_M1) ? lane : (lane <= SL_ALPHA) ? lane - (O.est_la() ? 0 : 2) : gq_D(gq) - 3 + (lane - SL_S);
        if (row) row[pos] = thv;
    }
    double lpl = 0.0;   // lane-local part of lp
    double tb = 0.0;    // d lp / d(constrained scalar of this lane)
    {
        // one log1p and one reciprocal for all lanes (lane-specific arguments); the per-slot branches below
        // then contain only a handful of multiply-adds each
        const double l1 = fast_log(1.0 + (is_la ? e : (lane == SL_S ? thv * thv : 0.0)));   // log1p, abs err 1e-16
        const double ie = fast_rcp(is_la ? (1.0 - thv) : (lane == SL_S ? fma(thv, thv, 1.0) : thv));
        double *const ith = smem + ws + C::OFF_ITH;
        if (lane < 16) ith[lane] = ie;
        __syncwarp();
        if (act) {
            if (is_la) {
                lpl += J * (-fabs(u) - 2.0 * l1);
                lpl += BSYN_BETA_C + 0.25 * (-fmax(u, 0.0) - l1);   // beta(1,1.25): 0.25*log1m(theta)
                tb += -0.25 * ie;
            } else if (lane == SL_M1 || lane == SL_M2) {
                const int k = lane - SL_M1;
                const double d = thv - th[SL_TH1 + k], is2k = ith[SL_S21 + k];
                lpl += -BSYN_LOG_2PI_HALF - 0.5 * ur[SL_S21 + k] - 0.5 * d * d * is2k;
                tb += -d * is2k;
            } else if (lane == SL_TH1 || lane == SL_TH2) {
                const int k = lane - SL_TH1;
                const double d = th[SL_M1 + k] - thv, is2k = ith[SL_S21 + k];
                lpl += -0.5 * thv * thv - BSYN_LOG_2PI_HALF;
                tb += -thv + d * is2k;
            } else if (lane == SL_SL1 || lane == SL_SL2) {
                lpl += -thv + J * u;
                tb += -1.0;
            } else if (lane == SL_B1 || lane == SL_B2) {
                const double rr = (thv - 1.0) * 10.0;
                lpl += -BSYN_LOG_2PI_HALF - BSYN_LOG_0_1 - 0.5 * rr * rr + J * u;
                tb += -(thv - 1.0) * 100.0;
            } else if (lane == SL_ELL) {
                if (O.pcprior()) { lpl += F.log_lam12 - 2.0 * u - F.lam1 * ie + J * u; tb += -2.0 * ie + F.lam1 * ie * ie; }
                else { lpl += BSYN_IG55_C - 6.0 * u - 5.0 * ie + J * u; tb += -6.0 * ie + 5.0 * ie * ie; }
            } else if (lane == SL_SIGF) {
                if (O.pcprior()) { const double sq = sqrt(thv); lpl += -F.lam2 * sq - 0.5 * u + J * u; tb += -0.5 * F.lam2 * sq * ie - 0.5 * ie; }
                else { lpl += -BSYN_LOG_2PI_HALF - u - 0.5 * (u - 1.0) * (u - 1.0) + J * u; tb += -u * ie; }
            } else if (lane == SL_ALPHA) {
                lpl += -thv + J * u;
                tb += -1.0;
            } else if (lane == SL_S) {
                lpl += BSYN_LOG_2 - BSYN_LOG_PI - l1 + J * u;
                tb += -2.0 * thv * ie;
            } else {   // SL_S21, SL_S22: inv_gamma(3,2) + the N(theta_k, sqrt(s2_k)) prior of log10_ec50_k
                const int k = lane - SL_S21;
                const double d = th[SL_M1 + k] - th[SL_TH1 + k];
                lpl += BSYN_IG32_C - 4.0 * u - 2.0 * ie + J * u;
                tb += -4.0 * ie + 2.0 * ie * ie - 0.5 * ie + 0.5 * d * d * ie * ie;
            }
        }
    }
    // ---- 2-4. GP forward -----------------------------------------------------------------------------
    double Gv[CPL];
    bool ok = true;
    if (gp) ok = gp_forward<CPL, SP>(O, P, ws, cell, q.z, Gv);
    // ---- 5. monotherapy curves (gp_grid.stan:154-157) ---------------------------------------------
    // gfin: the reference's reverse sweep yields NaN (0 * inf) exactly where 10^(...) overflows or where the Bliss
    // product rounds to 0 or 1 (log(p0 / (1 - p0)) = +-inf); value and lp stay finite there.  The formulas below have
    // finite limits, so those points are flagged and the gradient is replaced by NaN at the end -- the callers then
    // do what Stan does with a non-finite gradient (divergent leapfrog / rejected initial point / ADVI error).
    bool gfin = true;
    double om = 0.0, pmv = 0.0, lam = 0.0;
    if (lane < nm) {
        const bool d1 = lane < n1;
        const double sl = th[d1 ? SL_SL1 : SL_SL2], mm = th[d1 ? SL_M1 : SL_M2];
        lam = th[d1 ? SL_LA1 : SL_LA2];
        const double earg = c_kern[2] * sl * (smem[P.xs + lane] - mm);
        gfin = earg <= 709.782712893384;
        const double Em = fast_exp(earg);
        om = fast_rcp(1.0 + Em);
        pmv = fma(1.0 - lam, om, lam);
        pm[lane] = pmv;
        ok = ok && (pmv >= 0.0) && (pmv <= 1.0);
    }
    __syncwarp();
    // ---- 6. cells: Bliss product, Delta transform, likelihood and their adjoints, fused ------------
    const double s = th[SL_S];
    const double is2 = fast_rcp(s * s);
    const double b1 = th[SL_B1], b2 = th[SL_B2];
    double b1bar = 0.0, b2bar = 0.0, slogbar = 0.0, vprod = 1.0;
    // generated quantities (only when gq.on): output offsets, VUS accumulators, f on its (n2+1) x (n1+1) grid
    const int G = NF ? NF * NF : P.G;
    // row layout: [params D][p0 G][p01 n1][p02 n2][Delta G][GP G][ec50 2][CPO N][dss 2][VUS 2 (+3)]
    double *const fg = Lb1;   // (n2+1)(n1+1) <= 2 NN doubles: spans Lb1..Lb2 (= L2c, dead by now), free until the adjoint products
    double vus[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int t = 0; t < CPL; ++t) {
        // no branch on cell validity: masked cells decode to (i, j) = (0, 0) and contribute exact zeros
        const bool live = (cell[t] >> 16) & 1;
        const int i = cell[t] & 255, j = (cell[t] >> 8) & 255;
        const double p0 = pm[j] * pm[n1 + i], q0 = 1.0 - p0;
        double f = p0, dfdG = 0.0, dfdb1 = 0.0, dfdb2 = 0.0, dfdp0 = 1.0;
        ok = ok && (p0 >= 0.0) && (p0 <= 1.0);
        if (gp) {
            gfin = gfin && (!live || (p0 > 0.0 && q0 > 0.0));
            const double B = Gv[t];
            const double e1 = fast_exp(b1 * B), e2 = fast_exp(-b2 * B);
            const double iD1 = fast_rcp(fma(p0, e1, q0)), iD2 = fast_rcp(fma(q0, e2, p0));
            const double t1 = q0 * iD1, u1 = p0 * e1 * iD1, t2 = p0 * iD2, u2 = q0 * e2 * iD2;
            const double Delta = q0 * t2 - p0 * t1;
            ok = ok && (Delta >= -1.0) && (Delta <= 1.0);
            f = p0 + Delta;
            const double w1 = p0 * t1 * u1, w2 = q0 * t2 * u2;
            dfdG = w1 * b1 + w2 * b2;
            dfdb1 = w1 * B;
            dfdb2 = w2 * B;
            dfdp0 = 1.0 - t1 - t2 + iD1 * u1 + iD2 * u2;
            if (DBG && live) { dbg[2 * C::NN + lane + 32 * t] = B; dbg[2 * C::NN + C::GM + lane + 32 * t] = f; }
            if (gq.on && live) {
                const double w = smem[P.vw + j] * smem[P.vw + n1 + i];
                vus[2] = fma(w, Delta, vus[2]); vus[3] = fma(w, fmin(Delta, 0.0), vus[3]); vus[4] = fma(w, fmax(Delta, 0.0), vus[4]);
                double *const row = gq_ptr<double>(gq, GQS_ROW);
                if (row) {
                    const int c = lane + 32 * t, o_delta = gq_D(gq) + G + nm;
                    row[npre + c] = q.z[t]; row[o_delta + c] = Delta; row[o_delta + G + c] = B;
                }
            }
        }
        const int fc = (i + 1) + (j + 1) * (n2 + 1);
        if (gq.on && live) {
            const double w = smem[P.vw + j] * smem[P.vw + n1 + i];
            vus[0] = fma(w, 1.0 - f, vus[0]); vus[1] = fma(w, 1.0 - p0, vus[1]);
            fg[fc] = f;
            double *const row = gq_ptr<double>(gq, GQS_ROW);
            if (row) row[gq_D(gq) + lane + 32 * t] = p0;
        }
        double fbar;
        if (!O.robust()) fbar = lik_cell_normal(O, P, fc, f, is2, live, lpl, slogbar, vprod, ok);
        else fbar = live ? lik_cell(O, P, fc, f, s, is2, lpl, slogbar, vprod, ok) : 0.0;
        if (gp && live) Gs[zoff<CPL, SP>(cell[t])] = fbar * dfdG;   // Gbar, in place of G
        b1bar = fma(fbar, dfdb1, b1bar);
        b2bar = fma(fbar, dfdb2, b2bar);
        if (live) Ws[lane + 32 * t] = fbar * dfdp0;   // p0bar, packed cell index i + j*n2
    }
    // boundary cells of f: f[0,a] = p01[a-1], f[b,0] = p02[b-1], f[0,0] = 1 (gp_grid.stan:170-172)
    double pmbar = 0.0;
    if (!O.robust()) {   // Normal likelihoods: the branch-free form on every lane (lanes beyond the edge cells masked off)
        const int fc = lane < n1 ? (lane + 1) * (n2 + 1) : (lane < nm ? lane - n1 + 1 : 0);
        pmbar = lik_cell_normal(O, P, fc, lane < nm ? pmv : 1.0, is2, lane <= nm, lpl, slogbar, vprod, ok);
    }
    if (lane <= nm) {
        const int fc = lane < n1 ? (lane + 1) * (n2 + 1) : (lane < nm ? lane - n1 + 1 : 0);
        if (O.robust()) pmbar = lik_cell(O, P, fc, lane < nm ? pmv : 1.0, s, is2, lpl, slogbar, vprod, ok);
        if (gq.on) {
            fg[fc] = lane < nm ? pmv : 1.0;
            double *const row = gq_ptr<double>(gq, GQS_ROW);
            if (row && lane < nm) row[gq_D(gq) + G + lane] = pmv;   // p01 then p02 are contiguous
        }
    }
    if (O.het()) lpl += -0.5 * fast_log(vprod);
    if (gq.on) {
        __syncwarp();
        double *const row = gq_ptr<double>(gq, GQS_ROW), *const qout = gq_ptr<double>(gq, GQS_QOUT);
        double *const cpo_acc = gq_ptr<double>(gq, GQS_CPO);
        const double *const yglob = gq_ptr<const double>(gq, GQS_Y);
        const int *const iiglob = gq_ptr<const int>(gq, GQS_II);
        const int o_delta = gq_D(gq) + G + nm, o_ec = gp ? o_delta + 2 * G : o_delta, o_cpo = o_ec + 2;
        const int o_dss = o_cpo + P.N, o_vus = o_dss + 2;
        // per-chain accumulator of the response surface f on its (n2+1) x (n1+1) grid -> posterior mean of f per cell
        if (double *const f_acc = gq_ptr<double>(gq, GQS_FACC)) {
            const int ncf = (n1 + 1) * (n2 + 1);
            for (int e = lane; e < ncf; e += 32) f_acc[e] += fg[e];
        }
        const double vr = warp_sum8(vus);                       // lane L holds the total of vus[L & 7]
        double v_f = __shfl_sync(FULL, vr, 0), v_p0 = __shfl_sync(FULL, vr, 1), v_d = __shfl_sync(FULL, vr, 2);
        double v_syn = __shfl_sync(FULL, vr, 3), v_ant = __shfl_sync(FULL, vr, 4);
        // CPO (always the Normal density, even when robust: gp_grid.stan:264-270)
        if (row || cpo_acc) {
            for (int n = lane; n < P.N; n += 32) {
                const double fv = fg[iiglob[n]];
                const double sd = O.het() ? s * sqrt(fv + F.lambda) : s;
                const double r = (yglob[n] - fv) / sd;
                const double cpo = cold_exp(BSYN_LOG_2PI_HALF + cold_log(sd) + 0.5 * r * r);
                if (row) row[o_cpo + n] = cpo;
                if (cpo_acc) cpo_acc[n] += cpo;
            }
        }
        // ec50, dss (include/dss.stan) on lanes 0 / 1
        double ec = 0.0, ds = 0.0;
        if (lane < 2) {
            const int nn = lane ? n2 : n1;
            const double *xx = smem + P.xs + (lane ? n1 : 0);
            double c1 = xx[0], c2 = xx[0];
            for (int k = 1; k < nn; ++k) { c1 = fmin(c1, xx[k]); c2 = fmax(c2, xx[k]); }
            const double la = th[SL_LA1 + lane], sl = th[SL_SL1 + lane], mm = th[SL_M1 + lane];
            ec = cold_exp(BSYN_LN10 * mm);
            const double val = (c2 - c1) + (la - 1.0) / sl * (cold_log(1.0 + cold_exp(BSYN_LN10 * sl * (c2 - mm))) -
                                                              cold_log(1.0 + cold_exp(BSYN_LN10 * sl * (c1 - mm)))) * (1.0 / BSYN_LN10);
            ds = 100.0 * (1.0 - val / (c2 - c1));
        }
        const double ec2 = __shfl_sync(FULL, ec, 1), ds2 = __shfl_sync(FULL, ds, 1);
        // zero-valued integrals are replaced by U(1e-6,1e-4) draws (gp_grid.stan:294-297)
        if (v_f == 0.0) v_f = 1e-6 + (1e-4 - 1e-6) * smem[gq.so + GQS_UZ + 0];
        if (v_p0 == 0.0) v_p0 = 1e-6 + (1e-4 - 1e-6) * smem[gq.so + GQS_UZ + 1];
        if (gp) {
            if (v_syn == 0.0) v_syn = 1e-6 + (1e-4 - 1e-6) * smem[gq.so + GQS_UZ + 2];
            if (v_ant == 0.0) v_ant = 1e-6 + (1e-4 - 1e-6) * smem[gq.so + GQS_UZ + 3];
        }
        if (lane == 0) {
            if (row) {
                row[o_ec] = ec; row[o_ec + 1] = ec2; row[o_dss] = ds; row[o_dss + 1] = ds2;
                row[o_vus] = v_f; row[o_vus + 1] = v_p0;
                if (gp) { row[o_vus + 2] = v_d; row[o_vus + 3] = v_syn; row[o_vus + 4] = v_ant; }
            }
            if (qout) {
                qout[0] = ec; qout[1] = ec2; qout[2] = ds; qout[3] = ds2; qout[4] = s;
                qout[5] = v_f; qout[6] = v_p0; qout[7] = v_d; qout[8] = v_syn; qout[9] = v_ant;
            }
        }
    }
    __syncwarp();
    if (gq.fwd) {   // lp without the reverse sweep: add the N(0,1) prior of z and reduce
        if (gp) {
#pragma unroll
            for (int t = 0; t < CPL; ++t)
                if (cell[t] >> 16) lpl = fma(-0.5 * q.z[t], q.z[t], lpl);
        }
        double lpf = warp_sum(lpl) - (double)P.N * (BSYN_LOG_2PI_HALF + ur[SL_S]);
        if (gp) lpf += -(double)G * BSYN_LOG_2PI_HALF;
        lp_out = lpf;
        ok = ok && (lpf == lpf);
        return __all_sync(FULL, ok);
    }
    // ---- 7. adjoints of the monotherapy curves: 6 sums delivered to their owner lanes by one butterfly ----
    {
        double v8[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
        if (lane < nm) {
            const bool d1 = lane < n1;
            if (d1) {
#pragma unroll
                for (int i = 0; i < n2; ++i) pmbar = fma(Ws[i + lane * n2], pm[n1 + i], pmbar);
            } else {
                const int i = lane - n1;
#pragma unroll
                for (int j = 0; j < n1; ++j) pmbar = fma(Ws[i + j * n2], pm[j], pmbar);
            }
            const double sl = th[d1 ? SL_SL1 : SL_SL2], mm = th[d1 ? SL_M1 : SL_M2];
            const double c = pmbar * (1.0 - lam) * c_kern[2] * (1.0 - om) * om;   // (1-la) E ln10/(1+E)^2
            const double dla = pmbar * (1.0 - om);            // d/d la = E/(1+E)
            const double dsl = -c * (smem[P.xs + lane] - mm); // d/d slope
            const double dm = c * sl;                         // d/d log10_ec50
            // value index = owner lane & 7: la1 0, la2 1, m1 2, m2 3, slope_1 6, slope_2 7
            v8[0] = d1 ? dla : 0.0; v8[1] = d1 ? 0.0 : dla;
            v8[2] = d1 ? dm : 0.0;  v8[3] = d1 ? 0.0 : dm;
            v8[6] = d1 ? dsl : 0.0; v8[7] = d1 ? 0.0 : dsl;
        }
        const double rsum = warp_sum8(v8);
        if (lane < 8) tb += rsum;   // lanes 4, 5 (theta_k) receive the zero sums of indices 4, 5
    }
    // ---- 8-10. adjoints of the Kronecker product, the Cholesky factors and the kernels --------------
    double ellbar = 0.0, sigfbar = 0.0, alphabar = 0.0;
    if (gp) {
        __syncwarp();   // p0bar (in Ws) has been consumed
        if constexpr (BSYN_USE_BLOCK_PRODUCT && NF > 0 && (NF & 1) == 0) {
            // the same six products on 2 x 2 blocks (fixed even grid); masks keep the triangles the adjoints need
            block_product<NM, 1, NM, NF, NF>(Gs, L1r, [&](int r0, int c0, double a00, double a10, double a01, double a11) {
                stb<NM>(Ws, r0, c0, a00, a10, a01, a11);                                    // W = Gbar L1
            });
            block_product<NM, NM, 1, NF, NF>(Tt, Gs, [&](int r0, int c0, double a00, double a10, double a01, double a11) {
                stb<NM>(Lb1, r0, c0, (r0 <= c0) ? a00 : 0.0, (r0 + 1 <= c0) ? a10 : 0.0, (r0 <= c0 + 1) ? a01 : 0.0,
                        (r0 + 1 <= c0 + 1) ? a11 : 0.0);                                    // L1bar[j = c][l = r], l <= j
            });
            __syncwarp();
            block_product<NM, 1, NM, NF, NF>(Ws, Zs, [&](int r0, int c0, double a00, double a10, double a01, double a11) {
                stb<NM>(Lb2, r0, c0, (r0 >= c0) ? a00 : 0.0, (r0 + 1 >= c0) ? a10 : 0.0, (r0 >= c0 + 1) ? a01 : 0.0,
                        (r0 + 1 >= c0 + 1) ? a11 : 0.0);                                    // L2bar[i = r][k = c], i >= k
            });
            if (DBG) {
#pragma unroll
                for (int t = 0; t < CPL; ++t)
                    if (cell[t] >> 16) dbg[2 * C::NN + 2 * C::GM + lane + 32 * t] = Gs[zoff<CPL, SP>(cell[t])];
            }
            __syncwarp();
            block_product<NM, NM, 1, NF, NF>(L2r, Ws, [&](int r0, int c0, double a00, double a10, double a01, double a11) {
                stb<NM>(Gs, r0, c0, a00, a10, a01, a11);                                    // Zbar = L2^T W
            });
            block_product<NM, 1, NM, NF, NF>(L1r, Lb1, [&](int r0, int c0, double a00, double a10, double a01, double a11) {
                stb<NM>(Zs, r0, c0, a00, a10, a01, a11);                                    // S1 = L1^T L1bar
            });
            block_product<NM, NM, 1, NF, NF>(L2r, Lb2, [&](int r0, int c0, double a00, double a10, double a01, double a11) {
                stb<NM>(Ts, r0, c0, a00, a10, a01, a11);                                    // S2 = L2^T L2bar
            });
            __syncwarp();
        } else {
        // W = Gbar * L1 -> Ws            W[i,l] = sum_j Gbar[i,j] L1[j,l]
            strip_product<NM, 1, NM, NF, NF>(Gs, L1r, n1, n1, [&](int r0, int c, double a0, double a1, double a2, double a3) {
                st4<NM>(Ws + r0 + c * NM, r0, a0, a1, a2, a3);
            });
            // L1bar = tril(Gbar^T T) -> Lb1 row-major: strip rows l of column j hold L1bar[j,l], kept for l <= j
            strip_product<NM, NM, 1, NF, NF>(Tt, Gs, n1, n2, [&](int r0, int c, double a0, double a1, double a2, double a3) {
                st4<NM>(Lb1 + r0 + c * NM, r0, (r0 <= c) ? a0 : 0.0, (r0 + 1 <= c) ? a1 : 0.0, (r0 + 2 <= c) ? a2 : 0.0,
                        (r0 + 3 <= c) ? a3 : 0.0);
            });
            __syncwarp();
            // L2bar = tril(W Z^T) -> Lb2 col-major [i + k*NM], kept for i >= k
            strip_product<NM, 1, NM, NF, NF>(Ws, Zs, n2, n1, [&](int r0, int c, double a0, double a1, double a2, double a3) {
                st4<NM>(Lb2 + r0 + c * NM, r0, (r0 >= c) ? a0 : 0.0, (r0 + 1 >= c) ? a1 : 0.0, (r0 + 2 >= c) ? a2 : 0.0,
                        (r0 + 3 >= c) ? a3 : 0.0);
            });
            if (DBG) {
#pragma unroll
                for (int t = 0; t < CPL; ++t)
                    if (cell[t] >> 16) dbg[2 * C::NN + 2 * C::GM + lane + 32 * t] = Gs[zoff<CPL, SP>(cell[t])];
            }
            __syncwarp();
            // Zbar = L2^T W -> Gs            Zbar[k,l] = sum_i L2[i,k] W[i,l]
            strip_product<NM, NM, 1, NF, NF>(L2r, Ws, n1, n2, [&](int r0, int c, double a0, double a1, double a2, double a3) {
                st4<NM>(Gs + r0 + c * NM, r0, a0, a1, a2, a3);
            });
            // Cholesky adjoint (Murray 2016): Xfull = L^-T (Phi + Phi^T) L^-1, Phi = tril(L^T Lbar), diag halved.
            // Stage A: P = L^T Lbar (only its lower triangle is used) -> S1 (over Zs), S2 (over Ts)
            strip_product<NM, 1, NM, NF, NF>(L1r, Lb1, n1, n1, [&](int r0, int c, double a0, double a1, double a2, double a3) {
                st4<NM>(Zs + r0 + c * NM, r0, a0, a1, a2, a3);
            });
            strip_product<NM, NM, 1, NF, NF>(L2r, Lb2, n2, n2, [&](int r0, int c, double a0, double a1, double a2, double a3) {
                st4<NM>(Ts + r0 + c * NM, r0, a0, a1, a2, a3);
            });
            __syncwarp();
        }
        // Zbar is complete: gradient wrt z (+ its N(0,1) prior)
#pragma unroll
        for (int t = 0; t < CPL; ++t) {
            const bool v = cell[t] >> 16;
            const double zb = v ? Gs[zoff<CPL, SP>(cell[t])] : 0.0;
            g.z[t] = v ? zb - q.z[t] : 0.0;
            if (v) lpl = fma(-0.5 * q.z[t], q.z[t], lpl);
            if (DBG && v) dbg[2 * C::NN + 3 * C::GM + lane + 32 * t] = zb;
        }
        {
            const int h = lane >> 4, r = lane & 15;
            const int n = h ? n2 : n1;
            const double *const Lr = L1r + h * (C::OFF_L2R - C::OFF_L1R);
            const double *const Sm = Zs + h * (C::OFF_TS - C::OFF_ZS);
            double *const Yb = Tt + h * (C::OFF_WS - C::OFF_TT);
            double *const Xb = Lb1 + h * (C::OFF_LB2 - C::OFF_LB1);
            const double *const idg = invd + h * 16;
            double sv[NM];
            // Stage B: column r of Y = L^-T S;  S[rr, r] = P[max, min] (P stored [a + b*NM], lower part valid)
            static_for<0, NM>([&](auto kc) {
                constexpr int rr = decltype(kc)::value;
                const int hi = rr > r ? rr : r, lo = rr > r ? r : rr;
                sv[rr] = (rr < n && r < n) ? Sm[hi + lo * NM] : 0.0;
            });
            solve_LT<NM>(Lr, idg, sv);
            __syncwarp();   // (Yb aliases buffers read by the strip products above)
            if (r < NM) static_for<0, NM>([&](auto kc) { constexpr int rr = decltype(kc)::value; Yb[rr * NM + r] = sv[rr]; });
            __syncwarp();
            // Stage C: row r of X = Y L^-1  (same solve applied to row r of Y)
            static_for<0, NM / 2>([&](auto pc) {
                constexpr int c = 2 * decltype(pc)::value;
                const double2 y2 = (r < NM) ? *reinterpret_cast<const double2 *>(Yb + r * NM + c) : make_double2(0.0, 0.0);
                sv[c] = (r < n) ? y2.x : 0.0;
                sv[c + 1] = (r < n) ? y2.y : 0.0;
            });
            solve_LT<NM>(Lr, idg, sv);
            if (r < NM) static_for<0, NM / 2>([&](auto pc) { constexpr int c = 2 * decltype(pc)::value; st2(Xb + r * NM + c, sv[c], sv[c + 1]); });
            __syncwarp();
        }
        if (DBG) {
            for (int e0 = lane; e0 < C::NN; e0 += 32) {
                dbg[e0] = L1r[e0]; dbg[C::NN + e0] = L2r[e0];
                dbg[4 * C::NN + 4 * C::GM + e0] = Lb1[e0]; dbg[5 * C::NN + 4 * C::GM + e0] = Lb2[e0];
            }
        }
        // contraction with d cov / d(ell, sigma_f, alpha): sum over all (a,b) of Sbar_sym * dcov
        //   = sum_{a>b} Xfull[a,b] * dcov[a,b] + 1/2 sum_a Xfull[a,a] * dcov[a,a]
        const double sigf = th[SL_SIGF], alpha = th[SL_ALPHA], iell = fast_rcp(th[SL_ELL]);
        const int P1 = NF ? NF * (NF - 1) / 2 : P.P1, Pn = NF ? NF * (NF - 1) : P.P1 + P.P2;
#pragma unroll 3
        for (int e0 = lane; e0 < Pn; e0 += 32) {
            const bool m1 = e0 < P1;
            const int ab = SMI(P.pr + e0), a = ab >> 8, b = ab & 255;
            const double X = (m1 ? Lb1 : Lb2)[a * NM + b];
            double kv, dk, da;
            kern_partials(O, smem[P.dist + e0], iell, alpha, exv[e0], kv, dk, da);
            const double sc = m1 ? sigf : 1.0;
            ellbar = fma(X * sc, dk, ellbar);
            alphabar = fma(X * sc, da, alphabar);
            if (m1) sigfbar = fma(X, kv, sigfbar);
        }
        if (lane < n1) sigfbar = fma(0.5, Lb1[lane * NM + lane], sigfbar);
    } else {
#pragma unroll
        for (int t = 0; t < CPL; ++t) g.z[t] = 0.0;
    }
    // ---- 11. one butterfly for the 7 remaining sums; chain rule through the constraining transforms ------
    // value index = owner lane & 7: b1 (lane 8) 0, b2 1, ell 2, sigma_f 3, alpha 4, s (lane 13) 5; 6 = lp
    double lp;
    {
        const double v8[8] = {b1bar, b2bar, ellbar, sigfbar, alphabar, slogbar, lpl, 0.0};
        const double rsum = warp_sum8(v8);
        if (lane >= SL_B1 && lane <= SL_ALPHA) tb += rsum;
        lp = __shfl_sync(FULL, rsum, 6);
        lp += -(double)P.N * (BSYN_LOG_2PI_HALF + ur[SL_S]);
        if (gp) lp += -(double)G * BSYN_LOG_2PI_HALF;
        double gs;
        if (is_la) gs = tb * thv * (1.0 - thv) + J * (1.0 - 2.0 * thv);
        else if (is_lb) gs = tb * thv + J;
        else gs = tb;
        if (lane == SL_S) gs += rsum;
        if (!act) gs = 0.0;
        g.s = gs;
    }
    lp_out = lp;
    ok = ok && (lp == lp);
    if (!__all_sync(FULL, gfin)) {
        const double qnan = __longlong_as_double(0x7ff8000000000000LL);
#pragma unroll
        for (int t = 0; t < CPL; ++t) g.z[t] = qnan;
        g.s = qnan;
    }
    return __all_sync(FULL, ok);
}

}  // namespace bsyn
