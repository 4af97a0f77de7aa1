// bsyn_common.cuh -- shared definitions for the sm_100a kernels of the gp_grid / nointeraction path.
//
// Mapping: ONE WARP = ONE CHAIN.  A D-vector (position, momentum, gradient ...) is
// spread over the 32 lanes as
//     zc[t]  (t < CPL)  : GP cell c = lane + 32*t  (cell c <-> z[i,j], i = c % n2, j = c / n2, the
//                         reference's column-major z, src/stanExports_gp_grid.h:1190-1196)
//     sc               : canonical scalar slot `lane` (lanes 0..15, see enum Slot)
// so leapfrog updates, dot products and the per-cell model algebra are register-only.
#pragma once
#include <cstdint>
#include <type_traits>
#include <cuda_runtime.h>
#include <math.h>

namespace bsyn {

constexpr unsigned FULL = 0xffffffffu;
constexpr int WPC = 4;   // warps per CTA of the stand-alone lp+grad / write_array kernels

// canonical scalar slots (lane index).  Order = the reference's parameter order with la_k / alpha
// always given a slot (inactive slots carry 0 and never move).
enum Slot {
    SL_LA1 = 0, SL_LA2, SL_M1, SL_M2, SL_TH1, SL_TH2, SL_SL1, SL_SL2, SL_B1, SL_B2, SL_ELL, SL_SIGF,
    SL_ALPHA, SL_S, SL_S21, SL_S22, SL_COUNT
};

// Options uniform over a batch (the scalar Stan data fields) + derived constants
// (gp_grid.stan:37-38, 43-44).
struct Flags {
    int model, est_la, robust, pcprior, het, kernel, nu_matern, est_alpha;
    double lambda, tau, lamL, lam1, lam2;
    double log_lam12; // log(lam1) + log(lam2)
    double lptn_c0;   // -tau^2/2 + log(tau) + (lamL+1) log(log(tau))  (tail branch constant, without -log(2pi)/2)
    unsigned active_mask;   // bit s set <=> scalar slot s is a parameter of this model
};

// Per-problem descriptor in global memory; all offsets index the batch-wide pools.
struct ProbDesc {
    int n1, n2, N, G;        // G = n1*n2
    int ncf;                 // (n1+1)*(n2+1) cells of f
    int D, n_out;
    int off_d;               // doubles pool: x1[n1] x2[n2] vw1[n1] vw2[n2] dist[P1+P2] cnt[ncf] ybar[ncf] ss[ncf] ysort[N] y[N]
    int off_i;               // ints pool: pr[P1+P2] cstart[ncf+1] ii0[N] (0-based f index, original order)
    int n_d_smem;            // doubles copied to shared memory per CTA (without y[N]; without ysort if !robust)
    int n_i_smem;            // ints copied to shared memory (cstart only, and only if robust)
};

// ---- compile-time specialisation of the model options -------------------------------------
// The kernels are templates over a SPEC type.  SpecRT reads every option from the runtime Flags (the fallback that
// covers all 5 kernels x 4 likelihoods x both models in one instruction stream); SpecCT<...> fixes them at compile
// time, so the instruction stream of a launch carries exactly one kernel function, one likelihood, one model -- and,
// with NF > 0, a fixed NF x NF dose grid with exact loop bounds and buffer strides.  OptV<SP> is the view the model
// code queries: `O.het()` folds to a constant under SpecCT and reads F.het under SpecRT.
struct SpecRT {
    static constexpr bool fixed = false;
    static constexpr int NF = 0;
};
template <int MODEL, int KERNEL, int NU, int HET, int ROBUST, int ESTLA, int PCP, int NF_> struct SpecCT {
    static constexpr bool fixed = true;
    static constexpr int model = MODEL, kernel = KERNEL, nu_matern = NU, het = HET, robust = ROBUST, est_la = ESTLA,
                         pcprior = PCP, est_alpha = (KERNEL == 3) ? 1 : 0;
    static constexpr int NF = NF_;   // 0: grid sizes are runtime values
    static constexpr unsigned active_mask =
        (ESTLA ? ((1u << SL_LA1) | (1u << SL_LA2)) : 0u) | (1u << SL_M1) | (1u << SL_M2) | (1u << SL_TH1) | (1u << SL_TH2) |
        (1u << SL_SL1) | (1u << SL_SL2) |
        (MODEL == 0 ? ((1u << SL_B1) | (1u << SL_B2) | (1u << SL_ELL) | (1u << SL_SIGF) | (KERNEL == 3 ? (1u << SL_ALPHA) : 0u)) : 0u) |
        (1u << SL_S) | (1u << SL_S21) | (1u << SL_S22);
};
template <class SP> struct OptV {
    const Flags &F;
    __device__ __forceinline__ explicit OptV(const Flags &f) : F(f) {}
#define BSYN_OPT(name)                                                                                                     \
    __device__ __forceinline__ int name() const                                                                            \
    {                                                                                                                      \
        if constexpr (SP::fixed) return SP::name; else return F.name;                                                      \
    }
    BSYN_OPT(model) BSYN_OPT(kernel) BSYN_OPT(nu_matern) BSYN_OPT(het) BSYN_OPT(robust) BSYN_OPT(est_la) BSYN_OPT(pcprior)
    BSYN_OPT(est_alpha)
#undef BSYN_OPT
    __device__ __forceinline__ unsigned active_mask() const
    {
        if constexpr (SP::fixed) return SP::active_mask; else return F.active_mask;
    }
};

template <int CPL, class SP = SpecRT> struct Cfg {
    // max dose levels per axis: a multiple of 4 for the runtime-size kernels (4-row strips), the exact size rounded up
    // to even (LDS.128 alignment of a column) for a fixed grid
    static constexpr int NM = SP::NF ? ((SP::NF + 1) & ~1) : ((CPL <= 2) ? 8 : (CPL == 4) ? 12 : 16);
    static constexpr int NN = NM * NM;
    static constexpr int NS = (NM + 3) / 4;             // 4-row strips per column
    static constexpr int GM = 32 * CPL;                 // max GP cells
    static constexpr int PK = NM * (NM - 1);            // max off-diagonal pairs, both axes
    static constexpr int TR = NM * (NM + 1) / 2;        // packed lower-triangular entries of one matrix
    // per-warp scratch, in doubles (every offset even: LDS.128 / STS.128 need 16-byte alignment).
    // Grid-shaped buffers hold element (i, j) at i + j*NM (rows = drug-2 levels, padded to NM).
    static constexpr int OFF_TH = 0;                    // th[16] constrained scalars, ur[16] raw (unconstrained)
    static constexpr int OFF_UR = 16;
    static constexpr int OFF_PM = 32;                   // pm[32]: p01 then p02
    static constexpr int OFF_INVD = 64;                 // 2*16 inverse Cholesky diagonals
    static constexpr int OFF_ITH = 96;                  // 16 per-slot reciprocals (1/theta, 1/(1-la), 1/(1+s^2))
    static constexpr int OFF_L1R = 112;                 // L1 row-major  [j*NM + l]
    static constexpr int OFF_L2R = OFF_L1R + NN;        // L2 row-major  [i*NM + k]
    static constexpr int OFF_ZS = OFF_L2R + NN;         // Z           ; later S1 = L1^T L1bar
    static constexpr int OFF_TS = OFF_ZS + NN;          // T = L2 Z    ; later S2 = L2^T L2bar
    static constexpr int OFF_TT = OFF_TS + NN;          // T^T         ; later Y1 transpose
    static constexpr int OFF_GS = OFF_TT + NN;          // G -> Gbar (in place) -> Zbar
    static constexpr int OFF_WS = OFF_GS + NN;          // L1 col-major (Cholesky) ; p0bar ; W = Gbar L1 ; later Y2 transpose
    static constexpr int OFF_LB1 = OFF_WS + NN;         // L1bar row-major [j*NM + l] ; later X1
    // L2 col-major [k*NM + i] is dead after T = L2 Z, before L2bar is formed: one buffer serves both (the Cholesky
    // publishes whole columns, zeros above the diagonal, so X2 left over from the previous evaluation does no harm)
    static constexpr int OFF_L2C = OFF_LB1 + NN;
    static constexpr int OFF_LB2 = OFF_L2C;             // L2bar col-major [i + k*NM] ; later X2
    static constexpr int OFF_EX = OFF_LB2 + NN;         // exp table per pair
    static constexpr int WS_DOUBLES = OFF_EX + PK + (PK & 1);
};

// compile-time loop: nvcc's "#pragma unroll" gives up on the nested triangular loops of the in-register
// Cholesky (and then the row array falls into local memory), so those loops are unrolled by recursion.
template <int I, int N, typename Fn> __device__ __forceinline__ void static_for(Fn &&f)
{
    if constexpr (I < N) {
        f(std::integral_constant<int, I>{});
        static_for<I + 1, N>(f);
    }
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
// Sum 8 per-lane values over the warp with 9 (instead of 40) 64-bit shuffles: three halving exchanges over
// lane bits 0,1,2 leave lane L with the 8-lane partial sum of value (L & 7), two plain butterflies finish.
// Returns on lane L the warp total of v[L & 7].
__device__ __forceinline__ double warp_sum8(const double (&v)[8])
{
    const int lane = threadIdx.x & 31;
    const bool b0 = lane & 1, b1 = lane & 2, b2 = lane & 4;
    double w[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {   // keep values with index bit0 == lane bit0
        const double keep = b0 ? v[2 * k + 1] : v[2 * k], send = b0 ? v[2 * k] : v[2 * k + 1];
        w[k] = keep + __shfl_xor_sync(FULL, send, 1);
    }
    double x[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {   // w[k] has index bit1 = k & 1: keep those matching lane bit1
        const double keep = b1 ? w[2 * k + 1] : w[2 * k], send = b1 ? w[2 * k] : w[2 * k + 1];
        x[k] = keep + __shfl_xor_sync(FULL, send, 2);
    }
    double r;
    {
        const double keep = b2 ? x[1] : x[0], send = b2 ? x[0] : x[1];
        r = keep + __shfl_xor_sync(FULL, send, 4);
    }
    r += __shfl_xor_sync(FULL, r, 8);
    r += __shfl_xor_sync(FULL, r, 16);
    return r;
}
__device__ __forceinline__ double half_sum(double v)   // sum within each half-warp
{
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

// All shared memory is addressed through ONE dynamic array and 32-bit element offsets, so every access
// compiles to LDS/STS (a pointer smuggled through a struct decays to a generic LD with 64-bit address math).
extern __shared__ double smem[];
// A warp-uniform, loop-invariant value (a shared-memory offset) the compiler must KEEP: without this ptxas
// rematerialises the whole S2R / LDC / IMAD chain at every use (12 % of the sampler's issue slots in r01 profiles).
__device__ __forceinline__ int pin_uniform(const int v) { return __shfl_sync(0xffffffffu, v, 0); }
#define SMI(off) (reinterpret_cast<int *>(smem)[(off)])

// Problem view in shared memory (one per CTA, shared by its warps).  Members are offsets into smem
// (doubles for the double tables, ints for the int tables).
struct ProbS {
    int n1, n2, N, G, ncf, P1, P2, T1, T2;
    int xs;        // x1 then x2
    int vw;        // VUS trapezoid weights: w1[n1] (incl. 100/area) then w2[n2]
    int cnt, ybar, ss;
    int ysort;     // robust only
    int dist;      // [P1+P2] pair distances
    int cstart;    // (int offset) robust only
    int pr;        // (int offset) [P1+P2] (a<<8)|b, a > b
    int tri;       // (int offset) [TR] (r<<8)|c, c <= r, packed row-major lower triangle
};

// ---- Philox4x32-10 counter RNG (Salmon et al. 2011), stateless: every draw is a pure function of
// (seed, chain, lane-or-warp tag, iteration, draw index), so no RNG state lives in registers. ----
__device__ __forceinline__ void philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                           uint32_t k1, uint32_t (&out)[4])
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
__device__ __forceinline__ double u01_from_bits(uint32_t a, uint32_t b)   // (0,1), 53 bits
{
    uint64_t v = ((uint64_t)a << 21) ^ (uint64_t)(b >> 11);
    return ((double)v + 0.5) * (1.0 / 9007199254740992.0);
}
// RNG stream tags (c2 low byte = lane, or 0xFF for warp-uniform draws)
enum RngTag { RT_INIT = 1, RT_MOM = 2, RT_DIR = 3, RT_LEAF = 4, RT_TOP = 5, RT_JIT = 6, RT_GQ = 7, RT_ISTEP = 8 };

struct RngKey {
    uint32_t k0, k1, chain;
};
// two uniforms for (tag, lane_field, iter, idx)
__device__ __forceinline__ void rng_u2(const RngKey &k, uint32_t tag, uint32_t lane_field, uint32_t iter,
                                       uint32_t idx, double &ua, double &ub)
{
    uint32_t o[4];
    philox4x32(idx, iter, lane_field | (tag << 8), k.chain, k.k0, k.k1, o);
    ua = u01_from_bits(o[0], o[1]);
    ub = u01_from_bits(o[2], o[3]);
}
// warp-uniform U(0,1): identical on every lane, no shuffle needed
__device__ __forceinline__ double rng_warp_uniform(const RngKey &k, uint32_t tag, uint32_t iter, uint32_t idx)
{
    double a, b;
    rng_u2(k, tag, 0xFFu, iter, idx, a, b);
    return a;
}
// two independent N(0,1) per lane (Box-Muller)
__device__ __forceinline__ void rng_lane_normal2(const RngKey &k, uint32_t tag, uint32_t iter, uint32_t idx,
                                                 double &na, double &nb)
{
    double u1, u2, s, c;
    rng_u2(k, tag, (uint32_t)(threadIdx.x & 31), iter, idx, u1, u2);
    sincospi(2.0 * u2, &s, &c);
    double r = sqrt(-2.0 * log(u1));
    na = r * c;
    nb = r * s;
}

}  // namespace bsyn
