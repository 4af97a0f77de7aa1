// bsyn_nuts.cuh -- device-resident adaptive NUTS, one warp per chain.
//
// Replaces the control loop rstan enters through call_sampler (src/stanExports_gp_grid.cc:15):
// stan::services::sample::hmc_nuts_diag_e_adapt -> base_nuts::transition / build_tree,
// expl_leapfrog, diag_e_metric, stepsize_adaptation (dual averaging), windowed var_adaptation,
// base_hmc::init_stepsize, services::util::initialize  [StanHeaders, un-vendored; SURVEY.md App. D].
//
// Not a port of the recursive build_tree: the subtree of 2^d leapfrogs is built ITERATIVELY with
//   * uniform progressive (reservoir) sampling of the subtree proposal -- the same multinomial law as the
//     recursion's pairwise merges;
//   * per-level checkpoints (p at the first leaf of the block, p at the last leaf of its left half, rho of
//     its left half) so that every U-turn test the recursion makes, including Stan >= 2.21's two extra
//     cross-subtree tests, is made at the leaf that completes the block;
//   * position / momentum / gradient of the active trajectory end in registers (lane layout), everything
//     touched once per doubling or per block in a per-warp arena in HBM/L2.
#pragma once
#include "bsyn_model.cuh"

namespace bsyn {

// shared-memory layout (offsets in doubles), computed on the host from the batch maxima
struct SmemLayout {
    int tri_off, pd_off, pi_off, ws_off, pd_stride, pi_stride;   // *_stride: per-warp problem blocks (logp kernel)
};

struct SamplerParams {
    int n_problems, chains, groups;       // groups = ceil(chains / warps per CTA)
    int iter, warmup, thin, max_depth, save_warmup, n_save;
    double init_r, delta, stepsize0, jitter, gamma, kappa, t0;
    int init_buffer, term_buffer, window;
    unsigned long long seed;
    int jacobian;
    // sinks (device, may be null)
    double *draws; long long ld_draws;
    double *udraws; long long ld_u;
    double *qdraws;
    double *sp;
    double *chain_stats;
    double *inv_metric; int ld_im;
    double *cpo_acc; int ld_cpo;     // [n_problems][chains][ld_cpo]
    // work
    double *arena; long long arena_stride;
    unsigned long long *grad_counter;
    int *work_counter;
    SmemLayout sl;
};

enum ArenaSlot {
    A_OQ = 0, A_OP, A_OG, A_SQ, A_SG, A_PQ, A_PG, A_RHO_OLD, A_P_OLD_NEAR, A_P_NEW_NEAR, A_WM, A_WM2, A_LVL
};
__host__ __device__ inline int arena_vectors(int max_depth) { return A_LVL + 3 * (max_depth > 2 ? max_depth - 1 : 1); }

template <int CPL> __device__ __forceinline__ void vstore(double *base, const WVec<CPL> &v)
{
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int t = 0; t < CPL; ++t) base[t * 32 + lane] = v.z[t];
    base[CPL * 32 + lane] = v.s;
}
template <int CPL> __device__ __forceinline__ void vload(const double *base, WVec<CPL> &v)
{
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int t = 0; t < CPL; ++t) v.z[t] = base[t * 32 + lane];
    v.s = base[CPL * 32 + lane];
}
template <int CPL> __device__ __forceinline__ void vcopy(double *dst, const double *src)
{
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int t = 0; t <= CPL; ++t) dst[t * 32 + lane] = src[t * 32 + lane];
}
// sum_i minv_i a_i r_i
template <int CPL> __device__ __forceinline__ double vdot3(const WVec<CPL> &m, const WVec<CPL> &a, const WVec<CPL> &r)
{
    double s = m.s * a.s * r.s;
#pragma unroll
    for (int t = 0; t < CPL; ++t) s = fma(m.z[t] * a.z[t], r.z[t], s);
    return warp_sum(s);
}
// Stan's compute_criterion(p_sharp_a, p_sharp_b, rho) with p_sharp = minv (.) p
template <int CPL>
__device__ __forceinline__ bool uturn_ok(const WVec<CPL> &m, const WVec<CPL> &a, const WVec<CPL> &b, const WVec<CPL> &r)
{
    double sa = m.s * a.s * r.s, sb = m.s * b.s * r.s;
#pragma unroll
    for (int t = 0; t < CPL; ++t) {
        const double mr = m.z[t] * r.z[t];
        sa = fma(mr, a.z[t], sa);
        sb = fma(mr, b.z[t], sb);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sa += __shfl_xor_sync(FULL, sa, o);
        sb += __shfl_xor_sync(FULL, sb, o);
    }
    return sa > 0.0 && sb > 0.0;
}
__device__ __forceinline__ double log_add_exp(double a, double b)
{
    if (a == -INFINITY) return b;
    if (b == -INFINITY) return a;
    return a > b ? a + log1p(exp(b - a)) : b + log1p(exp(a - b));
}

template <int CPL> struct Chain {
    // state of the active trajectory end / current point
    WVec<CPL> q, p, g;   // g = gradient of lp (not of the potential)
    double V;            // potential = -lp
    WVec<CPL> minv;
    double eps;
};

template <int CPL> struct Ctx {
    const Flags *F;
    const ProbS *P;
    int ws;              // offset (doubles) of this warp's scratch in smem
    int cell[CPL];
    bool lane_active;    // scalar slot of this lane is a parameter
    double *arena;
    int jacobian;
    unsigned long long ngrad;
};

// potential energy and gradient at c.q; on a domain error V = +inf, g = 0 (Stan: exception => rejected)
template <int CPL> __device__ __forceinline__ void eval_pg(Ctx<CPL> &X, Chain<CPL> &c)
{
    double lp;
    const bool ok = warp_logp_grad<CPL, false>(*X.F, *X.P, X.ws, X.cell, c.q, X.jacobian, lp, c.g, nullptr);
    X.ngrad++;
    if (!ok || !(fabs(lp) < INFINITY)) {
        c.V = INFINITY;
#pragma unroll
        for (int t = 0; t < CPL; ++t) c.g.z[t] = 0.0;
        c.g.s = 0.0;
    } else {
        c.V = -lp;
    }
}
template <int CPL> __device__ __forceinline__ double kinetic(const Chain<CPL> &c)
{
    return 0.5 * vdot3<CPL>(c.minv, c.p, c.p);
}
template <int CPL> __device__ __forceinline__ void leapfrog(Ctx<CPL> &X, Chain<CPL> &c, double eps)
{
    const double he = 0.5 * eps;
#pragma unroll
    for (int t = 0; t < CPL; ++t) {
        c.p.z[t] = fma(he, c.g.z[t], c.p.z[t]);
        c.q.z[t] = fma(eps * c.minv.z[t], c.p.z[t], c.q.z[t]);
    }
    c.p.s = fma(he, c.g.s, c.p.s);
    c.q.s = fma(eps * c.minv.s, c.p.s, c.q.s);
    eval_pg<CPL>(X, c);
#pragma unroll
    for (int t = 0; t < CPL; ++t) c.p.z[t] = fma(he, c.g.z[t], c.p.z[t]);
    c.p.s = fma(he, c.g.s, c.p.s);
}
template <int CPL>
__device__ __forceinline__ void sample_momentum(const Ctx<CPL> &X, Chain<CPL> &c, const RngKey &rk, uint32_t tag,
                                                uint32_t iter, uint32_t idx0)
{
    // p_i = N(0,1) / sqrt(minv_i)  (diag_e_metric::sample_p)
    double n0, n1;
#pragma unroll
    for (int t = 0; t < CPL; t += 2) {
        rng_lane_normal2(rk, tag, iter, idx0 + t, n0, n1);
        c.p.z[t] = ((X.cell[t] >> 17) & 1) ? n0 * rsqrt(c.minv.z[t]) : 0.0;
        if (t + 1 < CPL) c.p.z[t + 1] = ((X.cell[t + 1] >> 17) & 1) ? n1 * rsqrt(c.minv.z[t + 1]) : 0.0;
    }
    rng_lane_normal2(rk, tag, iter, idx0 + CPL + 1, n0, n1);
    c.p.s = X.lane_active ? n0 * rsqrt(c.minv.s) : 0.0;
}

struct TransInfo {
    double accept, energy;
    int depth, nleap, divergent;
};

// One NUTS transition (base_nuts::transition).  On entry c.q/c.g/c.V = current state; on exit = new state.
template <int CPL>
__device__ __forceinline__ void nuts_transition(Ctx<CPL> &X, Chain<CPL> &c, const double eps, const int max_depth,
                                             const RngKey &rk, const uint32_t iter, TransInfo &info)
{
    constexpr int VS = (CPL + 1) * 32;
    double *A = X.arena;
    sample_momentum<CPL>(X, c, rk, RT_MOM, iter, 0);
    const double H0 = c.V + kinetic<CPL>(c);
    vstore<CPL>(A + A_OQ * VS, c.q); vstore<CPL>(A + A_OP * VS, c.p); vstore<CPL>(A + A_OG * VS, c.g);
    vstore<CPL>(A + A_SQ * VS, c.q); vstore<CPL>(A + A_SG * VS, c.g);
    vstore<CPL>(A + A_RHO_OLD * VS, c.p);
    double otherV = c.V, sampleV = c.V, sampleH = H0, proposeV = c.V, proposeH = H0;
    int curdir = 1;
    double lsw = 0.0, summetro = 0.0;
    int depth = 0, nleap = 0;
    bool divergent = false;
    uint32_t rctr = 0;
    while (depth < max_depth) {
        const int dir = rng_warp_uniform(rk, RT_DIR, iter, depth) > 0.5 ? 1 : -1;
        if (dir != curdir) {   // make the end we extend the register-resident one
            WVec<CPL> t;
            vload<CPL>(A + A_OQ * VS, t); vstore<CPL>(A + A_OQ * VS, c.q); c.q = t;
            vload<CPL>(A + A_OP * VS, t); vstore<CPL>(A + A_OP * VS, c.p); c.p = t;
            vload<CPL>(A + A_OG * VS, t); vstore<CPL>(A + A_OG * VS, c.g); c.g = t;
            const double tv = otherV; otherV = c.V; c.V = tv;
            curdir = dir;
        }
        vstore<CPL>(A + A_P_OLD_NEAR * VS, c.p);
        // ---- build the new subtree: 2^depth leapfrogs away from the old trajectory ----
        WVec<CPL> S, pprev;
#pragma unroll
        for (int t = 0; t < CPL; ++t) S.z[t] = 0.0;
        S.s = 0.0;
        double lsw_sub = -INFINITY;
        bool valid = true;
        const int nleaf = 1 << depth;
        const double seps = dir * eps;
        for (int k = 0; k < nleaf; ++k) {
            pprev = c.p;
            leapfrog<CPL>(X, c, seps);
            ++nleap;
            double h = c.V + kinetic<CPL>(c);
            if (!(h == h)) h = INFINITY;
            const double w = H0 - h;
            if (h - H0 > 1000.0) divergent = true;
            lsw_sub = log_add_exp(lsw_sub, w);
            summetro += (w > 0.0) ? 1.0 : exp(w);
            if (divergent) { valid = false; break; }
            // uniform progressive sampling of the subtree's proposal
            bool take = (k == 0);
            if (!take) take = rng_warp_uniform(rk, RT_LEAF, iter, rctr++) < exp(w - lsw_sub);
            if (take) {
                vstore<CPL>(A + A_PQ * VS, c.q); vstore<CPL>(A + A_PG * VS, c.g);
                proposeV = c.V; proposeH = h;
            }
#pragma unroll
            for (int t = 0; t < CPL; ++t) S.z[t] += c.p.z[t];
            S.s += c.p.s;
            if (k == 0) vstore<CPL>(A + A_P_NEW_NEAR * VS, c.p);
            // leaf k opens blocks at every level l >= 2 with k % 2^l == 0
            if ((k & 3) == 0) {
                for (int l = 2; l <= depth && (k & ((1 << l) - 1)) == 0; ++l)
                    vstore<CPL>(A + (A_LVL + 3 * (l - 2)) * VS, c.p);
            }
            if (k & 1) {
                // level 1: left = leaf k-1, right = leaf k (the three tests coincide)
                WVec<CPL> car;
#pragma unroll
                for (int t = 0; t < CPL; ++t) car.z[t] = pprev.z[t] + c.p.z[t];
                car.s = pprev.s + c.p.s;
                if (!uturn_ok<CPL>(c.minv, pprev, c.p, car)) { valid = false; break; }
                int l = 2;
                const int kk = k + 1;
                while (l <= depth && (kk & ((1 << l) - 1)) == 0) {
                    double *Lv = A + (A_LVL + 3 * (l - 2)) * VS;
                    WVec<CPL> pbegL, pendL, rhoL, ext;
                    vload<CPL>(Lv, pbegL); vload<CPL>(Lv + 2 * VS, rhoL);
                    // rho_L + rho_R around the merged block
#pragma unroll
                    for (int t = 0; t < CPL; ++t) ext.z[t] = rhoL.z[t] + car.z[t];
                    ext.s = rhoL.s + car.s;
                    bool okm = uturn_ok<CPL>(c.minv, pbegL, c.p, ext);
                    // rho_L + p(first leaf of R)
                    WVec<CPL> pbegR;
                    if (l == 2) pbegR = pprev; else vload<CPL>(A + (A_LVL + 3 * (l - 3)) * VS, pbegR);
                    {
                        WVec<CPL> e2;
#pragma unroll
                        for (int t = 0; t < CPL; ++t) e2.z[t] = rhoL.z[t] + pbegR.z[t];
                        e2.s = rhoL.s + pbegR.s;
                        okm = okm && uturn_ok<CPL>(c.minv, pbegL, pbegR, e2);
                    }
                    // rho_R + p(last leaf of L)
                    vload<CPL>(Lv + VS, pendL);
                    {
                        WVec<CPL> e3;
#pragma unroll
                        for (int t = 0; t < CPL; ++t) e3.z[t] = car.z[t] + pendL.z[t];
                        e3.s = car.s + pendL.s;
                        okm = okm && uturn_ok<CPL>(c.minv, pendL, c.p, e3);
                    }
                    if (!okm) { valid = false; break; }
                    car = ext;
                    ++l;
                }
                if (!valid) break;
                if (l <= depth) {   // the completed block is the left half of a level-l block
                    double *Lv = A + (A_LVL + 3 * (l - 2)) * VS;
                    vstore<CPL>(Lv + VS, c.p);
                    vstore<CPL>(Lv + 2 * VS, car);
                }
            }
        }
        if (!valid) break;
        ++depth;
        // ---- multinomial sample across subtrees (biased progressive sampling at the top level) ----
        {
            bool take = lsw_sub > lsw;
            if (!take) take = rng_warp_uniform(rk, RT_TOP, iter, depth) < exp(lsw_sub - lsw);
            if (take) {
                vcopy<CPL>(A + A_SQ * VS, A + A_PQ * VS); vcopy<CPL>(A + A_SG * VS, A + A_PG * VS);
                sampleV = proposeV; sampleH = proposeH;
            }
        }
        lsw = log_add_exp(lsw, lsw_sub);
        // ---- U-turn across the whole trajectory + the two cross-subtree tests ----
        WVec<CPL> rho_old, pfar, t1;
        vload<CPL>(A + A_RHO_OLD * VS, rho_old);
        vload<CPL>(A + A_OP * VS, pfar);
        WVec<CPL> rho;
#pragma unroll
        for (int t = 0; t < CPL; ++t) rho.z[t] = rho_old.z[t] + S.z[t];
        rho.s = rho_old.s + S.s;
        bool persist = uturn_ok<CPL>(c.minv, pfar, c.p, rho);
        vstore<CPL>(A + A_RHO_OLD * VS, rho);
        {
            WVec<CPL> pnn, e2;
            if (nleaf == 1) pnn = c.p; else vload<CPL>(A + A_P_NEW_NEAR * VS, pnn);
#pragma unroll
            for (int t = 0; t < CPL; ++t) e2.z[t] = rho_old.z[t] + pnn.z[t];
            e2.s = rho_old.s + pnn.s;
            persist = persist && uturn_ok<CPL>(c.minv, pfar, pnn, e2);
        }
        {
            vload<CPL>(A + A_P_OLD_NEAR * VS, t1);
            WVec<CPL> e3;
#pragma unroll
            for (int t = 0; t < CPL; ++t) e3.z[t] = S.z[t] + t1.z[t];
            e3.s = S.s + t1.s;
            persist = persist && uturn_ok<CPL>(c.minv, t1, c.p, e3);
        }
        if (!persist) break;
    }
    info.accept = summetro / (double)(nleap > 0 ? nleap : 1);
    info.depth = depth;
    info.nleap = nleap;
    info.divergent = divergent ? 1 : 0;
    info.energy = sampleH;
    vload<CPL>(A + A_SQ * VS, c.q);
    vload<CPL>(A + A_SG * VS, c.g);
    c.V = sampleV;
}

// base_hmc::init_stepsize: double / halve eps until the one-step acceptance crosses 0.8
template <int CPL>
__device__ __forceinline__ double init_stepsize(Ctx<CPL> &X, Chain<CPL> &c, double eps, const RngKey &rk, uint32_t iter)
{
    constexpr int VS = (CPL + 1) * 32;
    double *A = X.arena;
    vstore<CPL>(A + A_SQ * VS, c.q); vstore<CPL>(A + A_SG * VS, c.g);
    const double V0 = c.V;
    int dir = 0;
    for (uint32_t it = 0; it < 64; ++it) {
        vload<CPL>(A + A_SQ * VS, c.q); vload<CPL>(A + A_SG * VS, c.g); c.V = V0;
        sample_momentum<CPL>(X, c, rk, RT_ISTEP, iter, it * (CPL + 4));
        const double H0 = c.V + kinetic<CPL>(c);
        leapfrog<CPL>(X, c, eps);
        double h = c.V + kinetic<CPL>(c);
        if (!(h == h)) h = INFINITY;
        const double dH = H0 - h;
        const double l08 = -0.22314355131420976;   // log(0.8)
        if (it == 0) { dir = dH > l08 ? 1 : -1; continue; }   // first probe only fixes the direction
        if (dir == 1 && !(dH > l08)) break;
        if (dir == -1 && !(dH < l08)) break;
        eps = dir == 1 ? 2.0 * eps : 0.5 * eps;
        if (eps > 1e7 || eps == 0.0) break;
    }
    vload<CPL>(A + A_SQ * VS, c.q); vload<CPL>(A + A_SG * VS, c.g); c.V = V0;
    return eps;
}

// ---------------------------------------------------------------------------------------------------
// problem block -> ProbS view
// s_pd: offset (doubles) of the problem's double block, s_pi / s_tri: offsets (ints) of its int block / tri table
__device__ __forceinline__ void bind_problem(ProbS &P, const ProbDesc &pd, const int s_pd, const int s_pi,
                                             const int s_tri)
{
    P.n1 = pd.n1; P.n2 = pd.n2; P.N = pd.N; P.G = pd.G; P.ncf = pd.ncf;
    P.P1 = pd.n1 * (pd.n1 - 1) / 2; P.P2 = pd.n2 * (pd.n2 - 1) / 2;
    P.T1 = pd.n1 * (pd.n1 + 1) / 2; P.T2 = pd.n2 * (pd.n2 + 1) / 2;
    const int nm = pd.n1 + pd.n2, PP = P.P1 + P.P2;
    P.xs = s_pd; P.vw = s_pd + nm; P.dist = s_pd + 2 * nm;
    P.cnt = s_pd + 2 * nm + PP; P.ybar = P.cnt + pd.ncf; P.ss = P.ybar + pd.ncf; P.ysort = P.ss + pd.ncf;
    P.pr = s_pi; P.cstart = s_pi + PP; P.tri = s_tri;
}
__host__ __device__ inline int pool_y_off(const ProbDesc &pd)   // offset of y[N] (original order) in the doubles block
{
    const int nm = pd.n1 + pd.n2, PP = pd.n1 * (pd.n1 - 1) / 2 + pd.n2 * (pd.n2 - 1) / 2;
    return 2 * nm + PP + 3 * pd.ncf + pd.N;
}
__host__ __device__ inline int pool_ii_off(const ProbDesc &pd)  // offset of ii0[N] in the ints block
{
    const int PP = pd.n1 * (pd.n1 - 1) / 2 + pd.n2 * (pd.n2 - 1) / 2;
    return PP + pd.ncf + 1;
}

// position of scalar slot `slot` in the reference's unconstrained vector (-1 = not a parameter)
__device__ __forceinline__ int slot_pos(const Flags &F, int D, int slot)
{
    if (slot >= SL_COUNT || !((F.active_mask >> slot) & 1u)) return -1;
    if (slot < SL_M1) return slot;
    if (slot <= SL_ALPHA) return slot - (F.est_la ? 0 : 2);
    return D - 3 + (slot - SL_S);
}
template <int CPL>
__device__ __forceinline__ void load_user_vec(const Flags &F, const ProbDesc &pd, const int (&cell)[CPL],
                                              const double *u, WVec<CPL> &q)
{
    const int lane = threadIdx.x & 31;
    const int npre = 2 * F.est_la + 6 + (F.model == 0 ? 4 + F.est_alpha : 0);
#pragma unroll
    for (int t = 0; t < CPL; ++t) q.z[t] = ((cell[t] >> 17) & 1) ? u[npre + lane + 32 * t] : 0.0;
    const int pos = slot_pos(F, pd.D, lane);
    q.s = pos >= 0 ? u[pos] : 0.0;
}
template <int CPL>
__device__ __forceinline__ void store_user_vec(const Flags &F, const ProbDesc &pd, const int (&cell)[CPL],
                                               const WVec<CPL> &q, double *u)
{
    const int lane = threadIdx.x & 31;
    const int npre = 2 * F.est_la + 6 + (F.model == 0 ? 4 + F.est_alpha : 0);
#pragma unroll
    for (int t = 0; t < CPL; ++t)
        if ((cell[t] >> 17) & 1) u[npre + lane + 32 * t] = q.z[t];
    const int pos = slot_pos(F, pd.D, lane);
    if (pos >= 0) u[pos] = q.s;
}

// ---------------------------------------------------------------------------------------------------
// The sampler kernel: persistent CTAs of WPC warps; a work item = (problem, group of WPC chains).
// MINB = minimum resident CTAs per SM the register allocation must allow (occupancy / spill trade-off)
template <int CPL, int WPC, int MINB>
__global__ void __launch_bounds__(32 * WPC, MINB) nuts_kernel(const Flags F, const ProbDesc *__restrict__ descs,
                                                        const double *__restrict__ dpool, const int *__restrict__ ipool,
                                                        const int *__restrict__ tri_tab, const SamplerParams sp)
{
    using C = Cfg<CPL>;
    constexpr int VS = (CPL + 1) * 32;
    __shared__ int s_item;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_items = sp.n_problems * sp.groups;
    const int s_tri = 2 * sp.sl.tri_off, s_pd = sp.sl.pd_off, s_pi = 2 * sp.sl.pi_off;
    for (int e = threadIdx.x; e < C::TR; e += blockDim.x) SMI(s_tri + e) = tri_tab[e];
    for (int e = lane; e < C::WS_DOUBLES; e += 32) smem[sp.sl.ws_off + warp * C::WS_DOUBLES + e] = 0.0;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_item = atomicAdd(sp.work_counter, 1);
        __syncthreads();
        const int item = s_item;
        if (item >= n_items) break;
        const int ip = item / sp.groups, grp = item - ip * sp.groups;
        const ProbDesc pd = descs[ip];
        for (int e = threadIdx.x; e < pd.n_d_smem; e += blockDim.x) smem[s_pd + e] = dpool[pd.off_d + e];
        for (int e = threadIdx.x; e < pd.n_i_smem; e += blockDim.x) SMI(s_pi + e) = ipool[pd.off_i + e];
        __syncthreads();
        const int chain = grp * WPC + warp;
        if (chain >= sp.chains) continue;
        ProbS P;
        bind_problem(P, pd, s_pd, s_pi, s_tri);
        Ctx<CPL> X;
        X.F = &F; X.P = &P;
        X.ws = sp.sl.ws_off + warp * C::WS_DOUBLES;
        make_cells<CPL>(F, P, X.cell);
        X.lane_active = (F.active_mask >> lane) & 1u;
        X.arena = sp.arena + (long long)(blockIdx.x * WPC + warp) * sp.arena_stride;
        X.jacobian = sp.jacobian;
        X.ngrad = 0;
        const long long cg = (long long)ip * sp.chains + chain;
        RngKey rk;
        rk.k0 = (uint32_t)sp.seed; rk.k1 = (uint32_t)(sp.seed >> 32); rk.chain = (uint32_t)cg;
        double *A = X.arena;
        Chain<CPL> c;
#pragma unroll
        for (int t = 0; t < CPL; ++t) { c.minv.z[t] = ((X.cell[t] >> 17) & 1) ? 1.0 : 0.0; c.p.z[t] = 0.0; }
        c.minv.s = X.lane_active ? 1.0 : 0.0;
        c.p.s = 0.0;
        double *cs = sp.chain_stats ? sp.chain_stats + cg * 8 : nullptr;
        // ---- services::util::initialize: u ~ U(-R, R) until lp and its gradient are finite ----
        bool init_ok = false;
        for (uint32_t tr = 0; tr < 100 && !init_ok; ++tr) {
            double ua, ub;
#pragma unroll
            for (int t = 0; t < CPL; t += 2) {
                rng_u2(rk, RT_INIT, lane, tr, t, ua, ub);
                c.q.z[t] = ((X.cell[t] >> 17) & 1) ? sp.init_r * (2.0 * ua - 1.0) : 0.0;
                if (t + 1 < CPL) c.q.z[t + 1] = ((X.cell[t + 1] >> 17) & 1) ? sp.init_r * (2.0 * ub - 1.0) : 0.0;
            }
            rng_u2(rk, RT_INIT, lane, tr, CPL + 1, ua, ub);
            c.q.s = X.lane_active ? sp.init_r * (2.0 * ua - 1.0) : 0.0;
            eval_pg<CPL>(X, c);
            bool fin = c.V < INFINITY && fabs(c.g.s) < INFINITY;
#pragma unroll
            for (int t = 0; t < CPL; ++t) fin = fin && fabs(c.g.z[t]) < INFINITY;
            init_ok = __all_sync(FULL, fin);
        }
        if (!init_ok) {
            if (cs && lane == 0) { for (int k = 0; k < 8; ++k) cs[k] = 0.0; cs[6] = 1.0; }
            if (lane == 0) atomicAdd(sp.grad_counter, X.ngrad);
            continue;
        }
        double eps = sp.stepsize0;
        bool need_istep = true;          // base_hmc::init_stepsize runs before the first transition ...
        uint32_t istep_tag = 0xFFFFFF00u;
        // ---- adaptation state (stepsize_adaptation, windowed_adaptation) ----
        double mu = 0.0, sbar = 0.0, xbar = 0.0, cnt = 0.0;
        int nw = sp.warmup, init_buffer = sp.init_buffer, term_buffer = sp.term_buffer, base_window = sp.window;
        if (nw < 20) { init_buffer = term_buffer = base_window = 0; }
        else if (init_buffer + base_window + term_buffer > nw) {
            init_buffer = (int)(0.15 * nw); term_buffer = (int)(0.1 * nw);
            base_window = nw - (init_buffer + term_buffer);
        }
        int win_size = base_window, win_end = init_buffer + base_window - 1;
        double wn = 0.0;
        {
            WVec<CPL> zero;
#pragma unroll
            for (int t = 0; t < CPL; ++t) zero.z[t] = 0.0;
            zero.s = 0.0;
            vstore<CPL>(A + A_WM * VS, zero); vstore<CPL>(A + A_WM2 * VS, zero);
        }
        double n_div = 0.0, n_maxd = 0.0, nleap_tot = 0.0, nleap_samp = 0.0, acc_sum = 0.0, n_samp_it = 0.0;
        int isave = 0;
        const double *yglob = dpool + pd.off_d + pool_y_off(pd);
        const int *iiglob = ipool + pd.off_i + pool_ii_off(pd);
        double *cpo_acc = sp.cpo_acc ? sp.cpo_acc + cg * sp.ld_cpo : nullptr;
        if (cpo_acc) for (int n = lane; n < pd.N; n += 32) cpo_acc[n] = 0.0;
        for (int it = 0; it < sp.iter; ++it) {
            const bool warm = it < nw;
            if (need_istep) {
                eps = init_stepsize<CPL>(X, c, eps, rk, istep_tag);
                mu = log(10.0 * eps); sbar = 0.0; xbar = 0.0; cnt = 0.0;
                need_istep = false;
            }
            double e = eps;
            if (sp.jitter > 0.0) e = eps * (1.0 + sp.jitter * (2.0 * rng_warp_uniform(rk, RT_JIT, it, 0) - 1.0));
            TransInfo info;
            nuts_transition<CPL>(X, c, e, sp.max_depth, rk, (uint32_t)it, info);
            nleap_tot += info.nleap;
            if (warm) {
                // stepsize_adaptation::learn_stepsize (dual averaging)
                cnt += 1.0;
                const double astat = info.accept > 1.0 ? 1.0 : info.accept;
                const double eta = 1.0 / (cnt + sp.t0);
                sbar = (1.0 - eta) * sbar + eta * (sp.delta - astat);
                const double x = mu - sbar * sqrt(cnt) / sp.gamma;
                const double xeta = pow(cnt, -sp.kappa);
                xbar = (1.0 - xeta) * xbar + xeta * x;
                eps = exp(x);
                // windowed variance adaptation (var_adaptation::learn_variance, Welford)
                const bool in_window = (it >= init_buffer) && (it < nw - term_buffer);
                if (in_window) {
                    wn += 1.0;
                    WVec<CPL> m, m2;
                    vload<CPL>(A + A_WM * VS, m); vload<CPL>(A + A_WM2 * VS, m2);
                    const double iw = 1.0 / wn;
#pragma unroll
                    for (int t = 0; t < CPL; ++t) {
                        const double d = c.q.z[t] - m.z[t];
                        m.z[t] = fma(d, iw, m.z[t]);
                        m2.z[t] = fma(c.q.z[t] - m.z[t], d, m2.z[t]);
                    }
                    {
                        const double d = c.q.s - m.s;
                        m.s = fma(d, iw, m.s);
                        m2.s = fma(c.q.s - m.s, d, m2.s);
                    }
                    if (it == win_end) {
                        const double a = wn / (wn + 5.0), b = 1e-3 * (5.0 / (wn + 5.0)), inm1 = 1.0 / (wn - 1.0);
#pragma unroll
                        for (int t = 0; t < CPL; ++t) {
                            if ((X.cell[t] >> 17) & 1) c.minv.z[t] = fma(a, m2.z[t] * inm1, b);
                            m.z[t] = 0.0; m2.z[t] = 0.0;
                        }
                        if (X.lane_active) c.minv.s = fma(a, m2.s * inm1, b);
                        m.s = 0.0; m2.s = 0.0;
                        wn = 0.0;
                        // compute_next_window
                        if (win_end != nw - term_buffer - 1) {
                            win_size *= 2;
                            win_end = it + win_size;
                            if (win_end != nw - term_buffer - 1) {
                                const int next_boundary = win_end + 2 * win_size;
                                if (next_boundary >= nw - term_buffer) win_end = nw - term_buffer - 1;
                            }
                        }
                        vstore<CPL>(A + A_WM * VS, m); vstore<CPL>(A + A_WM2 * VS, m2);
                        need_istep = true;   // ... and after every metric update (restart of the dual averaging)
                        istep_tag = 0xFFFF0000u + (uint32_t)it;
                    } else {
                        vstore<CPL>(A + A_WM * VS, m); vstore<CPL>(A + A_WM2 * VS, m2);
                    }
                }
                if (it == nw - 1) eps = exp(xbar);   // complete_adaptation
            } else {
                n_div += info.divergent; n_maxd += (info.depth >= sp.max_depth); nleap_samp += info.nleap;
                acc_sum += info.accept; n_samp_it += 1.0;
            }
            const int base = warm ? 0 : nw;
            if ((!warm || sp.save_warmup) && ((it - base) % sp.thin == 0) && isave < sp.n_save) {
                const long long rowi = cg * sp.n_save + isave;
                double *row = sp.draws ? sp.draws + rowi * sp.ld_draws : nullptr;
                double *qout = sp.qdraws ? sp.qdraws + rowi * 12 : nullptr;
                if (row || qout || (cpo_acc && !warm)) {
                    double uz[4];
                    rng_u2(rk, RT_GQ, 0xFFu, it, 0, uz[0], uz[1]);
                    rng_u2(rk, RT_GQ, 0xFFu, it, 1, uz[2], uz[3]);
                    warp_write_array<CPL>(F, P, pd, X.ws, X.cell, c.q, yglob, iiglob, uz, row, qout,
                                          warm ? nullptr : cpo_acc);
                    if (qout && lane == 0) { qout[10] = -c.V; qout[11] = 0.0; }
                }
                if (sp.udraws) store_user_vec<CPL>(F, pd, X.cell, c.q, sp.udraws + rowi * sp.ld_u);
                if (sp.sp && lane == 0) {
                    double *o = sp.sp + rowi * 7;
                    o[0] = info.accept; o[1] = e; o[2] = info.depth; o[3] = info.nleap; o[4] = info.divergent;
                    o[5] = info.energy; o[6] = -c.V;
                }
                ++isave;
            }
        }
        if (cs && lane == 0) {
            cs[0] = eps; cs[1] = n_div; cs[2] = n_maxd; cs[3] = nleap_tot; cs[4] = nleap_samp;
            cs[5] = n_samp_it > 0 ? acc_sum / n_samp_it : 0.0; cs[6] = 0.0; cs[7] = (double)isave;
        }
        if (sp.inv_metric) store_user_vec<CPL>(F, pd, X.cell, c.minv, sp.inv_metric + cg * sp.ld_im);
        if (lane == 0) atomicAdd(sp.grad_counter, X.ngrad);
    }
}

}  // namespace bsyn
