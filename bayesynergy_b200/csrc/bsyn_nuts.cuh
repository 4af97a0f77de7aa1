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
    int park_off;                                                // sampler: per-warp chain-state block (PARK_DOUBLES each)
    int vpark_off;                                               // sampler: per-warp vector parking (3 D-vectors), or -1
};
// per-warp chain-state block: the warp-uniform chain state (struct ChainSt, <= PARK_GQ doubles) lives here for the
// whole life of a chain, followed by the generated-quantities sink block (GQ_DOUBLES) the evaluation reads when gq.on
constexpr int PARK_GQ = 52, PARK_DOUBLES = 64;

struct SamplerParams {
    int n_problems, chains;
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
    double *init_u;                  // [n_problems][chains][ld_im]: the initial point of every chain (unconstrained)
    double *cpo_acc; int ld_cpo;     // [n_problems][chains][ld_cpo]
    double *f_acc; int ld_f;         // [n_problems][chains][ld_f]: sum over the saved sampling draws of f per grid cell
    // dense_e metric (metric == 1): per resident warp  M^-1 (float, column-major DI x DI),  its Cholesky factor
    // (double, column-major lower) and the Welford cross-product accumulator (double)
    int metric;
    float *dense_minv; double *dense_L, *dense_m2; long long dense_stride;
    // extension launches (bsyn_sample_to_ess): the chains of the listed problems resume from their adapted state and
    // append their draws behind the ones already saved
    const int *prob_list; int n_list;      // null / 0: every problem of the batch
    int isave0, ld_save;                   // first draw slot written by this launch, draw slots per chain in the sinks
    double *resume; long long resume_stride;   // [n_problems * chains][q (VS), M^-1 diagonal (VS), step size]
    int resume_in, resume_out;
    int tick_period;   // the CTA barrier of the tick state machine is taken every tick_period-th tick (>= 1)
    // work
    double *arena; long long arena_stride;
    unsigned long long *grad_counter;
    int *work_counter;
    SmemLayout sl;
};

enum ArenaSlot {
    A_OQ = 0, A_OP, A_OG, A_SQ, A_SG, A_PQ, A_PG, A_RHO_OLD, A_P_OLD_NEAR, A_P_NEW_NEAR, A_WM, A_WM2, A_PREV,
    A_PARK_S, A_PARK_PREV, A_PARK_MINV, A_PARK_P, A_PARK_YV,   // chain vectors parked in the arena while the evaluation runs
    A_LVL
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
// compute_criterion with the sharp momenta given explicitly (p_sharp = M^-1 p): a_sharp . rho > 0 && b_sharp . rho > 0
template <int CPL>
__device__ __forceinline__ bool uturn_sharp_ok(const WVec<CPL> &as, const WVec<CPL> &bs, const WVec<CPL> &r)
{
    double sa = as.s * r.s, sb = bs.s * r.s;
#pragma unroll
    for (int t = 0; t < CPL; ++t) {
        sa = fma(as.z[t], r.z[t], sa);
        sb = fma(bs.z[t], r.z[t], sb);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sa += __shfl_xor_sync(FULL, sa, o);
        sb += __shfl_xor_sync(FULL, sb, o);
    }
    return sa > 0.0 && sb > 0.0;
}
// Three compute_criterion tests at once: (a1, b1 | r1), (a2, b2 | r2), (a3, b3 | r3) -> 6 dot products reduced with ONE
// 8-value butterfly (9 shuffles, one latency chain) instead of three pairs of reductions.
template <int CPL>
__device__ __forceinline__ bool uturn3_ok(const WVec<CPL> &a1, const WVec<CPL> &b1, const WVec<CPL> &r1,
                                          const WVec<CPL> &a2, const WVec<CPL> &b2, const WVec<CPL> &r2,
                                          const WVec<CPL> &a3, const WVec<CPL> &b3, const WVec<CPL> &r3)
{
    double v[8] = {a1.s * r1.s, b1.s * r1.s, a2.s * r2.s, b2.s * r2.s, a3.s * r3.s, b3.s * r3.s, 0.0, 0.0};
#pragma unroll
    for (int t = 0; t < CPL; ++t) {
        v[0] = fma(a1.z[t], r1.z[t], v[0]); v[1] = fma(b1.z[t], r1.z[t], v[1]);
        v[2] = fma(a2.z[t], r2.z[t], v[2]); v[3] = fma(b2.z[t], r2.z[t], v[3]);
        v[4] = fma(a3.z[t], r3.z[t], v[4]); v[5] = fma(b3.z[t], r3.z[t], v[5]);
    }
    const double tot = warp_sum8(v);                       // lane L holds the total of v[L & 7]
    const bool mine = ((threadIdx.x & 7) >= 6) || (tot > 0.0);
    return __all_sync(FULL, mine);
}
template <int CPL> __device__ __forceinline__ double vdot2(const WVec<CPL> &a, const WVec<CPL> &b)
{
    double s = a.s * b.s;
#pragma unroll
    for (int t = 0; t < CPL; ++t) s = fma(a.z[t], b.z[t], s);
    return warp_sum(s);
}
__device__ __forceinline__ double log_add_exp(double a, double b)
{
    if (a == -INFINITY) return b;
    if (b == -INFINITY) return a;
    const double hi = fmax(a, b), d = -fabs(a - b);
    return hi + fast_log(1.0 + fast_exp(d));   // fast_exp clamps d at -708: exp(d) = 3e-308 ~ 0
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
// The sampler kernel.
//
// A CTA is NW warps, each warp runs one chain at a time (work items = (problem, chain), pulled from an atomic
// counter).  The kernel is a TICK state machine: every pass of the main loop contains exactly ONE leapfrog (one
// lp+gradient evaluation) per warp, followed by that chain's bookkeeping (tree logic, adaptation, fetching the next
// chain ...), and every tick_period-th pass (default 2) starts with a CTA barrier.  Two reasons:
//   * one copy of the model code (5-10 k instructions) in the kernel instead of one per call site;
//   * the warps of a CTA run the evaluation in step, so they share instruction-cache lines.  With free-running
//     warps the sampler spent a third of its issue slots waiting for instructions (profiles/r01_nuts_v3*); with the
//     compile-time specialised model code (35 % smaller) free-running warps are still 11 % slower than the tick, two
//     independently synchronised halves of the CTA 9 % slower, and barriers INSIDE the evaluation 5 % slower
//     (profiles/r02_ab_nuts_v1 / v6 / v7): one barrier per two ticks is the measured optimum.
// The phases map onto Stan's control flow: PH_INIT = services::util::initialize, PH_ISTEP =
// base_hmc::init_stepsize, PH_TREE = base_nuts::transition / build_tree (iteratively, see the header comment),
// end-of-transition = stepsize_adaptation + windowed var_adaptation, PH_GQ = write_array of the saved draw.
struct TransInfo {
    double accept, energy;
    int depth, nleap, divergent;
};
// Warp-uniform state of one chain, resident in shared memory (every lane stores the same value to the same address).
// Only what the evaluation itself touches lives in registers; the bookkeeping after an evaluation loads what it needs
// where it needs it.  (Round 1 kept all of this in registers and parked it with 25 STS.128 + 25 LDS.128 around every
// evaluation; the bulk reload then spilled to local memory: 74 LDL + 45 STL per evaluation, profiles/r02_nuts_v1*.)
struct ChainSt {
    double eps, mu, sbar, xbar, cnt, wn, n_div, n_maxd, nleap_tot, nleap_samp, acc_sum, n_samp_it;      // chain level
    double is_H0, is_V0, eps_try;                                                                       // init_stepsize
    double e_used, H0, otherV, sampleV, sampleH, proposeV, proposeH, lsw, summetro, lsw_sub, u_leaf;    // transition
    TransInfo info;
    long long cg;
    unsigned long long ngrad;
    double *cs, *cpo_acc;
    const double *yglob;
    const int *iiglob;
    int it, isave, win_size, win_end, init_try, is_it, is_dir, curdir, depth, nleap, k, nleaf;
    int need_istep, divergent;
    uint32_t istep_tag;
    int pad_;
    ProbDesc pd;
};
static_assert(sizeof(ChainSt) <= PARK_GQ * sizeof(double), "ChainSt must fit in front of the GQ sink block");
enum Phase { PH_FETCH = 0, PH_INIT, PH_ISTEP, PH_TREE, PH_GQ, PH_DONE };
enum Next { N_NONE = 0, N_ENDTRANS, N_ITER, N_TRANS, N_DOUBLE };

// Template parameters: CPL cells per lane; NW warps per CTA; MINB register cap per thread; DENSE dense_e metric;
// SP the compile-time model specialisation (SpecRT = every option read from F); TICK = the barrier-aligned tick
// described above (false: warps free-run -- only sensible when the specialised instruction stream is small).
template <int CPL, int NW, int MINB, bool DENSE, class SP = SpecRT, bool TICK = true>
__global__ void __launch_bounds__(32 * NW) __maxnreg__(MINB) nuts_kernel(const Flags F, const ProbDesc *__restrict__ descs,
                                                         const double *__restrict__ dpool, const int *__restrict__ ipool,
                                                         const int *__restrict__ tri_tab, const SamplerParams sp)
{
    using C = Cfg<CPL, SP>;
    const OptV<SP> O(F);
    constexpr int VS = (CPL + 1) * 32;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_items = (sp.prob_list ? sp.n_list : sp.n_problems) * sp.chains;
    const int s_tri = 2 * sp.sl.tri_off;
    const int s_pd = pin_uniform(sp.sl.pd_off + warp * sp.sl.pd_stride);
    const int s_pi = pin_uniform(2 * (sp.sl.pi_off + warp * sp.sl.pi_stride));
    const int ws = pin_uniform(sp.sl.ws_off + warp * C::WS_DOUBLES);
    const int s_park = pin_uniform(sp.sl.park_off + warp * PARK_DOUBLES);
    for (int e = threadIdx.x; e < C::TR; e += blockDim.x) SMI(s_tri + e) = tri_tab[e];
    for (int e = lane; e < C::WS_DOUBLES; e += 32) smem[ws + e] = 0.0;   // the factor's upper triangle must stay zero
    double *const A = sp.arena + (long long)(blockIdx.x * NW + warp) * sp.arena_stride;
    // dense_e: sharp momenta (M^-1 p) cannot be recomputed cheaply, so every stored p has a mirror slot for it
    constexpr int DI = (CPL + 1) * 32;                 // internal slots of a D-vector: cell c at c, scalar s at 32 CPL + s
    const int SOFF = arena_vectors(sp.max_depth);
    float *const Mf = DENSE ? sp.dense_minv + (long long)(blockIdx.x * NW + warp) * sp.dense_stride : nullptr;
    double *const Lc = DENSE ? sp.dense_L + (long long)(blockIdx.x * NW + warp) * sp.dense_stride : nullptr;
    double *const M2 = DENSE ? sp.dense_m2 + (long long)(blockIdx.x * NW + warp) * sp.dense_stride : nullptr;
    double *const pv = smem + ws + C::OFF_ZS;           // DI <= 2 NN doubles: spans Zs..Ts, free outside the evaluation
    const bool lane_active = (O.active_mask() >> lane) & 1u;
    const double l08 = -0.22314355131420976;   // log(0.8)
    // warm-up schedule (windowed_adaptation), identical for every chain
    const int nw = sp.warmup;
    int init_buffer = sp.init_buffer, term_buffer = sp.term_buffer, base_window = sp.window;
    if (nw < 20) { init_buffer = term_buffer = base_window = 0; }
    else if (init_buffer + base_window + term_buffer > nw) {
        init_buffer = (int)(0.15 * nw); term_buffer = (int)(0.1 * nw);
        base_window = nw - (init_buffer + term_buffer);
    }
    // ---- chain state: D-vectors and what the evaluation touches in registers, everything warp-uniform in ChainSt ----
    int phase = PH_FETCH;
    ProbS P;
    int cell[CPL];
    RngKey rk;
    WVec<CPL> q, p, g, minv, S, pprev;
    // dense_e: M^-1 p and M^-1 g of the state in registers, so that a leapfrog costs ONE matrix-vector product:
    // M^-1 (p + e/2 g) = M^-1 p + e/2 M^-1 g before the evaluation, and the same update with the new g after it
    WVec<CPL> ps_c, mg_c, yv_keep;
    bool sharp_valid = false;
    double V = 0.0, step = 0.0;
    ChainSt &st = *reinterpret_cast<ChainSt *>(smem + s_park);
    ProbDesc &pd = st.pd;
    long long &cg = st.cg;
    unsigned long long &ngrad = st.ngrad;
    int &it = st.it, &isave = st.isave, &win_size = st.win_size, &win_end = st.win_end, &init_try = st.init_try;
    double &eps = st.eps, &mu = st.mu, &sbar = st.sbar, &xbar = st.xbar, &cnt = st.cnt, &wn = st.wn;
    int &need_istep = st.need_istep;
    uint32_t &istep_tag = st.istep_tag;
    double &n_div = st.n_div, &n_maxd = st.n_maxd, &nleap_tot = st.nleap_tot, &nleap_samp = st.nleap_samp,
           &acc_sum = st.acc_sum, &n_samp_it = st.n_samp_it;
    int &is_it = st.is_it, &is_dir = st.is_dir;
    double &is_H0 = st.is_H0, &is_V0 = st.is_V0, &eps_try = st.eps_try;
    double &e_used = st.e_used, &H0 = st.H0, &otherV = st.otherV, &sampleV = st.sampleV, &sampleH = st.sampleH,
           &proposeV = st.proposeV, &proposeH = st.proposeH;
    double &lsw = st.lsw, &summetro = st.summetro, &lsw_sub = st.lsw_sub;
    int &curdir = st.curdir, &depth = st.depth, &nleap = st.nleap, &k = st.k, &nleaf = st.nleaf;
    int &divergent = st.divergent;
    double &u_leaf = st.u_leaf;
    TransInfo &info = st.info;
    const double *&yglob = st.yglob;
    const int *&iiglob = st.iiglob;
    double *&cpo_acc = st.cpo_acc, *&cs = st.cs;
    if (lane == 0) { st.ngrad = 0; st.cg = 0; }
    __syncwarp();
    // vectors parked across the evaluation: in shared memory when the layout has room for them (sl.vpark_off >= 0),
    // else as coalesced rows of the arena
    const int s_vp = pin_uniform(sp.sl.vpark_off >= 0 ? sp.sl.vpark_off + warp * 3 * VS : -1);

    // y = M^-1 x.  diag_e: elementwise; dense_e: column sweep over the active slots, x broadcast from shared memory,
    // M^-1 read coalesced (column g, 32 consecutive rows per load) from the per-warp float matrix.
    auto sharp = [&](const WVec<CPL> &x, WVec<CPL> &y) {
        if constexpr (!DENSE) {
#pragma unroll
            for (int t = 0; t < CPL; ++t) y.z[t] = minv.z[t] * x.z[t];
            y.s = minv.s * x.s;
        } else {
            __syncwarp();
#pragma unroll
            for (int t = 0; t < CPL; ++t) pv[t * 32 + lane] = x.z[t];
            pv[CPL * 32 + lane] = x.s;
            __syncwarp();
#pragma unroll
            for (int t = 0; t < CPL; ++t) y.z[t] = 0.0;
            y.s = 0.0;
            const int nz = (O.model() == 0) ? P.G : 0;
#pragma unroll 8
            for (int gg = 0; gg < nz; ++gg) {
                const double xg = pv[gg];
                const float *col = Mf + (size_t)gg * DI + lane;
#pragma unroll
                for (int t = 0; t < CPL; ++t) y.z[t] = fma((double)col[t * 32], xg, y.z[t]);
                y.s = fma((double)col[CPL * 32], xg, y.s);
            }
            for (int sidx = 0; sidx < SL_COUNT; ++sidx) {
                if (!((O.active_mask() >> sidx) & 1u)) continue;
                const double xg = pv[CPL * 32 + sidx];
                const float *col = Mf + (size_t)(CPL * 32 + sidx) * DI + lane;
#pragma unroll
                for (int t = 0; t < CPL; ++t) y.z[t] = fma((double)col[t * 32], xg, y.z[t]);
                y.s = fma((double)col[CPL * 32], xg, y.s);
            }
            __syncwarp();
        }
    };
    // sharp momentum of a stored p: mirror slot (dense) or recomputed (diag)
    auto sh_load = [&](int slot, const WVec<CPL> &pvec, WVec<CPL> &out) {
        if constexpr (DENSE) vload<CPL>(A + (long long)(slot + SOFF) * VS, out);
        else sharp(pvec, out);
    };
    auto sh_store = [&](int slot, const WVec<CPL> &psv) {
        if constexpr (DENSE) vstore<CPL>(A + (long long)(slot + SOFF) * VS, psv);
    };
    auto slot_active = [&](int f) -> bool {   // is internal slot f a parameter?
        return f < CPL * 32 ? (O.model() == 0 && f < P.G) : ((f - CPL * 32) < SL_COUNT && ((O.active_mask() >> (f - CPL * 32)) & 1u));
    };
    auto sample_momentum = [&](uint32_t tag, uint32_t iter, uint32_t idx0) {
        // diag_e_metric::sample_p: p_i = N(0,1) / sqrt(minv_i);  dense_e_metric::sample_p: p = L^-T z, M^-1 = L L^T
        double n0, n1;
#pragma unroll
        for (int t = 0; t < CPL; t += 2) {
            rng_lane_normal2(rk, tag, iter, idx0 + t, n0, n1);
            const double s0 = DENSE ? 1.0 : rsqrt(minv.z[t]);
            p.z[t] = ((cell[t] >> 17) & 1) ? n0 * s0 : 0.0;
            if (t + 1 < CPL) {
                const double s1 = DENSE ? 1.0 : rsqrt(minv.z[t + 1]);
                p.z[t + 1] = ((cell[t + 1] >> 17) & 1) ? n1 * s1 : 0.0;
            }
        }
        rng_lane_normal2(rk, tag, iter, idx0 + CPL + 1, n0, n1);
        p.s = lane_active ? n0 * (DENSE ? 1.0 : rsqrt(minv.s)) : 0.0;
        if constexpr (DENSE) {
            // back substitution on L^T (left-looking, L column-major lower): p_g = (z_g - sum_{f>g} L[f][g] p_f) / L[g][g]
            __syncwarp();
#pragma unroll
            for (int t = 0; t < CPL; ++t) pv[t * 32 + lane] = p.z[t];
            pv[CPL * 32 + lane] = p.s;
            __syncwarp();
            for (int gg = DI - 1; gg >= 0; --gg) {
                if (!slot_active(gg)) continue;
                const double *col = Lc + (size_t)gg * DI;
                double part = 0.0;
#pragma unroll
                for (int t = 0; t <= CPL; ++t) {
                    const int f = t * 32 + lane;
                    if (f > gg) part = fma(col[f], pv[f], part);
                }
                const double tot = warp_sum(part);
                const double pg = (pv[gg] - tot) / col[gg];
                __syncwarp();
                if (lane == 0) pv[gg] = pg;
                __syncwarp();
            }
#pragma unroll
            for (int t = 0; t < CPL; ++t) p.z[t] = pv[t * 32 + lane];
            p.s = pv[CPL * 32 + lane];
            __syncwarp();
        }
    };
    auto draw_init = [&](uint32_t tr) {
        double ua, ub;
#pragma unroll
        for (int t = 0; t < CPL; t += 2) {
            rng_u2(rk, RT_INIT, lane, tr, t, ua, ub);
            q.z[t] = ((cell[t] >> 17) & 1) ? sp.init_r * (2.0 * ua - 1.0) : 0.0;
            if (t + 1 < CPL) q.z[t + 1] = ((cell[t + 1] >> 17) & 1) ? sp.init_r * (2.0 * ub - 1.0) : 0.0;
        }
        rng_u2(rk, RT_INIT, lane, tr, CPL + 1, ua, ub);
        q.s = lane_active ? sp.init_r * (2.0 * ua - 1.0) : 0.0;
#pragma unroll
        for (int t = 0; t < CPL; ++t) { p.z[t] = 0.0; g.z[t] = 0.0; }
        p.s = 0.0; g.s = 0.0;
        step = 0.0;
    };
    auto istep_probe = [&]() {   // restore the starting point, draw a fresh momentum, aim one leapfrog of eps_try
        vload<CPL>(A + A_SQ * VS, q); vload<CPL>(A + A_SG * VS, g); V = is_V0;
        sample_momentum(RT_ISTEP, istep_tag, (uint32_t)is_it * (CPL + 4));
        WVec<CPL> ts;
        sharp(p, ts);
        is_H0 = V + 0.5 * vdot2<CPL>(p, ts);
        step = eps_try;
    };

    if constexpr (!TICK) __syncthreads();   // the tri table is complete (the tick barrier covers it otherwise)
    int tickn = 0;
    for (;;) {
        if constexpr (TICK) {
            if (tickn == 0 && __syncthreads_and(phase == PH_DONE)) break;
            if (++tickn >= sp.tick_period) tickn = 0;
        } else {
            if (phase == PH_DONE) break;
        }
        // ---- fetch the next chain -----------------------------------------------------------------------
        if (phase == PH_FETCH) {
            int item = 0;
            if (lane == 0) item = atomicAdd(sp.work_counter, 1);
            item = __shfl_sync(FULL, item, 0);
            if (item >= n_items) {
                phase = PH_DONE;
            } else {
                const int ipl = item / sp.chains;
                const int ip = sp.prob_list ? sp.prob_list[ipl] : ipl;
                cg = (long long)ip * sp.chains + (item - ipl * sp.chains);
                pd = descs[ip];
                __syncwarp();
                for (int e = lane; e < pd.n_d_smem; e += 32) smem[s_pd + e] = dpool[pd.off_d + e];
                for (int e = lane; e < pd.n_i_smem; e += 32) SMI(s_pi + e) = ipool[pd.off_i + e];
                __syncwarp();
                bind_problem(P, pd, s_pd, s_pi, s_tri);
                make_cells<CPL, SP>(F, P, cell);
                rk.k0 = (uint32_t)sp.seed; rk.k1 = (uint32_t)(sp.seed >> 32); rk.chain = (uint32_t)cg;
#pragma unroll
                for (int t = 0; t < CPL; ++t) minv.z[t] = ((cell[t] >> 17) & 1) ? 1.0 : 0.0;
                minv.s = lane_active ? 1.0 : 0.0;
                if constexpr (DENSE) {   // M^-1 = L = I on the parameter slots, Welford accumulator = 0
                    for (int gg = 0; gg < DI; ++gg) {
                        const bool ag = slot_active(gg);
#pragma unroll
                        for (int t = 0; t <= CPL; ++t) {
                            const int f = t * 32 + lane;
                            const double v = (f == gg && ag) ? 1.0 : 0.0;
                            Mf[(size_t)gg * DI + f] = (float)v;
                            Lc[(size_t)gg * DI + f] = (f == gg) ? 1.0 : 0.0;
                            M2[(size_t)gg * DI + f] = 0.0;
                        }
                    }
                    __syncwarp();
                }
                cs = sp.chain_stats ? sp.chain_stats + cg * 8 : nullptr;
                yglob = dpool + pd.off_d + pool_y_off(pd);
                iiglob = ipool + pd.off_i + pool_ii_off(pd);
                cpo_acc = sp.cpo_acc ? sp.cpo_acc + cg * sp.ld_cpo : nullptr;
                if (cpo_acc && !sp.resume_in) for (int e = lane; e < pd.N; e += 32) cpo_acc[e] = 0.0;
                init_try = 0;
                draw_init(0);
                if (sp.resume_in) {   // continue a finished chain: its last position and adapted (diagonal) metric
                    const double *R = sp.resume + cg * sp.resume_stride;
                    vload<CPL>(R, q);
                    if constexpr (!DENSE) vload<CPL>(R + VS, minv);
                }
                if constexpr (!DENSE) vstore<CPL>(A + A_PARK_MINV * VS, minv);
                phase = PH_INIT;
            }
        }
        if (phase == PH_DONE) continue;
        // ---- the one leapfrog (expl_leapfrog::evolve): half-step p, full-step q, gradient, half-step p ----
        {
            const double he = 0.5 * step;
#pragma unroll
            for (int t = 0; t < CPL; ++t) p.z[t] = fma(he, g.z[t], p.z[t]);
            p.s = fma(he, g.s, p.s);
            {
                WVec<CPL> yv;   // dtau/dp = M^-1 p
                if (DENSE && sharp_valid) {
#pragma unroll
                    for (int t = 0; t < CPL; ++t) yv.z[t] = fma(he, mg_c.z[t], ps_c.z[t]);
                    yv.s = fma(he, mg_c.s, ps_c.s);
                } else {
                    sharp(p, yv);
                }
#pragma unroll
                for (int t = 0; t < CPL; ++t) q.z[t] = fma(step, yv.z[t], q.z[t]);
                q.s = fma(step, yv.s, q.s);
                if constexpr (DENSE) yv_keep = yv;
            }
            GQ gq;
            gq.on = (phase == PH_GQ);
            gq.so = s_park + PARK_GQ; gq.fwd = false;
            if (gq.on) {
                const long long rowi = cg * sp.ld_save + sp.isave0 + isave;
                double uz[4];
                rng_u2(rk, RT_GQ, 0xFFu, it, 0, uz[0], uz[1]);
                rng_u2(rk, RT_GQ, 0xFFu, it, 1, uz[2], uz[3]);
                gq_publish(gq.so, pd.D, yglob, iiglob, uz, sp.draws ? sp.draws + rowi * sp.ld_draws : nullptr,
                           sp.qdraws ? sp.qdraws + rowi * 12 : nullptr, (it >= nw) ? cpo_acc : nullptr,
                           (sp.f_acc && it >= nw) ? sp.f_acc + cg * sp.ld_f : nullptr);
            }
            // The evaluation needs every register it can get; S, pprev and p are dead weight across it: they are parked
            // in shared memory (or, when it has no room, in the arena: L2-resident, coalesced).  M^-1 only changes at a
            // metric update, where it is written to its arena row; it is re-read after the evaluation.
            if (s_vp >= 0) {
                vstore<CPL>(smem + s_vp, S); vstore<CPL>(smem + s_vp + VS, pprev); vstore<CPL>(smem + s_vp + 2 * VS, p);
            } else {
                vstore<CPL>(A + A_PARK_S * VS, S); vstore<CPL>(A + A_PARK_PREV * VS, pprev); vstore<CPL>(A + A_PARK_P * VS, p);
            }
            if constexpr (DENSE) vstore<CPL>(A + A_PARK_YV * VS, yv_keep);
            double lp;
            const bool okv = warp_logp_grad<CPL, false, SP>(F, P, ws, cell, q, sp.jacobian, lp, g, nullptr, gq);
            if (s_vp >= 0) {
                vload<CPL>(smem + s_vp, S); vload<CPL>(smem + s_vp + VS, pprev); vload<CPL>(smem + s_vp + 2 * VS, p);
            } else {
                vload<CPL>(A + A_PARK_S * VS, S); vload<CPL>(A + A_PARK_PREV * VS, pprev); vload<CPL>(A + A_PARK_P * VS, p);
            }
            if constexpr (!DENSE) vload<CPL>(A + A_PARK_MINV * VS, minv);
            else vload<CPL>(A + A_PARK_YV * VS, yv_keep);
            rk.k0 = (uint32_t)sp.seed; rk.k1 = (uint32_t)(sp.seed >> 32); rk.chain = (uint32_t)cg;
            if (!gq.on) ++ngrad;
            if (!okv || !(fabs(lp) < INFINITY)) {   // domain error in the reference => infinite potential
                V = INFINITY;
#pragma unroll
                for (int t = 0; t < CPL; ++t) g.z[t] = 0.0;
                g.s = 0.0;
            } else {
                V = -lp;
            }
#pragma unroll
            for (int t = 0; t < CPL; ++t) p.z[t] = fma(he, g.z[t], p.z[t]);
            p.s = fma(he, g.s, p.s);
        }
        WVec<CPL> ps;   // sharp momentum of the state just reached
        if constexpr (DENSE) {
            const double he = 0.5 * step;
            sharp(g, mg_c);
#pragma unroll
            for (int t = 0; t < CPL; ++t) ps.z[t] = fma(he, mg_c.z[t], yv_keep.z[t]);
            ps.s = fma(he, mg_c.s, yv_keep.s);
            ps_c = ps;
            sharp_valid = true;   // cleared below wherever the bookkeeping replaces p, g or the metric
        } else {
            sharp(p, ps);
        }
        auto kinetic = [&]() { return 0.5 * vdot2<CPL>(p, ps); };
        // ---- bookkeeping of this chain ---------------------------------------------------------------
        int next = N_NONE;
        bool keep_sharp = false;   // set where the next leapfrog starts from exactly this (p, g): the next leaf of a subtree
        if (phase == PH_INIT) {
            // services::util::initialize: u ~ U(-R, R) until lp and its gradient are finite (<= 100 tries)
            bool fin = V < INFINITY && fabs(g.s) < INFINITY;
#pragma unroll
            for (int t = 0; t < CPL; ++t) fin = fin && fabs(g.z[t]) < INFINITY;
            if (__all_sync(FULL, fin)) {
                eps = sp.stepsize0;
                need_istep = 1;            // base_hmc::init_stepsize runs before the first transition ...
                if (sp.init_u) store_user_vec<CPL>(F, pd, cell, q, sp.init_u + cg * sp.ld_im);
                if (sp.resume_in && init_try == 0) {   // ... unless the chain resumes with its adapted step size; a resumed
                    eps = sp.resume[cg * sp.resume_stride + 2 * VS];   // chain that re-adapts (warmup > 0: a stricter adapt_delta)
                    need_istep = (nw > 0) ? 1 : 0;                     // restarts the dual averaging from it, as Stan does
                }
                istep_tag = 0xFFFFFF00u;
                mu = 0.0; sbar = 0.0; xbar = 0.0; cnt = 0.0; wn = 0.0;
                win_size = base_window; win_end = init_buffer + base_window - 1;
                it = 0; isave = 0;
                n_div = n_maxd = nleap_tot = nleap_samp = acc_sum = n_samp_it = 0.0;
                WVec<CPL> zero;
#pragma unroll
                for (int t = 0; t < CPL; ++t) zero.z[t] = 0.0;
                zero.s = 0.0;
                vstore<CPL>(A + A_WM * VS, zero); vstore<CPL>(A + A_WM2 * VS, zero);
                next = N_ITER;
            } else if (++init_try < 100) {
                draw_init((uint32_t)init_try);
            } else {
                if (cs && lane == 0) { for (int e = 0; e < 8; ++e) cs[e] = 0.0; cs[6] = 1.0; }
                phase = PH_FETCH;
            }
        } else if (phase == PH_ISTEP) {
            // base_hmc::init_stepsize: double / halve eps until the one-step acceptance crosses 0.8
            double h = V + kinetic();
            if (!(h == h)) h = INFINITY;
            const double dH = is_H0 - h;
            bool done = false;
            if (is_it == 0) {
                is_dir = dH > l08 ? 1 : -1;   // the first probe only fixes the direction
            } else {
                if ((is_dir == 1 && !(dH > l08)) || (is_dir == -1 && !(dH < l08))) done = true;
                else {
                    eps_try = is_dir == 1 ? 2.0 * eps_try : 0.5 * eps_try;
                    if (eps_try > 1e7 || eps_try == 0.0 || is_it >= 63) done = true;
                }
            }
            if (!done) {
                ++is_it;
                istep_probe();
            } else {
                vload<CPL>(A + A_SQ * VS, q); vload<CPL>(A + A_SG * VS, g); V = is_V0;
                eps = eps_try;
                mu = log(10.0 * eps); sbar = 0.0; xbar = 0.0; cnt = 0.0;
                need_istep = 0;
                next = N_TRANS;
            }
        } else if (phase == PH_TREE) {
            // ---- leaf k of the subtree being built (base_nuts::build_tree, depth 0) ----
            ++nleap;
            double h = V + kinetic();
            if (!(h == h)) h = INFINITY;
            const double w = H0 - h;
            if (h - H0 > 1000.0) divergent = 1;
            // ONE exponential serves the running log-sum of the subtree's weights and the reservoir probability of this
            // leaf: with a = w - lsw_sub(old) and e = exp(-|a|),  lsw_sub(new) = max + log(1 + e)  and
            // pr = exp(w - lsw_sub(new)) = 1 / (1 + exp(-a)).
            double pr = 1.0;
            if (lsw_sub == -INFINITY) {
                lsw_sub = w;                       // first weight of the subtree
            } else if (w == -INFINITY) {
                pr = 0.0;                          // H = inf: weight zero
            } else {
                const double a = w - lsw_sub, e = fast_exp(-fabs(a)), r1 = fast_rcp(1.0 + e);
                pr = (a >= 0.0) ? r1 : e * r1;
                lsw_sub = fmax(lsw_sub, w) + fast_log(1.0 + e);
            }
            summetro += (w > 0.0) ? 1.0 : fast_exp(w);
            bool valid = !divergent;
            if (valid) {
                // uniform progressive sampling of the subtree's proposal.  ONE uniform per subtree is recycled:
                // after a Bernoulli(pr) decision taken as (u < pr), u/pr resp. (u-pr)/(1-pr) is again U(0,1).
                bool take = (k == 0);
                if (!take) {
                    take = u_leaf < pr;
                    u_leaf = (take ? u_leaf : u_leaf - pr) * fast_rcp(take ? pr : 1.0 - pr);
                    u_leaf = fmin(fmax(u_leaf, 0.0), 1.0);
                }
                if (take) {
                    vstore<CPL>(A + A_PQ * VS, q); vstore<CPL>(A + A_PG * VS, g);
                    proposeV = V; proposeH = h;
                }
#pragma unroll
                for (int t = 0; t < CPL; ++t) S.z[t] += p.z[t];
                S.s += p.s;
                if (k == 0) { vstore<CPL>(A + A_P_NEW_NEAR * VS, p); sh_store(A_P_NEW_NEAR, ps); }
                // leaf k opens blocks at every level l >= 2 with k % 2^l == 0
                if ((k & 3) == 0) {
                    for (int l = 2; l <= depth && (k & ((1 << l) - 1)) == 0; ++l) {
                        vstore<CPL>(A + (A_LVL + 3 * (l - 2)) * VS, p);
                        sh_store(A_LVL + 3 * (l - 2), ps);
                    }
                }
                if (k & 1) {
                    // level 1: left = leaf k-1, right = leaf k (the three tests coincide)
                    WVec<CPL> car;
#pragma unroll
                    for (int t = 0; t < CPL; ++t) car.z[t] = pprev.z[t] + p.z[t];
                    car.s = pprev.s + p.s;
                    WVec<CPL> pprev_s;
                    sh_load(A_PREV, pprev, pprev_s);
                    valid = uturn_sharp_ok<CPL>(pprev_s, ps, car);
                    int l = 2;
                    const int kk = k + 1;
                    while (valid && l <= depth && (kk & ((1 << l) - 1)) == 0) {
                        double *Lv = A + (A_LVL + 3 * (l - 2)) * VS;
                        WVec<CPL> pbegL, rhoL, pendL, pbegR, ext, e2, e3, pbegL_s, pbegR_s, pendL_s;
                        // issue all loads of this merge together: L's first / last leaf and rho, R's first leaf
                        vload<CPL>(Lv, pbegL); vload<CPL>(Lv + 2 * VS, rhoL); vload<CPL>(Lv + VS, pendL);
                        if (l == 2) pbegR = pprev; else vload<CPL>(A + (A_LVL + 3 * (l - 3)) * VS, pbegR);
                        sh_load(A_LVL + 3 * (l - 2), pbegL, pbegL_s);
                        sh_load(A_LVL + 3 * (l - 2) + 1, pendL, pendL_s);
                        if (l == 2) pbegR_s = pprev_s; else sh_load(A_LVL + 3 * (l - 3), pbegR, pbegR_s);
#pragma unroll
                        for (int t = 0; t < CPL; ++t) {
                            ext.z[t] = rhoL.z[t] + car.z[t];      // rho_L + rho_R around the merged block
                            e2.z[t] = rhoL.z[t] + pbegR.z[t];     // rho_L + p(first leaf of R)
                            e3.z[t] = car.z[t] + pendL.z[t];      // rho_R + p(last leaf of L)
                        }
                        ext.s = rhoL.s + car.s; e2.s = rhoL.s + pbegR.s; e3.s = car.s + pendL.s;
                        const bool okm = uturn3_ok<CPL>(pbegL_s, ps, ext, pbegL_s, pbegR_s, e2, pendL_s, ps, e3);
                        valid = okm;
                        car = ext;
                        ++l;
                    }
                    if (valid && l <= depth) {   // the completed block is the left half of a level-l block
                        double *Lv = A + (A_LVL + 3 * (l - 2)) * VS;
                        vstore<CPL>(Lv + VS, p);
                        sh_store(A_LVL + 3 * (l - 2) + 1, ps);
                        vstore<CPL>(Lv + 2 * VS, car);
                    }
                }
            }
            if (!valid) {
                next = N_ENDTRANS;           // invalid subtree: the trajectory stops, depth is not incremented
            } else if (k + 1 < nleaf) {
                ++k;
                pprev = p;
                sh_store(A_PREV, ps);
                keep_sharp = true;
            } else {
                // ---- subtree complete: multinomial sample across subtrees (biased progressive sampling) ----
                ++depth;
                bool take = lsw_sub > lsw;
                if (!take) take = rng_warp_uniform(rk, RT_TOP, it, depth) < fast_exp(lsw_sub - lsw);
                if (take) {
                    vcopy<CPL>(A + A_SQ * VS, A + A_PQ * VS); vcopy<CPL>(A + A_SG * VS, A + A_PG * VS);
                    sampleV = proposeV; sampleH = proposeH;
                }
                lsw = log_add_exp(lsw, lsw_sub);
                // U-turn across the whole trajectory + the two cross-subtree tests
                WVec<CPL> rho_old, pfar, rho;
                vload<CPL>(A + A_RHO_OLD * VS, rho_old);
                vload<CPL>(A + A_OP * VS, pfar);
#pragma unroll
                for (int t = 0; t < CPL; ++t) rho.z[t] = rho_old.z[t] + S.z[t];
                rho.s = rho_old.s + S.s;
                WVec<CPL> pfar_s, pnn, pnn_s, pon, pon_s, e2, e3;
                sh_load(A_OP, pfar, pfar_s);
                if (nleaf == 1) { pnn = p; pnn_s = ps; }
                else { vload<CPL>(A + A_P_NEW_NEAR * VS, pnn); sh_load(A_P_NEW_NEAR, pnn, pnn_s); }
                vload<CPL>(A + A_P_OLD_NEAR * VS, pon);
                sh_load(A_P_OLD_NEAR, pon, pon_s);
                vstore<CPL>(A + A_RHO_OLD * VS, rho);
#pragma unroll
                for (int t = 0; t < CPL; ++t) {
                    e2.z[t] = rho_old.z[t] + pnn.z[t];
                    e3.z[t] = S.z[t] + pon.z[t];
                }
                e2.s = rho_old.s + pnn.s; e3.s = S.s + pon.s;
                const bool persist = uturn3_ok<CPL>(pfar_s, ps, rho, pfar_s, pnn_s, e2, pon_s, ps, e3);
                next = (persist && depth < sp.max_depth) ? N_DOUBLE : N_ENDTRANS;
            }
        } else if (phase == PH_GQ) {
            // the saved draw's generated quantities have just been written by the evaluation
            const long long rowi = cg * sp.ld_save + sp.isave0 + isave;
            if (sp.qdraws && lane == 0) { double *qo = sp.qdraws + rowi * 12; qo[10] = -V; qo[11] = 0.0; }
            ++isave;
            ++it;
            next = N_ITER;
        }
        // ---- end of a transition: adaptation, statistics, saving -------------------------------------------
        if (next == N_ENDTRANS) {
            info.accept = summetro / (double)(nleap > 0 ? nleap : 1);
            info.depth = depth; info.nleap = nleap; info.divergent = divergent ? 1 : 0; info.energy = sampleH;
            vload<CPL>(A + A_SQ * VS, q); vload<CPL>(A + A_SG * VS, g); V = sampleV;
            const bool warm = it < nw;
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
                if ((it >= init_buffer) && (it < nw - term_buffer)) {
                    wn += 1.0;
                    WVec<CPL> m, m2, dold;
                    vload<CPL>(A + A_WM * VS, m); vload<CPL>(A + A_WM2 * VS, m2);
                    const double iw = 1.0 / wn;
#pragma unroll
                    for (int t = 0; t < CPL; ++t) {
                        const double d = q.z[t] - m.z[t];
                        dold.z[t] = d;
                        m.z[t] = fma(d, iw, m.z[t]);
                        m2.z[t] = fma(q.z[t] - m.z[t], d, m2.z[t]);
                    }
                    {
                        const double d = q.s - m.s;
                        dold.s = d;
                        m.s = fma(d, iw, m.s);
                        m2.s = fma(q.s - m.s, d, m2.s);
                    }
                    if constexpr (DENSE) {
                        // covar_adaptation::learn_covariance (Welford): M2 += (q - mean_new) (q - mean_old)^T
                        __syncwarp();
#pragma unroll
                        for (int t = 0; t < CPL; ++t) pv[t * 32 + lane] = q.z[t] - m.z[t];
                        pv[CPL * 32 + lane] = q.s - m.s;
                        __syncwarp();
                        for (int gg = 0; gg < DI; ++gg) {
                            if (!slot_active(gg)) continue;
                            const double dn = pv[gg];
                            double *col = M2 + (size_t)gg * DI + lane;
#pragma unroll
                            for (int t = 0; t < CPL; ++t) col[t * 32] = fma(dn, dold.z[t], col[t * 32]);
                            col[CPL * 32] = fma(dn, dold.s, col[CPL * 32]);
                        }
                        __syncwarp();
                    }
                    if (it == win_end) {
                        const double aa = wn / (wn + 5.0), bb = 1e-3 * (5.0 / (wn + 5.0)), inm1 = 1.0 / (wn - 1.0);
                        if constexpr (!DENSE) {
#pragma unroll
                            for (int t = 0; t < CPL; ++t)
                                if ((cell[t] >> 17) & 1) minv.z[t] = fma(aa, m2.z[t] * inm1, bb);
                            if (lane_active) minv.s = fma(aa, m2.s * inm1, bb);
                            vstore<CPL>(A + A_PARK_MINV * VS, minv);
                        } else {
                            // M^-1 = n/(n+5) cov + 1e-3 5/(n+5) I, symmetrised from the lower triangle, rounded to float;
                            // L = chol(M^-1) of exactly that rounded matrix (column-major lower, right-looking, in place)
                            for (int gg = 0; gg < DI; ++gg) {
                                const bool ag = slot_active(gg);
#pragma unroll
                                for (int t = 0; t <= CPL; ++t) {
                                    const int f = t * 32 + lane;
                                    double v = (f == gg) ? 1.0 : 0.0;
                                    if (ag && slot_active(f)) {
                                        const int hi = f > gg ? f : gg, lo = f > gg ? gg : f;
                                        v = aa * M2[(size_t)hi * DI + lo] * inm1 + ((f == gg) ? bb : 0.0);
                                    }
                                    const float vf = (float)v;
                                    Mf[(size_t)gg * DI + f] = vf;
                                    Lc[(size_t)gg * DI + f] = (double)vf;
                                }
                            }
                            __syncwarp();
                            for (int gg = 0; gg < DI; ++gg) {
#pragma unroll
                                for (int t = 0; t <= CPL; ++t) M2[(size_t)gg * DI + t * 32 + lane] = 0.0;
                            }
                            for (int kk = 0; kk < DI; ++kk) {
                                if (!slot_active(kk)) continue;
                                double *colk = Lc + (size_t)kk * DI;
                                const double inv = rsqrt(fmax(colk[kk], 1e-300));
                                double ck[CPL + 1];
                                __syncwarp();
#pragma unroll
                                for (int t = 0; t <= CPL; ++t) {
                                    const int f = t * 32 + lane;
                                    ck[t] = (f >= kk) ? colk[f] * inv : 0.0;
                                    if (f >= kk) colk[f] = ck[t];
                                }
                                __syncwarp();
                                for (int jj = kk + 1; jj < DI; ++jj) {
                                    if (!slot_active(jj)) continue;
                                    const double ljk = colk[jj];
                                    double *colj = Lc + (size_t)jj * DI;
#pragma unroll
                                    for (int t = 0; t <= CPL; ++t) {
                                        const int f = t * 32 + lane;
                                        if (f >= jj) colj[f] = fma(-ck[t], ljk, colj[f]);
                                    }
                                }
                            }
                            __syncwarp();
                        }
#pragma unroll
                        for (int t = 0; t < CPL; ++t) { m.z[t] = 0.0; m2.z[t] = 0.0; }
                        m.s = 0.0; m2.s = 0.0;
                        wn = 0.0;
                        if (win_end != nw - term_buffer - 1) {   // compute_next_window
                            win_size *= 2;
                            win_end = it + win_size;
                            if (win_end != nw - term_buffer - 1) {
                                const int next_boundary = win_end + 2 * win_size;
                                if (next_boundary >= nw - term_buffer) win_end = nw - term_buffer - 1;
                            }
                        }
                        need_istep = 1;   // ... and after every metric update (restart of the dual averaging)
                        istep_tag = 0xFFFF0000u + (uint32_t)it;
                    }
                    vstore<CPL>(A + A_WM * VS, m); vstore<CPL>(A + A_WM2 * VS, m2);
                }
                if (it == nw - 1) eps = exp(xbar);   // complete_adaptation
            } else {
                n_div += info.divergent; n_maxd += (info.depth >= sp.max_depth); nleap_samp += info.nleap;
                acc_sum += info.accept; n_samp_it += 1.0;
            }
            const int base = warm ? 0 : nw;
            const bool save = (!warm || sp.save_warmup) && ((it - base) % sp.thin == 0) && isave < sp.n_save;
            if (save) {
                const long long rowi = cg * sp.ld_save + sp.isave0 + isave;
                if (sp.udraws) store_user_vec<CPL>(F, pd, cell, q, sp.udraws + rowi * sp.ld_u);
                if (sp.sp && lane == 0) {
                    double *o = sp.sp + rowi * 7;
                    o[0] = info.accept; o[1] = e_used; o[2] = info.depth; o[3] = info.nleap; o[4] = info.divergent;
                    o[5] = info.energy; o[6] = -V;
                }
            }
            if (save && (sp.draws || sp.qdraws || ((cpo_acc || sp.f_acc) && !warm))) {
                // one evaluation tick at the saved draw with the generated-quantities sink switched on
#pragma unroll
                for (int t = 0; t < CPL; ++t) p.z[t] = 0.0;
                p.s = 0.0;
                step = 0.0;
                phase = PH_GQ;
                next = N_NONE;
            } else {
                if (save) ++isave;
                ++it;
                next = N_ITER;
            }
        }
        if (next == N_ITER) {
            if (it >= sp.iter) {   // chain finished
                if (cs && lane == 0) {
                    if (sp.resume_in) {   // an extension launch adds to the statistics of the chain's earlier launches
                        const double n_old = cs[7], n_new = n_old + (double)isave;
                        cs[0] = eps; cs[1] += n_div; cs[2] += n_maxd; cs[3] += nleap_tot; cs[4] += nleap_samp;
                        cs[5] = n_new > 0 ? (cs[5] * n_old + acc_sum) / n_new : 0.0; cs[7] = n_new;
                    } else {
                        cs[0] = eps; cs[1] = n_div; cs[2] = n_maxd; cs[3] = nleap_tot; cs[4] = nleap_samp;
                        cs[5] = n_samp_it > 0 ? acc_sum / n_samp_it : 0.0; cs[6] = 0.0; cs[7] = (double)isave;
                    }
                }
                if (sp.resume_out) {
                    double *R = sp.resume + cg * sp.resume_stride;
                    vstore<CPL>(R, q);
                    if constexpr (!DENSE) vstore<CPL>(R + VS, minv);
                    if (lane == 0) R[2 * VS] = eps;
                }
                if (sp.inv_metric) {
                    if constexpr (DENSE) {   // the sink is a D-vector: the diagonal of the adapted dense M^-1
                        WVec<CPL> dgm;
#pragma unroll
                        for (int t = 0; t < CPL; ++t) dgm.z[t] = (double)Mf[(size_t)(t * 32 + lane) * (DI + 1)];
                        dgm.s = (double)Mf[(size_t)(CPL * 32 + lane) * (DI + 1)];
                        store_user_vec<CPL>(F, pd, cell, dgm, sp.inv_metric + cg * sp.ld_im);
                    } else {
                        store_user_vec<CPL>(F, pd, cell, minv, sp.inv_metric + cg * sp.ld_im);
                    }
                }
                phase = PH_FETCH;
                next = N_NONE;
            } else if (need_istep) {
                vstore<CPL>(A + A_SQ * VS, q); vstore<CPL>(A + A_SG * VS, g);
                is_V0 = V; is_it = 0; is_dir = 0; eps_try = eps;
                istep_probe();
                phase = PH_ISTEP;
                next = N_NONE;
            } else {
                next = N_TRANS;
            }
        }
        if (next == N_TRANS) {
            // ---- base_nuts::transition: fresh momentum, initial trajectory = the current point ----
            e_used = eps;
            if (sp.jitter > 0.0) e_used = eps * (1.0 + sp.jitter * (2.0 * rng_warp_uniform(rk, RT_JIT, it, 0) - 1.0));
            sample_momentum(RT_MOM, (uint32_t)it, 0);
            sharp(p, ps);
            H0 = V + kinetic();
            vstore<CPL>(A + A_OQ * VS, q); vstore<CPL>(A + A_OP * VS, p); vstore<CPL>(A + A_OG * VS, g);
            sh_store(A_OP, ps);
            vstore<CPL>(A + A_SQ * VS, q); vstore<CPL>(A + A_SG * VS, g);
            vstore<CPL>(A + A_RHO_OLD * VS, p);
            otherV = V; sampleV = V; sampleH = H0; proposeV = V; proposeH = H0;
            curdir = 1; lsw = 0.0; summetro = 0.0; depth = 0; nleap = 0; divergent = 0;
            next = N_DOUBLE;
        }
        if (next == N_DOUBLE) {
            // ---- extend the trajectory by a subtree of 2^depth leapfrogs in a random direction ----
            double u_dir;
            rng_u2(rk, RT_DIR, 0xFFu, it, depth, u_dir, u_leaf);   // direction + this subtree's sampling uniform
            const int dir = u_dir > 0.5 ? 1 : -1;
            if (dir != curdir) {   // make the end being extended the register-resident one
                WVec<CPL> tv;
                vload<CPL>(A + A_OQ * VS, tv); vstore<CPL>(A + A_OQ * VS, q); q = tv;
                vload<CPL>(A + A_OP * VS, tv); vstore<CPL>(A + A_OP * VS, p); p = tv;
                if constexpr (DENSE) { vload<CPL>(A + (long long)(A_OP + SOFF) * VS, tv); vstore<CPL>(A + (long long)(A_OP + SOFF) * VS, ps); ps = tv; }
                else sharp(p, ps);
                vload<CPL>(A + A_OG * VS, tv); vstore<CPL>(A + A_OG * VS, g); g = tv;
                const double t2 = otherV; otherV = V; V = t2;
                curdir = dir;
            }
            vstore<CPL>(A + A_P_OLD_NEAR * VS, p);
            sh_store(A_P_OLD_NEAR, ps);
            sh_store(A_PREV, ps);
#pragma unroll
            for (int t = 0; t < CPL; ++t) S.z[t] = 0.0;
            S.s = 0.0;
            lsw_sub = -INFINITY;
            nleaf = 1 << depth;
            k = 0;
            step = dir * e_used;
            pprev = p;
            phase = PH_TREE;
        }
        if (!keep_sharp) sharp_valid = false;
    }
    if (lane == 0 && ngrad) atomicAdd(sp.grad_counter, ngrad);
}

}  // namespace bsyn
