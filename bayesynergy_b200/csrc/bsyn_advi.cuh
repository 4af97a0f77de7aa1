// bsyn_advi.cuh -- full-rank ADVI (rstan::vb(algorithm = "fullrank"), R/bayesynergy.R:225-231; the default method of
// synergyscreen, R/synergyscreen.R:90-93) on the same lp+gradient evaluation the sampler uses.
//
// The algorithm is Stan's (un-vendored StanHeaders; stan/variational/advi.hpp: run / adapt_eta /
// stochastic_gradient_ascent / calc_ELBO / calc_ELBO_grad, families/normal_fullrank.hpp: sample / calc_grad /
// entropy), restated for one warp per problem:
//   * variational state of a problem -- mu[D], L[D x D] lower, and the two adaGrad histories -- lives in a per-warp
//     slab in HBM/L2 (row-major, row stride DP = D rounded up to 32 so every row sweep is coalesced); it is ~0.2 MB
//     for D = 115, far more than a warp's share of shared memory;
//   * one TICK = draw eta, zeta = mu + L eta (row sweep, 8 rows per butterfly), ONE lp+gradient evaluation, then the
//     bookkeeping of whatever the evaluation was for (initial point, ELBO sample, gradient sample -> adaGrad sweep,
//     output draw -> generated quantities).  As in the sampler there is a single copy of the model code in the
//     kernel.
//   * grad_samples = 1 (rstan's default) is the only supported value: the gradient of L is then the rank-one
//     g eta^T + diag(1 / L_ii), which the sweep forms on the fly instead of storing.
#pragma once
#include <cfloat>

#include "bsyn_nuts.cuh"

namespace bsyn {

struct AdviParams {
    int n_problems;
    int max_iter, elbo_samples, eval_elbo, adapt_engaged, adapt_iter, output_samples, cb_size;
    double eta, tol_rel_obj, init_r;
    unsigned long long seed;
    SmemLayout sl;
    int DP;                                   // row stride of L / H (multiple of 32, >= max D)
    double *state; long long state_stride;    // per resident warp: 6 vectors [DP] + L [DP*DP] (double) + H [DP*DP] (float)
    // per-problem sinks (device; draws / udraws / cpo may be null)
    double *mean; int ld_mean;                // variational mean, unconstrained, Stan order
    double *info;                             // [n][8]: status, eta, iterations, last ELBO, convergence bits, n_grad, n_lp, 0
    double *draws; long long ld_draws;        // [n][output_samples][ld_draws] write_array rows of the approximate posterior draws
    double *udraws; long long ld_u;           // [n][output_samples][ld_u]
    double *qdraws;                           // [n][output_samples][12]
    double *cpo_acc; int ld_cpo;              // [n][ld_cpo]
    int *work_counter;
    unsigned long long *grad_counter;
};

enum AdviStage { VS_FETCH = 0, VS_INIT, VS_ELBO, VS_GRAD, VS_DRAW, VS_DONE };
enum AdviElboFor { VE_INIT = 0, VE_ADAPT, VE_MAIN };
constexpr uint32_t RT_VB = 9;   // RNG stream tag of the eta draws
constexpr int ADVI_NW = 12;     // warps (experiments) per CTA

template <int CPL, int NW>
__global__ void __launch_bounds__(32 * NW, 12 / NW) advi_kernel(const Flags F, const ProbDesc *__restrict__ descs,
                                                                       const double *__restrict__ dpool, const int *__restrict__ ipool,
                                                                       const int *__restrict__ tri_tab, const AdviParams vp)
{
    using C = Cfg<CPL>;
    constexpr int ES = CPL + 1;               // eta / mu entries per lane: index lane + 32 s
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int s_tri = 2 * vp.sl.tri_off;
    const int s_pd = pin_uniform(vp.sl.pd_off + warp * vp.sl.pd_stride);
    const int s_pi = pin_uniform(2 * (vp.sl.pi_off + warp * vp.sl.pi_stride));
    const int ws = pin_uniform(vp.sl.ws_off + warp * C::WS_DOUBLES);
    const int s_gq = pin_uniform(vp.sl.park_off + warp * PARK_DOUBLES + PARK_GQ);
    for (int e = threadIdx.x; e < C::TR; e += blockDim.x) SMI(s_tri + e) = tri_tab[e];
    for (int e = lane; e < C::WS_DOUBLES; e += 32) smem[ws + e] = 0.0;   // the factor's upper triangle must stay zero
    const int DP = vp.DP;
    double *const St = vp.state + (long long)(blockIdx.x * NW + warp) * vp.state_stride;
    double *const mu = St, *const hmu = St + DP, *const init = St + 2 * DP, *const zeta = St + 3 * DP, *const gu = St + 4 * DP;
    double *const etav = St + 5 * DP;
    double *__restrict__ const Lm = St + 6 * DP;
    float *__restrict__ const Hm = reinterpret_cast<float *>(Lm + (long long)DP * DP);   // adaGrad history of L: FP32 is plenty
    const bool lane_active = (F.active_mask >> lane) & 1u;
    __syncthreads();

    int stage = VS_FETCH, efor = VE_INIT;
    ProbS P;
    ProbDesc pd;
    int cell[CPL];
    RngKey rk;
    int ip = 0, D = 0;
    WVec<CPL> q, g;
    const double *yglob = nullptr;
    const int *iiglob = nullptr;
    double *cpo_acc = nullptr;
    uint32_t tick = 0;                        // counts eta draws of this problem: RNG iteration field
    int init_try = 0, it = 0, k_eta = 0, e_done = 0, e_drop = 0, idraw = 0, status = 0, conv = 0, iters = 0;
    bool in_adapt = false;
    double eta = 0.0, e_sum = 0.0, elbo = 0.0, elbo_best = 0.0, elbo_init = 0.0, eta_best = 0.0;
    double cbv = 0.0;                         // lane l holds entry l of the circular buffer of relative ELBO changes
    int cb_n = 0, cb_head = 0;
    double n_grad = 0.0, n_lp = 0.0;

    // Q(cont_params): mu = init, L = I, histories = 0
    auto reset_q = [&]() {
        for (int s = 0; s < ES; ++s) {
            const int i = lane + 32 * s;
            if (i < DP) { mu[i] = init[i]; hmu[i] = 0.0; }
        }
        for (int i = 0; i < D; ++i)
            for (int s = 0; s < ES; ++s) {
                const int j = lane + 32 * s;
                if (j <= i) { Lm[(long long)i * DP + j] = (j == i) ? 1.0 : 0.0; Hm[(long long)i * DP + j] = 0.0f; }
            }
        __syncwarp();
    };
    // normal_fullrank::sample: eta ~ N(0, I), zeta = mu + L eta  -> q (lane layout)
    auto draw_sample = [&]() {
        double er[ES];
#pragma unroll
        for (int s = 0; s < ES; s += 2) {
            double n0, n1;
            rng_lane_normal2(rk, RT_VB, tick, (uint32_t)s, n0, n1);
            er[s] = (lane + 32 * s < D) ? n0 : 0.0;
            if (s + 1 < ES) er[s + 1] = (lane + 32 * (s + 1) < D) ? n1 : 0.0;
        }
        ++tick;
#pragma unroll
        for (int s = 0; s < ES; ++s)
            if (lane + 32 * s < DP) etav[lane + 32 * s] = er[s];
        for (int i0 = 0; i0 < D; i0 += 8) {
            // all loads of the 8-row block are issued before the first use (the sweep is DRAM-latency bound)
            double lv[8][ES];
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const int i = i0 + r;
                const double *row = Lm + (long long)i * DP;
#pragma unroll
                for (int s = 0; s < ES; ++s) {
                    const int j = lane + 32 * s;
                    lv[r][s] = (i < D && j <= i) ? row[j] : 0.0;
                }
            }
            double v[8];
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                double part = 0.0;
#pragma unroll
                for (int s = 0; s < ES; ++s) part = fma(lv[r][s], er[s], part);
                v[r] = part;
            }
            const double tot = warp_sum8(v);
            if (lane < 8 && i0 + lane < D) zeta[i0 + lane] = mu[i0 + lane] + tot;
        }
        __syncwarp();
        load_user_vec<CPL>(F, pd, cell, zeta, q);
    };
    // variational += eta_scaled * elbo_grad / (tau + sqrt(history))  with  elbo_grad = (g, g eta^T + diag(1/L_ii));
    // zero == the robust branch of adapt_eta (gradient failed -> elbo_grad.set_to_zero())
    auto sga_step = [&](const int iter, const double eta_cur, const bool zero) {
        const double es = eta_cur / sqrt((double)iter);
        const bool first = iter == 1;
        for (int s = 0; s < ES; ++s) {
            const int i = lane + 32 * s;
            if (i < D) {
                const double gm = zero ? 0.0 : gu[i];
                const double h = first ? hmu[i] + gm * gm : 0.9 * gm * gm + 0.1 * hmu[i];
                hmu[i] = h;
                mu[i] += es * gm / (1.0 + sqrt(h));
            }
        }
        double er[ES];
#pragma unroll
        for (int s = 0; s < ES; ++s) er[s] = (lane + 32 * s < DP) ? etav[lane + 32 * s] : 0.0;
        const float esf = (float)es;
        constexpr int RB = 4;               // rows per block: RB * ES loads of L and of its history in flight per lane
        for (int i0 = 0; i0 < D; i0 += RB) {
            double lv[RB][ES], gi[RB];
            float hv[RB][ES];
#pragma unroll
            for (int r = 0; r < RB; ++r) {
                const int i = i0 + r;
                gi[r] = (i < D && !zero) ? gu[i] : 0.0;
                const double *row = Lm + (long long)i * DP;
                const float *hrow = Hm + (long long)i * DP;
#pragma unroll
                for (int s = 0; s < ES; ++s) {
                    const int j = lane + 32 * s;
                    const bool in = i < D && j <= i;
                    lv[r][s] = in ? row[j] : 1.0;
                    hv[r][s] = in ? hrow[j] : 0.0f;
                }
            }
#pragma unroll
            for (int r = 0; r < RB; ++r) {
                const int i = i0 + r;
#pragma unroll
                for (int s = 0; s < ES; ++s) {
                    const int j = lane + 32 * s;
                    if (i < D && j <= i) {
                        const double l = lv[r][s];
                        double gl = gi[r] * er[s];
                        if (j == i && !zero) gl += 1.0 / l;
                        const float gf = (float)gl;
                        const float h = first ? fmaf(gf, gf, hv[r][s]) : fmaf(0.9f * gf, gf, 0.1f * hv[r][s]);
                        Hm[(long long)i * DP + j] = h;
                        Lm[(long long)i * DP + j] = fma(gl, (double)__fdividef(esf, 1.0f + sqrtf(h)), l);
                    }
                }
            }
        }
        __syncwarp();
    };
    auto entropy = [&]() -> double {
        double part = 0.0;
        for (int s = 0; s < ES; ++s) {
            const int i = lane + 32 * s;
            if (i < D) { const double t = Lm[(long long)i * DP + i]; if (t != 0.0) part += log(fabs(t)); }
        }
        return warp_sum(part) + 0.5 * D * (1.0 + 1.8378770664093453);   // log(2 pi)
    };
    // The sweeps are inlined ONCE, at the top of the tick: the bookkeeping only raises these flags
    bool need_reset = false, need_sample = false;
    auto start_elbo = [&](int purpose) { efor = purpose; e_done = 0; e_drop = 0; e_sum = 0.0; stage = VS_ELBO; need_sample = true; };
    auto start_grad = [&]() { stage = VS_GRAD; need_sample = true; };
    auto start_main = [&]() {
        need_reset = true;
        in_adapt = false; it = 1; elbo = 0.0; elbo_best = -DBL_MAX; cbv = 0.0; cb_n = 0; cb_head = 0;
        start_grad();
    };
    auto finish = [&]() {
        if (lane == 0) {
            double *o = vp.info + (long long)ip * 8;
            o[0] = status; o[1] = eta; o[2] = iters; o[3] = elbo; o[4] = conv; o[5] = n_grad; o[6] = n_lp; o[7] = 0.0;
            atomicAdd(vp.grad_counter, (unsigned long long)(n_grad + n_lp));
        }
    };
    auto start_draws = [&]() {
        for (int s = 0; s < ES; ++s) {
            const int i = lane + 32 * s;
            if (i < D && vp.mean) vp.mean[(long long)ip * vp.ld_mean + i] = mu[i];
        }
        idraw = 0;
        if (vp.output_samples > 0) { stage = VS_DRAW; need_sample = true; }
        else { finish(); stage = VS_FETCH; }
    };

    for (;;) {
        // tick barrier as in the sampler: free-running warps spent half their time waiting for instructions
        // (no_instruction 5.5 cycles per issue, profiles/r01_advi_v1_summary.txt)
        if (__syncthreads_and(stage == VS_DONE)) break;
        if (stage == VS_FETCH) {
            int item = 0;
            if (lane == 0) item = atomicAdd(vp.work_counter, 1);
            item = __shfl_sync(FULL, item, 0);
            if (item >= vp.n_problems) {
                stage = VS_DONE;
            } else {
                ip = item;
                pd = descs[ip];
                D = pd.D;
                __syncwarp();
                for (int e = lane; e < pd.n_d_smem; e += 32) smem[s_pd + e] = dpool[pd.off_d + e];
                for (int e = lane; e < pd.n_i_smem; e += 32) SMI(s_pi + e) = ipool[pd.off_i + e];
                __syncwarp();
                bind_problem(P, pd, s_pd, s_pi, s_tri);
                make_cells<CPL>(F, P, cell);
                rk.k0 = (uint32_t)vp.seed; rk.k1 = (uint32_t)(vp.seed >> 32); rk.chain = (uint32_t)ip;
                yglob = dpool + pd.off_d + pool_y_off(pd);
                iiglob = ipool + pd.off_i + pool_ii_off(pd);
                cpo_acc = vp.cpo_acc ? vp.cpo_acc + (long long)ip * vp.ld_cpo : nullptr;
                if (cpo_acc) for (int e = lane; e < pd.N; e += 32) cpo_acc[e] = 0.0;
                need_reset = false; need_sample = false;
                tick = 0; init_try = 0; status = 0; conv = 0; iters = 0; n_grad = 0.0; n_lp = 0.0; elbo = 0.0; eta = vp.eta;
                stage = VS_INIT;
            }
        }
        if (stage == VS_DONE) continue;
        if (stage == VS_INIT) {   // services::util::initialize: u ~ U(-R, R)
            double ua, ub;
#pragma unroll
            for (int t = 0; t < CPL; t += 2) {
                rng_u2(rk, RT_INIT, lane, (uint32_t)init_try, t, ua, ub);
                q.z[t] = ((cell[t] >> 17) & 1) ? vp.init_r * (2.0 * ua - 1.0) : 0.0;
                if (t + 1 < CPL) q.z[t + 1] = ((cell[t + 1] >> 17) & 1) ? vp.init_r * (2.0 * ub - 1.0) : 0.0;
            }
            rng_u2(rk, RT_INIT, lane, (uint32_t)init_try, CPL + 1, ua, ub);
            q.s = lane_active ? vp.init_r * (2.0 * ua - 1.0) : 0.0;
        }
        if (need_reset) { reset_q(); need_reset = false; }
        if (need_sample) { draw_sample(); need_sample = false; }
        // ---- the one evaluation of this tick ----
        GQ gq;
        gq.on = (stage == VS_DRAW);
        gq.so = s_gq; gq.fwd = (stage == VS_ELBO);
        if (gq.on) {
            const long long rowi = (long long)ip * vp.output_samples + idraw;
            double uz[4];
            rng_u2(rk, RT_GQ, 0xFFu, (uint32_t)idraw, 0, uz[0], uz[1]);
            rng_u2(rk, RT_GQ, 0xFFu, (uint32_t)idraw, 1, uz[2], uz[3]);
            gq_publish(gq.so, pd.D, yglob, iiglob, uz, vp.draws ? vp.draws + rowi * vp.ld_draws : nullptr,
                       vp.qdraws ? vp.qdraws + rowi * 12 : nullptr, cpo_acc);
            if (vp.udraws) store_user_vec<CPL>(F, pd, cell, q, vp.udraws + rowi * vp.ld_u);
        }
        double lp;
        const bool okv = warp_logp_grad<CPL, false>(F, P, ws, cell, q, 1, lp, g, nullptr, gq);
        const bool lp_ok = okv && (fabs(lp) < INFINITY);
        // ---- bookkeeping ----
        if (stage == VS_INIT) {
            bool fin = lp_ok && fabs(g.s) < INFINITY;
#pragma unroll
            for (int t = 0; t < CPL; ++t) fin = fin && fabs(g.z[t]) < INFINITY;
            if (__all_sync(FULL, fin)) {
                for (int s = 0; s < ES; ++s) if (lane + 32 * s < DP) init[lane + 32 * s] = 0.0;
                __syncwarp();
                store_user_vec<CPL>(F, pd, cell, q, init);
                __syncwarp();
                need_reset = true;
                if (vp.adapt_engaged) {
                    in_adapt = true; k_eta = 0; elbo_best = -DBL_MAX; eta_best = 0.0;
                    start_elbo(VE_INIT);
                } else {
                    start_main();
                }
            } else if (++init_try >= 100) {
                status = 1; finish(); stage = VS_FETCH;
            }
        } else if (stage == VS_ELBO) {
            n_lp += 1.0;
            bool done = false, failed = false;
            if (lp_ok) { e_sum += lp; done = (++e_done == vp.elbo_samples); }
            else failed = (++e_drop >= vp.elbo_samples);
            if (!done && !failed) {
                need_sample = true;
            } else {
                const double el = failed ? -DBL_MAX : e_sum / vp.elbo_samples + entropy();
                if (efor == VE_INIT) {
                    if (failed) { status = 2; finish(); stage = VS_FETCH; }
                    else { elbo_init = el; it = 1; start_grad(); }
                } else if (efor == VE_ADAPT) {
                    // advi::adapt_eta: stop at the first eta that is worse than the best so far (if that beat the start)
                    const double e_cur = (k_eta == 0) ? 100.0 : (k_eta == 1) ? 10.0 : (k_eta == 2) ? 1.0 : (k_eta == 3) ? 0.1 : 0.01;
                    bool more = true;
                    if (el < elbo_best && elbo_best > elbo_init) {
                        more = false;
                    } else {
                        if (k_eta < 4) { elbo_best = el; eta_best = e_cur; }
                        else if (el > elbo_init) { eta_best = e_cur; more = false; }
                        else { status = 3; }
                    }
                    if (status) { finish(); stage = VS_FETCH; }
                    else if (more) { ++k_eta; need_reset = true; it = 1; start_grad(); }
                    else { eta = eta_best; start_main(); }
                } else {
                    if (failed) { status = 5; finish(); stage = VS_FETCH; }
                    else {
                        // advi::stochastic_gradient_ascent convergence heuristics
                        const double prev = elbo;
                        elbo = el;
                        if (elbo > elbo_best) elbo_best = elbo;
                        const double delta = fabs((prev - elbo) / elbo);
                        if (lane == cb_head) cbv = delta;
                        cb_head = (cb_head + 1) % vp.cb_size;
                        if (cb_n < vp.cb_size) ++cb_n;
                        const bool have = lane < cb_n;
                        const double ave = warp_sum(have ? cbv : 0.0) / cb_n;
                        int rank = 0;   // position of this lane's entry in the sorted buffer (ties by index)
                        for (int o = 0; o < 32; ++o) {
                            const double other = __shfl_sync(FULL, cbv, o);
                            if (o < cb_n && (other < cbv || (other == cbv && o < lane))) ++rank;
                        }
                        const unsigned pick = __ballot_sync(FULL, have && rank == cb_n / 2);
                        const double med = __shfl_sync(FULL, cbv, pick ? __ffs(pick) - 1 : 0);
                        bool more = true;
                        if (ave < vp.tol_rel_obj) { conv |= 1; more = false; }
                        if (med < vp.tol_rel_obj) { conv |= 2; more = false; }
                        if (it > 10 * vp.eval_elbo && (med > 0.5 || ave > 0.5)) conv |= 4;   // "MAY BE DIVERGING": message only
                        if (it == vp.max_iter) { if (more) conv |= 8; more = false; }
                        if (more) { ++it; start_grad(); }
                        else start_draws();
                    }
                }
            }
        } else if (stage == VS_GRAD) {
            n_grad += 1.0;
            bool fin = lp_ok && fabs(g.s) < INFINITY;
#pragma unroll
            for (int t = 0; t < CPL; ++t) fin = fin && fabs(g.z[t]) < INFINITY;
            fin = __all_sync(FULL, fin);
            if (!fin && !in_adapt) {
                status = 4; finish(); stage = VS_FETCH;     // calc_grad throws outside the adaptation phase
            } else {
                if (fin) { store_user_vec<CPL>(F, pd, cell, g, gu); __syncwarp(); }
                const double e_cur = in_adapt ? ((k_eta == 0) ? 100.0 : (k_eta == 1) ? 10.0 : (k_eta == 2) ? 1.0 : (k_eta == 3) ? 0.1 : 0.01) : eta;
                sga_step(it, e_cur, !fin);
                if (in_adapt) {
                    if (it == vp.adapt_iter) start_elbo(VE_ADAPT);
                    else { ++it; start_grad(); }
                } else {
                    iters = it;
                    if (it % vp.eval_elbo == 0) start_elbo(VE_MAIN);
                    else if (it == vp.max_iter) { conv |= 8; start_draws(); }
                    else { ++it; start_grad(); }
                }
            }
        } else if (stage == VS_DRAW) {
            if (vp.qdraws && lane == 0) { double *qo = vp.qdraws + ((long long)ip * vp.output_samples + idraw) * 12; qo[10] = lp; qo[11] = 0.0; }
            if (++idraw == vp.output_samples) { finish(); stage = VS_FETCH; }
            else need_sample = true;
        }
    }
}

}  // namespace bsyn
