// bsyn_kernels.cuh -- the __global__ entry points around the warp-cooperative device functions, and the
// per-variant launch table (one translation unit per cells-per-lane variant, see bsyn_v*.cu).
#pragma once
#include "bsyn_nuts.cuh"
#include "bsyn_kernels_decl.h"

namespace bsyn {

// stand-alone log_prob / grad_log_prob and write_array kernels: one warp per evaluation point
template <int CPL, bool DBG, class SP = SpecRT>
__global__ void __launch_bounds__(32 * WPC) logp_kernel(const Flags F, const ProbDesc *__restrict__ descs,
                                                        const double *__restrict__ dpool, const int *__restrict__ ipool,
                                                        const int *__restrict__ tri_tab, const SmemLayout sl, const int B,
                                                        const int *__restrict__ prob_idx, const double *__restrict__ u,
                                                        const int ldu, const int jacobian, double *__restrict__ lp,
                                                        double *__restrict__ grad, int *__restrict__ status,
                                                        double *__restrict__ dbg, const int dbg_stride)
{
    using C = Cfg<CPL, SP>;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int s_tri = 2 * sl.tri_off;
    for (int e = threadIdx.x; e < C::TR; e += blockDim.x) SMI(s_tri + e) = tri_tab[e];
    __syncthreads();
    const int s_pd = sl.pd_off + warp * sl.pd_stride;
    const int s_pi = 2 * (sl.pi_off + warp * sl.pi_stride);
    const int ws = sl.ws_off + warp * C::WS_DOUBLES;
    for (int e = lane; e < C::WS_DOUBLES; e += 32) smem[ws + e] = 0.0;   // the factor's upper triangle must stay zero
    __syncwarp();
    int loaded = -1;
    ProbS P;
    int cell[CPL];
    for (int pt = blockIdx.x * WPC + warp; pt < B; pt += gridDim.x * WPC) {
        const int ip = prob_idx[pt];
        const ProbDesc pd = descs[ip];
        if (ip != loaded) {
            __syncwarp();
            for (int e = lane; e < pd.n_d_smem; e += 32) smem[s_pd + e] = dpool[pd.off_d + e];
            for (int e = lane; e < pd.n_i_smem; e += 32) SMI(s_pi + e) = ipool[pd.off_i + e];
            __syncwarp();
            bind_problem(P, pd, s_pd, s_pi, s_tri);
            make_cells<CPL, SP>(F, P, cell);
            loaded = ip;
        }
        WVec<CPL> q, g;
        load_user_vec<CPL>(F, pd, cell, u + (size_t)pt * ldu, q);
        double lpv;
        GQ gq{};
        gq.on = false; gq.so = 0; gq.fwd = (grad == nullptr) && !DBG;   // value-only calls (log_prob(upar, jacobian, gradient = FALSE)) skip the reverse sweep
        const bool ok = warp_logp_grad<CPL, DBG, SP>(F, P, ws, cell, q, jacobian, lpv, g, DBG ? dbg + (size_t)pt * dbg_stride : nullptr, gq);
        if (!ok) {
            lpv = -INFINITY;
#pragma unroll
            for (int t = 0; t < CPL; ++t) g.z[t] = 0.0;
            g.s = 0.0;
        }
        if (lane == 0) { lp[pt] = lpv; status[pt] = ok ? 0 : 1; }
        if (grad) store_user_vec<CPL>(F, pd, cell, g, grad + (size_t)pt * ldu);
    }
}

template <int CPL>
__global__ void __launch_bounds__(32 * WPC) write_array_kernel(const Flags F, const ProbDesc *__restrict__ descs,
                                                               const double *__restrict__ dpool, const int *__restrict__ ipool,
                                                               const int *__restrict__ tri_tab, const SmemLayout sl, const int B,
                                                               const int *__restrict__ prob_idx, const double *__restrict__ u,
                                                               const int ldu, double *__restrict__ out, const int ldo,
                                                               const unsigned long long seed, int *__restrict__ status)
{
    using C = Cfg<CPL>;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int s_tri = 2 * sl.tri_off;
    for (int e = threadIdx.x; e < C::TR; e += blockDim.x) SMI(s_tri + e) = tri_tab[e];
    __syncthreads();
    const int s_pd = sl.pd_off + warp * sl.pd_stride;
    const int s_pi = 2 * (sl.pi_off + warp * sl.pi_stride);
    const int ws = sl.ws_off + warp * C::WS_DOUBLES;
    for (int e = lane; e < C::WS_DOUBLES; e += 32) smem[ws + e] = 0.0;   // the factor's upper triangle must stay zero
    __syncwarp();
    int loaded = -1;
    ProbS P;
    int cell[CPL];
    for (int pt = blockIdx.x * WPC + warp; pt < B; pt += gridDim.x * WPC) {
        const int ip = prob_idx[pt];
        const ProbDesc pd = descs[ip];
        if (ip != loaded) {
            __syncwarp();
            for (int e = lane; e < pd.n_d_smem; e += 32) smem[s_pd + e] = dpool[pd.off_d + e];
            for (int e = lane; e < pd.n_i_smem; e += 32) SMI(s_pi + e) = ipool[pd.off_i + e];
            __syncwarp();
            bind_problem(P, pd, s_pd, s_pi, s_tri);
            make_cells<CPL>(F, P, cell);
            loaded = ip;
        }
        WVec<CPL> q;
        load_user_vec<CPL>(F, pd, cell, u + (size_t)pt * ldu, q);
        RngKey rk;
        rk.k0 = (uint32_t)seed; rk.k1 = (uint32_t)(seed >> 32); rk.chain = (uint32_t)pt;
        GQ gq{};
        gq.on = true; gq.so = sl.park_off + warp * PARK_DOUBLES + PARK_GQ; gq.fwd = false;
        double uz[4];
        rng_u2(rk, RT_GQ, 0xFFu, 0, 0, uz[0], uz[1]);
        rng_u2(rk, RT_GQ, 0xFFu, 0, 1, uz[2], uz[3]);
        __syncwarp();
        gq_publish(gq.so, pd.D, dpool + pd.off_d + pool_y_off(pd), ipool + pd.off_i + pool_ii_off(pd), uz,
                   out + (size_t)pt * ldo, nullptr, nullptr);
        WVec<CPL> gtmp;
        double lpv;
        const bool ok = warp_logp_grad<CPL, false>(F, P, ws, cell, q, 1, lpv, gtmp, nullptr, gq);
        if (lane == 0) status[pt] = ok ? 0 : 1;
    }
}


template <int CPL, class SP = SpecRT> struct VariantImpl {
    static cudaError_t logp(bool dbg, int grid, size_t bytes, cudaStream_t st, const Flags &F, const ProbDesc *descs,
                            const double *dpool, const int *ipool, const int *tri, SmemLayout sl, int B, const int *idx,
                            const double *u, int ldu, int jac, double *lp, double *grad, int *status, double *dbgbuf,
                            int dbg_stride)
    {
        cudaError_t e;
        if (dbg) {
            e = cudaFuncSetAttribute(logp_kernel<CPL, true, SP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
            if (e != cudaSuccess) return e;
            logp_kernel<CPL, true, SP><<<grid, 32 * WPC, bytes, st>>>(F, descs, dpool, ipool, tri, sl, B, idx, u, ldu, jac, lp, grad,
                                                                    status, dbgbuf, dbg_stride);
        } else {
            e = cudaFuncSetAttribute(logp_kernel<CPL, false, SP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
            if (e != cudaSuccess) return e;
            logp_kernel<CPL, false, SP><<<grid, 32 * WPC, bytes, st>>>(F, descs, dpool, ipool, tri, sl, B, idx, u, ldu, jac, lp, grad,
                                                                     status, nullptr, 0);
        }
        return cudaGetLastError();
    }
    static cudaError_t write_array(int grid, size_t bytes, cudaStream_t st, const Flags &F, const ProbDesc *descs,
                                   const double *dpool, const int *ipool, const int *tri, SmemLayout sl, int B,
                                   const int *idx, const double *u, int ldu, double *out, int ldo,
                                   unsigned long long seed, int *status)
    {
        cudaError_t e = cudaFuncSetAttribute(write_array_kernel<CPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) return e;
        write_array_kernel<CPL><<<grid, 32 * WPC, bytes, st>>>(F, descs, dpool, ipool, tri, sl, B, idx, u, ldu, out, ldo, seed, status);
        return cudaGetLastError();
    }
    // sampler instantiations <warps per CTA, register cap, dense metric, spec, tick barrier>.
    // SpecRT: 12 warps per SM as 1 x 12, 2 x 6, 3 x 4 or 12 x 1 CTAs for diag_e; dense_e (robust models,
    // R/bayesynergy.R:112-115) as 1 x 12 or 3 x 4.  Specialised (SpecCT) variants: 12 x 168 registers and 16 x 128
    // registers, each with and without the tick barrier (`tick` = 0: free-running warps), diag_e only.
    template <int NW, int MINB, bool DENSE, bool TICK>
    static cudaError_t nuts_launch(int grid, size_t bytes, cudaStream_t st, const Flags &F, const ProbDesc *descs,
                                   const double *dpool, const int *ipool, const int *tri, const SamplerParams &sp)
    {
        nuts_kernel<CPL, NW, MINB, DENSE, SP, TICK><<<grid, 32 * NW, bytes, st>>>(F, descs, dpool, ipool, tri, sp);
        return cudaGetLastError();
    }
    template <int NW, int MINB, bool DENSE, bool TICK> static cudaError_t nuts_occ(size_t bytes, int *per_sm)
    {
        cudaError_t e = cudaFuncSetAttribute(nuts_kernel<CPL, NW, MINB, DENSE, SP, TICK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) return e;
        return cudaOccupancyMaxActiveBlocksPerMultiprocessor(per_sm, nuts_kernel<CPL, NW, MINB, DENSE, SP, TICK>, 32 * NW, bytes);
    }
    // one table for launch and occupancy query: op = 0 launch, 1 occupancy
    static cudaError_t nuts_do(int op, int nw, int metric, int tick, int grid, size_t bytes, cudaStream_t st, const Flags *F,
                               const ProbDesc *descs, const double *dpool, const int *ipool, const int *tri,
                               const SamplerParams *sp, int *per_sm)
    {
#define BSYN_NUTS_CASE(NW_, MINB_, DENSE_, TICK_)                                                                          \
    return op == 0 ? nuts_launch<NW_, MINB_, DENSE_, TICK_>(grid, bytes, st, *F, descs, dpool, ipool, tri, *sp)            \
                   : nuts_occ<NW_, MINB_, DENSE_, TICK_>(bytes, per_sm)
        if constexpr (SP::fixed) {
            if constexpr (SP::robust != 0) {
                // robust models keep their sorted observations in shared memory: on the O'Neil grids 12 warps do not fit,
                // so 3 CTAs of 4 warps are built as well; dense_e is what bayesynergy() selects for robust = TRUE
                if (metric == 1 && nw == 12 && tick) { BSYN_NUTS_CASE(12, 168, true, true); }
                if (metric == 1 && nw == 4 && tick) { BSYN_NUTS_CASE(4, 168, true, true); }
                if (metric == 0 && nw == 4 && tick) { BSYN_NUTS_CASE(4, 168, false, true); }
            }
            if (metric == 1) return cudaErrorInvalidValue;
            if (nw == 12 && tick) { BSYN_NUTS_CASE(12, 168, false, true); }
            if constexpr (SP::NF > 0) {   // the A/B variants are only built for the fixed-grid specialisation
                if (nw == 12 && !tick) { BSYN_NUTS_CASE(12, 168, false, false); }
                if (nw == 14 && tick) { BSYN_NUTS_CASE(14, 144, false, true); }
                if (nw == 16 && tick) { BSYN_NUTS_CASE(16, 128, false, true); }
            }
            return cudaErrorInvalidValue;
        } else {
            if (!tick) return cudaErrorInvalidValue;
            if (metric == 1) {
                if (nw == 12) { BSYN_NUTS_CASE(12, 168, true, true); }
                if (nw == 4) { BSYN_NUTS_CASE(4, 168, true, true); }
                return cudaErrorInvalidValue;
            }
            if (nw == 12) { BSYN_NUTS_CASE(12, 168, false, true); }
            if (nw == 6) { BSYN_NUTS_CASE(6, 168, false, true); }
            if (nw == 4) { BSYN_NUTS_CASE(4, 168, false, true); }
            if (nw == 1) { BSYN_NUTS_CASE(1, 168, false, true); }
            return cudaErrorInvalidValue;
        }
#undef BSYN_NUTS_CASE
    }
    static cudaError_t nuts(int nw, int tick, int grid, size_t bytes, cudaStream_t st, const Flags &F, const ProbDesc *descs,
                            const double *dpool, const int *ipool, const int *tri, const SamplerParams &sp)
    {
        return nuts_do(0, nw, sp.metric, tick, grid, bytes, st, &F, descs, dpool, ipool, tri, &sp, nullptr);
    }
    static cudaError_t nuts_occupancy(int nw, int metric, int tick, size_t bytes, int *per_sm)
    {
        return nuts_do(1, nw, metric, tick, 0, bytes, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, per_sm);
    }
    // full-rank ADVI: ADVI_NW warps (problems) per CTA, one CTA per SM
    static cudaError_t advi(int grid, size_t bytes, cudaStream_t st, const Flags &F, const ProbDesc *descs,
                            const double *dpool, const int *ipool, const int *tri, const AdviParams &vp)
    {
        advi_kernel<CPL, ADVI_NW><<<grid, 32 * ADVI_NW, bytes, st>>>(F, descs, dpool, ipool, tri, vp);
        return cudaGetLastError();
    }
    static cudaError_t advi_occupancy(size_t bytes, int *per_sm)
    {
        cudaError_t e = cudaFuncSetAttribute(advi_kernel<CPL, ADVI_NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) return e;
        return cudaOccupancyMaxActiveBlocksPerMultiprocessor(per_sm, advi_kernel<CPL, ADVI_NW>, 32 * ADVI_NW, bytes);
    }
    static const VariantOps *ops()
    {
        if constexpr (SP::fixed) {   // specialised variants carry the lp+gradient kernel and the sampler only
            static const VariantOps o = {CPL, Cfg<CPL, SP>::WS_DOUBLES, Cfg<CPL, SP>::TR, Cfg<CPL, SP>::NM, &logp, nullptr, &nuts,
                                         &nuts_occupancy, nullptr, nullptr};
            return &o;
        } else {
            static const VariantOps o = {CPL, Cfg<CPL>::WS_DOUBLES, Cfg<CPL>::TR, Cfg<CPL>::NM, &logp, &write_array, &nuts,
                                         &nuts_occupancy, &advi, &advi_occupancy};
            return &o;
        }
    }
};

}  // namespace bsyn
