// bsyn_kernels_decl.h -- launch tables of the compiled kernel variants (defined in bsyn_v*.cu)
#pragma once
#include "bsyn_advi.cuh"

namespace bsyn {
// launch table of one compiled variant
struct VariantOps {
    int cpl, ws_doubles, tr, nm;
    cudaError_t (*logp)(bool dbg, int grid, size_t smem_bytes, cudaStream_t st, const Flags &F, const ProbDesc *descs,
                        const double *dpool, const int *ipool, const int *tri, SmemLayout sl, int B, const int *idx,
                        const double *u, int ldu, int jac, double *lp, double *grad, int *status, double *dbgbuf,
                        int dbg_stride);
    cudaError_t (*write_array)(int grid, size_t smem_bytes, cudaStream_t st, const Flags &F, const ProbDesc *descs,
                               const double *dpool, const int *ipool, const int *tri, SmemLayout sl, int B,
                               const int *idx, const double *u, int ldu, double *out, int ldo, unsigned long long seed,
                               int *status);
    cudaError_t (*nuts)(int nw, int tick, int grid, size_t smem_bytes, cudaStream_t st, const Flags &F, const ProbDesc *descs,
                        const double *dpool, const int *ipool, const int *tri, const SamplerParams &sp);
    cudaError_t (*nuts_occupancy)(int nw, int metric, int tick, size_t smem_bytes, int *per_sm);
    cudaError_t (*advi)(int grid, size_t smem_bytes, cudaStream_t st, const Flags &F, const ProbDesc *descs,
                        const double *dpool, const int *ipool, const int *tri, const AdviParams &vp);
    cudaError_t (*advi_occupancy)(size_t smem_bytes, int *per_sm);
};

}  // namespace bsyn

const bsyn::VariantOps *bsyn_variant_1();
const bsyn::VariantOps *bsyn_variant_2();
const bsyn::VariantOps *bsyn_variant_4();
const bsyn::VariantOps *bsyn_variant_8();
// compile-time specialisations (bsyn_common.cuh: SpecCT): the package-default options on a fixed 10 x 10 dose grid
// (cfg5, the bench workload) for the gp_grid model and for its H0, the nointeraction model
// the package-default options (gp_grid, Matern 3/2, heteroscedastic Normal likelihood, lower asymptotes, default priors)
// with runtime grid sizes, one per cells-per-lane class (bsyn_d.cu)
const bsyn::VariantOps *bsyn_variant_default(int cpl);
const bsyn::VariantOps *bsyn_variant_gp10();
const bsyn::VariantOps *bsyn_variant_h0_10();
