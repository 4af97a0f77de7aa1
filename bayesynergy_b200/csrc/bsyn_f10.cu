// bsyn_f10.cu -- compile-time specialised kernels for the package-default options on a fixed 10 x 10 dose grid:
// gp_grid with the Matern 3/2 kernel, heteroscedastic Normal likelihood, lower asymptotes estimated, default priors
// (the bench workload).
#include "bsyn_kernels.cuh"
using SpecGp10 = bsyn::SpecCT</*model*/ 0, /*kernel*/ 2, /*nu*/ 2, /*het*/ 1, /*robust*/ 0, /*est_la*/ 1, /*pcprior*/ 0, /*NF*/ 10>;
const bsyn::VariantOps *bsyn_variant_gp10() { return bsyn::VariantImpl<4, SpecGp10>::ops(); }
