// bsyn_d8.cu -- the package-default model options fixed at compile time (gp_grid, Matern 3/2, heteroscedastic Normal
// likelihood, lower asymptotes estimated, default priors), grid sizes at run time, 8 GP cell(s) per lane.
#include "bsyn_kernels.cuh"
using SpecDefault = bsyn::SpecCT</*model*/ 0, /*kernel*/ 2, /*nu*/ 2, /*het*/ 1, /*robust*/ 0, /*est_la*/ 1, /*pcprior*/ 0, /*NF*/ 0>;
const bsyn::VariantOps *bsyn_variant_default_8() { return bsyn::VariantImpl<8, SpecDefault>::ops(); }
