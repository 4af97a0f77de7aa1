// bsyn_r4.cu -- cfg3 of BASELINE.json fixed at compile time: robust (LPTN) likelihood + PC prior on the Matern 3/2
// kernel, heteroscedastic, lower asymptotes estimated; grid sizes at run time, 4 GP cells per lane (the 11 x 11 O'Neil
// grids).  Sampler variants: dense_e (what bayesynergy() selects when robust = TRUE, R/bayesynergy.R:112-115) and diag_e.
#include "bsyn_kernels.cuh"
using SpecRobustPC = bsyn::SpecCT</*model*/ 0, /*kernel*/ 2, /*nu*/ 2, /*het*/ 1, /*robust*/ 1, /*est_la*/ 1, /*pcprior*/ 1, /*NF*/ 0>;
const bsyn::VariantOps *bsyn_variant_robust_pc_4() { return bsyn::VariantImpl<4, SpecRobustPC>::ops(); }
