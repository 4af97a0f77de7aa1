// bsyn_h10.cu -- the nointeraction model (H0 of the Bayes factor) with the package-default options on a fixed 10 x 10
// dose grid, specialised at compile time (see bsyn_f10.cu).
#include "bsyn_kernels.cuh"
using SpecH0_10 = bsyn::SpecCT</*model*/ 1, /*kernel*/ 0, /*nu*/ 0, /*het*/ 1, /*robust*/ 0, /*est_la*/ 1, /*pcprior*/ 0, /*NF*/ 10>;
const bsyn::VariantOps *bsyn_variant_h0_10() { return bsyn::VariantImpl<4, SpecH0_10>::ops(); }
