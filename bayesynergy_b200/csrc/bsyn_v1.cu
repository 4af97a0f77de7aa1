// bsyn_v1.cu -- kernel variant with 1 GP cell(s) per lane (grids up to 32 cells).
#include "bsyn_kernels.cuh"
const bsyn::VariantOps *bsyn_variant_1() { return bsyn::VariantImpl<1>::ops(); }
