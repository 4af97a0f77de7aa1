// bsyn_v2.cu -- kernel variant with 2 GP cell(s) per lane (grids up to 64 cells).
#include "bsyn_kernels.cuh"
const bsyn::VariantOps *bsyn_variant_2() { return bsyn::VariantImpl<2>::ops(); }
