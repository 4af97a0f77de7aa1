// bsyn_v8.cu -- kernel variant with 8 GP cell(s) per lane (grids up to 256 cells).
#include "bsyn_kernels.cuh"
const bsyn::VariantOps *bsyn_variant_8() { return bsyn::VariantImpl<8>::ops(); }
