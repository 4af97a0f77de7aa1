// bsyn_v4.cu -- kernel variant with 4 GP cell(s) per lane (grids up to 128 cells).
#include "bsyn_kernels.cuh"
const bsyn::VariantOps *bsyn_variant_4() { return bsyn::VariantImpl<4>::ops(); }
