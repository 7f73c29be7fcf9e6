// Gallery-search tensor-core pass, precision mode 3 (3xBF16): instantiations of nt_gemm_rowscan_kernel
// with the fused top-k epilogue (see gallery_epi.cuh, nt_gemm.cuh).
#include "gallery_epi.cuh"

namespace dif {
int launch_search_bf16x3(int metric, int ctas, int ares, const CUtensorMap* maps, const GemmShape& shape,
                         const TopkEpi<1>::Params& ep, int n_units, cudaStream_t st) {
  return launch_search_prec<3>(metric, ctas, ares, maps, shape, ep, n_units, st);
}
}  // namespace dif
