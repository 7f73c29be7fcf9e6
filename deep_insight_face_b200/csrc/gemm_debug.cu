// dif_debug_nt_gemm: the tcgen05/TMA row-scan skeleton with a plain "store the tile" epilogue, so the
// tensor-core plumbing can be validated against a CPU matmul independently of any fused epilogue.
#include <algorithm>

#include "dif_canon.cuh"
#include "nt_gemm.cuh"

namespace dif {

struct StoreEpi {
  struct Params {
    float* C;
    int M, N;
  };
  static constexpr int kSmemBytes = 16;
  const Params& p;
  int m_row;
  __device__ StoreEpi(const Params& pp, uint8_t*, int) : p(pp), m_row(0) {}
  __device__ void begin_item(int m, int, int) { m_row = m; }
  __device__ void consume(int col0, const uint32_t (&acc)[32]) {
    if (m_row >= p.M) return;
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (col0 + i < p.N) p.C[(size_t)m_row * p.N + col0 + i] = __uint_as_float(acc[i]);
  }
  __device__ void end_item(int, int) {}
};

__global__ void split_planes_kernel(const float* x, int64_t n, float* hi, float* lo, __nv_bfloat16* bf) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = x[i];
    if (hi) {
      const float h = tf32_round(v);
      hi[i] = h;
      lo[i] = __fsub_rn(v, h);
    }
    if (bf) bf[i] = __float2bfloat16_rn(v);
  }
}

template <int PREC, int CTAS>
static int run_debug(const void* a0, const void* a1, const void* b0, const void* b1, int M, int N, int K, float* C,
                     int n_splits, cudaStream_t st) {
  constexpr int BN = 256;
  const bool bf = PREC == 1;
  const int esz = bf ? 2 : 4;
  const uint32_t bcols = 128 / esz;
  CUtensorMap maps[4];
  if (int rc = make_tmap_2d(&maps[0], a0, M, K, (uint64_t)K * esz, GEMM_BM, bcols, bf)) return rc;
  if (int rc = make_tmap_2d(&maps[1], a1, M, K, (uint64_t)K * esz, GEMM_BM, bcols, bf)) return rc;
  if (int rc = make_tmap_2d(&maps[2], b0, N, K, (uint64_t)K * esz, BN / CTAS, bcols, bf)) return rc;
  if (int rc = make_tmap_2d(&maps[3], b1, N, K, (uint64_t)K * esz, BN / CTAS, bcols, bf)) return rc;
  GemmShape shape{};
  shape.m_blocks = (M + GEMM_BM * CTAS - 1) / (GEMM_BM * CTAS);
  shape.n_tiles = (N + BN - 1) / BN;
  shape.k_chunks = (K + (int)bcols - 1) / (int)bcols;
  shape.n_splits = std::max(1, std::min(n_splits, shape.n_tiles));
  shape.tiles_per_split = (shape.n_tiles + shape.n_splits - 1) / shape.n_splits;
  shape.n_splits = (shape.n_tiles + shape.tiles_per_split - 1) / shape.tiles_per_split;
  StoreEpi::Params ep{C, M, N};
  const int units = std::max(1, device_sm_count() / CTAS);
  return launch_nt_gemm<PREC, BN, CTAS, StoreEpi>(maps, shape, ep, units, st);
}

}  // namespace dif

using namespace dif;

extern "C" int dif_debug_nt_gemm(const float* A, const float* B, int M, int N, int K, float* C, int precision, int ctas,
                                 int n_splits, void* stream) {
  DIF_REQUIRE(A && B && C && M > 0 && N > 0 && K >= 32 && K % 4 == 0, DIF_ERR_INVALID, "dif_debug_nt_gemm: bad shape");
  DIF_REQUIRE(precision >= 0 && precision <= 2 && (ctas == 1 || ctas == 2), DIF_ERR_INVALID, "bad precision/ctas");
  DIF_REQUIRE(precision != DIF_PREC_BF16 || (K % 8 == 0 && K >= 64), DIF_ERR_INVALID, "bf16 needs K %% 8 == 0, K >= 64");
  DIF_REQUIRE(device_sm_count() > 0, DIF_ERR_STATE, "dif_init has not succeeded on this process");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float *ah = nullptr, *al = nullptr, *bh = nullptr, *bl = nullptr;
  __nv_bfloat16 *ab = nullptr, *bb = nullptr;
  const size_t na = (size_t)M * K, nb = (size_t)N * K;
  int rc = DIF_OK;
  auto cleanup = [&] { cudaFree(ah); cudaFree(al); cudaFree(bh); cudaFree(bl); cudaFree(ab); cudaFree(bb); };
  if (precision == DIF_PREC_TF32X3) {
    if (cudaMalloc((void**)&ah, na * 4) || cudaMalloc((void**)&al, na * 4) || cudaMalloc((void**)&bh, nb * 4) ||
        cudaMalloc((void**)&bl, nb * 4)) {
      cleanup();
      set_error("dif_debug_nt_gemm: allocation failed");
      return DIF_ERR_CUDA;
    }
    split_planes_kernel<<<148 * 4, 256, 0, st>>>(A, (int64_t)na, ah, al, nullptr);
    split_planes_kernel<<<148 * 4, 256, 0, st>>>(B, (int64_t)nb, bh, bl, nullptr);
    count_launch(2);
    rc = ctas == 2 ? run_debug<0, 2>(ah, al, bh, bl, M, N, K, C, n_splits, st)
                   : run_debug<0, 1>(ah, al, bh, bl, M, N, K, C, n_splits, st);
  } else if (precision == DIF_PREC_BF16) {
    if (cudaMalloc((void**)&ab, na * 2) || cudaMalloc((void**)&bb, nb * 2)) {
      cleanup();
      set_error("dif_debug_nt_gemm: allocation failed");
      return DIF_ERR_CUDA;
    }
    split_planes_kernel<<<148 * 4, 256, 0, st>>>(A, (int64_t)na, nullptr, nullptr, ab);
    split_planes_kernel<<<148 * 4, 256, 0, st>>>(B, (int64_t)nb, nullptr, nullptr, bb);
    count_launch(2);
    rc = ctas == 2 ? run_debug<1, 2>(ab, ab, bb, bb, M, N, K, C, n_splits, st)
                   : run_debug<1, 1>(ab, ab, bb, bb, M, N, K, C, n_splits, st);
  } else {
    rc = ctas == 2 ? run_debug<2, 2>(A, A, B, B, M, N, K, C, n_splits, st)
                   : run_debug<2, 1>(A, A, B, B, M, N, K, C, n_splits, st);
  }
  cudaError_t e = cudaStreamSynchronize(st);
  cleanup();
  if (rc) return rc;
  if (e != cudaSuccess) {
    set_error("dif_debug_nt_gemm: %s", cudaGetErrorString(e));
    return DIF_ERR_CUDA;
  }
  return DIF_OK;
}
