// dif_debug_nt_gemm: the tcgen05/TMA row-scan skeleton with a plain "store the tile" epilogue, so the
// tensor-core plumbing can be validated against a CPU matmul independently of any fused epilogue.
#include <algorithm>

#include "dif_canon.cuh"
#include "nt_gemm.cuh"
#include "store_epi.cuh"

namespace dif {

// Epilogue that only folds the tile into one checksum per row: isolates main-loop speed.
struct SumEpi {
  struct Params {
    float* row_sum;  // [rows padded]
  };
  static int smem_bytes(const Params&) { return 16; }
  const Params& p;
  float acc_sum;
  __device__ SumEpi(const Params& pp, uint8_t*, int) : p(pp), acc_sum(0.f) {}
  __device__ void begin_item(int, int, int) { acc_sum = 0.f; }
  __device__ void begin_tile(int) {}
  __device__ void consume(int, const uint32_t (&acc)[32], uint32_t, uint32_t (&)[32]) {
    float m = __uint_as_float(acc[0]);
#pragma unroll
    for (int i = 1; i < 32; ++i) m = fmaxf(m, __uint_as_float(acc[i]));
    acc_sum += m;
  }
  __device__ void end_item(int m_row, int) { atomicAdd(p.row_sum + m_row, acc_sum); }
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
                     int n_splits, int ares, cudaStream_t st) {
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
  StoreEpi::Params ep{C, M, N, N, shape.n_splits, 0};
  const int units = std::max(1, device_sm_count() / CTAS);
  if (ares) {
    if constexpr (PREC != 0 && CTAS == 2) return launch_nt_gemm<PREC, BN, CTAS, 1, StoreEpi>(maps, shape, ep, units, st);
  }
  return launch_nt_gemm<PREC, BN, CTAS, 0, StoreEpi>(maps, shape, ep, units, st);
}

template <int PREC, int CTAS>
static int run_time(const void* a0, const void* a1, const void* b0, const void* b1, int M, int N, int K, float* rs,
                    int n_splits, int ares, cudaStream_t st, int bn = 256) {
  constexpr int BN = 256;
  const bool bf = PREC == 1;
  const int esz = bf ? 2 : 4;
  const uint32_t bcols = 128 / esz;
  CUtensorMap maps[4];
  if (int rc = make_tmap_2d(&maps[0], a0, M, K, (uint64_t)K * esz, GEMM_BM, bcols, bf)) return rc;
  if (int rc = make_tmap_2d(&maps[1], a1, M, K, (uint64_t)K * esz, GEMM_BM, bcols, bf)) return rc;
  if (int rc = make_tmap_2d(&maps[2], b0, N, K, (uint64_t)K * esz, bn / CTAS, bcols, bf)) return rc;
  if (int rc = make_tmap_2d(&maps[3], b1, N, K, (uint64_t)K * esz, bn / CTAS, bcols, bf)) return rc;
  GemmShape shape{};
  shape.bn = bn;
  shape.m_blocks = (M + GEMM_BM * CTAS - 1) / (GEMM_BM * CTAS);
  shape.n_tiles = (N + bn - 1) / bn;
  shape.k_chunks = (K + (int)bcols - 1) / (int)bcols;
  shape.n_splits = std::max(1, std::min(n_splits, shape.n_tiles));
  shape.tiles_per_split = (shape.n_tiles + shape.n_splits - 1) / shape.n_splits;
  shape.n_splits = (shape.n_tiles + shape.tiles_per_split - 1) / shape.tiles_per_split;
  SumEpi::Params ep{rs};
  const int units = std::max(1, device_sm_count() / CTAS);
  if (ares) {
    if constexpr (PREC != 0 && CTAS == 2) return launch_nt_gemm<PREC, BN, CTAS, 1, SumEpi>(maps, shape, ep, units, st);
  }
  return launch_nt_gemm<PREC, BN, CTAS, 0, SumEpi>(maps, shape, ep, units, st);
}

// planes of a matrix for the two-plane modes: TF32 hi + exact residual, or bf16 b0 + bf16(x - b0)
__global__ void split_planes2_kernel(const float* x, int64_t n, float* hi, float* lo, __nv_bfloat16* b0, __nv_bfloat16* b1) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = x[i];
    if (hi) {
      const float h = tf32_round(v);
      hi[i] = h;
      lo[i] = __fsub_rn(v, h);
    } else {
      const __nv_bfloat16 b = __float2bfloat16_rn(v);
      b0[i] = b;
      b1[i] = __float2bfloat16_rn(__fsub_rn(v, __bfloat162float(b)));
    }
  }
}

// C [M][N] = op(A) op(B)^T on CTA pairs with a runtime tile width and either operand MN-major
template <int PREC, int AMN, int BMN, class T>
static int run_layout(const T* a0, const T* a1, const T* b0, const T* b1, int M, int N, int K, float* C, int bn, int n_splits,
                      cudaStream_t st) {
  constexpr int esz = (int)sizeof(T), bf = esz == 2 ? 1 : 0;
  constexpr uint32_t cols = 128 / esz;
  CUtensorMap maps[4];
  if (AMN) {
    if (int rc = make_tmap_2d(&maps[0], a0, K, M, (uint64_t)M * esz, cols, cols, bf, !bf)) return rc;
    if (int rc = make_tmap_2d(&maps[1], a1, K, M, (uint64_t)M * esz, cols, cols, bf, !bf)) return rc;
  } else {
    if (int rc = make_tmap_2d(&maps[0], a0, M, K, (uint64_t)K * esz, GEMM_BM, cols, bf)) return rc;
    if (int rc = make_tmap_2d(&maps[1], a1, M, K, (uint64_t)K * esz, GEMM_BM, cols, bf)) return rc;
  }
  if (BMN) {
    if (int rc = make_tmap_2d(&maps[2], b0, K, N, (uint64_t)N * esz, cols, cols, bf, !bf)) return rc;
    if (int rc = make_tmap_2d(&maps[3], b1, K, N, (uint64_t)N * esz, cols, cols, bf, !bf)) return rc;
  } else {
    if (int rc = make_tmap_2d(&maps[2], b0, N, K, (uint64_t)K * esz, bn / 2, cols, bf)) return rc;
    if (int rc = make_tmap_2d(&maps[3], b1, N, K, (uint64_t)K * esz, bn / 2, cols, bf)) return rc;
  }
  GemmShape shape{};
  shape.bn = bn;
  shape.m_blocks = (M + GEMM_BM * 2 - 1) / (GEMM_BM * 2);
  shape.n_tiles = (N + bn - 1) / bn;
  shape.k_chunks = (K + (int)cols - 1) / (int)cols;
  shape.n_splits = std::max(1, std::min(n_splits, shape.n_tiles));
  shape.tiles_per_split = (shape.n_tiles + shape.n_splits - 1) / shape.n_splits;
  shape.n_splits = (shape.n_tiles + shape.tiles_per_split - 1) / shape.tiles_per_split;
  StoreEpi::Params ep{C, M, N, N, shape.n_splits, 0};
  return launch_nt_gemm<PREC, 256, 2, 0, StoreEpi, AMN, BMN>(maps, shape, ep, std::max(1, device_sm_count() / 2), st);
}

}  // namespace dif

using namespace dif;

// diagnostic for the operand-major / tile-width options of the skeleton (tests/test_gemm_gpu.py):
// C [M][N] = A B^T where A is given as [M][K] (a_mn = 0) or [K][M] (a_mn = 1) and B as [N][K] or [K][N];
// precision DIF_PREC_TF32X3 | DIF_PREC_BF16X3, CTA pairs, tile width bn (multiple of 32; of 64 | 128 with b_mn).
extern "C" int dif_debug_gemm_layout(const float* A, const float* B, int M, int N, int K, float* C, int precision, int a_mn,
                                     int b_mn, int bn, int n_splits, void* stream) {
  DIF_REQUIRE(A && B && C && M > 0 && N > 0 && K > 0, DIF_ERR_INVALID, "dif_debug_gemm_layout: bad shape");
  DIF_REQUIRE(precision == DIF_PREC_TF32X3 || precision == DIF_PREC_BF16X3, DIF_ERR_INVALID, "precision 0 or 3");
  const int al = precision == DIF_PREC_BF16X3 ? 8 : 4;
  DIF_REQUIRE((a_mn ? M : K) % al == 0 && (b_mn ? N : K) % al == 0, DIF_ERR_INVALID, "row pitches must be 16-byte multiples");
  DIF_REQUIRE(device_sm_count() > 0, DIF_ERR_STATE, "dif_init has not succeeded on this process");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t na = (size_t)M * K, nb = (size_t)N * K;
  void *a0 = nullptr, *a1 = nullptr, *b0 = nullptr, *b1 = nullptr;
  auto cleanup = [&] { cudaFree(a0); cudaFree(a1); cudaFree(b0); cudaFree(b1); };
  const size_t esz = precision == DIF_PREC_BF16X3 ? 2 : 4;
  if (cudaMalloc(&a0, na * esz) || cudaMalloc(&a1, na * esz) || cudaMalloc(&b0, nb * esz) || cudaMalloc(&b1, nb * esz)) {
    cleanup();
    set_error("dif_debug_gemm_layout: allocation failed");
    return DIF_ERR_CUDA;
  }
  int rc = DIF_OK;
  if (precision == DIF_PREC_TF32X3) {
    split_planes2_kernel<<<148 * 4, 256, 0, st>>>(A, (int64_t)na, (float*)a0, (float*)a1, nullptr, nullptr);
    split_planes2_kernel<<<148 * 4, 256, 0, st>>>(B, (int64_t)nb, (float*)b0, (float*)b1, nullptr, nullptr);
    const float *p0 = (const float*)a0, *p1 = (const float*)a1, *q0 = (const float*)b0, *q1 = (const float*)b1;
    rc = a_mn ? (b_mn ? run_layout<0, 1, 1>(p0, p1, q0, q1, M, N, K, C, bn, n_splits, st)
                      : run_layout<0, 1, 0>(p0, p1, q0, q1, M, N, K, C, bn, n_splits, st))
              : (b_mn ? run_layout<0, 0, 1>(p0, p1, q0, q1, M, N, K, C, bn, n_splits, st)
                      : run_layout<0, 0, 0>(p0, p1, q0, q1, M, N, K, C, bn, n_splits, st));
  } else {
    split_planes2_kernel<<<148 * 4, 256, 0, st>>>(A, (int64_t)na, nullptr, nullptr, (__nv_bfloat16*)a0, (__nv_bfloat16*)a1);
    split_planes2_kernel<<<148 * 4, 256, 0, st>>>(B, (int64_t)nb, nullptr, nullptr, (__nv_bfloat16*)b0, (__nv_bfloat16*)b1);
    const __nv_bfloat16 *p0 = (const __nv_bfloat16*)a0, *p1 = (const __nv_bfloat16*)a1, *q0 = (const __nv_bfloat16*)b0,
                        *q1 = (const __nv_bfloat16*)b1;
    rc = a_mn ? (b_mn ? run_layout<3, 1, 1>(p0, p1, q0, q1, M, N, K, C, bn, n_splits, st)
                      : run_layout<3, 1, 0>(p0, p1, q0, q1, M, N, K, C, bn, n_splits, st))
              : (b_mn ? run_layout<3, 0, 1>(p0, p1, q0, q1, M, N, K, C, bn, n_splits, st)
                      : run_layout<3, 0, 0>(p0, p1, q0, q1, M, N, K, C, bn, n_splits, st));
  }
  count_launch(2);
  cudaError_t e = cudaStreamSynchronize(st);
  cleanup();
  if (rc) return rc;
  if (e != cudaSuccess) {
    set_error("dif_debug_gemm_layout: %s", cudaGetErrorString(e));
    return DIF_ERR_CUDA;
  }
  return DIF_OK;
}

// diagnostic: time the NT-GEMM main loop (checksum epilogue, nothing written per element) on synthetic
// operands; ms_out = average over `iters` launches (CUDA events on `stream`).
extern "C" int dif_debug_gemm_time(int M, int N, int K, int precision, int ctas, int n_splits, int iters,
                                   float* ms_out) {
  const int ares = (ctas & 16) ? 1 : 0;
  const int bn = (ctas >> 8) ? (ctas >> 8) : 256;   // tile width rides in the upper bits of `ctas` (0 = 256)
  ctas &= 15;
  DIF_REQUIRE(M > 0 && N > 0 && K >= 64 && K % 8 == 0 && ms_out && iters > 0, DIF_ERR_INVALID, "dif_debug_gemm_time: bad shape");
  DIF_REQUIRE(precision >= 0 && precision <= 2 && (ctas == 1 || ctas == 2), DIF_ERR_INVALID, "bad precision/ctas");
  DIF_REQUIRE(!ares || (ctas == 2 && precision != DIF_PREC_TF32X3), DIF_ERR_INVALID, "resident-A needs ctas 2, bf16/tf32");
  DIF_REQUIRE(device_sm_count() > 0, DIF_ERR_STATE, "dif_init has not succeeded on this process");
  const size_t na = (size_t)M * K, nb = (size_t)N * K;
  float *a = nullptr, *b = nullptr, *rs = nullptr;
  __nv_bfloat16 *ab = nullptr, *bb = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  auto cleanup = [&] {
    cudaFree(a); cudaFree(b); cudaFree(rs); cudaFree(ab); cudaFree(bb);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
  };
  if (cudaMalloc((void**)&a, na * 4) || cudaMalloc((void**)&b, nb * 4) || cudaMalloc((void**)&rs, ((size_t)M + 512) * 4) ||
      cudaMalloc((void**)&ab, na * 2) || cudaMalloc((void**)&bb, nb * 2) || cudaEventCreate(&e0) || cudaEventCreate(&e1)) {
    cleanup();
    set_error("dif_debug_gemm_time: allocation failed");
    return DIF_ERR_CUDA;
  }
  cudaMemset(rs, 0, ((size_t)M + 512) * 4);
  if (int rc = dif_synth_fill(a, 11, 0, nullptr, M, K, nullptr)) { cleanup(); return rc; }
  if (int rc = dif_synth_fill(b, 12, 0, nullptr, N, K, nullptr)) { cleanup(); return rc; }
  split_planes_kernel<<<148 * 4, 256>>>(a, (int64_t)na, nullptr, nullptr, ab);
  split_planes_kernel<<<148 * 4, 256>>>(b, (int64_t)nb, nullptr, nullptr, bb);
  int rc = DIF_OK;
  for (int i = 0; i < iters + 2 && rc == DIF_OK; ++i) {
    if (i == 2) cudaEventRecord(e0, nullptr);
    if (precision == DIF_PREC_BF16)
      rc = ctas == 2 ? run_time<1, 2>(ab, ab, bb, bb, M, N, K, rs, n_splits, ares, nullptr, bn)
                     : run_time<1, 1>(ab, ab, bb, bb, M, N, K, rs, n_splits, ares, nullptr, bn);
    else if (precision == DIF_PREC_TF32X1)
      rc = ctas == 2 ? run_time<2, 2>(a, a, b, b, M, N, K, rs, n_splits, ares, nullptr, bn)
                     : run_time<2, 1>(a, a, b, b, M, N, K, rs, n_splits, ares, nullptr, bn);
    else  // 3xTF32 timing uses the same plane for hi and lo: identical instruction stream and traffic
      rc = ctas == 2 ? run_time<0, 2>(a, a, b, b, M, N, K, rs, n_splits, ares, nullptr, bn)
                     : run_time<0, 1>(a, a, b, b, M, N, K, rs, n_splits, ares, nullptr, bn);
  }
  cudaEventRecord(e1, nullptr);
  cudaError_t e = cudaDeviceSynchronize();
  float ms = 0.f;
  if (e == cudaSuccess) cudaEventElapsedTime(&ms, e0, e1);
  *ms_out = ms / (float)iters;
  cleanup();
  if (rc) return rc;
  if (e != cudaSuccess) {
    set_error("dif_debug_gemm_time: %s", cudaGetErrorString(e));
    return DIF_ERR_CUDA;
  }
  return DIF_OK;
}

extern "C" int dif_debug_nt_gemm(const float* A, const float* B, int M, int N, int K, float* C, int precision, int ctas,
                                 int n_splits, void* stream) {
  DIF_REQUIRE(A && B && C && M > 0 && N > 0 && K >= 32 && K % 4 == 0, DIF_ERR_INVALID, "dif_debug_nt_gemm: bad shape");
  // ctas = 1 | 2 selects the CTA-pair mode; ctas + 16 additionally asks for the resident-A schedule
  const int ares = (ctas & 16) ? 1 : 0;
  ctas &= 15;
  DIF_REQUIRE(precision >= 0 && precision <= 2 && (ctas == 1 || ctas == 2), DIF_ERR_INVALID, "bad precision/ctas");
  DIF_REQUIRE(!ares || (ctas == 2 && precision != DIF_PREC_TF32X3), DIF_ERR_INVALID, "resident-A needs ctas 2, bf16/tf32");
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
    rc = ctas == 2 ? run_debug<0, 2>(ah, al, bh, bl, M, N, K, C, n_splits, ares, st)
                   : run_debug<0, 1>(ah, al, bh, bl, M, N, K, C, n_splits, ares, st);
  } else if (precision == DIF_PREC_BF16) {
    if (cudaMalloc((void**)&ab, na * 2) || cudaMalloc((void**)&bb, nb * 2)) {
      cleanup();
      set_error("dif_debug_nt_gemm: allocation failed");
      return DIF_ERR_CUDA;
    }
    split_planes_kernel<<<148 * 4, 256, 0, st>>>(A, (int64_t)na, nullptr, nullptr, ab);
    split_planes_kernel<<<148 * 4, 256, 0, st>>>(B, (int64_t)nb, nullptr, nullptr, bb);
    count_launch(2);
    rc = ctas == 2 ? run_debug<1, 2>(ab, ab, bb, bb, M, N, K, C, n_splits, ares, st)
                   : run_debug<1, 1>(ab, ab, bb, bb, M, N, K, C, n_splits, ares, st);
  } else {
    rc = ctas == 2 ? run_debug<2, 2>(A, A, B, B, M, N, K, C, n_splits, ares, st)
                   : run_debug<2, 1>(A, A, B, B, M, N, K, C, n_splits, ares, st);
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
