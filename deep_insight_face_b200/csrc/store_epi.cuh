// "Store the tile" epilogue of the NT-GEMM skeleton: C[slot][m][col] = acc (row-major, leading dim ldc).
// slot = k_split * n_splits + n_split lets split-K partial products land in separate planes that a later
// kernel sums in a fixed order (deterministic, no atomics).
#pragma once
#include "nt_gemm.cuh"

namespace dif {

struct StoreEpi {
  struct Params {
    float* C;
    int M, N, ldc;
    int n_splits;          // slots that belong to one K split (= shape.n_splits)
    size_t slot_stride;    // elements between K-split planes (0: single plane)
  };
  static int smem_bytes(const Params&) { return 16; }
  const Params& p;
  float* row_ptr;
  __device__ StoreEpi(const Params& pp, uint8_t*, int) : p(pp), row_ptr(nullptr) {}
  __device__ void begin_item(int m, int slot, int) {
    row_ptr = m < p.M ? p.C + (size_t)(slot / p.n_splits) * p.slot_stride + (size_t)m * p.ldc : nullptr;
  }
  __device__ void begin_tile(int) {}
  __device__ void consume(int col0, const uint32_t (&acc)[32], uint32_t, uint32_t (&)[32]) {
    if (!row_ptr) return;
    if (col0 + 32 <= p.N && (p.ldc & 3) == 0) {
      float4* dst = reinterpret_cast<float4*>(row_ptr + col0);
#pragma unroll
      for (int i = 0; i < 8; ++i)
        dst[i] = make_float4(__uint_as_float(acc[4 * i]), __uint_as_float(acc[4 * i + 1]), __uint_as_float(acc[4 * i + 2]),
                             __uint_as_float(acc[4 * i + 3]));
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (col0 + i < p.N) row_ptr[col0 + i] = __uint_as_float(acc[i]);
    }
  }
  __device__ void end_item(int, int) {}
};

}  // namespace dif
