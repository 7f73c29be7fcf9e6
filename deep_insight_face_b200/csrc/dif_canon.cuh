// Canonical fp32 reductions.  "Exact" results of this library (re-ranked gallery scores,
// row norms, pair distances) are DEFINED by these functions and mirrored operation for operation
// in oracle/dif_oracle.c, so the CPU oracle reproduces them bit for bit:
//   * 32 strided chains: chain l accumulates elements d = l, l+32, l+64, ... with one IEEE fma
//     (or sub+fma) each, starting from +0;
//   * the 32 partial sums are combined by the balanced tree t[i] += t[i ^ o], o = 16, 8, 4, 2, 1
//     (round-to-nearest adds; a + b == b + a so every lane ends with the same value);
//   * normalisation multiplies by inv = 1 / sqrt(max(ss, 1e-12)) using IEEE sqrt and divide
//     (the reference's tf.nn.l2_normalize is x * rsqrt(max(ss, 1e-12)); same value to ~1 ulp).
// Only intrinsics with fixed rounding are used, so -fmad contraction cannot change results.
#pragma once
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace dif {

constexpr float kNormEps = 1e-12f;

__device__ __forceinline__ float canon_tree(float acc) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) acc = __fadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, o));
  return acc;
}

// All 32 lanes of the warp call this; every lane returns the same value.
__device__ __forceinline__ float canon_dot_warp(const float* __restrict__ a, const float* __restrict__ b, int D) {
  float acc = 0.f;
  for (int d = (int)(threadIdx.x & 31u); d < D; d += 32) acc = __fmaf_rn(a[d], b[d], acc);
  return canon_tree(acc);
}

// Operands stored as TF32 hi plane + exact residual lo plane: value = hi + lo (exact).
__device__ __forceinline__ float canon_dot_warp_planes(const float* __restrict__ a_hi, const float* __restrict__ a_lo,
                                                       const float* __restrict__ b_hi, const float* __restrict__ b_lo,
                                                       int D) {
  float acc = 0.f;
  for (int d = (int)(threadIdx.x & 31u); d < D; d += 32)
    acc = __fmaf_rn(__fadd_rn(a_hi[d], a_lo[d]), __fadd_rn(b_hi[d], b_lo[d]), acc);
  return canon_tree(acc);
}

__device__ __forceinline__ float canon_sqdist_warp(const float* __restrict__ a, const float* __restrict__ b, int D) {
  float acc = 0.f;
  for (int d = (int)(threadIdx.x & 31u); d < D; d += 32) {
    const float t = __fsub_rn(a[d], b[d]);
    acc = __fmaf_rn(t, t, acc);
  }
  return canon_tree(acc);
}

__device__ __forceinline__ float canon_sqdist_warp_planes(const float* __restrict__ a_hi,
                                                          const float* __restrict__ a_lo,
                                                          const float* __restrict__ b_hi,
                                                          const float* __restrict__ b_lo, int D) {
  float acc = 0.f;
  for (int d = (int)(threadIdx.x & 31u); d < D; d += 32) {
    const float t = __fsub_rn(__fadd_rn(a_hi[d], a_lo[d]), __fadd_rn(b_hi[d], b_lo[d]));
    acc = __fmaf_rn(t, t, acc);
  }
  return canon_tree(acc);
}

__device__ __forceinline__ float canon_inv_norm(float ss) {
  return __fdiv_rn(1.0f, __fsqrt_rn(fmaxf(ss, kNormEps)));
}

// Synthetic data: element (row, col) of the seed's [*, dim] matrix, an Irwin-Hall(4) approximation
// of N(0,1) built from integer arithmetic and a single fp32 multiply, hence bit-identical on CPU and GPU.
__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
__host__ __device__ __forceinline__ float synth_value(uint64_t seed, uint64_t row, uint64_t col, uint64_t dim) {
  const uint64_t h = splitmix64(seed * 0xD1342543DE82EF95ull + row * dim + col);
  const int s = (int)(h & 0xFFFF) + (int)((h >> 16) & 0xFFFF) + (int)((h >> 32) & 0xFFFF) + (int)((h >> 48) & 0xFFFF) -
                131070;
  return (float)s * 2.6428998e-05f;  // 1 / (65536 / sqrt(3))
}

// Order-preserving map fp32 -> u32 (larger float -> larger unsigned).
__host__ __device__ __forceinline__ uint32_t float_orderable(float f) {
#ifdef __CUDA_ARCH__
  uint32_t u = __float_as_uint(f);
#else
  uint32_t u;
  memcpy(&u, &f, 4);
#endif
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
// Selection key: higher `better` wins, ties go to the lower index.  key 0 == invalid.
__host__ __device__ __forceinline__ uint64_t make_key(float better, uint32_t idx) {
  return ((uint64_t)float_orderable(better) << 32) | (uint64_t)(0xFFFFFFFFu - idx);
}
__host__ __device__ __forceinline__ uint32_t key_index(uint64_t key) { return 0xFFFFFFFFu - (uint32_t)key; }

}  // namespace dif
