// Rows -> canonical operand planes (normalise, TF32 hi/lo split, bf16), shared by gallery.cu and arcface.cu.
#pragma once
#include <algorithm>

#include "dif_canon.cuh"
#include "dif_common.cuh"
#include "dif_ptx.cuh"

namespace dif {

// ------------------------------------------------------------------------------------------
// K5: rows -> canonical planes.  One warp per row.
//   SRC 0: rows read from `src`; SRC 1: synthetic rows (seed, row0 + r).
//   p0/p1: fp32 planes (TF32x3: p0 = tf32(x), p1 = x - p0 exactly; otherwise p0 = x (may be NULL), p1 unused)
//   pb   : bf16 plane (bf16 modes);  pb1: second bf16 plane, bf16(x - pb) (3xBF16 mode)
//   sq[r] = canonical sum of squares of the STORED row
//   gmax : running max of sq (orderable uint), may be NULL
// ------------------------------------------------------------------------------------------
struct PrepParams {
  const float* src;
  uint64_t seed;
  int64_t row0;
  int64_t n;
  int D;
  int normalize;
  int split;  // 1: write hi/lo planes
  float* p0;
  float* p1;
  __nv_bfloat16* pb;
  __nv_bfloat16* pb1;
  float* sq;
  unsigned int* gmax;
  float* inv;   // inverse norm applied to each row (normalize), may be NULL
};

template <int SRC>
__device__ __forceinline__ void prep_one_row(const PrepParams& p, int64_t r, int lane) {
  const float* s = SRC == 0 ? p.src + r * p.D : nullptr;
  float acc = 0.f;
#pragma unroll 8   // the loads of a row are independent of the fma chain: keep several in flight
  for (int d = lane; d < p.D; d += 32) {
    const float x = SRC == 0 ? s[d] : synth_value(p.seed, (uint64_t)(p.row0 + r), (uint64_t)d, (uint64_t)p.D);
    acc = __fmaf_rn(x, x, acc);
  }
  const float ss = canon_tree(acc);
  const float inv = p.normalize ? canon_inv_norm(ss) : 1.0f;
  float acc2 = 0.f;
#pragma unroll 4
  for (int d = lane; d < p.D; d += 32) {
    float x = SRC == 0 ? s[d] : synth_value(p.seed, (uint64_t)(p.row0 + r), (uint64_t)d, (uint64_t)p.D);
    if (p.normalize) x = __fmul_rn(x, inv);
    acc2 = __fmaf_rn(x, x, acc2);
    if (p.split) {
      const float hi = tf32_round(x);
      p.p0[r * p.D + d] = hi;
      p.p1[r * p.D + d] = __fsub_rn(x, hi);
    } else if (p.p0) {
      p.p0[r * p.D + d] = x;
    }
    if (p.pb) {
      const __nv_bfloat16 b0 = __float2bfloat16_rn(x);
      p.pb[r * p.D + d] = b0;
      if (p.pb1) p.pb1[r * p.D + d] = __float2bfloat16_rn(__fsub_rn(x, __bfloat162float(b0)));
    }
  }
  const float ss2 = canon_tree(acc2);
  if (lane == 0) {
    if (p.inv) p.inv[r] = inv;
    if (p.sq) p.sq[r] = ss2;
    if (p.gmax) atomicMax(p.gmax, float_orderable(ss2));
  }
}

template <int SRC>
__global__ void __launch_bounds__(256) prep_rows_kernel(PrepParams p) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < p.n; r += warps)
    prep_one_row<SRC>(p, r, lane);
}

// Two row sets in one launch (ArcFace: the batch and the class weights): rows [0, a.n) belong to `a`, the rest to `b`.
static __global__ void __launch_bounds__(256) prep_rows_pair_kernel(PrepParams a, PrepParams b) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < a.n + b.n; r += warps) {
    if (r < a.n) prep_one_row<0>(a, r, lane);
    else prep_one_row<0>(b, r - a.n, lane);
  }
}


inline int prep_launch(const PrepParams& pp, bool synth, cudaStream_t st) {
  if (pp.n == 0) return DIF_OK;
  const int warps_per_block = 8;
  int64_t blocks = (pp.n + warps_per_block - 1) / warps_per_block;
  blocks = std::min<int64_t>(blocks, 148 * 16);
  if (synth) prep_rows_kernel<1><<<(unsigned)blocks, 256, 0, st>>>(pp);
  else prep_rows_kernel<0><<<(unsigned)blocks, 256, 0, st>>>(pp);
  DIF_LAUNCH_OK();
  return DIF_OK;
}

inline int prep_launch_pair(const PrepParams& a, const PrepParams& b, cudaStream_t st) {
  const int64_t n = a.n + b.n;
  if (n == 0) return DIF_OK;
  const int64_t blocks = std::min<int64_t>((n + 7) / 8, 148 * 16);
  prep_rows_pair_kernel<<<(unsigned)blocks, 256, 0, st>>>(a, b);
  DIF_LAUNCH_OK();
  return DIF_OK;
}


}  // namespace dif
