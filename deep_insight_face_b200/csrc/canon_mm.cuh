// Canonical fp32 B x B products on CUDA cores, one thread per entry (D <= 128) - shared by the tfa pairwise-distance
// matrix (tfa_triplet.cu) and the batch-all similarity matrix (batch_hard.cu).
#pragma once
#include "dif_canon.cuh"
#include "dif_common.cuh"

namespace dif {

// The canonical dot product (dif_canon.cuh) is 32 strided fma chains plus a fixed pairing tree.  Nothing in that
// definition needs 32 lanes: here ONE thread runs all 32 chains of its 8 x 2 entries one after the other - visiting
// them in bit-reversed order (0, 16, 8, 24, ...) turns the butterfly t[i] += t[i ^ o], o = 16 .. 1, into a binary
// counter over at most six pending partial sums - so an entry costs 128 fma + 31 add, no shuffles, and the operands
// come out of shared memory as one 16-byte load per (chain, row): rows are staged chain-major, [chain][row][k], with
// the row index XOR-swizzled by the chain so that the staging stores spread over all banks.  Bit-identical to the
// warp-tile kernels of bh_tile.cuh (same products, same order; padding terms are fma(0, 0, acc) = acc).
constexpr int PF_T = 64;          // block tile: 64 x 64 entries; 8 warps x 32 lanes, 8 rows x 2 columns per thread
constexpr int PF_THREADS = 256;

__device__ __forceinline__ int pf_slot(int l, int row) { return (l * PF_T + (row ^ ((l >> 2) & 7))) * 4; }

// rows [row0, row0 + 64) -> dst [32 chains][64 rows][4]: element d = l + 32 k of a row lands at (l, row, k)
__device__ __forceinline__ void pf_stage(const float* __restrict__ x, int B, int D, int row0, float* __restrict__ dst,
                                         bool vec) {
  if (vec) {   // D % 4 == 0, 16-byte aligned rows: one float4 per lane, four conflict-free scalar stores
    for (int idx = threadIdx.x; idx < PF_T * 32; idx += PF_THREADS) {
      const int row = idx >> 5, q = idx & 31, d0 = q * 4, gr = row0 + row;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gr < B && d0 < D) v = *reinterpret_cast<const float4*>(x + (size_t)gr * D + d0);
      const int k = d0 >> 5, l0 = d0 & 31;
      dst[pf_slot(l0, row) + k] = v.x;
      dst[pf_slot(l0 + 1, row) + k] = v.y;
      dst[pf_slot(l0 + 2, row) + k] = v.z;
      dst[pf_slot(l0 + 3, row) + k] = v.w;
    }
  } else {
    for (int idx = threadIdx.x; idx < PF_T * 128; idx += PF_THREADS) {
      const int row = idx >> 7, d = idx & 127, gr = row0 + row;
      dst[pf_slot(d & 31, row) + (d >> 5)] = (gr < B && d < D) ? x[(size_t)gr * D + d] : 0.f;
    }
  }
}

// Epi: `float operator()(int gi, int gj, float dot) const` maps the canonical dot product of rows gi and gj to the stored
// entry; it must be symmetric in (gi, gj) bit for bit, because only tiles on or above the diagonal are computed.
template <typename Epi>
__global__ void __launch_bounds__(PF_THREADS, 2) canon_mm_kernel(const float* __restrict__ x, int B, int D, int tiles_per_block,
                                                                 int vec, Epi epi, float* __restrict__ P, int ldp) {
  extern __shared__ __align__(16) float pf_sm[];
  float* sA = pf_sm;                    // [32][64][4]
  float* sB = pf_sm + 32 * PF_T * 4;    // the same for the column tile; afterwards the 64 x 65 transpose buffer
  // warp w owns rows 8w .. 8w + 7 (its A loads are whole-warp broadcasts), lane t columns t and t + 32 (its B loads
  // are 512 distinct contiguous bytes): 10 shared-memory loads per 64 fma, none of them redundant
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int bi = blockIdx.y, row0 = bi * PF_T;
  const int n_tiles = (B + PF_T - 1) / PF_T;
  // P is symmetric bit for bit (products and the two norms commute): only tiles on or above the diagonal are
  // computed, each off-diagonal tile is also written transposed
  if ((int)(blockIdx.x + 1) * tiles_per_block <= bi) return;
  pf_stage(x, B, D, row0, sA, vec != 0);
  for (int t = 0; t < tiles_per_block; ++t) {
    const int bj = blockIdx.x * tiles_per_block + t;
    if (bj >= n_tiles) break;   // block-uniform
    if (bj < bi) continue;
    const int c0 = bj * PF_T;
    __syncthreads();            // the previous tile's readers are done
    pf_stage(x, B, D, c0, sB, vec != 0);
    __syncthreads();
    float st[6][8][2];
#pragma unroll
    for (int n = 0; n < 32; ++n) {
      const int l = ((n & 1) << 4) | ((n & 2) << 2) | (n & 4) | ((n & 8) >> 2) | ((n & 16) >> 4);   // bit reversal
      float4 b[2];
#pragma unroll
      for (int j = 0; j < 2; ++j) b[j] = *reinterpret_cast<const float4*>(sB + pf_slot(l, tx + 32 * j));
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 a = *reinterpret_cast<const float4*>(sA + pf_slot(l, ty * 8 + i));
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          float s = __fmaf_rn(a.x, b[j].x, 0.f);
          s = __fmaf_rn(a.y, b[j].y, s);
          s = __fmaf_rn(a.z, b[j].z, s);
          s = __fmaf_rn(a.w, b[j].w, s);
          int lvl = 0;
#pragma unroll
          for (int m = n; m & 1; m >>= 1) {
            s = __fadd_rn(st[lvl][i][j], s);
            ++lvl;
          }
          st[lvl][i][j] = s;
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int gi = row0 + ty * 8 + i, gj = c0 + tx + 32 * j;
        const float d = (gi < B && gj < B) ? epi(gi, gj, st[5][i][j]) : 0.f;
        st[5][i][j] = d;
        if (gi < B && gj < B) P[(size_t)gi * ldp + gj] = d;
      }
    if (bj > bi) {
      __syncthreads();          // sB is free
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) sB[(tx + 32 * j) * (PF_T + 1) + ty * 8 + i] = st[5][i][j];
      __syncthreads();
      for (int r = ty; r < PF_T; r += PF_THREADS / 32) {
        const int gj = c0 + r;
        if (gj >= B) break;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int gi = row0 + tx + 32 * h;
          if (gi < B) P[(size_t)gj * ldp + gi] = sB[r * (PF_T + 1) + tx + 32 * h];
        }
      }
    }
  }
}

// Host: P [B][ldp] = epi(canonical x_i . x_j) for all i, j; x [B][D], D <= 128.
template <typename Epi>
int canon_mm_launch(const float* x, int B, int D, const Epi& epi, float* P, int ldp, cudaStream_t st) {
  const size_t smem = 2 * 32 * PF_T * 4 * sizeof(float);   // 64 KB
  static bool configured = false;   // per Epi instantiation
  if (!configured) {
    DIF_CUDA_OK(cudaFuncSetAttribute(canon_mm_kernel<Epi>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  const int sms = device_sm_count() > 0 ? device_sm_count() : 148;
  const int tiles = (B + PF_T - 1) / PF_T;
  // the row tile stays staged while a block walks `tpb` column tiles (only those on or above the diagonal do work)
  const int tpb = tiles * tiles / 2 >= 8 * sms ? 2 : 1;
  const int vec = (D % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15u) == 0) ? 1 : 0;
  canon_mm_kernel<Epi><<<dim3((tiles + tpb - 1) / tpb, tiles), PF_THREADS, smem, st>>>(x, B, D, tpb, vec, epi, P, ldp);
  DIF_LAUNCH_OK();
  return DIF_OK;
}

}  // namespace dif
