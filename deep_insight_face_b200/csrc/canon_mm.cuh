// Canonical fp32 B x B products on CUDA cores, one thread per entry (D <= 512) - shared by the tfa pairwise-distance
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
constexpr int PF_THREADS = 256;   // 8 warps x 32 lanes; a thread owns TM rows x TN columns of the T x T block tile

// Geometry by row length: D <= 128 -> one 16-byte quad per chain and row (KG = 1), 64 x 64 tile, 8 x 2 entries per thread;
// D <= 256 / 512 -> KG = 2 / 4 quads per chain and row, 32 x 32 tile, 4 x 1 entries per thread (the operand tiles are
// KG x 16 KB per 32 rows, so the tile shrinks as the rows grow).
template <int KG>
struct CanonMmGeom {
  static constexpr int TM = KG == 1 ? 8 : 4;
  static constexpr int TN = KG == 1 ? 2 : 1;
  static constexpr int T = 8 * TM;                 // == 32 * TN
  static constexpr int kOperandFloats = KG * 32 * T * 4;
  static constexpr int kMinBlocks = KG == 4 ? 1 : 2;
};

// operand slot of (quad group kg, chain l, row): [KG][32 chains][T rows][4]
template <int T>
__device__ __forceinline__ int pf_slot(int kg, int l, int row) { return ((kg * 32 + l) * T + (row ^ ((l >> 2) & 7))) * 4; }
// Row operand, PAIRED layout [KG][32 chains][T / 2 row pairs][4 k][2 rows]: the two rows of a pair sit side by side for
// every k, so a 16-byte load yields two ready-made (row 2p, row 2p + 1) operand pairs of the packed fp32 pipe below.
template <int T>
__device__ __forceinline__ int pf_pair_slot(int kg, int l, int pair) {
  return ((kg * 32 + l) * (T / 2) + (pair ^ ((l >> 2) & 7))) * 8;
}

// Packed fp32 arithmetic (sm_100: FFMA2 / FADD2).  Each half is an ordinary IEEE fma.rn / add.rn, so results are the
// bits of the scalar instructions - and the fp32 pipe retires two of them per issue slot (a scalar FFMA occupies its
// SMSP's fma pipe for two cycles: the scalar version of this kernel was bound by exactly that).
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ unsigned long long dup2(float v) {
  unsigned long long d;
  asm("mov.b64 %0, {%1, %1};" : "=l"(d) : "f"(v));
  return d;
}
__device__ __forceinline__ void unpack2(unsigned long long p, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p));
}

// rows [row0, row0 + T) -> dst: element d = l + 32 k of a row lands at (k / 4, l, row) [k % 4]
template <int KG, int T, bool PAIR = false>
__device__ __forceinline__ void pf_stage(const float* __restrict__ x, int B, int D, int row0, float* __restrict__ dst,
                                         bool vec) {
  constexpr int DP = 128 * KG;   // padded row length
  auto at = [](int kg, int l, int row, int k) {
    return PAIR ? pf_pair_slot<T>(kg, l, row >> 1) + 2 * k + (row & 1) : pf_slot<T>(kg, l, row) + k;
  };
  if (vec) {   // D % 4 == 0, 16-byte aligned rows: one float4 per lane, four conflict-free scalar stores
    for (int idx = threadIdx.x; idx < T * (DP / 4); idx += PF_THREADS) {
      const int row = idx / (DP / 4), q = idx % (DP / 4), d0 = q * 4, gr = row0 + row;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gr < B && d0 < D) v = *reinterpret_cast<const float4*>(x + (size_t)gr * D + d0);
      const int kg = d0 >> 7, k = (d0 & 127) >> 5, l0 = d0 & 31;
      dst[at(kg, l0, row, k)] = v.x;
      dst[at(kg, l0 + 1, row, k)] = v.y;
      dst[at(kg, l0 + 2, row, k)] = v.z;
      dst[at(kg, l0 + 3, row, k)] = v.w;
    }
  } else {
    for (int idx = threadIdx.x; idx < T * DP; idx += PF_THREADS) {
      const int row = idx / DP, d = idx % DP, gr = row0 + row;
      dst[at(d >> 7, d & 31, row, (d & 127) >> 5)] = (gr < B && d < D) ? x[(size_t)gr * D + d] : 0.f;
    }
  }
}

// Epi: `float operator()(int gi, int gj, float dot) const` maps the canonical dot product of rows gi and gj to the stored
// entry; it must be symmetric in (gi, gj) bit for bit, because only tiles on or above the diagonal are computed.
template <typename Epi, int KG>
__global__ void __launch_bounds__(PF_THREADS, CanonMmGeom<KG>::kMinBlocks)
canon_mm_kernel(const float* __restrict__ x, int B, int D, int tiles_per_block, int vec, Epi epi, float* __restrict__ P, int ldp) {
  using Geo = CanonMmGeom<KG>;
  constexpr int T = Geo::T, TM = Geo::TM, TN = Geo::TN;
  extern __shared__ __align__(16) float pf_sm[];
  float* sA = pf_sm;                          // [KG][32][T][4]
  float* sB = pf_sm + Geo::kOperandFloats;    // the same for the column tile; afterwards the T x (T + 1) transpose buffer
  // warp w owns rows TM w .. TM w + TM - 1 (its A loads are whole-warp broadcasts), lane t columns t (and t + 32: its B
  // loads are 512 distinct contiguous bytes): for D <= 128, 10 shared-memory loads per 64 fma, none of them redundant
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int bi = blockIdx.y, row0 = bi * T;
  const int n_tiles = (B + T - 1) / T;
  // P is symmetric bit for bit (products and the two norms commute): only tiles on or above the diagonal are
  // computed, each off-diagonal tile is also written transposed
  if ((int)(blockIdx.x + 1) * tiles_per_block <= bi) return;
  pf_stage<KG, T, true>(x, B, D, row0, sA, vec != 0);   // row operand: paired layout
  constexpr int TP = TM / 2;                            // row pairs per thread
  for (int t = 0; t < tiles_per_block; ++t) {
    const int bj = blockIdx.x * tiles_per_block + t;
    if (bj >= n_tiles) break;   // block-uniform
    if (bj < bi) continue;
    const int c0 = bj * T;
    __syncthreads();            // the previous tile's readers are done
    pf_stage<KG, T>(x, B, D, c0, sB, vec != 0);
    __syncthreads();
    unsigned long long st[6][TP][TN];   // (row 2p, row 2p + 1) pairs of pending partial sums
#pragma unroll
    for (int n = 0; n < 32; ++n) {
      const int l = ((n & 1) << 4) | ((n & 2) << 2) | (n & 4) | ((n & 8) >> 2) | ((n & 16) >> 4);   // bit reversal
      unsigned long long s2[TP][TN];
#pragma unroll
      for (int kg = 0; kg < KG; ++kg) {   // chain l: k = 0, 1, 2, ... in ascending order
        unsigned long long bd[TN][4];     // the column operand, each value duplicated into both halves
#pragma unroll
        for (int j = 0; j < TN; ++j) {
          const float4 b = *reinterpret_cast<const float4*>(sB + pf_slot<T>(kg, l, tx + 32 * j));
          bd[j][0] = dup2(b.x);
          bd[j][1] = dup2(b.y);
          bd[j][2] = dup2(b.z);
          bd[j][3] = dup2(b.w);
        }
#pragma unroll
        for (int p2 = 0; p2 < TP; ++p2) {
          const ulonglong2* ap = reinterpret_cast<const ulonglong2*>(sA + pf_pair_slot<T>(kg, l, ty * TP + p2));
          const ulonglong2 a01 = ap[0], a23 = ap[1];   // (k0 pair, k1 pair), (k2 pair, k3 pair)
#pragma unroll
          for (int j = 0; j < TN; ++j) {
            unsigned long long v = fma2(a01.x, bd[j][0], kg == 0 ? 0ull : s2[p2][j]);
            v = fma2(a01.y, bd[j][1], v);
            v = fma2(a23.x, bd[j][2], v);
            s2[p2][j] = fma2(a23.y, bd[j][3], v);
          }
        }
      }
#pragma unroll
      for (int p2 = 0; p2 < TP; ++p2)
#pragma unroll
        for (int j = 0; j < TN; ++j) {
          unsigned long long v = s2[p2][j];
          int lvl = 0;
#pragma unroll
          for (int m = n; m & 1; m >>= 1) {
            v = add2(st[lvl][p2][j], v);
            ++lvl;
          }
          st[lvl][p2][j] = v;
        }
    }
    float out[TM][TN];
#pragma unroll
    for (int p2 = 0; p2 < TP; ++p2)
#pragma unroll
      for (int j = 0; j < TN; ++j) unpack2(st[5][p2][j], out[2 * p2][j], out[2 * p2 + 1][j]);
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        const int gi = row0 + ty * TM + i, gj = c0 + tx + 32 * j;
        const float d = (gi < B && gj < B) ? epi(gi, gj, out[i][j]) : 0.f;
        out[i][j] = d;
        if (gi < B && gj < B) P[(size_t)gi * ldp + gj] = d;
      }
    if (bj > bi) {
      __syncthreads();          // sB is free
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) sB[(tx + 32 * j) * (T + 1) + ty * TM + i] = out[i][j];
      __syncthreads();
      for (int r = ty; r < T; r += PF_THREADS / 32) {
        const int gj = c0 + r;
        if (gj >= B) break;
#pragma unroll
        for (int h = 0; h < TN; ++h) {
          const int gi = row0 + tx + 32 * h;
          if (gi < B) P[(size_t)gj * ldp + gi] = sB[r * (T + 1) + tx + 32 * h];
        }
      }
    }
  }
}

template <typename Epi, int KG>
int canon_mm_launch_kg(const float* x, int B, int D, const Epi& epi, float* P, int ldp, cudaStream_t st) {
  using Geo = CanonMmGeom<KG>;
  const size_t smem = 2 * (size_t)Geo::kOperandFloats * sizeof(float);   // 64 KB (D <= 256), 128 KB (D <= 512)
  static bool configured = false;   // per instantiation
  if (!configured) {
    DIF_CUDA_OK(cudaFuncSetAttribute(canon_mm_kernel<Epi, KG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  const int sms = device_sm_count() > 0 ? device_sm_count() : 148;
  const int tiles = (B + Geo::T - 1) / Geo::T;
  // the row tile stays staged while a block walks `tpb` column tiles (only those on or above the diagonal do work)
  const int tpb = tiles * tiles / 2 >= 8 * sms ? 2 : 1;
  const int vec = (D % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15u) == 0) ? 1 : 0;
  canon_mm_kernel<Epi, KG><<<dim3((tiles + tpb - 1) / tpb, tiles), PF_THREADS, smem, st>>>(x, B, D, tpb, vec, epi, P, ldp);
  DIF_LAUNCH_OK();
  return DIF_OK;
}

// Host: P [B][ldp] = epi(canonical x_i . x_j) for all i, j; x [B][D], D <= 512.
constexpr int kCanonMmMaxD = 512;
template <typename Epi>
int canon_mm_launch(const float* x, int B, int D, const Epi& epi, float* P, int ldp, cudaStream_t st) {
  if (D <= 128) return canon_mm_launch_kg<Epi, 1>(x, B, D, epi, P, ldp, st);
  if (D <= 256) return canon_mm_launch_kg<Epi, 2>(x, B, D, epi, P, ldp, st);
  return canon_mm_launch_kg<Epi, 4>(x, B, D, epi, P, ldp, st);
}

}  // namespace dif
