// C[m, n] = sum_k A[m, k] * B[n, k]  ("NT" contraction of two K-major row sets) on the sm_100a
// tensor cores, with the accumulator tile consumed straight out of TMEM by a row-owner epilogue:
// epilogue thread r of a CTA owns row r of the 128-row A block and sees every column of its
// row, 32 at a time, so per-row reductions (top-k, masked min/max, online softmax) never touch
// HBM.  This is the one GEMM skeleton behind gallery search, batch-hard mining and ArcFace.
//
// Precision modes (PREC):
//   0  3xTF32: A = Ahi + Alo, B = Bhi + Blo pre-split into TF32 planes; Alo*Bhi + Ahi*Blo + Ahi*Bhi
//   1  bf16  : single bf16 plane per operand
//   2  1xTF32: hi planes only
//
// Structure: persistent CTAs (or cta_group::2 CTA pairs), 6 warps: warps 0-3 epilogue (TMEM lane
// quarter = warp index), warp 4 TMA producer, warp 5 TMEM allocator + single-thread MMA issuer.
// smem ring of STAGES K-chunks (one 128-byte swizzle row of K per chunk), NBUF accumulator
// buffers in TMEM so the epilogue of tile t overlaps the MMAs of tile t+1.
#pragma once
#include <cuda.h>

#include "dif_common.cuh"
#include "dif_ptx.cuh"

namespace dif {

constexpr int GEMM_BM = 128;          // A rows per CTA == epilogue threads == TMEM lanes
constexpr int GEMM_THREADS = 192;     // 4 epilogue warps + producer + MMA
constexpr int GEMM_SWZ = 128;         // bytes of K per smem row (SWIZZLE_128B)
constexpr int GEMM_SMEM_MAX = 232448; // 227 KB opt-in limit per CTA

template <int PREC>
struct PrecTraits {
  static constexpr bool kTf32 = (PREC != 1);
  static constexpr int kElemBytes = kTf32 ? 4 : 2;
  static constexpr int kPlanes = (PREC == 0) ? 2 : 1;
  static constexpr int kChunkElems = GEMM_SWZ / kElemBytes;  // K elements per stage
  static constexpr int kKSteps = GEMM_SWZ / 32;               // MMAs (per product) per stage: 32 B of K each
  static constexpr uint32_t kFmt = kTf32 ? 2u : 1u;
};

struct GemmShape {
  int m_blocks;         // number of (128*CTAS)-row blocks of A
  int n_tiles;          // number of BN-row tiles of B
  int k_chunks;         // K / chunk elems
  int n_splits;         // B tile range is cut into n_splits pieces -> items = n_splits * m_blocks
  int tiles_per_split;  // ceil(n_tiles / n_splits)
};

template <int PREC, int BN, int CTAS, int EPI_SMEM>
struct GemmLayout {
  using PT = PrecTraits<PREC>;
  static constexpr int kATile = GEMM_BM * GEMM_SWZ;       // bytes per plane per stage
  static constexpr int kBTile = (BN / CTAS) * GEMM_SWZ;   // each CTA of a pair stages half of B's rows
  static constexpr int kStage = PT::kPlanes * (kATile + kBTile);
  static constexpr int kAccBufs = 512 / BN;
  static constexpr int kBarBytes = 1024;
  static constexpr int kAvail = GEMM_SMEM_MAX - 1024 /*align slack*/ - kBarBytes - EPI_SMEM;
  static constexpr int kStagesRaw = kAvail / kStage;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr int kSmemBytes = 1024 + kStages * kStage + kBarBytes + EPI_SMEM;
  static_assert(kStages >= 2, "not enough shared memory for a 2-stage ring");
  static_assert(BN == 128 || BN == 256, "BN");
  static_assert(kAccBufs >= 2, "need two accumulator buffers");
};

// Epilogue concept:
//   struct Epi {
//     struct Params;                       // POD, passed by value to the kernel
//     static constexpr int kSmemBytes;     // CTA-shared scratch, 16-byte aligned
//     __device__ Epi(const Params&, uint8_t* smem, int row_in_block);
//     __device__ void begin_item(int m_row /*global A row of this thread*/, int split, int col_begin);
//     __device__ void consume(int col0, const uint32_t (&acc)[32]);   // fp32 bits, columns col0..col0+31
//     __device__ void end_item(int m_row, int split);
//   };

template <int PREC, int BN, int CTAS, class Epi>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
nt_gemm_rowscan_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                       const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo,
                       GemmShape shape, typename Epi::Params ep) {
  using PT = PrecTraits<PREC>;
  using L = GemmLayout<PREC, BN, CTAS, Epi::kSmemBytes>;
  constexpr int STAGES = L::kStages;
  constexpr int NBUF = L::kAccBufs;

  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B operand tiles need 1024-byte alignment.
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* stage_base = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * L::kStage);
  uint64_t* full_bar = bars;                    // [STAGES]
  uint64_t* empty_bar = bars + STAGES;          // [STAGES]
  uint64_t* acc_full = bars + 2 * STAGES;       // [NBUF]
  uint64_t* acc_empty = bars + 2 * STAGES + NBUF;  // [NBUF]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 2 * NBUF);
  uint8_t* epi_smem = smem + STAGES * L::kStage + L::kBarBytes;

  const int warp = threadIdx.x >> 5;
  const uint32_t cta_rank = (CTAS == 2) ? cluster_ctarank() : 0u;
  const bool leader = (cta_rank == 0);
  const int unit = (CTAS == 2) ? (blockIdx.x >> 1) : blockIdx.x;   // persistent work unit (CTA or pair)
  const int n_units = (CTAS == 2) ? (gridDim.x >> 1) : gridDim.x;
  const int n_items = shape.n_splits * shape.m_blocks;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], CTAS);   // one arrive(+tx) per producing CTA
      mbar_init(&empty_bar[s], 1);     // one tcgen05.commit
    }
    for (int b = 0; b < NBUF; ++b) {
      mbar_init(&acc_full[b], 1);                 // one tcgen05.commit
      mbar_init(&acc_empty[b], GEMM_BM * CTAS);   // every epilogue thread of the unit
    }
    fence_mbar_init();
  }
  if (warp == 4 && lane_id() == 0) {
    tma_prefetch_desc(&tm_a_hi);
    tma_prefetch_desc(&tm_b_hi);
    if (PT::kPlanes == 2) {
      tma_prefetch_desc(&tm_a_lo);
      tma_prefetch_desc(&tm_b_lo);
    }
  }
  if (warp == 5) {
    tmem_alloc<CTAS>(tmem_slot, 512);
    tmem_relinquish<CTAS>();
  }
  tc_fence_before_sync();
  if (CTAS == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    // ===================== TMA producer (one lane) =====================
    if (lane_id() == 0) {
      uint32_t it = 0;
      for (int item = unit; item < n_items; item += n_units) {
        const int split = item / shape.m_blocks;
        const int mb = item - split * shape.m_blocks;
        const int a_row = (mb * CTAS + (int)cta_rank) * GEMM_BM;
        const int t0 = split * shape.tiles_per_split;
        const int t1 = min(t0 + shape.tiles_per_split, shape.n_tiles);
        for (int t = t0; t < t1; ++t) {
          const int b_row = t * BN + (int)cta_rank * (BN / CTAS);
          for (int kc = 0; kc < shape.k_chunks; ++kc, ++it) {
            const int s = it % STAGES;
            const uint32_t ph = (it / STAGES) & 1u;
            mbar_wait(&empty_bar[s], ph ^ 1u);
            uint8_t* st = stage_base + s * L::kStage;
            const int kx = kc * PT::kChunkElems;
            if (CTAS == 1) {
              mbar_arrive_expect_tx(&full_bar[s], L::kStage);
              tma_load_2d(st, &tm_a_hi, &full_bar[s], kx, a_row);
              tma_load_2d(st + L::kATile * PT::kPlanes, &tm_b_hi, &full_bar[s], kx, b_row);
              if (PT::kPlanes == 2) {
                tma_load_2d(st + L::kATile, &tm_a_lo, &full_bar[s], kx, a_row);
                tma_load_2d(st + L::kATile * 2 + L::kBTile, &tm_b_lo, &full_bar[s], kx, b_row);
              }
            } else {
              // Both CTAs' bytes are credited to the leader's barrier.
              if (leader) mbar_arrive_expect_tx(&full_bar[s], L::kStage * 2);
              else mbar_arrive_cluster(&full_bar[s], 0);
              tma_load_2d_pair(st, &tm_a_hi, &full_bar[s], kx, a_row);
              tma_load_2d_pair(st + L::kATile * PT::kPlanes, &tm_b_hi, &full_bar[s], kx, b_row);
              if (PT::kPlanes == 2) {
                tma_load_2d_pair(st + L::kATile, &tm_a_lo, &full_bar[s], kx, a_row);
                tma_load_2d_pair(st + L::kATile * 2 + L::kBTile, &tm_b_lo, &full_bar[s], kx, b_row);
              }
            }
          }
        }
      }
    }
  } else if (warp == 5) {
    // ===================== MMA issuer (one lane of the leader CTA) =====================
    if (leader && lane_id() == 0) {
      constexpr uint32_t idesc = make_idesc(PT::kFmt, GEMM_BM * CTAS, BN);
      uint32_t it = 0, tc = 0;
      for (int item = unit; item < n_items; item += n_units) {
        const int split = item / shape.m_blocks;
        const int t0 = split * shape.tiles_per_split;
        const int t1 = min(t0 + shape.tiles_per_split, shape.n_tiles);
        for (int t = t0; t < t1; ++t, ++tc) {
          const int buf = tc % NBUF;
          const uint32_t aph = (tc / NBUF) & 1u;
          mbar_wait(&acc_empty[buf], aph ^ 1u);
          tc_fence_after_sync();
          const uint32_t d_tmem = tmem_base + (uint32_t)(buf * BN);
          for (int kc = 0; kc < shape.k_chunks; ++kc, ++it) {
            const int s = it % STAGES;
            const uint32_t ph = (it / STAGES) & 1u;
            mbar_wait(&full_bar[s], ph);
            tc_fence_after_sync();
            const uint32_t st = smem_u32(stage_base + s * L::kStage);
            const uint64_t a_hi = make_kmajor_desc<GEMM_SWZ>(st);
            const uint64_t b_hi = make_kmajor_desc<GEMM_SWZ>(st + L::kATile * PT::kPlanes);
#pragma unroll
            for (int ks = 0; ks < PT::kKSteps; ++ks) {
              const uint64_t adv = (uint64_t)(ks * 2);  // 32 bytes of K, in 16-byte units
              if (PT::kPlanes == 2) {
                const uint64_t a_lo = make_kmajor_desc<GEMM_SWZ>(st + L::kATile);
                const uint64_t b_lo = make_kmajor_desc<GEMM_SWZ>(st + L::kATile * 2 + L::kBTile);
                // small cross terms first, dominant term last
                tc_mma<CTAS, PT::kTf32>(d_tmem, a_lo + adv, b_hi + adv, idesc, (kc | ks) != 0);
                tc_mma<CTAS, PT::kTf32>(d_tmem, a_hi + adv, b_lo + adv, idesc, 1u);
                tc_mma<CTAS, PT::kTf32>(d_tmem, a_hi + adv, b_hi + adv, idesc, 1u);
              } else {
                tc_mma<CTAS, PT::kTf32>(d_tmem, a_hi + adv, b_hi + adv, idesc, (kc | ks) != 0);
              }
            }
            tc_commit<CTAS>(&empty_bar[s]);   // smem slot reusable once these MMAs retire
          }
          tc_commit<CTAS>(&acc_full[buf]);    // accumulator tile complete
        }
      }
    }
  } else {
    // ===================== epilogue warps 0..3 =====================
    const int row = threadIdx.x;  // 0..127 == TMEM lane
    Epi epi(ep, epi_smem, row);
    const uint32_t lane_base = ((uint32_t)(warp * 32)) << 16;
    uint32_t tc = 0;
    for (int item = unit; item < n_items; item += n_units) {
      const int split = item / shape.m_blocks;
      const int mb = item - split * shape.m_blocks;
      const int m_row = (mb * CTAS + (int)cta_rank) * GEMM_BM + row;
      const int t0 = split * shape.tiles_per_split;
      const int t1 = min(t0 + shape.tiles_per_split, shape.n_tiles);
      epi.begin_item(m_row, split, t0 * BN);
      for (int t = t0; t < t1; ++t, ++tc) {
        const int buf = tc % NBUF;
        const uint32_t aph = (tc / NBUF) & 1u;
        mbar_wait(&acc_full[buf], aph);
        tc_fence_after_sync();
        const uint32_t taddr = tmem_base + lane_base + (uint32_t)(buf * BN);
        uint32_t va[32], vb[32];
        tmem_ld32(taddr, va);
#pragma unroll 1
        for (int c = 0; c < BN / 32; c += 2) {
          tmem_ld_wait(va);
          tmem_ld32(taddr + (uint32_t)((c + 1) * 32), vb);
          epi.consume(t * BN + c * 32, va);
          tmem_ld_wait(vb);
          if (c + 2 < BN / 32) tmem_ld32(taddr + (uint32_t)((c + 2) * 32), va);
          epi.consume(t * BN + (c + 1) * 32, vb);
        }
        // all of this thread's TMEM reads of `buf` have completed (wait::ld above)
        tc_fence_before_sync();
        if (CTAS == 1 || leader) mbar_arrive(&acc_empty[buf]);
        else mbar_arrive_cluster(&acc_empty[buf], 0);
      }
      epi.end_item(m_row, split);
    }
  }

  __syncwarp();  // single-lane roles rejoin their warp before the aligned barriers below
  tc_fence_before_sync();
  if (CTAS == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 5) {
    tc_fence_after_sync();
    tmem_dealloc<CTAS>(tmem_base, 512);
  }
}

// Host-side launch of one instantiation.  maps = {A hi, A lo, B hi, B lo}; single-plane modes pass
// the hi map twice.  n_units = persistent CTAs (CTAS == 1) or CTA pairs (CTAS == 2).
template <int PREC, int BN, int CTAS, class Epi>
int launch_nt_gemm(const CUtensorMap* maps, const GemmShape& shape, const typename Epi::Params& ep, int n_units,
                   cudaStream_t stream) {
  using L = GemmLayout<PREC, BN, CTAS, Epi::kSmemBytes>;
  auto kern = nt_gemm_rowscan_kernel<PREC, BN, CTAS, Epi>;
  static bool configured = false;   // per instantiation
  if (!configured) {
    DIF_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kSmemBytes));
    configured = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(n_units * CTAS));
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = L::kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CTAS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  DIF_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, maps[0], maps[1], maps[2], maps[3], shape, ep));
  DIF_LAUNCH_OK();
  return DIF_OK;
}

}  // namespace dif
