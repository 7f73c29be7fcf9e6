// C[m, n] = sum_k A[m, k] * B[n, k]  ("NT" contraction of two K-major row sets) on the sm_100a
// tensor cores, with the accumulator tile consumed straight out of TMEM by a row-owner epilogue:
// epilogue thread r of a CTA owns row r of the 128-row A block and sees every column of its
// row, 32 at a time, so per-row reductions (top-k, masked min/max, online softmax) never touch
// HBM.  This is the one GEMM skeleton behind gallery search, batch-hard mining and ArcFace.
//
// Precision modes (PREC):
//   0  3xTF32: A = Ahi + Alo, B = Bhi + Blo pre-split into TF32 planes; Alo*Bhi + Ahi*Blo + Ahi*Bhi
//   1  bf16  : single bf16 plane per operand
//   2  1xTF32: fp32 storage read as TF32 by the tensor cores
//   3  3xBF16: A = A0 + A1, B = B0 + B1 pre-split into bf16 planes (x - bf16(x) rounded to bf16 again);
//              A1*B0 + A0*B1 + A0*B0 - the mode-0 schedule on the bf16 pipe (twice the MMA rate)
//
// Structure: persistent CTAs (or cta_group::2 CTA pairs), 6 warps: warps 0-3 epilogue (TMEM lane
// quarter = warp index), warp 4 TMA producer, warp 5 TMEM allocator + single-thread MMA issuer.
// A work item is (A row block, range of B tiles).  Two operand schedules:
//   ARES = 0  A and B K-chunks stream together through a ring of smem stages;
//   ARES = 1  the item's whole A block (all K) is loaded once and stays resident in smem while only
//             B streams through the ring: halves the L2->smem traffic per MMA, which is what bounds a
//             128x256 tile at tensor-core speed ("operand feed").
// NBUF accumulator buffers in TMEM let the epilogue of tile t overlap the MMAs of tile t+1.
//
// Operand majors (AMN / BMN): 0 = the operand's global matrix is [rows][K] with K contiguous (K-major, "NT");
// 1 = it is [K][rows] with the ROW index contiguous (MN-major): the same matrix can then serve a second GEMM
// that contracts over its other dimension without a transposed copy in HBM (ArcFace backward: dcos and the
// normalised weights are each read in both roles).  An MN-major K-chunk is staged as 128-byte-wide slabs
// ([K rows][32 fp32 | 64 bf16], one TMA box each) and described to the tensor core with the MN-major canonical
// layout (LBO = slab stride, idesc major bits 15 / 16): SWIZZLE_128B for bf16, SWIZZLE_128B_BASE32B for TF32 - the
// only MN-major layout 32-bit operands have - with the matching TMA mode (make_tmap_2d, atom32).
//
// Tile width: `shape.bn` (a multiple of 32, <= BN, runtime) columns of B per tile - the MMA's N and the TMA box
// height come from it - so the host can pick the width that fills the SMs with whole waves (10 000 classes over
// 74 CTA pairs is two waves of 256-wide tiles but one wave of three 96-wide ones).
#pragma once
#include <cuda.h>

#include <type_traits>

#include "dif_common.cuh"
#include "dif_ptx.cuh"

namespace dif {

constexpr int GEMM_BM = 128;          // A rows per CTA == epilogue threads == TMEM lanes
constexpr int GEMM_THREADS = 192;     // 4 epilogue warps + producer + MMA (8 epilogue warps: 320, see EpiWarps)
constexpr int GEMM_SWZ = 128;         // bytes of K per smem row (SWIZZLE_128B)
constexpr int GEMM_SMEM_MAX = 232448; // 227 KB opt-in limit per CTA
constexpr int GEMM_MAX_STAGES = 8;
constexpr int GEMM_BAR_BYTES = 1024;

template <int PREC>
struct PrecTraits {
  static constexpr bool kTf32 = (PREC == 0 || PREC == 2);
  static constexpr int kElemBytes = kTf32 ? 4 : 2;
  static constexpr int kPlanes = (PREC == 0 || PREC == 3) ? 2 : 1;
  static constexpr int kChunkElems = GEMM_SWZ / kElemBytes;  // K elements per chunk
  static constexpr int kKSteps = GEMM_SWZ / 32;               // MMAs (per product) per chunk: 32 B of K each
  static constexpr uint32_t kFmt = kTf32 ? 2u : 1u;
};

struct GemmShape {
  int m_blocks;         // number of (128*CTAS)-row blocks of A
  int n_tiles;          // number of BN-row tiles of B
  int k_chunks;         // ceil(K / chunk elems)
  int n_splits;         // B tile range is cut into n_splits pieces -> items = n_splits * m_blocks
  int tiles_per_split;  // ceil(n_tiles / n_splits)
  int stages;           // smem ring depth (host-computed from the shared-memory budget)
  int k_splits;         // the K range is cut into k_splits pieces -> items *= k_splits (0/1: no split)
  int chunks_per_ksplit;
  int l2_prefetch;      // 1: the producer prefetches B tiles into L2 two tiles ahead
  int bn;               // columns of B per tile (multiple of 32, <= BN); 0 = BN.  n_tiles counts tiles of this width
  unsigned long long* trace;   // development aid: [CTA][8] globaltimer stamps of the kernel's phases (NULL: off)
};

template <int PREC, int BN, int CTAS>
struct GemmTiles {
  using PT = PrecTraits<PREC>;
  static constexpr int kATile = GEMM_BM * GEMM_SWZ;       // bytes per plane per K chunk
  static constexpr int kBTile = (BN / CTAS) * GEMM_SWZ;   // each CTA of a pair stages half of B's rows
  static constexpr int kAccBufs = 512 / BN;
  static_assert(BN == 128 || BN == 256, "BN");
  static_assert(kAccBufs >= 2, "need two accumulator buffers");
};

// Host: shared-memory plan of one launch.  Returns false if nothing fits.
struct GemmSmemPlan {
  int stages = 0;
  int a_res_bytes = 0;   // resident A block (ARES = 1), else 0
  int stage_bytes = 0;
  int total = 0;
};
template <int PREC, int BN, int CTAS, int ARES>
inline bool plan_gemm_smem(int k_chunks, int epi_smem, GemmSmemPlan* out, int bn = BN) {
  using T = GemmTiles<PREC, BN, CTAS>;
  using PT = PrecTraits<PREC>;
  GemmSmemPlan p;
  p.a_res_bytes = ARES ? PT::kPlanes * k_chunks * T::kATile : 0;
  // a stage holds the tile width actually used: narrower tiles buy a deeper ring (more TMA bytes in flight)
  p.stage_bytes = PT::kPlanes * ((ARES ? 0 : T::kATile) + (bn / CTAS) * GEMM_SWZ);
  const int avail = GEMM_SMEM_MAX - 1024 /*align slack*/ - GEMM_BAR_BYTES - epi_smem - p.a_res_bytes;
  if (avail < 2 * p.stage_bytes) return false;
  p.stages = avail / p.stage_bytes;
  if (p.stages > GEMM_MAX_STAGES) p.stages = GEMM_MAX_STAGES;
  p.total = 1024 + p.a_res_bytes + p.stages * p.stage_bytes + GEMM_BAR_BYTES + epi_smem;
  *out = p;
  return true;
}

// Epilogue concept:
//   struct Epi {
//     struct Params;                       // POD, passed by value to the kernel
//     static int smem_bytes(const Params&); // host: CTA-shared scratch the epilogue needs, 16-byte multiple
//     __device__ Epi(const Params&, uint8_t* smem, int row_in_block);
//     // split = index of the (K split, B-tile range) slot of this item: k_split * n_splits + n_split
//     __device__ void begin_item(int m_row /*global A row of this thread*/, int split, int col_begin);
//     __device__ void begin_tile(int col_begin);   // called by all epilogue threads (128, or 256 with kEpiWarps = 8) before the tile is ready
//     // fp32 bits of columns col0..col0+31 of this thread's row; taddr = TMEM address of column col0
//     // for this warp (a slow path may re-read single columns with tmem_ld1).  `pending` is the
//     // register block of the NEXT chunk, whose tcgen05.ld may still be in flight: code that could
//     // make the compiler move or spill registers (any slow path) must tmem_ld_wait(pending) first.
//     __device__ void consume(int col0, const uint32_t (&acc)[32], uint32_t taddr, uint32_t (&pending)[32]);
//     __device__ void end_item(int m_row, int split);
//   };

// v[i] for a run-time i: registers cannot be indexed dynamically, five levels of selects can.  Slow paths of the row
// epilogues use it to fetch "the column that hit" from the chunk they already hold - re-reading one accumulator
// column from TMEM instead costs a few hundred cycles of latency per column.
__device__ __forceinline__ float pick32(const float (&v)[32], int i) {
  float a[16], b[8], c[4];
#pragma unroll
  for (int k = 0; k < 16; ++k) a[k] = (i & 16) ? v[16 + k] : v[k];
#pragma unroll
  for (int k = 0; k < 8; ++k) b[k] = (i & 8) ? a[8 + k] : a[k];
#pragma unroll
  for (int k = 0; k < 4; ++k) c[k] = (i & 4) ? b[4 + k] : b[k];
  const float d0 = (i & 2) ? c[2] : c[0], d1 = (i & 2) ? c[3] : c[1];
  return (i & 1) ? d1 : d0;
}

// An epilogue that declares `static constexpr int kEpiWarps = 8` gets TWO warps per TMEM lane quarter (warps 0-3
// and 6-9): both own the same 32 rows and each consumes half of a tile's column chunks, so a latency-bound
// epilogue has two warps per scheduler to hide behind.  Such an Epi is constructed with (params, smem, row, half)
// and is handed the slot index 2 * slot + half: its per-row results are per (slot, half), merged downstream.
template <class E, class = void>
struct EpiWarps {
  static constexpr int value = 4;
};
template <class E>
struct EpiWarps<E, std::void_t<decltype(E::kEpiWarps)>> {
  static constexpr int value = E::kEpiWarps;
};
template <class Epi, int EW>
__device__ __forceinline__ Epi make_epi(const typename Epi::Params& p, uint8_t* smem, int row, int half) {
  if constexpr (EW == 8) return Epi(p, smem, row, half);
  else return Epi(p, smem, row);
}

// MN-major operand: K-chunk kc of rows [row0, row0 + n_rows) -> n_rows / slab 128-byte-wide slabs at dst
template <int PREC, int CTAS>
__device__ __forceinline__ void tma_load_mn_slabs(uint8_t* dst, const CUtensorMap* tm, uint64_t* bar, int row0, int n_rows,
                                                  int k0) {
  using PT = PrecTraits<PREC>;
  constexpr int kSlabElems = GEMM_SWZ / PT::kElemBytes;          // rows (MN index) per slab
  constexpr int kSlabBytes = PT::kChunkElems * GEMM_SWZ;         // [K chunk rows][128 B]
  for (int j = 0; j * kSlabElems < n_rows; ++j) {
    if (CTAS == 1) tma_load_2d(dst + j * kSlabBytes, tm, bar, row0 + j * kSlabElems, k0);
    else tma_load_2d_pair(dst + j * kSlabBytes, tm, bar, row0 + j * kSlabElems, k0);
  }
}

template <int PREC, int BN, int CTAS, int ARES, class Epi, int AMN = 0, int BMN = 0>
__global__ void __launch_bounds__((2 + EpiWarps<Epi>::value) * 32, 1)
nt_gemm_rowscan_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                       const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo,
                       GemmShape shape, typename Epi::Params ep) {
  using PT = PrecTraits<PREC>;
  using T = GemmTiles<PREC, BN, CTAS>;
  constexpr int NBUF = T::kAccBufs;
  constexpr int EW = EpiWarps<Epi>::value;
  static_assert(EW == 4 || EW == 8, "4 or 8 epilogue warps");
  static_assert(!(ARES && AMN), "the resident-A schedule stages A K-major");
  const int bn = shape.bn;                    // tile width (host: 0 < bn <= BN, multiple of 32, even chunk count if EW == 8)
  const int CPW = (bn / 32) / (EW / 4);       // 32-column chunks of a tile per epilogue warp
  const int b_rows = bn / CTAS;               // rows of B this CTA stages per tile
  const int kBTileRt = b_rows * GEMM_SWZ;     // bytes of one B plane of a stage (a multiple of 1024: b_rows % 16 == 0)
  const int kStage = PT::kPlanes * ((ARES ? 0 : T::kATile) + kBTileRt);
  const uint32_t stage_tx = (uint32_t)kStage;

  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B operand tiles need 1024-byte alignment.
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int STAGES = shape.stages;
  const int a_res_bytes = ARES ? PT::kPlanes * shape.k_chunks * T::kATile : 0;
  uint8_t* a_res = smem;                       // [plane][k_chunk][128 rows x 128 B]
  uint8_t* stage_base = smem + a_res_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(stage_base + STAGES * kStage);
  uint64_t* full_bar = bars;                                   // [8]
  uint64_t* empty_bar = bars + GEMM_MAX_STAGES;                // [8]
  uint64_t* acc_full = bars + 2 * GEMM_MAX_STAGES;             // [NBUF]
  uint64_t* acc_empty = acc_full + NBUF;                       // [NBUF]
  uint64_t* a_full = acc_empty + NBUF;                         // [1]
  uint64_t* a_empty = a_full + 1;                              // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_empty + 1);
  uint8_t* epi_smem = reinterpret_cast<uint8_t*>(bars) + GEMM_BAR_BYTES;

  constexpr int kBOff = ARES ? 0 : PT::kPlanes * T::kATile;  // offset of the B planes inside a stage

  auto stamp = [&](int slot) {
    if (shape.trace) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      shape.trace[(size_t)blockIdx.x * 8 + slot] = t;
    }
  };
  if (threadIdx.x == 0) stamp(0);
  const int warp = threadIdx.x >> 5;
  const uint32_t cta_rank = (CTAS == 2) ? cluster_ctarank() : 0u;
  const bool leader = (cta_rank == 0);
  const int unit = (CTAS == 2) ? (blockIdx.x >> 1) : blockIdx.x;   // persistent work unit (CTA or pair)
  const int n_units = (CTAS == 2) ? (gridDim.x >> 1) : gridDim.x;
  const int items_per_ks = shape.n_splits * shape.m_blocks;
  const int n_items = items_per_ks * shape.k_splits;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);      // the leader's arrive.expect_tx; both CTAs' TMA bytes land on it
      mbar_init(&empty_bar[s], 1);     // one tcgen05.commit
    }
    for (int b = 0; b < NBUF; ++b) {
      mbar_init(&acc_full[b], 1);                 // one tcgen05.commit
      mbar_init(&acc_empty[b], EW * CTAS);        // one arrive per epilogue warp of the unit
    }
    mbar_init(a_full, 1);
    mbar_init(a_empty, 1);
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
  if (threadIdx.x == 0) stamp(1);

  if (warp == 4) {
    // ===================== TMA producer (one lane) =====================
    if (lane_id() == 0) {
      uint32_t it = 0, n_item = 0;
      for (int item = unit; item < n_items; item += n_units, ++n_item) {
        const int ks = item / items_per_ks;
        const int rem = item - ks * items_per_ks;
        const int split = rem / shape.m_blocks;
        const int mb = rem - split * shape.m_blocks;
        const int kc0 = ks * shape.chunks_per_ksplit, kc1 = min(shape.k_chunks, kc0 + shape.chunks_per_ksplit);
        const int a_row = (mb * CTAS + (int)cta_rank) * GEMM_BM;
        const int t0 = split * shape.tiles_per_split;
        const int t1 = min(t0 + shape.tiles_per_split, shape.n_tiles);
        if (ARES) {
          // the previous item's MMAs must have finished reading the resident block
          mbar_wait(a_empty, (n_item & 1u) ^ 1u);
          // Only the leader arrives (with the byte count of BOTH CTAs).  The peer's TMA may complete_tx
          // before that arrive: the phase cannot complete until the pending arrival is in.  A peer-side
          // remote arrive would cost a cluster-scope release fence per stage, which serialises its TMAs.
          if (leader) mbar_arrive_expect_tx(a_full, (uint32_t)a_res_bytes * (uint32_t)CTAS);
          for (int kc = 0; kc < shape.k_chunks; ++kc) {
            const int kx = kc * PT::kChunkElems;
            uint8_t* dst = a_res + kc * T::kATile;
            if (CTAS == 1) {
              tma_load_2d(dst, &tm_a_hi, a_full, kx, a_row);
              if (PT::kPlanes == 2) tma_load_2d(dst + shape.k_chunks * T::kATile, &tm_a_lo, a_full, kx, a_row);
            } else {
              tma_load_2d_pair(dst, &tm_a_hi, a_full, kx, a_row);
              if (PT::kPlanes == 2) tma_load_2d_pair(dst + shape.k_chunks * T::kATile, &tm_a_lo, a_full, kx, a_row);
            }
          }
        }
        for (int t = t0; t < t1; ++t) {
          const int b_row = t * bn + (int)cta_rank * b_rows;
          // warm L2 with this CTA's rows of the B tile two tiles ahead: the ring only buffers ~1.3 us of MMA work,
          // less than a DRAM round trip under load, so first-touch tiles would otherwise stall the tensor pipe
          if (shape.l2_prefetch && t + 2 < t1) {
            const int p_row = (t + 2) * bn + (int)cta_rank * b_rows;
            for (int kc = kc0; kc < kc1; ++kc) {
              tma_prefetch_l2_2d(&tm_b_hi, kc * PT::kChunkElems, p_row);
              if (PT::kPlanes == 2) tma_prefetch_l2_2d(&tm_b_lo, kc * PT::kChunkElems, p_row);
            }
          }
          for (int kc = kc0; kc < kc1; ++kc, ++it) {
            const int s = it % STAGES;
            const uint32_t ph = (it / STAGES) & 1u;
            mbar_wait(&empty_bar[s], ph ^ 1u);
            uint8_t* st = stage_base + s * kStage;
            const int kx = kc * PT::kChunkElems;
            // Pairs: both CTAs' bytes are credited to the leader's barrier (see a_full above).
            if (CTAS == 1) mbar_arrive_expect_tx(&full_bar[s], stage_tx);
            else if (leader) mbar_arrive_expect_tx(&full_bar[s], stage_tx * 2);
            for (int pl = 0; pl < PT::kPlanes; ++pl) {
              const CUtensorMap* ta = pl ? &tm_a_lo : &tm_a_hi;
              const CUtensorMap* tb = pl ? &tm_b_lo : &tm_b_hi;
              if (!ARES) {
                uint8_t* da = st + pl * T::kATile;
                if (AMN) tma_load_mn_slabs<PREC, CTAS>(da, ta, &full_bar[s], a_row, GEMM_BM, kx);
                else if (CTAS == 1) tma_load_2d(da, ta, &full_bar[s], kx, a_row);
                else tma_load_2d_pair(da, ta, &full_bar[s], kx, a_row);
              }
              uint8_t* db = st + kBOff + pl * kBTileRt;
              if (BMN) tma_load_mn_slabs<PREC, CTAS>(db, tb, &full_bar[s], b_row, b_rows, kx);
              else if (CTAS == 1) tma_load_2d(db, tb, &full_bar[s], kx, b_row);
              else tma_load_2d_pair(db, tb, &full_bar[s], kx, b_row);
            }
          }
        }
      }
    }
  } else if (warp == 5) {
    // ===================== MMA issuer (one lane of the leader CTA) =====================
    if (leader && lane_id() == 0) {
      const uint32_t idesc = make_idesc(PT::kFmt, GEMM_BM * CTAS, (uint32_t)bn) | ((uint32_t)AMN << 15) | ((uint32_t)BMN << 16);
      constexpr uint32_t kSlabBytes = PT::kChunkElems * GEMM_SWZ;
      // descriptor start-address step per MMA, in 16-byte units: 32 B of K inside the swizzle row (K-major) or one
      // MMA's worth of K rows of 128 B each (MN-major: 8 rows for TF32, 16 for bf16)
      constexpr uint64_t kAdvA = AMN ? (uint64_t)((32 / PT::kElemBytes) * GEMM_SWZ / 16) : 2ull;
      constexpr uint64_t kAdvB = BMN ? (uint64_t)((32 / PT::kElemBytes) * GEMM_SWZ / 16) : 2ull;
      uint32_t it = 0, tc = 0, n_item = 0;
      for (int item = unit; item < n_items; item += n_units, ++n_item) {
        const int ks = item / items_per_ks;
        const int split = (item - ks * items_per_ks) / shape.m_blocks;
        const int kc0 = ks * shape.chunks_per_ksplit, kc1 = min(shape.k_chunks, kc0 + shape.chunks_per_ksplit);
        const int t0 = split * shape.tiles_per_split;
        const int t1 = min(t0 + shape.tiles_per_split, shape.n_tiles);
        if (ARES) {
          mbar_wait(a_full, n_item & 1u);
          tc_fence_after_sync();
        }
        for (int t = t0; t < t1; ++t, ++tc) {
          const int buf = tc % NBUF;
          const uint32_t aph = (tc / NBUF) & 1u;
          mbar_wait(&acc_empty[buf], aph ^ 1u);
          tc_fence_after_sync();
          const uint32_t d_tmem = tmem_base + (uint32_t)(buf * BN);
          for (int kc = kc0; kc < kc1; ++kc, ++it) {
            const int s = it % STAGES;
            const uint32_t ph = (it / STAGES) & 1u;
            mbar_wait(&full_bar[s], ph);
            tc_fence_after_sync();
            if (it == 0) stamp(2);
            const uint32_t st = smem_u32(stage_base + s * kStage);
            const uint32_t a_addr = ARES ? smem_u32(a_res + kc * T::kATile) : st;
            const uint32_t a_lo_addr = ARES ? a_addr + (uint32_t)(shape.k_chunks * T::kATile) : st + T::kATile;
            const uint64_t a_hi = AMN ? make_mnmajor_desc<PT::kTf32>(a_addr, kSlabBytes) : make_kmajor_desc<GEMM_SWZ>(a_addr);
            const uint64_t b_hi = BMN ? make_mnmajor_desc<PT::kTf32>(st + kBOff, kSlabBytes) : make_kmajor_desc<GEMM_SWZ>(st + kBOff);
#pragma unroll
            for (int kstep = 0; kstep < PT::kKSteps; ++kstep) {
              const uint64_t adva = (uint64_t)kstep * kAdvA, advb = (uint64_t)kstep * kAdvB;
              if (PT::kPlanes == 2) {
                const uint64_t a_lo = AMN ? make_mnmajor_desc<PT::kTf32>(a_lo_addr, kSlabBytes) : make_kmajor_desc<GEMM_SWZ>(a_lo_addr);
                const uint64_t b_lo = BMN ? make_mnmajor_desc<PT::kTf32>(st + kBOff + kBTileRt, kSlabBytes)
                                          : make_kmajor_desc<GEMM_SWZ>(st + kBOff + kBTileRt);
                // small cross terms first, dominant term last
                tc_mma<CTAS, PT::kTf32>(d_tmem, a_lo + adva, b_hi + advb, idesc, ((kc - kc0) | kstep) != 0);
                tc_mma<CTAS, PT::kTf32>(d_tmem, a_hi + adva, b_lo + advb, idesc, 1u);
                tc_mma<CTAS, PT::kTf32>(d_tmem, a_hi + adva, b_hi + advb, idesc, 1u);
              } else {
                tc_mma<CTAS, PT::kTf32>(d_tmem, a_hi + adva, b_hi + advb, idesc, ((kc - kc0) | kstep) != 0);
              }
            }
            tc_commit<CTAS>(&empty_bar[s]);   // smem slot reusable once these MMAs retire
          }
          tc_commit<CTAS>(&acc_full[buf]);    // accumulator tile complete
        }
        if (ARES) tc_commit<CTAS>(a_empty);   // resident A block reusable once the item's MMAs retire
      }
      stamp(3);
    }
  } else {
    // ===================== epilogue warps 0..3 (and 6..9 when EW == 8) =====================
    const int quarter = warp & 3;                      // the TMEM lane quarter a warp may read is warp id % 4
    const int half = (EW == 8 && warp >= 6) ? 1 : 0;   // which half of every tile's chunks this warp consumes
    const int row = quarter * 32 + (int)lane_id();     // 0..127 == TMEM lane
    Epi epi = make_epi<Epi, EW>(ep, epi_smem, row, half);
    const uint32_t lane_base = ((uint32_t)(quarter * 32)) << 16;
    const int cb = half * CPW;
    uint32_t tc = 0;
    for (int item = unit; item < n_items; item += n_units) {
      const int ks = item / items_per_ks;
      const int rem = item - ks * items_per_ks;
      const int slot = rem / shape.m_blocks + ks * shape.n_splits;    // slot index: K split major
      const int split = EW == 8 ? 2 * slot + half : slot;
      const int mb = rem % shape.m_blocks;
      const int m_row = (mb * CTAS + (int)cta_rank) * GEMM_BM + row;
      const int t0 = (rem / shape.m_blocks) * shape.tiles_per_split;
      const int t1 = min(t0 + shape.tiles_per_split, shape.n_tiles);
      epi.begin_item(m_row, split, t0 * bn);
      for (int t = t0; t < t1; ++t, ++tc) {
        const int buf = tc % NBUF;
        const uint32_t aph = (tc / NBUF) & 1u;
        epi.begin_tile(t * bn);
        mbar_wait(&acc_full[buf], aph);
        tc_fence_after_sync();
        if (tc == 0 && threadIdx.x == 0) stamp(4);
        const uint32_t taddr = tmem_base + lane_base + (uint32_t)(buf * BN);
        uint32_t va[32], vb[32];
        tmem_ld32(taddr + (uint32_t)(cb * 32), va);
        int c = cb;
#pragma unroll 1
        for (; c + 1 < cb + CPW; c += 2) {
          tmem_ld_wait(va);
          tmem_ld32(taddr + (uint32_t)((c + 1) * 32), vb);
          epi.consume(t * bn + c * 32, va, taddr + (uint32_t)(c * 32), vb);
          tmem_ld_wait(vb);
          if (c + 2 < cb + CPW) tmem_ld32(taddr + (uint32_t)((c + 2) * 32), va);
          epi.consume(t * bn + (c + 1) * 32, vb, taddr + (uint32_t)((c + 1) * 32), va);
        }
        if (c < cb + CPW) {   // odd chunk count (narrow tiles): the last chunk is already in flight in `va`
          tmem_ld_wait(va);
          epi.consume(t * bn + c * 32, va, taddr + (uint32_t)(c * 32), vb);
        }
        // all of this warp's TMEM reads of `buf` have completed (wait::ld above)
        tc_fence_before_sync();
        __syncwarp();
        if (lane_id() == 0) {
          if (CTAS == 1 || leader) mbar_arrive(&acc_empty[buf]);
          else mbar_arrive_cluster_relaxed(&acc_empty[buf], 0);
        }
      }
      epi.end_item(m_row, split);
    }
    if (threadIdx.x == 0) stamp(5);
  }

  __syncwarp();  // single-lane roles rejoin their warp before the aligned barriers below
  tc_fence_before_sync();
  if (threadIdx.x == 0) stamp(6);
  if (CTAS == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 5) {
    tc_fence_after_sync();
    tmem_dealloc<CTAS>(tmem_base, 512);
  }
  if (threadIdx.x == 0) stamp(7);
}

// Host-side launch of one instantiation.  maps = {A hi, A lo, B hi, B lo}; single-plane modes pass
// the hi map twice.  n_units = persistent CTAs (CTAS == 1) or CTA pairs (CTAS == 2).
// shape.stages is filled in here from the shared-memory plan.
template <int PREC, int BN, int CTAS, int ARES, class Epi, int AMN = 0, int BMN = 0>
int launch_nt_gemm(const CUtensorMap* maps, GemmShape shape, const typename Epi::Params& ep, int n_units,
                   cudaStream_t stream) {
  if (shape.bn <= 0) shape.bn = BN;
  {
    constexpr int kSlab = GEMM_SWZ / PrecTraits<PREC>::kElemBytes;
    const int chunks = shape.bn / 32;
    DIF_REQUIRE(shape.bn <= BN && shape.bn % 32 == 0 && (EpiWarps<Epi>::value == 4 || chunks % 2 == 0) &&
                    (!BMN || (shape.bn / CTAS) % kSlab == 0) && (CTAS == 1 || shape.bn % 16 == 0),
                DIF_ERR_INVALID, "nt_gemm: tile width %d is not valid for this instantiation (BN %d, %d CTAs, MN-major B %d)",
                shape.bn, BN, CTAS, BMN);
  }
  GemmSmemPlan plan;
  const bool fits = plan_gemm_smem<PREC, BN, CTAS, ARES>(shape.k_chunks, Epi::smem_bytes(ep), &plan, shape.bn);
  DIF_REQUIRE(fits, DIF_ERR_CAPACITY, "nt_gemm: K = %d chunks does not fit the shared-memory plan (ARES=%d)",
              shape.k_chunks, ARES);
  shape.stages = plan.stages;
  if (shape.k_splits <= 1) {
    shape.k_splits = 1;
    shape.chunks_per_ksplit = shape.k_chunks;
  }
  DIF_REQUIRE(!(ARES && shape.k_splits > 1), DIF_ERR_INVALID, "nt_gemm: the resident-A schedule does not split K");
  auto kern = nt_gemm_rowscan_kernel<PREC, BN, CTAS, ARES, Epi, AMN, BMN>;
  static bool configured = false;   // per instantiation
  if (!configured) {
    DIF_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_MAX));
    configured = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(n_units * CTAS));
  cfg.blockDim = dim3((2 + EpiWarps<Epi>::value) * 32);
  cfg.dynamicSmemBytes = (size_t)plan.total;
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
