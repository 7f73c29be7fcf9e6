// ArcFace additive-angular-margin logits + softmax cross-entropy, forward + backward.
// ABSENT from the reference; specification in DESIGN.md and oracle/losses_oracle.py:arcface (arXiv 1801.07698):
//   xh = l2_normalize(x), wh = l2_normalize(w), cos = clip(xh . wh, -1, 1)
//   target logit s * phi(cos_y), phi = cos(theta + m) if cos_y > cos(pi - m) else cos_y - m sin(pi - m); others s * cos
//   loss_b = logsumexp(logits_b) - logits_b[y_b];   dX, dW = gradients of sum_b dloss_b * loss_b (default 1/B)
//
// All three contractions run on the tcgen05/TMA GEMM skeleton (nt_gemm.cuh) in 3xTF32 / 3xBF16 (fp32-exact to ~1e-6);
// a step is 6 launches and NO operand is transposed in HBM - the backward GEMMs read the planes the forward pass
// made through MN-major shared-memory descriptors:
//   1. prep        x and w rows -> normalised operand planes (one launch for both)
//   2. arc_fwd     cos = xh wh^T with a fused epilogue: clip, margin on the target column, scale, online softmax
//                  (running max / sum per class split) in registers; cos is also written out once.  The tile
//                  width is chosen per shape so the class tiles fill the SMs in whole waves (plan_tiles)
//   3. arc_dcos    merges the per-split (max, sum) pairs -> logZ, loss, d phi / d cos, then p - onehot -> d cos as
//                  operand planes [B][C], written once
//   4. dxh = dcos wh        A = dcos K-major, B = wh planes [C][D] read MN-major; split over K = classes
//      dwh = dcos^T xh      A = dcos read MN-major, B = xh planes [B][D] read MN-major
//   5. arc_norm_bwd  l2_normalize backward for the X rows and the W rows (one launch)
#include <algorithm>

#include "prep_rows.cuh"
#include "store_epi.cuh"

namespace dif {

constexpr int kArcBN = 256;
constexpr int kArcCtas = 2;   // CTA pairs: 64 KB stages -> a 3-deep ring instead of 2 (the 1-CTA 3xTF32 loop was TMA-latency bound)
constexpr float kLog2e = 1.4426950408889634f;

struct ArcMargin {
  float s, cos_m, sin_m, th, mm;
};

__device__ __forceinline__ float arc_phi(float ct, const ArcMargin& a) {
  return ct > a.th ? ct * a.cos_m - sqrtf(fmaxf(1.f - ct * ct, 0.f)) * a.sin_m : ct - a.mm;
}
__device__ __forceinline__ float arc_dphi(float ct, const ArcMargin& a) {
  if (!(ct > a.th)) return 1.f;
  const float st = sqrtf(fmaxf(1.f - ct * ct, 0.f));
  return a.cos_m + (st > 0.f ? ct / st : 0.f) * a.sin_m;
}

// Forward epilogue: thread = one sample, columns = classes.
struct ArcFwdEpi {
  struct Params {
    float* cosbuf;          // [B][ldc] raw (unclipped) cosines
    const int32_t* y;       // [B]
    float2* part;           // [B][n_slots] (running max, running sum) in log2 units of the scaled logits
    int B, C, ldc, n_slots;
    ArcMargin mg;
  };
  static int smem_bytes(const Params&) { return 16; }
  const Params& p;
  float* row_ptr;
  float m2, l2;   // running max / sum of 2^(logit * log2e - m2)
  int yb;
  __device__ ArcFwdEpi(const Params& pp, uint8_t*, int) : p(pp), row_ptr(nullptr), m2(-INFINITY), l2(0.f), yb(-1) {}
  __device__ void begin_item(int m, int, int) {
    row_ptr = m < p.B ? p.cosbuf + (size_t)m * p.ldc : nullptr;
    yb = m < p.B ? p.y[m] : -1;
    m2 = -INFINITY;
    l2 = 0.f;
  }
  __device__ void begin_tile(int) {}
  __device__ void consume(int col0, const uint32_t (&acc)[32], uint32_t, uint32_t (&)[32]) {
    if (!row_ptr) return;
    float z[32];
    const float k2 = p.mg.s * kLog2e;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const float raw = __uint_as_float(acc[i]);
      const float c = fminf(fmaxf(raw, -1.f), 1.f);
      z[i] = (col0 + i < p.C) ? c * k2 : -INFINITY;
    }
    if (yb >= col0 && yb < col0 + 32) {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (col0 + i == yb) z[i] = arc_phi(fminf(fmaxf(__uint_as_float(acc[i]), -1.f), 1.f), p.mg) * k2;
    }
    if (col0 + 32 <= p.ldc) {
      float4* dst = reinterpret_cast<float4*>(row_ptr + col0);
#pragma unroll
      for (int i = 0; i < 8; ++i)
        dst[i] = make_float4(__uint_as_float(acc[4 * i]), __uint_as_float(acc[4 * i + 1]), __uint_as_float(acc[4 * i + 2]),
                             __uint_as_float(acc[4 * i + 3]));
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (col0 + i < p.ldc) row_ptr[col0 + i] = __uint_as_float(acc[i]);
    }
    float mx = z[0];
#pragma unroll
    for (int i = 1; i < 32; ++i) mx = fmaxf(mx, z[i]);
    if (mx == -INFINITY) return;
    const float mn = fmaxf(m2, mx);
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) sum += exp2f(z[i] - mn);
    l2 = l2 * exp2f(m2 - mn) + sum;
    m2 = mn;
  }
  __device__ void end_item(int m, int slot) {
    if (m < p.B) p.part[(size_t)m * p.n_slots + slot] = make_float2(m2, l2);
  }
};

// Operand planes: T = float -> TF32 hi + exact residual lo (mode 0); T = bf16 -> b0 = bf16(v), b1 = bf16(v - b0) (mode 3)
__device__ __forceinline__ void split_store(float v, float* hi, float* lo, size_t i) {
  const float h = tf32_round(v);
  hi[i] = h;
  lo[i] = __fsub_rn(v, h);
}
__device__ __forceinline__ void split_store(float v, __nv_bfloat16* p0, __nv_bfloat16* p1, size_t i) {
  const __nv_bfloat16 b0 = __float2bfloat16_rn(v);
  p0[i] = b0;
  p1[i] = __float2bfloat16_rn(__fsub_rn(v, __bfloat162float(b0)));
}

__device__ __forceinline__ void split_store4(const float (&v)[4], float* hi, float* lo, size_t i) {
  float4 h, l;
  h.x = tf32_round(v[0]); h.y = tf32_round(v[1]); h.z = tf32_round(v[2]); h.w = tf32_round(v[3]);
  l.x = __fsub_rn(v[0], h.x); l.y = __fsub_rn(v[1], h.y); l.z = __fsub_rn(v[2], h.z); l.w = __fsub_rn(v[3], h.w);
  *reinterpret_cast<float4*>(hi + i) = h;
  *reinterpret_cast<float4*>(lo + i) = l;
}
__device__ __forceinline__ void split_store4(const float (&v)[4], __nv_bfloat16* p0, __nv_bfloat16* p1, size_t i) {
  __nv_bfloat16 a[4], b[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    a[k] = __float2bfloat16_rn(v[k]);
    b[k] = __float2bfloat16_rn(__fsub_rn(v[k], __bfloat162float(a[k])));
  }
  *reinterpret_cast<uint2*>(p0 + i) = *reinterpret_cast<const uint2*>(a);
  *reinterpret_cast<uint2*>(p1 + i) = *reinterpret_cast<const uint2*>(b);
}

// Loss + d cos.  Block = kDcosRows samples x a chunk of classes.  One warp per sample first merges the forward
// pass's per-split (max, sum) pairs -> logZ, target phi / d phi (and the loss, written by the blocks of class chunk 0);
// then the block streams its class chunk: cos -> p - onehot -> d cos, split into the two operand planes.
constexpr int kDcosRows = 8;
template <class T>
__global__ void __launch_bounds__(256) arc_dcos_kernel(const float2* __restrict__ part, int n_slots,
                                                       const float* __restrict__ cosbuf, int ldc,
                                                       const int32_t* __restrict__ y, const float* __restrict__ dloss,
                                                       int B, int C, int chunk, ArcMargin mg, float* __restrict__ loss,
                                                       T* __restrict__ d_hi, T* __restrict__ d_lo) {
  __shared__ float s_logz[kDcosRows], s_dphi[kDcosRows], s_g[kDcosRows];
  __shared__ int s_y[kDcosRows];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b0 = blockIdx.y * kDcosRows;
  if (warp < kDcosRows) {
    const int b = b0 + warp;
    if (b < B) {
      float m2 = -INFINITY, l2 = 0.f;
      for (int s = lane; s < n_slots; s += 32) {
        const float2 q = part[(size_t)b * n_slots + s];
        if (q.x == -INFINITY) continue;
        const float mn = fmaxf(m2, q.x);
        l2 = l2 * exp2f(m2 - mn) + q.y * exp2f(q.x - mn);
        m2 = mn;
      }
      for (int o = 16; o >= 1; o >>= 1) {
        const float om = __shfl_xor_sync(0xffffffffu, m2, o), ol = __shfl_xor_sync(0xffffffffu, l2, o);
        const float mn = fmaxf(m2, om);
        l2 = (m2 == -INFINITY ? 0.f : l2 * exp2f(m2 - mn)) + (om == -INFINITY ? 0.f : ol * exp2f(om - mn));
        m2 = mn;
      }
      if (lane == 0) {
        const float lz = (m2 + log2f(l2)) / kLog2e;
        const int yb = y[b];
        const float ct = fminf(fmaxf(cosbuf[(size_t)b * ldc + yb], -1.f), 1.f);
        if (blockIdx.x == 0) loss[b] = lz - mg.s * arc_phi(ct, mg);
        s_logz[warp] = lz;
        s_dphi[warp] = arc_dphi(ct, mg);
        s_g[warp] = dloss ? dloss[b] : 1.f / (float)B;
        s_y[warp] = yb;
      }
    } else if (lane == 0) {
      s_y[warp] = -1;
    }
  }
  __syncthreads();
  if (!d_hi) return;
  // Four classes per thread and step, the kDcosRows row loads issued together: 8 x 16 B in flight per thread (with one
  // 4-byte load at a time the 20 MB of cosines arrive at ~1 TB/s - the kernel was a chain of DRAM round trips).
  const int c_begin = blockIdx.x * chunk, c_end = min(ldc, c_begin + chunk);   // ldc and chunk are multiples of 4
  const int n_rows = min(kDcosRows, B - b0);
  for (int c = c_begin + 4 * (int)threadIdx.x; c < c_end; c += 4 * 256) {
    float4 raw[kDcosRows];
#pragma unroll
    for (int r = 0; r < kDcosRows; ++r)
      raw[r] = r < n_rows ? *reinterpret_cast<const float4*>(cosbuf + (size_t)(b0 + r) * ldc + c) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int r = 0; r < kDcosRows; ++r) {
      if (r >= n_rows) break;
      const float in[4] = {raw[r].x, raw[r].y, raw[r].z, raw[r].w};
      float v[4];
      const float sg = mg.s * s_g[r], lz = s_logz[r];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float ct = fminf(fmaxf(in[i], -1.f), 1.f);
        v[i] = sg * expf(mg.s * ct - lz);                         // every class but the target: s g p
        if (c + i >= C || in[i] < -1.f || in[i] > 1.f) v[i] = 0.f;   // padding; clip_by_value passes no gradient outside
      }
      const int yt = s_y[r] - c;
      if (yt >= 0 && yt < 4) {                                     // the target column: margin logit, d phi / d cos
        const float ct = fminf(fmaxf(in[yt], -1.f), 1.f);
        const float pr = expf(mg.s * arc_phi(ct, mg) - lz);
        v[yt] = (in[yt] < -1.f || in[yt] > 1.f) ? 0.f : sg * (pr - 1.f) * s_dphi[r];
      }
      split_store4(v, d_hi, d_lo, (size_t)(b0 + r) * ldc + c);
    }
  }
}

struct NormBwdJob {
  const float* g;        // [planes][R][D] partial gradients wrt the normalised rows
  int planes;
  size_t plane_stride;
  const float* src;      // [R][D] the ORIGINAL rows: the normalised row is src * inv, exactly what prep stored (x * inv
                         // rounded once; its hi + lo planes add back to the same float), at half the bytes of two planes
  const float* inv;      // [R] inverse norms
  int R;
  float* out;            // [R][D]
};

template <int TPR>
__global__ void __launch_bounds__(256) arc_norm_bwd_kernel(NormBwdJob j0, NormBwdJob j1, int blocks0, int D) {
  __shared__ float red[8];
  constexpr int kRows = 256 / TPR;
  const bool first = (int)blockIdx.x < blocks0;
  const NormBwdJob& j = first ? j0 : j1;
  const int r = ((int)blockIdx.x - (first ? 0 : blocks0)) * kRows + (int)threadIdx.x / TPR;
  const int lane = (int)threadIdx.x % TPR;
  const bool live = r < j.R;
  const size_t base = (size_t)r * D;
  float dot = 0.f;
  float4 a[4], h[4];
  const float iv = live ? j.inv[r] : 0.f;
  // all eight 16-byte loads of a thread are issued before the first use (the class-weight rows are 60 MB of traffic)
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const int d = (lane + t * TPR) * 4;
    const bool on = live && d < D;
    a[t] = on ? *reinterpret_cast<const float4*>(j.g + base + d) : make_float4(0.f, 0.f, 0.f, 0.f);
    h[t] = on ? *reinterpret_cast<const float4*>(j.src + base + d) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const int d = (lane + t * TPR) * 4;
    if (live && d < D) {
      for (int s = 1; s < j.planes; ++s) {   // split-K partial planes, summed in a fixed order
        const float4 v = *reinterpret_cast<const float4*>(j.g + (size_t)s * j.plane_stride + base + d);
        a[t].x += v.x; a[t].y += v.y; a[t].z += v.z; a[t].w += v.w;
      }
      h[t] = make_float4(__fmul_rn(h[t].x, iv), __fmul_rn(h[t].y, iv), __fmul_rn(h[t].z, iv), __fmul_rn(h[t].w, iv));
      dot += a[t].x * h[t].x + a[t].y * h[t].y + a[t].z * h[t].z + a[t].w * h[t].w;
    }
  }
  for (int o = 16; o >= 1; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
  if (TPR > 32) {
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = dot;
    __syncthreads();
    dot = 0.f;
    for (int w = 0; w < 8; ++w) dot += red[w];
  }
  if (!live) return;
  const bool clamped = iv >= 0.99e6f;   // 1 / sqrt(1e-12)
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const int d = (lane + t * TPR) * 4;
    if (d < D) {
      float4 o4;
      o4.x = clamped ? iv * a[t].x : iv * (a[t].x - h[t].x * dot);
      o4.y = clamped ? iv * a[t].y : iv * (a[t].y - h[t].y * dot);
      o4.z = clamped ? iv * a[t].z : iv * (a[t].z - h[t].z * dot);
      o4.w = clamped ? iv * a[t].w : iv * (a[t].w - h[t].w * dot);
      *reinterpret_cast<float4*>(j.out + base + d) = o4;
    }
  }
}

static void launch_norm_bwd(const NormBwdJob& j0, const NormBwdJob& j1, int D, cudaStream_t st) {
  if (D <= 512) {
    const int b0 = (j0.R + 7) / 8, b1 = (j1.R + 7) / 8;
    arc_norm_bwd_kernel<32><<<b0 + b1, 256, 0, st>>>(j0, j1, b0, D);
  } else {
    arc_norm_bwd_kernel<256><<<j0.R + j1.R, 256, 0, st>>>(j0, j1, j0.R, D);
  }
}

// Tile width and split for one GEMM: columns `N` in tiles of `bn`, m_blocks row blocks, `units` persistent CTA pairs.
// An item = (row block, tiles_per_split consecutive tiles).  Cost model per K step of one SM of a pair (clocks):
// max(tensor time 1.59 bn, shared-memory port time 160 + 0.625 bn) - below ~166 columns a tile is bound by the
// 128 B/clk port (the A block is re-read for every tile), so narrower is not automatically better.  The plan with the
// lowest waves x tiles_per_split x cost(bn) wins.
struct TilePlan {
  int bn, n_tiles, n_splits, tiles_per_split;
};
static TilePlan plan_tiles(int N, int m_blocks, int units, int bn_step, bool single_tile_items) {
  TilePlan best{256, (N + 255) / 256, 1, (N + 255) / 256};
  double best_cost = 1e30;
  for (int bn = 256; bn >= 64; bn -= bn_step) {
    const int n_tiles = (N + bn - 1) / bn;
    for (int tps = 1; tps <= n_tiles; ++tps) {
      if (single_tile_items && tps > 1) break;
      const int n_splits = (n_tiles + tps - 1) / tps;
      const int waves = (m_blocks * n_splits + units - 1) / units;
      const double cyc = std::max(1.59 * bn, 160.0 + 0.625 * bn);
      const double cost = (double)waves * tps * cyc * (1.0 + 0.02 * n_splits / 40.0);   // mild preference for fewer slots
      if (cost < best_cost) {
        best_cost = cost;
        best = TilePlan{bn, n_tiles, n_splits, tps};
      }
    }
  }
  return best;
}

struct ArcWorkspace {
  char* base = nullptr;
  size_t bytes = 0;
  int ensure(size_t need) {
    if (need <= bytes) return DIF_OK;
    retire_device_block(base);
    base = nullptr;
    bytes = 0;
    DIF_CUDA_OK(cudaMalloc((void**)&base, need));
    bytes = need;
    return DIF_OK;
  }
};
static thread_local ArcWorkspace g_arc;

// One operand's tensor maps (hi, lo).  K-major: global [rows][K], box [box_rows][one 128-byte K chunk].
// MN-major: global [K][rows] (row index contiguous), box [K chunk rows][one 128-byte slab of rows].
template <class T>
static int operand_maps(CUtensorMap* hi, CUtensorMap* lo, const T* p_hi, const T* p_lo, int rows, int K, int pitch,
                        int box_rows, bool mn_major) {
  constexpr int esz = (int)sizeof(T), bf = esz == 2 ? 1 : 0;
  constexpr uint32_t cols = 128 / esz;
  if (mn_major) {   // fp32 planes: the 32-byte-atom swizzle (see make_mnmajor_desc)
    if (int rc = make_tmap_2d(hi, p_hi, K, rows, (uint64_t)pitch * esz, cols, cols, bf, !bf)) return rc;
    return make_tmap_2d(lo, p_lo, K, rows, (uint64_t)pitch * esz, cols, cols, bf, !bf);
  }
  if (int rc = make_tmap_2d(hi, p_hi, rows, K, (uint64_t)pitch * esz, box_rows, cols, bf)) return rc;
  return make_tmap_2d(lo, p_lo, rows, K, (uint64_t)pitch * esz, box_rows, cols, bf);
}

// PREC 0: TF32 hi/lo planes (3xTF32);  PREC 3: bf16 b0/b1 planes (3xBF16, half the plane bytes and MMA time)
template <int PREC>
static int run_arcface(const float* X, const float* W, const int32_t* y, int B, int C, int D, float s, float m,
                       float* loss, const float* dloss, float* dX, float* dW, cudaStream_t st) {
  using T = typename std::conditional<PREC == 3, __nv_bfloat16, float>::type;
  constexpr int esz = (int)sizeof(T), kChunk = 128 / esz;
  const bool bwd = dX != nullptr;
  const int sms = device_sm_count();
  const int Cp = (C + 7) & ~7;   // d cos / cos pitch: a 16-byte multiple for TMA in both modes
  ArcMargin mg{s, cosf(m), sinf(m), cosf(3.14159265358979323846f - m), sinf(3.14159265358979323846f - m) * m};

  // ---- GEMM shapes
  const int bm = GEMM_BM * kArcCtas, units = std::max(1, sms / kArcCtas);
  GemmShape fwd{};
  fwd.m_blocks = (B + bm - 1) / bm;
  const TilePlan pf = plan_tiles(C, fwd.m_blocks, units, 32, false);
  fwd.bn = pf.bn;
  fwd.n_tiles = pf.n_tiles;
  fwd.n_splits = pf.n_splits;
  fwd.tiles_per_split = pf.tiles_per_split;
  fwd.k_chunks = (D + kChunk - 1) / kChunk;
  constexpr int kMnStep = 2 * kChunk;   // MN-major B in a CTA pair: each CTA's half tile is a whole number of slabs
  // The two backward GEMMs are independent (both read dcos): they run SIDE BY SIDE on disjoint sets of CTA pairs, on
  // two streams, so the idle pairs and the ramp / drain of one are covered by the other.  The split follows the work:
  // dxh is 2 B C D over a long K (classes), dwh the same flops over many row blocks.
  static const bool serial = getenv("DIF_ARC_SERIAL") != nullptr;   // development aid: one after the other, all pairs each
  const bool side_by_side = bwd && !serial && units >= 8;
  // (46 % of the pairs to dxh measured best at C2: 36 % 131.5 us, 41 % 125.6, 46 % 123.5, 52 % 150.5 - there dwh's 120
  // items need a fourth wave on the remaining 36 pairs)
  static const int x_pct = getenv("DIF_ARC_XPCT") ? atoi(getenv("DIF_ARC_XPCT")) : 46;   // development aid
  const int units_x = side_by_side ? std::max(2, (units * x_pct) / 100) : units, units_w = side_by_side ? units - units_x : units;
  GemmShape gx{};   // dxh [B, D] = dcos [B, C] x wh [C, D], split over K = classes
  gx.m_blocks = fwd.m_blocks;
  gx.bn = D <= 128 ? 128 : 256;
  gx.n_tiles = (D + gx.bn - 1) / gx.bn;
  gx.k_chunks = (C + kChunk - 1) / kChunk;
  gx.n_splits = gx.n_tiles;
  gx.tiles_per_split = 1;
  gx.k_splits = std::max(1, std::min(gx.k_chunks, units_x / std::max(1, gx.m_blocks * gx.n_tiles)));
  gx.chunks_per_ksplit = (gx.k_chunks + gx.k_splits - 1) / gx.k_splits;
  gx.k_splits = (gx.k_chunks + gx.chunks_per_ksplit - 1) / gx.chunks_per_ksplit;
  GemmShape gw{};   // dwh [C, D] = dcos^T [C, B] x xh [B, D]
  gw.m_blocks = (C + bm - 1) / bm;
  const TilePlan pw = plan_tiles(D, gw.m_blocks, units_w, kMnStep, false);
  gw.bn = pw.bn;
  gw.n_tiles = pw.n_tiles;
  gw.n_splits = pw.n_splits;
  gw.tiles_per_split = pw.tiles_per_split;
  gw.k_chunks = (B + kChunk - 1) / kChunk;

  // ---- workspace carve-up (256-byte aligned pieces)
  size_t off = 0;
  auto take = [&](size_t bytes) {
    const size_t at = off;
    off += (bytes + 255) & ~(size_t)255;
    return at;
  };
  // mode 0: xh / xl = TF32 hi / lo planes.  mode 3: xh / xl = bf16 b0 / b1 planes
  const size_t o_xh = take((size_t)B * D * esz), o_xl = take((size_t)B * D * esz), o_xi = take((size_t)B * 4);
  const size_t o_wh = take((size_t)C * D * esz), o_wl = take((size_t)C * D * esz), o_wi = take((size_t)C * 4);
  const size_t o_cos = take((size_t)B * Cp * 4);
  const size_t o_part = take((size_t)B * fwd.n_splits * 8);
  size_t o_dh = 0, o_dl = 0, o_gx = 0, o_gw = 0;
  if (bwd) {
    o_dh = take((size_t)B * Cp * esz); o_dl = take((size_t)B * Cp * esz);
    o_gx = take((size_t)gx.k_splits * B * D * 4);
    o_gw = take((size_t)C * D * 4);
  }
  if (int rc = g_arc.ensure(off)) return rc;
  char* ws = g_arc.base;
  auto F = [&](size_t o) { return reinterpret_cast<float*>(ws + o); };
  auto P = [&](size_t o) { return reinterpret_cast<T*>(ws + o); };

  // ---- 1. normalise + operand planes, X and W in one launch
  PrepParams px{};
  px.src = X; px.n = B; px.D = D; px.normalize = 1; px.inv = F(o_xi);
  PrepParams pw2 = px;
  pw2.src = W; pw2.n = C; pw2.inv = F(o_wi);
  if (PREC == 3) {
    px.split = 0; px.p0 = nullptr;
    px.pb = reinterpret_cast<__nv_bfloat16*>(ws + o_xh); px.pb1 = reinterpret_cast<__nv_bfloat16*>(ws + o_xl);
    pw2.split = 0; pw2.p0 = nullptr;
    pw2.pb = reinterpret_cast<__nv_bfloat16*>(ws + o_wh); pw2.pb1 = reinterpret_cast<__nv_bfloat16*>(ws + o_wl);
  } else {
    px.split = 1; px.p0 = F(o_xh); px.p1 = F(o_xl);
    pw2.split = 1; pw2.p0 = F(o_wh); pw2.p1 = F(o_wl);
  }
  // DIF_ARC_PROFILE=1 (development aid): CUDA events between the launches, printed after a synchronise
  static const bool profile = getenv("DIF_ARC_PROFILE") != nullptr;
  static cudaEvent_t pev[8];
  static bool pev_made = false;
  if (profile && !pev_made) {
    for (cudaEvent_t& e : pev) cudaEventCreate(&e);
    pev_made = true;
  }
  int pn = 0;
  auto mark = [&] {
    if (profile) cudaEventRecord(pev[pn++], st);
  };
  static unsigned long long* trace_d = nullptr;
  if (profile && !trace_d) cudaMalloc((void**)&trace_d, 3 * 148 * 8 * 8);
  if (profile) {
    cudaMemsetAsync(trace_d, 0, 3 * 148 * 8 * 8, st);
    fwd.trace = trace_d;
    gx.trace = trace_d + 148 * 8;
    gw.trace = trace_d + 2 * 148 * 8;
  }
  mark();
  if (int rc = prep_launch_pair(px, pw2, st)) return rc;
  mark();

  // ---- 2. forward GEMM + online softmax
  CUtensorMap maps[4];
  if (int rc = operand_maps<T>(&maps[0], &maps[1], P(o_xh), P(o_xl), B, D, D, GEMM_BM, false)) return rc;
  if (int rc = operand_maps<T>(&maps[2], &maps[3], P(o_wh), P(o_wl), C, D, D, fwd.bn / kArcCtas, false)) return rc;
  ArcFwdEpi::Params fp{F(o_cos), y, reinterpret_cast<float2*>(ws + o_part), B, C, Cp, fwd.n_splits, mg};
  if (int rc = launch_nt_gemm<PREC, kArcBN, kArcCtas, 0, ArcFwdEpi>(maps, fwd, fp, units, st)) return rc;
  mark();

  // ---- 3. loss (+ d cos planes)
  {
    const int row_groups = (B + kDcosRows - 1) / kDcosRows;
    int chunks = bwd ? std::max(1, std::min((Cp + 255) / 256, (4 * sms + row_groups - 1) / row_groups)) : 1;
    const int chunk = ((Cp + chunks - 1) / chunks + 255) & ~255;
    chunks = (Cp + chunk - 1) / chunk;
    arc_dcos_kernel<T><<<dim3(chunks, row_groups), 256, 0, st>>>(reinterpret_cast<const float2*>(ws + o_part), fwd.n_splits,
                                                                 F(o_cos), Cp, y, dloss, B, C, chunk, mg, loss,
                                                                 bwd ? P(o_dh) : nullptr, bwd ? P(o_dl) : nullptr);
    DIF_LAUNCH_OK();
  }
  mark();
  if (!bwd) return DIF_OK;

  // ---- 4. dxh (split-K planes): A = d cos K-major, B = the W planes read MN-major.  dwh forks onto the side stream.
  static thread_local cudaStream_t side = nullptr;
  static thread_local cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  if (side_by_side && !side) {
    DIF_CUDA_OK(cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking));
    DIF_CUDA_OK(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
    DIF_CUDA_OK(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
  }
  cudaStream_t st_w = side_by_side ? side : st;
  if (side_by_side) {
    DIF_CUDA_OK(cudaEventRecord(ev_fork, st));
    DIF_CUDA_OK(cudaStreamWaitEvent(st_w, ev_fork, 0));
  }
  if (int rc = operand_maps<T>(&maps[0], &maps[1], P(o_dh), P(o_dl), B, Cp, Cp, GEMM_BM, false)) return rc;
  if (int rc = operand_maps<T>(&maps[2], &maps[3], P(o_wh), P(o_wl), D, C, D, 0, true)) return rc;
  StoreEpi::Params sx{F(o_gx), B, D, D, gx.n_splits, (size_t)B * D};
  if (int rc = launch_nt_gemm<PREC, kArcBN, kArcCtas, 0, StoreEpi, 0, 1>(maps, gx, sx, units_x, st)) return rc;
  mark();
  //      dwh: A = d cos read MN-major (classes are the rows of the product), B = the X planes read MN-major
  if (int rc = operand_maps<T>(&maps[0], &maps[1], P(o_dh), P(o_dl), C, B, Cp, 0, true)) return rc;
  if (int rc = operand_maps<T>(&maps[2], &maps[3], P(o_xh), P(o_xl), D, B, D, 0, true)) return rc;
  StoreEpi::Params sw{F(o_gw), C, D, D, gw.n_splits, 0};
  if (int rc = launch_nt_gemm<PREC, kArcBN, kArcCtas, 0, StoreEpi, 1, 1>(maps, gw, sw, units_w, st_w)) return rc;

  // ---- 5. l2_normalize backward.  Side by side: each half follows its own GEMM on its own stream (the 10 000 W rows are
  //         60 MB of traffic: they start when dwh is done, under the tail of dxh, instead of after both); else one launch.
  NormBwdJob jx{F(o_gx), gx.k_splits, (size_t)B * D, X, F(o_xi), B, dX};
  NormBwdJob jw{F(o_gw), 1, 0, W, F(o_wi), C, dW};
  static const bool joint_norm = getenv("DIF_ARC_JOINT_NORM") != nullptr;   // A/B switch
  if (side_by_side && !joint_norm) {
    NormBwdJob none{nullptr, 0, 0, nullptr, nullptr, 0, nullptr};
    launch_norm_bwd(none, jw, D, st_w);
    DIF_LAUNCH_OK();
    DIF_CUDA_OK(cudaEventRecord(ev_join, st_w));
    launch_norm_bwd(jx, none, D, st);
    DIF_LAUNCH_OK();
    DIF_CUDA_OK(cudaStreamWaitEvent(st, ev_join, 0));
    mark();
  } else {
    if (side_by_side) {
      DIF_CUDA_OK(cudaEventRecord(ev_join, st_w));
      DIF_CUDA_OK(cudaStreamWaitEvent(st, ev_join, 0));
    }
    mark();
    launch_norm_bwd(jx, jw, D, st);
    DIF_LAUNCH_OK();
  }
  mark();
  if (profile) {
    cudaStreamSynchronize(st);
    static const char* names[] = {"prep", "fwd", "dcos", "dx", "dw", "norm_bwd"};
    fprintf(stderr, "arcface phases (us): fwd bn %d tps %d splits %d | dw bn %d tps %d splits %d | dx ksplits %d |", fwd.bn,
            fwd.tiles_per_split, fwd.n_splits, gw.bn, gw.tiles_per_split, gw.n_splits, gx.k_splits);
    for (int i = 0; i + 1 < pn; ++i) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, pev[i], pev[i + 1]);
      fprintf(stderr, " %s %.1f", names[i], ms * 1e3f);
    }
    fprintf(stderr, "\n");
    static unsigned long long th[3 * 148 * 8];
    cudaMemcpy(th, trace_d, sizeof(th), cudaMemcpyDeviceToHost);
    const char* gn[3] = {"fwd", "dx", "dw"};
    for (int g = 0; g < 3; ++g) {
      const unsigned long long* t = th + g * 148 * 8;
      unsigned long long t00 = ~0ull;
      for (int c = 0; c < 148; ++c)
        if (t[c * 8]) t00 = std::min(t00, t[c * 8]);
      fprintf(stderr, "  %s trace (us after the first CTA started; avg / max over CTAs that did work):", gn[g]);
      static const char* sn[8] = {"entry", "setup", "1st_full", "mma_done", "1st_acc", "epi_done", "pre_sync", "exit"};
      for (int k = 0; k < 8; ++k) {
        double sum = 0, mx = 0;
        int n = 0;
        for (int c = 0; c < 148; ++c) {
          if (!t[c * 8 + k] || !t[c * 8 + 5] || !t[c * 8 + 4]) continue;
          const double v = (double)(t[c * 8 + k] - t00) * 1e-3;
          sum += v;
          mx = std::max(mx, v);
          ++n;
        }
        if (n) fprintf(stderr, " %s %.1f/%.1f", sn[k], sum / n, mx);
      }
      fprintf(stderr, "\n");
    }
  }
  return DIF_OK;
}

}  // namespace dif

using namespace dif;

extern "C" int dif_arcface(const float* X, const float* W, const int32_t* y, int B, int C, int D, float s, float m,
                           float* loss, const float* dloss, float* dX, float* dW, int precision, void* stream) {
  DIF_REQUIRE(X && W && y && loss, DIF_ERR_INVALID, "dif_arcface: null argument");
  DIF_REQUIRE(precision == DIF_PREC_TF32X3 || precision == DIF_PREC_BF16X3, DIF_ERR_INVALID,
              "dif_arcface: precision %d (0 = 3xTF32, 3 = 3xBF16; both fp32-class)", precision);
  const int dmul = precision == DIF_PREC_BF16X3 ? 8 : 4, dmin = precision == DIF_PREC_BF16X3 ? 64 : 32;
  DIF_REQUIRE(B >= 1 && C >= 2 && D >= dmin && D % dmul == 0 && D <= 4096, DIF_ERR_INVALID,
              "dif_arcface: B %d, C %d, D %d (D a multiple of %d in %d..4096)", B, C, D, dmul, dmin);
  DIF_REQUIRE((dX == nullptr) == (dW == nullptr), DIF_ERR_INVALID, "dif_arcface: pass both dX and dW, or neither");
  DIF_REQUIRE(device_sm_count() > 0, DIF_ERR_STATE, "dif_init has not succeeded on this process");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return precision == DIF_PREC_BF16X3 ? run_arcface<3>(X, W, y, B, C, D, s, m, loss, dloss, dX, dW, st)
                                      : run_arcface<0>(X, W, y, B, C, D, s, m, loss, dloss, dX, dW, st);
}

extern "C" int dif_arcface_host(const float* X_host, const float* W_host, const int32_t* y_host, int B, int C, int D,
                                float s, float m, float* loss_host, const float* dloss_host, float* dX_host,
                                float* dW_host, int precision) {
  DIF_REQUIRE(X_host && W_host && y_host && loss_host && B >= 1 && C >= 2 && D >= 1, DIF_ERR_INVALID,
              "dif_arcface_host: invalid argument");
  static thread_local struct {
    void* h = nullptr; void* d = nullptr; size_t bytes = 0; cudaStream_t st = nullptr;
  } stage;
  auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
  const size_t xb = al((size_t)B * D * 4), wb = al((size_t)C * D * 4), yb = al((size_t)B * 4);
  const size_t in_bytes = xb + wb + 2 * yb, out_bytes = yb + xb + wb;
  if (!stage.st) DIF_CUDA_OK(cudaStreamCreateWithFlags(&stage.st, cudaStreamNonBlocking));
  if (in_bytes + out_bytes > stage.bytes) {
    if (stage.h) cudaFreeHost(stage.h);
    if (stage.d) cudaFree(stage.d);
    stage.h = stage.d = nullptr;
    stage.bytes = 0;
    DIF_CUDA_OK(cudaMallocHost(&stage.h, in_bytes + out_bytes));
    DIF_CUDA_OK(cudaMalloc(&stage.d, in_bytes + out_bytes));
    stage.bytes = in_bytes + out_bytes;
  }
  char* h = (char*)stage.h;
  char* d = (char*)stage.d;
  memcpy(h, X_host, (size_t)B * D * 4);
  memcpy(h + xb, W_host, (size_t)C * D * 4);
  memcpy(h + xb + wb, y_host, (size_t)B * 4);
  if (dloss_host) memcpy(h + xb + wb + yb, dloss_host, (size_t)B * 4);
  DIF_CUDA_OK(cudaMemcpyAsync(d, h, in_bytes, cudaMemcpyHostToDevice, stage.st));
  char* o = d + in_bytes;
  const bool bwd = dX_host && dW_host;
  if (int rc = dif_arcface((const float*)d, (const float*)(d + xb), (const int32_t*)(d + xb + wb), B, C, D, s, m, (float*)o,
                           dloss_host ? (const float*)(d + xb + wb + yb) : nullptr, bwd ? (float*)(o + yb) : nullptr,
                           bwd ? (float*)(o + yb + xb) : nullptr, precision, stage.st))
    return rc;
  DIF_CUDA_OK(cudaMemcpyAsync(h + in_bytes, o, bwd ? out_bytes : yb, cudaMemcpyDeviceToHost, stage.st));
  DIF_CUDA_OK(cudaStreamSynchronize(stage.st));
  memcpy(loss_host, h + in_bytes, (size_t)B * 4);
  if (bwd) {
    memcpy(dX_host, h + in_bytes + yb, (size_t)B * D * 4);
    memcpy(dW_host, h + in_bytes + yb + xb, (size_t)C * D * 4);
  }
  return DIF_OK;
}
