// ArcFace additive-angular-margin logits + softmax cross-entropy, forward + backward.
// ABSENT from the reference; specification in DESIGN.md and oracle/losses_oracle.py:arcface (arXiv 1801.07698):
//   xh = l2_normalize(x), wh = l2_normalize(w), cos = clip(xh . wh, -1, 1)
//   target logit s * phi(cos_y), phi = cos(theta + m) if cos_y > cos(pi - m) else cos_y - m sin(pi - m); others s * cos
//   loss_b = logsumexp(logits_b) - logits_b[y_b];   dX, dW = gradients of sum_b dloss_b * loss_b (default 1/B)
//
// All three contractions run on the tcgen05/TMA NT-GEMM skeleton in 3xTF32 (fp32-exact to ~1e-6):
//   1. arc_fwd     cos = xh wh^T with a fused epilogue: clip, margin on the target column, scale, online
//                  softmax (running max / sum per class split) in registers; cos is also written out once
//   2. arc_loss    merges the per-split (max, sum) pairs -> logZ, loss, d phi / d cos
//   3. arc_dcos    p - onehot -> d cos, written as TF32 hi/lo planes in both orientations ([B,C] and [C,B])
//   4. dxh = dcos wh        (NT GEMM over K = C, split-K planes summed in a fixed order)
//      dwh = dcos^T xh      (NT GEMM over K = B)
//   5. arc_norm_bwd  l2_normalize backward for X rows and W rows
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

// one thread per sample: merge the split partials, loss, logZ (natural log units), target phi and d phi
__global__ void arc_loss_kernel(const float2* __restrict__ part, int n_slots, const float* __restrict__ cosbuf, int ldc,
                                const int32_t* __restrict__ y, int B, ArcMargin mg, float* __restrict__ loss,
                                float* __restrict__ logz, float* __restrict__ dphi) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float m2 = -INFINITY, l2 = 0.f;
  for (int s = 0; s < n_slots; ++s) {
    const float2 q = part[(size_t)b * n_slots + s];
    if (q.x == -INFINITY) continue;
    const float mn = fmaxf(m2, q.x);
    l2 = l2 * exp2f(m2 - mn) + q.y * exp2f(q.x - mn);
    m2 = mn;
  }
  const float lz = (m2 + log2f(l2)) / kLog2e;
  const float ct = fminf(fmaxf(cosbuf[(size_t)b * ldc + y[b]], -1.f), 1.f);
  loss[b] = lz - mg.s * arc_phi(ct, mg);
  logz[b] = lz;
  dphi[b] = arc_dphi(ct, mg);
}

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

// d cos in both orientations as operand planes.  Tile 32 samples x 32 classes per block (32 x 8 threads).
template <class T>
__global__ void __launch_bounds__(256) arc_dcos_kernel(const float* __restrict__ cosbuf, int ldc,
                                                       const int32_t* __restrict__ y, const float* __restrict__ logz,
                                                       const float* __restrict__ dphi, const float* __restrict__ dloss,
                                                       int B, int C, ArcMargin mg, T* __restrict__ d_hi,
                                                       T* __restrict__ d_lo,    // [Bp][ldc]
                                                       T* __restrict__ t_hi, T* __restrict__ t_lo, int ldt,
                                                       int Cp, int Bp) {            // [Cp][ldt]
  __shared__ float tile[32][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int b = blockIdx.y * 32 + r;
    float v = 0.f;
    if (b < B && c < C) {
      const float raw = cosbuf[(size_t)b * ldc + c];
      const float ct = fminf(fmaxf(raw, -1.f), 1.f);
      const float g = dloss ? dloss[b] : 1.f / (float)B;
      const bool tgt = (c == y[b]);
      const float logit = mg.s * (tgt ? arc_phi(ct, mg) : ct);
      const float pr = expf(logit - logz[b]);
      v = mg.s * g * (pr - (tgt ? 1.f : 0.f)) * (tgt ? dphi[b] : 1.f);
      if (raw < -1.f || raw > 1.f) v = 0.f;   // clip_by_value passes no gradient outside the bounds
    }
    tile[r][threadIdx.x] = v;
    if (b < Bp && c < ldc) split_store(v, d_hi, d_lo, (size_t)b * ldc + c);
  }
  __syncthreads();
  const int b2 = blockIdx.y * 32 + threadIdx.x;
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int c2 = blockIdx.x * 32 + r;
    if (c2 < Cp && b2 < ldt) split_store(tile[threadIdx.x][r], t_hi, t_lo, (size_t)c2 * ldt + b2);
  }
}

// src [R][Cc] (two planes) -> transposed planes [Cc][ldt] (zero padded to ldt)
template <class T>
__global__ void __launch_bounds__(256) transpose_planes_kernel(const T* __restrict__ s_hi,
                                                               const T* __restrict__ s_lo, int R, int Cc,
                                                               T* __restrict__ t_hi, T* __restrict__ t_lo,
                                                               int ldt) {
  __shared__ T th[32][33 + (sizeof(T) == 2 ? 1 : 0)], tl[32][33 + (sizeof(T) == 2 ? 1 : 0)];
  const T zero = T(0.f);
  const int c = blockIdx.x * 32 + threadIdx.x;
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int row = blockIdx.y * 32 + r;
    const bool ok = row < R && c < Cc;
    th[r][threadIdx.x] = ok ? s_hi[(size_t)row * Cc + c] : zero;
    tl[r][threadIdx.x] = ok ? s_lo[(size_t)row * Cc + c] : zero;
  }
  __syncthreads();
  const int row2 = blockIdx.y * 32 + threadIdx.x;
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int c2 = blockIdx.x * 32 + r;
    if (c2 < Cc && row2 < ldt) {
      t_hi[(size_t)c2 * ldt + row2] = th[threadIdx.x][r];
      t_lo[(size_t)c2 * ldt + row2] = tl[threadIdx.x][r];
    }
  }
}

// out_r = inv_r * (g_r - h_r (h_r . g_r)) with g = sum over `planes` partial planes (fixed order), h = hi + lo
// normalised row (h_lo NULL: h_hi is the fp32 normalised row itself); rows whose squared norm was clamped (inv == 1e6) are a plain scaling.  TPR threads own one row
// (TPR = 32: a warp per row, 8 rows per block, D <= 512; TPR = 256: a block per row, D <= 4096), one float4 group
// per thread and pass: every load is independent and coalesced and the planes are read exactly once.
template <int TPR>
__global__ void __launch_bounds__(256) arc_norm_bwd_kernel(const float* __restrict__ g, int planes, size_t plane_stride,
                                                           const float* __restrict__ h_hi, const float* __restrict__ h_lo,
                                                           const float* __restrict__ inv, int R, int D,
                                                           float* __restrict__ out) {
  __shared__ float red[8];
  constexpr int kRows = 256 / TPR;
  const int r = blockIdx.x * kRows + (int)threadIdx.x / TPR;
  const int lane = (int)threadIdx.x % TPR;
  const bool live = r < R;
  const size_t base = (size_t)r * D;
  float dot = 0.f;
  float4 a[4], h[4];
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const int d = (lane + t * TPR) * 4;
    a[t] = make_float4(0.f, 0.f, 0.f, 0.f);
    h[t] = a[t];
    if (live && d < D) {
      for (int s = 0; s < planes; ++s) {
        const float4 v = *reinterpret_cast<const float4*>(g + (size_t)s * plane_stride + base + d);
        a[t].x += v.x; a[t].y += v.y; a[t].z += v.z; a[t].w += v.w;
      }
      const float4 hh = *reinterpret_cast<const float4*>(h_hi + base + d);
      const float4 hl = h_lo ? *reinterpret_cast<const float4*>(h_lo + base + d) : make_float4(0.f, 0.f, 0.f, 0.f);
      h[t] = make_float4(hh.x + hl.x, hh.y + hl.y, hh.z + hl.z, hh.w + hl.w);
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
  const float iv = inv[r];
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
      *reinterpret_cast<float4*>(out + base + d) = o4;
    }
  }
}

static void launch_norm_bwd(const float* g, int planes, size_t plane_stride, const float* h_hi, const float* h_lo,
                            const float* inv, int R, int D, float* out, cudaStream_t st) {
  if (D <= 512)
    arc_norm_bwd_kernel<32><<<(R + 7) / 8, 256, 0, st>>>(g, planes, plane_stride, h_hi, h_lo, inv, R, D, out);
  else
    arc_norm_bwd_kernel<256><<<R, 256, 0, st>>>(g, planes, plane_stride, h_hi, h_lo, inv, R, D, out);
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

template <class T>
static int tmaps(CUtensorMap* maps, const T* a_hi, const T* a_lo, int M, const T* b_hi, const T* b_lo, int N, int K,
                 int ldk) {
  constexpr int esz = (int)sizeof(T), bf = esz == 2 ? 1 : 0;
  constexpr uint32_t cols = 128 / esz;
  if (int rc = make_tmap_2d(&maps[0], a_hi, M, K, (uint64_t)ldk * esz, GEMM_BM, cols, bf)) return rc;
  if (int rc = make_tmap_2d(&maps[1], a_lo, M, K, (uint64_t)ldk * esz, GEMM_BM, cols, bf)) return rc;
  if (int rc = make_tmap_2d(&maps[2], b_hi, N, K, (uint64_t)ldk * esz, kArcBN / kArcCtas, cols, bf)) return rc;
  if (int rc = make_tmap_2d(&maps[3], b_lo, N, K, (uint64_t)ldk * esz, kArcBN / kArcCtas, cols, bf)) return rc;
  return DIF_OK;
}

// PREC 0: TF32 hi/lo planes (3xTF32);  PREC 3: bf16 b0/b1 planes (3xBF16, half the plane bytes and MMA time)
template <int PREC>
static int run_arcface(const float* X, const float* W, const int32_t* y, int B, int C, int D, float s, float m,
                       float* loss, const float* dloss, float* dX, float* dW, cudaStream_t st) {
  using T = typename std::conditional<PREC == 3, __nv_bfloat16, float>::type;
  constexpr int esz = (int)sizeof(T), kChunk = 128 / esz;
  const bool bwd = dX != nullptr;
  const int sms = device_sm_count();
  const int Bp = (B + 7) & ~7, Cp = (C + 7) & ~7;   // plane pitches stay 16-byte multiples for TMA in both modes
  ArcMargin mg{s, cosf(m), sinf(m), cosf(3.14159265358979323846f - m), sinf(3.14159265358979323846f - m) * m};

  // ---- GEMM shapes
  GemmShape fwd{};
  const int bm = GEMM_BM * kArcCtas, units = std::max(1, sms / kArcCtas);
  fwd.m_blocks = (B + bm - 1) / bm;
  fwd.n_tiles = (C + kArcBN - 1) / kArcBN;
  fwd.k_chunks = (D + kChunk - 1) / kChunk;
  fwd.n_splits = std::max(1, std::min(fwd.n_tiles, 2 * units / std::max(1, fwd.m_blocks)));
  fwd.tiles_per_split = (fwd.n_tiles + fwd.n_splits - 1) / fwd.n_splits;
  fwd.n_splits = (fwd.n_tiles + fwd.tiles_per_split - 1) / fwd.tiles_per_split;
  GemmShape gx{};   // dxh [B, D] = dcos [B, Cp] x whT [D, Cp]^T, split over K = classes
  gx.m_blocks = fwd.m_blocks;
  gx.n_tiles = (D + kArcBN - 1) / kArcBN;
  gx.k_chunks = (Cp + kChunk - 1) / kChunk;
  gx.n_splits = gx.n_tiles;
  gx.tiles_per_split = 1;
  gx.k_splits = std::max(1, std::min(gx.k_chunks, units / std::max(1, gx.m_blocks * gx.n_tiles)));
  gx.chunks_per_ksplit = (gx.k_chunks + gx.k_splits - 1) / gx.k_splits;
  gx.k_splits = (gx.k_chunks + gx.chunks_per_ksplit - 1) / gx.chunks_per_ksplit;
  GemmShape gw{};   // dwh [C, D] = dcosT [Cp, Bp] x xhT [D, Bp]^T
  gw.m_blocks = (C + bm - 1) / bm;
  gw.n_tiles = gx.n_tiles;
  gw.k_chunks = (Bp + kChunk - 1) / kChunk;
  gw.n_splits = gw.n_tiles;
  gw.tiles_per_split = 1;

  // ---- workspace carve-up (256-byte aligned pieces)
  size_t off = 0;
  auto take = [&](size_t bytes) {
    const size_t at = off;
    off += (bytes + 255) & ~(size_t)255;
    return at;
  };
  // mode 0: xh / xl = hi / lo planes.  mode 3: xn = fp32 normalised rows (for the normalisation backward),
  // xh / xl = bf16 planes
  const size_t o_xn = PREC == 3 ? take((size_t)B * D * 4) : 0, o_wn = PREC == 3 ? take((size_t)C * D * 4) : 0;
  const size_t o_xh = take((size_t)B * D * esz), o_xl = take((size_t)B * D * esz), o_xi = take((size_t)B * 4);
  const size_t o_wh = take((size_t)C * D * esz), o_wl = take((size_t)C * D * esz), o_wi = take((size_t)C * 4);
  const size_t o_cos = take((size_t)B * Cp * 4);
  const size_t o_part = take((size_t)B * fwd.n_splits * 8);
  const size_t o_logz = take((size_t)B * 4), o_dphi = take((size_t)B * 4);
  size_t o_dh = 0, o_dl = 0, o_th = 0, o_tl = 0, o_wth = 0, o_wtl = 0, o_xth = 0, o_xtl = 0, o_gx = 0, o_gw = 0;
  if (bwd) {
    o_dh = take((size_t)Bp * Cp * esz); o_dl = take((size_t)Bp * Cp * esz);
    o_th = take((size_t)Cp * Bp * esz); o_tl = take((size_t)Cp * Bp * esz);
    o_wth = take((size_t)D * Cp * esz); o_wtl = take((size_t)D * Cp * esz);
    o_xth = take((size_t)D * Bp * esz); o_xtl = take((size_t)D * Bp * esz);
    o_gx = take((size_t)gx.k_splits * B * D * 4);
    o_gw = take((size_t)C * D * 4);
  }
  if (int rc = g_arc.ensure(off)) return rc;
  char* ws = g_arc.base;
  auto F = [&](size_t o) { return reinterpret_cast<float*>(ws + o); };
  auto P = [&](size_t o) { return reinterpret_cast<T*>(ws + o); };

  // ---- 0. normalise + TF32 planes
  PrepParams px{};
  px.src = X; px.n = B; px.D = D; px.normalize = 1; px.inv = F(o_xi);
  PrepParams pw = px;
  pw.src = W; pw.n = C; pw.inv = F(o_wi);
  if (PREC == 3) {
    px.split = 0; px.p0 = F(o_xn);
    px.pb = reinterpret_cast<__nv_bfloat16*>(ws + o_xh); px.pb1 = reinterpret_cast<__nv_bfloat16*>(ws + o_xl);
    pw.split = 0; pw.p0 = F(o_wn);
    pw.pb = reinterpret_cast<__nv_bfloat16*>(ws + o_wh); pw.pb1 = reinterpret_cast<__nv_bfloat16*>(ws + o_wl);
  } else {
    px.split = 1; px.p0 = F(o_xh); px.p1 = F(o_xl);
    pw.split = 1; pw.p0 = F(o_wh); pw.p1 = F(o_wl);
  }
  if (int rc = prep_launch(px, false, st)) return rc;
  if (int rc = prep_launch(pw, false, st)) return rc;

  // ---- 1. forward GEMM + online softmax
  CUtensorMap maps[4];
  if (int rc = tmaps<T>(maps, P(o_xh), P(o_xl), B, P(o_wh), P(o_wl), C, D, D)) return rc;
  ArcFwdEpi::Params fp{F(o_cos), y, reinterpret_cast<float2*>(ws + o_part), B, C, Cp, fwd.n_splits, mg};
  if (int rc = launch_nt_gemm<PREC, kArcBN, kArcCtas, 0, ArcFwdEpi>(maps, fwd, fp, units, st)) return rc;
  // ---- 2. loss
  arc_loss_kernel<<<(B + 127) / 128, 128, 0, st>>>(reinterpret_cast<const float2*>(ws + o_part), fwd.n_splits, F(o_cos), Cp,
                                                  y, B, mg, loss, F(o_logz), F(o_dphi));
  DIF_LAUNCH_OK();
  if (!bwd) return DIF_OK;

  // ---- 3. d cos planes (both orientations) and transposed operand planes
  arc_dcos_kernel<T><<<dim3((Cp + 31) / 32, (Bp + 31) / 32), dim3(32, 8), 0, st>>>(
      F(o_cos), Cp, y, F(o_logz), F(o_dphi), dloss, B, C, mg, P(o_dh), P(o_dl), P(o_th), P(o_tl), Bp, Cp, Bp);
  DIF_LAUNCH_OK();
  transpose_planes_kernel<T><<<dim3((D + 31) / 32, (C + 31) / 32), dim3(32, 8), 0, st>>>(P(o_wh), P(o_wl), C, D, P(o_wth),
                                                                                       P(o_wtl), Cp);
  DIF_LAUNCH_OK();
  transpose_planes_kernel<T><<<dim3((D + 31) / 32, (B + 31) / 32), dim3(32, 8), 0, st>>>(P(o_xh), P(o_xl), B, D, P(o_xth),
                                                                                       P(o_xtl), Bp);
  DIF_LAUNCH_OK();
  // ---- 4. dxh (split-K planes) and dwh
  if (int rc = tmaps<T>(maps, P(o_dh), P(o_dl), B, P(o_wth), P(o_wtl), D, Cp, Cp)) return rc;
  StoreEpi::Params sx{F(o_gx), B, D, D, gx.n_splits, (size_t)B * D};
  if (int rc = launch_nt_gemm<PREC, kArcBN, kArcCtas, 0, StoreEpi>(maps, gx, sx, units, st)) return rc;
  if (int rc = tmaps<T>(maps, P(o_th), P(o_tl), C, P(o_xth), P(o_xtl), D, Bp, Bp)) return rc;
  StoreEpi::Params sw{F(o_gw), C, D, D, gw.n_splits, 0};
  if (int rc = launch_nt_gemm<PREC, kArcBN, kArcCtas, 0, StoreEpi>(maps, gw, sw, units, st)) return rc;
  // ---- 5. l2_normalize backward
  if (PREC == 3) launch_norm_bwd(F(o_gx), gx.k_splits, (size_t)B * D, F(o_xn), nullptr, F(o_xi), B, D, dX, st);
  else launch_norm_bwd(F(o_gx), gx.k_splits, (size_t)B * D, F(o_xh), F(o_xl), F(o_xi), B, D, dX, st);
  DIF_LAUNCH_OK();
  if (PREC == 3) launch_norm_bwd(F(o_gw), 1, 0, F(o_wn), nullptr, F(o_wi), C, D, dW, st);
  else launch_norm_bwd(F(o_gw), 1, 0, F(o_wh), F(o_wl), F(o_wi), C, D, dW, st);
  DIF_LAUNCH_OK();
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
