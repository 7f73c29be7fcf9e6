// The gallery handle behind dif_gallery_t, shared by gallery.cu (single-GPU search) and shard.cu (row-sharded search).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/dif_b200.h"

struct dif_shard_state;   // shard.cu

struct dif_gallery {
  int device = 0;
  int64_t capacity = 0;
  int64_t size = 0;
  int D = 0;
  int metric = 1;
  int precision = 0;
  int64_t id_base = 0;
  bool has_ids = false;    // `ids` holds the id of every stored row (explicit ids, or default ids frozen by a remove)
  bool user_ids = false;   // the caller supplied ids at least once: every later add needs them too
  int64_t enrolled = 0;    // rows ever appended since the last reset (default id of the next row = id_base + enrolled)
  float* g0 = nullptr;
  float* g1 = nullptr;
  __nv_bfloat16* gb = nullptr;
  __nv_bfloat16* gb1 = nullptr;   // second bf16 plane (3xBF16)
  float* gsq = nullptr;  // [capacity + kGalBN]
  unsigned int* gmax = nullptr;
  int64_t* ids = nullptr;
  // query-side workspace (grown on demand)
  int q_cap = 0;
  float* q0 = nullptr;
  float* q1 = nullptr;
  __nv_bfloat16* qb = nullptr;
  __nv_bfloat16* qb1 = nullptr;
  float* qsq = nullptr;
  uint64_t* cand = nullptr;
  size_t cand_elems = 0;
  int* flagged = nullptr;  // [0] = count, [1..] = list
  unsigned int* bound = nullptr;  // [q_cap] shared per-query lower bound on the k-th best score
  unsigned int* maxima = nullptr; // [splits][q_cap] best score of each split's list
  size_t maxima_elems = 0;
  uint64_t* ex_keys = nullptr;
  size_t ex_elems = 0;
  // host staging for the *_host entry points
  void* h_pin = nullptr;
  size_t h_pin_bytes = 0;
  void* d_stage = nullptr;
  size_t d_stage_bytes = 0;
  cudaStream_t own_stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;   // around the tensor-core pass
  cudaEvent_t ev_phase[4] = {nullptr, nullptr, nullptr, nullptr};   // search start, after re-rank, after exact path, after exchange + merge
  bool phase_sharded = false;   // the last search recorded ev_phase[3]
  int64_t stats[6] = {0, 0, 0, 0, 0, 0};
  int opt_ctas = 2;
  int opt_force_fallback = 0;
  int opt_resident = -1;   // -1 auto, 0 never, 1 whenever it fits
  int opt_splits = 0;      // 0 auto
  int opt_l2_prefetch = 0;   // measured: no gain at C3 (tiles are L2 hits already), so off by default
  dif_shard_state* shard = nullptr;   // set by dif_gallery_shard_attach
};

namespace dif {
// dif_gallery_search with optional outputs: ids == NULL skips the id lookup (packed chunks of id-less shards)
int gallery_search_impl(dif_gallery* g, const float* queries, int n_queries, int k, float* scores, int64_t* ids,
                        int32_t* rows, cudaStream_t st);
int gallery_ensure_stage(dif_gallery* g, size_t bytes);
void shard_state_destroy(dif_shard_state* s);
}  // namespace dif
