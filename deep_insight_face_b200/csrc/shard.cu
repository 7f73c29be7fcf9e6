// Row-sharded 1:N search over several B200s, one process per GPU (the reference is single-device:
// deep_insight_face/predictions.py:104-150 compares one pair per call and has no distributed code at all).
//
//   local search (gallery.cu)  ->  packed chunk: scores f32 [Q*k] | local rows i32 [Q*k] | ids i64 [Q*k] (explicit ids)
//   exchange + merge           ->  every rank ends with the same global top-k, key (score best-first, global row asc)
//
// Two transports for the exchange:
//   DIF_TRANSPORT_NCCL  one in-place ncclAllGather of the chunks on the caller's stream, then shard_merge_kernel;
//   DIF_TRANSPORT_PEER  shard_peer_exchange_merge_kernel: ONE kernel stores this rank's chunk into every peer's
//                       exchange buffer over NVLink (CUDA-IPC-mapped at attach time), publishes an epoch flag on
//                       each peer with a system-scope release, waits for the peers' flags and merges.  The buffers
//                       are double-buffered by epoch parity: a rank can run at most one step ahead of a peer (its
//                       step e+1 merge needs the peer's step e+1 chunk, which the peer sends after its step e merge).
// NCCL is bound with dlopen/dlsym so libdif_b200.so still loads where NCCL is absent (the CPU build box).
#include <dlfcn.h>

#include <algorithm>
#include <mutex>
#include <new>
#include <vector>

#include "dif_canon.cuh"
#include "dif_common.cuh"
#include "gallery_state.cuh"

namespace dif {

// ------------------------------------------------------------------------------------------ NCCL binding
typedef struct ncclComm* ncclComm_t;
struct NcclUniqueId {
  char internal[DIF_NCCL_ID_BYTES];
};
constexpr int kNcclUint8 = 1;   // ncclDataType_t::ncclUint8
struct NcclApi {
  int (*GetUniqueId)(NcclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, NcclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*CommCount)(const ncclComm_t, int*) = nullptr;
  int (*CommUserRank)(const ncclComm_t, int*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok = false;
};

static const NcclApi* nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    void* h = nullptr;
    if (const char* path = getenv("DIF_NCCL_LIB")) h = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
    // the copy the process already uses (torch's bundled libnccl.so.2) wins: an ncclComm_t borrowed from torch
    // must be driven by the library that created it
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return;
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(dlsym(h, "ncclAllGather"));
    api.Broadcast = reinterpret_cast<decltype(api.Broadcast)>(dlsym(h, "ncclBroadcast"));
    api.CommCount = reinterpret_cast<decltype(api.CommCount)>(dlsym(h, "ncclCommCount"));
    api.CommUserRank = reinterpret_cast<decltype(api.CommUserRank)>(dlsym(h, "ncclCommUserRank"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllGather && api.Broadcast && api.CommCount &&
             api.CommUserRank && api.GetErrorString;
  });
  return api.ok ? &api : nullptr;
}

#define DIF_NCCL_OK(api, expr)                                                                     \
  do {                                                                                             \
    const int _r = (expr);                                                                         \
    if (_r != 0) {                                                                                 \
      dif::set_error("%s failed: %s (%s:%d)", #expr, (api)->GetErrorString(_r), __FILE__, __LINE__); \
      return DIF_ERR_CUDA;                                                                         \
    }                                                                                              \
  } while (0)

// ------------------------------------------------------------------------------------------ packed chunks
constexpr int kMaxPeers = 16;        // DIF_TRANSPORT_PEER: peer base pointers travel in the kernel parameters
constexpr int kMergeThreads = 128;
constexpr int kFlagBytes = 256;      // epoch flags at the head of every exchange region: u32 [world], world <= 64

__host__ __device__ inline size_t chunk_bytes(int n_queries, int k, int with_ids) {
  const size_t nk = (size_t)n_queries * k;
  return (nk * (with_ids ? 16 : 8) + 15) & ~(size_t)15;
}

struct ShardMergeParams {
  const char* chunks;       // chunk of rank w at chunks + w * chunk_stride
  size_t chunk_stride;
  const int64_t* info;      // [world][2] = {shard_row0, id_base}
  int world, n_queries, k, metric, with_ids;
  float* scores;
  int64_t* ids;
  int64_t* grows;
};

// One block per query (grid-stride): world * k candidates -> the k best by (better score, global row ascending).
__device__ __forceinline__ void shard_merge_queries(const ShardMergeParams& p, uint8_t* sm_raw) {
  const int n = p.world * p.k;
  const size_t nk = (size_t)p.n_queries * p.k;
  uint32_t* so = reinterpret_cast<uint32_t*>(sm_raw);                        // orderable "better" score
  int64_t* sr = reinterpret_cast<int64_t*>(sm_raw + (((size_t)n * 4 + 7) & ~(size_t)7));   // global row, -1 = empty
  int* s_valid = reinterpret_cast<int*>(sr + n);
  for (int q = blockIdx.x; q < p.n_queries; q += gridDim.x) {
    __syncthreads();
    if (threadIdx.x == 0) *s_valid = 0;
    __syncthreads();
    int mine = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const int w = i / p.k, j = i - w * p.k;
      const char* c = p.chunks + (size_t)w * p.chunk_stride;
      const size_t at = (size_t)q * p.k + j;
      const float s = reinterpret_cast<const float*>(c)[at];
      const int32_t row = reinterpret_cast<const int32_t*>(c + nk * 4)[at];
      so[i] = float_orderable(p.metric == 1 ? s : -s);
      sr[i] = row >= 0 ? p.info[2 * w] + (int64_t)row : -1;
      mine += row >= 0;
    }
    if (mine) atomicAdd(s_valid, mine);
    __syncthreads();
    const int n_valid = *s_valid;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      if (sr[i] < 0) continue;
      int rank = 0;
      for (int j = 0; j < n; ++j)
        rank += sr[j] >= 0 && (so[j] > so[i] || (so[j] == so[i] && (sr[j] < sr[i] || (sr[j] == sr[i] && j < i))));
      if (rank < p.k) {
        const int w = i / p.k, jj = i - w * p.k;
        const char* c = p.chunks + (size_t)w * p.chunk_stride;
        const size_t at = (size_t)q * p.k + jj;
        const size_t out = (size_t)q * p.k + rank;
        p.scores[out] = reinterpret_cast<const float*>(c)[at];
        if (p.grows) p.grows[out] = sr[i];
        if (p.ids) {
          const int32_t row = reinterpret_cast<const int32_t*>(c + nk * 4)[at];
          p.ids[out] = p.with_ids ? reinterpret_cast<const int64_t*>(c + nk * 8)[at] : p.info[2 * w + 1] + (int64_t)row;
        }
      }
    }
    for (int i = n_valid + threadIdx.x; i < p.k; i += blockDim.x) {
      p.scores[(size_t)q * p.k + i] = 0.f;
      if (p.grows) p.grows[(size_t)q * p.k + i] = -1;
      if (p.ids) p.ids[(size_t)q * p.k + i] = -1;
    }
  }
}

__global__ void __launch_bounds__(kMergeThreads) shard_merge_kernel(ShardMergeParams p) {
  extern __shared__ uint8_t sm_merge[];
  shard_merge_queries(p, sm_merge);
}

struct PeerParams {
  char* region[kMaxPeers];   // base of every rank's exchange region as mapped in THIS process ([rank] = own)
  size_t slot_off;           // offset of this epoch's slot (parity) inside a region
  const char* my_chunk;      // this rank's packed chunk (local memory)
  size_t my_bytes;           // multiple of 16
  int rank;
  unsigned int epoch;
  unsigned int* done;        // block counter of phase 1 (self-resetting)
};

__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint64_t global_timer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// The grid must be co-resident (host: <= 4 blocks per SM of 128 threads): blocks that reach phase 2 spin on flags
// that the PEERS' phase 1 sets, and a peer's phase 1 ends when all of ITS blocks have run.
__global__ void __launch_bounds__(kMergeThreads) shard_peer_exchange_merge_kernel(ShardMergeParams mp, PeerParams pp) {
  extern __shared__ uint8_t sm_merge[];
  // ---- phase 1: my chunk -> slot [rank] of every rank's region (16-byte stores over NVLink; own region included)
  {
    const uint4* src = reinterpret_cast<const uint4*>(pp.my_chunk);
    const size_t n16 = pp.my_bytes >> 4;
    const size_t off = pp.slot_off + (size_t)pp.rank * mp.chunk_stride;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
      const uint4 v = src[i];
      for (int w = 0; w < mp.world; ++w) reinterpret_cast<uint4*>(pp.region[w] + off)[i] = v;
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int prev = atomicAdd(pp.done, 1u);
    if (prev == gridDim.x - 1) {   // every block's stores are fenced: publish the epoch on every rank
      atomicExch(pp.done, 0u);
      __threadfence_system();
      for (int w = 0; w < mp.world; ++w)
        st_release_sys(reinterpret_cast<unsigned int*>(pp.region[w]) + pp.rank, pp.epoch);
    }
  }
  // ---- phase 2: wait until every rank's chunk of this epoch has landed here
  if ((int)threadIdx.x < mp.world) {
    const unsigned int* flag = reinterpret_cast<const unsigned int*>(pp.region[pp.rank]) + threadIdx.x;
    const uint64_t t0 = global_timer_ns();
    while ((int)(ld_acquire_sys(flag) - pp.epoch) < 0) {
      __nanosleep(64);
      if (global_timer_ns() - t0 > 30ull * 1000000000ull) {
        printf("dif: rank %d waited 30 s for the candidates of rank %d (epoch %u)\n", pp.rank, (int)threadIdx.x, pp.epoch);
        __trap();
      }
    }
  }
  __syncthreads();
  shard_merge_queries(mp, sm_merge);
}

}  // namespace dif

using namespace dif;

struct dif_shard_state {
  void* comm = nullptr;
  int rank = 0, world = 1;
  int64_t row0 = 0;
  int transport = DIF_TRANSPORT_NCCL;
  int with_ids = 0;
  int64_t id_base = 0;      // the gallery's state when it was attached (search refuses a gallery that changed)
  bool has_ids = false;
  int max_q = 0, max_k = 0;
  size_t chunk_cap = 0;     // bytes reserved per rank chunk
  int64_t* info = nullptr;  // device [world][2]
  char* region = nullptr;   // NCCL: world chunks.  PEER: flags | slot 0 (world chunks) | slot 1
  size_t region_bytes = 0;
  char* send = nullptr;     // PEER: this rank's chunk before it is pushed
  unsigned int* done = nullptr;
  unsigned int epoch = 0;
  char* peer_region[kMaxPeers] = {};
  // *_host entry point: device query / result staging
  char* d_out = nullptr;     // ids i64 [nk] | scores f32 [nk] | (8-byte aligned) global rows i64 [nk]
};

void dif::shard_state_destroy(dif_shard_state* s) {
  if (!s) return;
  for (int w = 0; w < kMaxPeers; ++w)
    if (s->peer_region[w] && w != s->rank) cudaIpcCloseMemHandle(s->peer_region[w]);
  cudaFree(s->info);
  cudaFree(s->region);
  cudaFree(s->send);
  cudaFree(s->done);
  cudaFree(s->d_out);
  delete s;
}

static size_t merge_smem(int world, int k) {
  const size_t n = (size_t)world * k;
  return ((n * 4 + 7) & ~(size_t)7) + n * 8 + 16;
}

extern "C" {

int dif_nccl_unique_id(void* id_out_host) {
  DIF_REQUIRE(id_out_host, DIF_ERR_INVALID, "dif_nccl_unique_id: null argument");
  const NcclApi* api = nccl_api();
  DIF_REQUIRE(api, DIF_ERR_STATE, "NCCL is not available in this process (libnccl.so.2 not found)");
  NcclUniqueId id;
  DIF_NCCL_OK(api, api->GetUniqueId(&id));
  memcpy(id_out_host, &id, sizeof(id));
  return DIF_OK;
}

int dif_nccl_comm_create(int world, int rank, const void* id_host, void** comm_out) {
  DIF_REQUIRE(id_host && comm_out && world >= 1 && rank >= 0 && rank < world, DIF_ERR_INVALID,
              "dif_nccl_comm_create: invalid argument (world %d, rank %d)", world, rank);
  DIF_REQUIRE(device_sm_count() > 0, DIF_ERR_STATE, "dif_init has not succeeded on this process");
  const NcclApi* api = nccl_api();
  DIF_REQUIRE(api, DIF_ERR_STATE, "NCCL is not available in this process (libnccl.so.2 not found)");
  NcclUniqueId id;
  memcpy(&id, id_host, sizeof(id));
  ncclComm_t comm = nullptr;
  DIF_NCCL_OK(api, api->CommInitRank(&comm, world, id, rank));
  *comm_out = comm;
  return DIF_OK;
}

int dif_nccl_comm_destroy(void* nccl_comm) {
  if (!nccl_comm) return DIF_OK;
  const NcclApi* api = nccl_api();
  DIF_REQUIRE(api, DIF_ERR_STATE, "NCCL is not available in this process");
  DIF_NCCL_OK(api, api->CommDestroy(static_cast<ncclComm_t>(nccl_comm)));
  return DIF_OK;
}

int64_t dif_shard_chunk_bytes(int n_queries, int k, int with_ids) {
  if (n_queries <= 0 || k < 1 || k > DIF_MAX_TOPK) return -1;
  return (int64_t)chunk_bytes(n_queries, k, with_ids != 0);
}

int dif_gallery_search_packed(dif_gallery_t* g, const float* queries, int n_queries, int k, int with_ids, void* chunk,
                              void* stream) {
  DIF_REQUIRE(g && queries && chunk, DIF_ERR_INVALID, "dif_gallery_search_packed: null argument");
  DIF_REQUIRE(n_queries > 0 && k >= 1 && k <= DIF_MAX_TOPK, DIF_ERR_INVALID, "dif_gallery_search_packed: n_queries %d, k %d",
              n_queries, k);
  DIF_REQUIRE((reinterpret_cast<uintptr_t>(chunk) & 15u) == 0, DIF_ERR_INVALID, "chunk must be 16-byte aligned");
  DIF_REQUIRE(with_ids || !g->has_ids, DIF_ERR_STATE,
              "this gallery carries explicit ids: its packed chunks need the id plane (with_ids = 1)");
  const size_t nk = (size_t)n_queries * k;
  char* c = static_cast<char*>(chunk);
  return gallery_search_impl(g, queries, n_queries, k, reinterpret_cast<float*>(c), with_ids ? reinterpret_cast<int64_t*>(c + nk * 8) : nullptr,
                             reinterpret_cast<int32_t*>(c + nk * 4), static_cast<cudaStream_t>(stream));
}

int dif_shard_merge(const void* chunks, const int64_t* shard_info, int world, int n_queries, int k, int metric,
                    int with_ids, float* scores, int64_t* ids, int64_t* grows, void* stream) {
  DIF_REQUIRE(chunks && shard_info && scores, DIF_ERR_INVALID, "dif_shard_merge: null argument");
  DIF_REQUIRE(world >= 1 && world <= 64 && n_queries > 0 && k >= 1 && k <= DIF_MAX_TOPK, DIF_ERR_INVALID,
              "dif_shard_merge: world %d, n_queries %d, k %d", world, n_queries, k);
  ShardMergeParams p{static_cast<const char*>(chunks), chunk_bytes(n_queries, k, with_ids != 0), shard_info, world, n_queries, k,
                     metric, with_ids != 0, scores, ids, grows};
  shard_merge_kernel<<<std::min(n_queries, 148 * 8), kMergeThreads, merge_smem(world, k), static_cast<cudaStream_t>(stream)>>>(p);
  DIF_LAUNCH_OK();
  return DIF_OK;
}

int dif_gallery_shard_attach(dif_gallery_t* g, void* nccl_comm, int rank, int world, int64_t shard_row0, int max_queries,
                             int max_k, int transport) {
  DIF_REQUIRE(g && nccl_comm, DIF_ERR_INVALID, "dif_gallery_shard_attach: null argument");
  DIF_REQUIRE(world >= 1 && world <= 64 && rank >= 0 && rank < world && shard_row0 >= 0 && max_queries > 0 && max_k >= 1 &&
                  max_k <= DIF_MAX_TOPK,
              DIF_ERR_INVALID, "dif_gallery_shard_attach: world %d, rank %d, row0 %lld, max_queries %d, max_k %d", world, rank,
              (long long)shard_row0, max_queries, max_k);
  DIF_REQUIRE(transport == DIF_TRANSPORT_NCCL || transport == DIF_TRANSPORT_PEER, DIF_ERR_INVALID, "unknown transport %d", transport);
  DIF_REQUIRE(transport != DIF_TRANSPORT_PEER || world <= kMaxPeers, DIF_ERR_INVALID,
              "DIF_TRANSPORT_PEER serves at most %d ranks", kMaxPeers);
  const NcclApi* api = nccl_api();
  DIF_REQUIRE(api, DIF_ERR_STATE, "NCCL is not available in this process (libnccl.so.2 not found)");
  ncclComm_t comm = static_cast<ncclComm_t>(nccl_comm);
  int n = 0, r = -1;
  DIF_NCCL_OK(api, api->CommCount(comm, &n));
  DIF_NCCL_OK(api, api->CommUserRank(comm, &r));
  DIF_REQUIRE(n == world && r == rank, DIF_ERR_INVALID, "the communicator has %d ranks and this is rank %d of it; got world %d, rank %d",
              n, r, world, rank);
  shard_state_destroy(g->shard);
  g->shard = nullptr;
  dif_shard_state* s = new (std::nothrow) dif_shard_state();
  DIF_REQUIRE(s, DIF_ERR_CUDA, "out of host memory");
  struct Guard {
    dif_shard_state* s;
    ~Guard() { shard_state_destroy(s); }
  } guard{s};
  s->comm = nccl_comm;
  s->rank = rank;
  s->world = world;
  s->row0 = shard_row0;
  s->transport = transport;
  s->max_q = max_queries;
  s->max_k = max_k;
  cudaStream_t st = g->own_stream;

  // {row0, id_base, explicit ids?} of every rank, once
  int64_t* xinfo = nullptr;
  DIF_CUDA_OK(cudaMalloc((void**)&xinfo, (size_t)world * 3 * 8));
  struct Free {
    void* p;
    ~Free() { cudaFree(p); }
  } fx{xinfo};
  const int64_t mine[3] = {shard_row0, g->id_base, g->has_ids ? 1 : 0};
  DIF_CUDA_OK(cudaMemcpyAsync(xinfo + 3 * rank, mine, sizeof(mine), cudaMemcpyHostToDevice, st));
  DIF_NCCL_OK(api, api->AllGather(xinfo + 3 * rank, xinfo, sizeof(mine), kNcclUint8, comm, st));
  std::vector<int64_t> all((size_t)world * 3);
  DIF_CUDA_OK(cudaMemcpyAsync(all.data(), xinfo, all.size() * 8, cudaMemcpyDeviceToHost, st));
  DIF_CUDA_OK(cudaStreamSynchronize(st));
  std::vector<int64_t> info((size_t)world * 2);
  for (int w = 0; w < world; ++w) {
    DIF_REQUIRE(all[3 * w + 2] == all[2], DIF_ERR_STATE,
                "rank %d %s explicit ids but rank 0 %s: the shards of one gallery must agree", w, all[3 * w + 2] ? "carries" : "has no",
                all[2] ? "does" : "does not");
    info[2 * w] = all[3 * w];
    info[2 * w + 1] = all[3 * w + 1];
  }
  s->with_ids = (int)all[2];
  s->id_base = g->id_base;
  s->has_ids = g->has_ids;
  DIF_CUDA_OK(cudaMalloc((void**)&s->info, info.size() * 8));
  DIF_CUDA_OK(cudaMemcpyAsync(s->info, info.data(), info.size() * 8, cudaMemcpyHostToDevice, st));

  s->chunk_cap = chunk_bytes(max_queries, max_k, s->with_ids);
  const size_t nk = (size_t)max_queries * max_k;
  DIF_CUDA_OK(cudaMalloc((void**)&s->d_out, nk * 20 + 8));
  if (transport == DIF_TRANSPORT_NCCL) {
    s->region_bytes = (size_t)world * s->chunk_cap;
    DIF_CUDA_OK(cudaMalloc((void**)&s->region, s->region_bytes));
  } else {
    s->region_bytes = kFlagBytes + 2 * (size_t)world * s->chunk_cap;
    DIF_CUDA_OK(cudaMalloc((void**)&s->region, s->region_bytes));
    DIF_CUDA_OK(cudaMemsetAsync(s->region, 0, kFlagBytes, st));
    DIF_CUDA_OK(cudaMalloc((void**)&s->send, s->chunk_cap));
    DIF_CUDA_OK(cudaMalloc((void**)&s->done, 4));
    DIF_CUDA_OK(cudaMemsetAsync(s->done, 0, 4, st));
    // every rank learns every region's IPC handle through the communicator, then maps the peers' regions
    cudaIpcMemHandle_t h{};
    if (world > 1) DIF_CUDA_OK(cudaIpcGetMemHandle(&h, s->region));
    char* xh = nullptr;
    DIF_CUDA_OK(cudaMalloc((void**)&xh, (size_t)world * sizeof(h)));
    Free fh{xh};
    DIF_CUDA_OK(cudaMemcpyAsync(xh + (size_t)rank * sizeof(h), &h, sizeof(h), cudaMemcpyHostToDevice, st));
    DIF_NCCL_OK(api, api->AllGather(xh + (size_t)rank * sizeof(h), xh, sizeof(h), kNcclUint8, comm, st));
    std::vector<cudaIpcMemHandle_t> hs((size_t)world);
    DIF_CUDA_OK(cudaMemcpyAsync(hs.data(), xh, (size_t)world * sizeof(h), cudaMemcpyDeviceToHost, st));
    DIF_CUDA_OK(cudaStreamSynchronize(st));
    int mapped = 1, bad_rank = -1;
    cudaError_t bad = cudaSuccess;
    for (int w = 0; w < world && mapped; ++w) {
      if (w == rank) {
        s->peer_region[w] = s->region;
        continue;
      }
      void* p = nullptr;
      const cudaError_t e = cudaIpcOpenMemHandle(&p, hs[(size_t)w], cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess) {
        cudaGetLastError();
        mapped = 0;
        bad = e;
        bad_rank = w;
      } else {
        s->peer_region[w] = static_cast<char*>(p);
      }
    }
    // One more collective: (a) nobody may push into a region before its owner has zeroed the flags, (b) every rank
    // learns whether EVERY rank mapped its peers, so all ranks return the same status and can fall back together.
    int* xok = reinterpret_cast<int*>(xh);
    DIF_CUDA_OK(cudaMemcpyAsync(xok + rank, &mapped, sizeof(int), cudaMemcpyHostToDevice, st));
    DIF_NCCL_OK(api, api->AllGather(xok + rank, xok, sizeof(int), kNcclUint8, comm, st));
    std::vector<int> oks((size_t)world);
    DIF_CUDA_OK(cudaMemcpyAsync(oks.data(), xok, (size_t)world * sizeof(int), cudaMemcpyDeviceToHost, st));
    DIF_CUDA_OK(cudaStreamSynchronize(st));
    for (int w = 0; w < world; ++w) {
      if (oks[(size_t)w]) continue;
      if (!mapped)
        set_error("cudaIpcOpenMemHandle of rank %d's exchange region failed: %s (DIF_TRANSPORT_PEER needs one process per GPU "
                  "on one node with peer access; use DIF_TRANSPORT_NCCL otherwise)", bad_rank, cudaGetErrorString(bad));
      else
        set_error("rank %d could not map its peers' exchange regions (DIF_TRANSPORT_PEER); use DIF_TRANSPORT_NCCL", w);
      return DIF_ERR_CUDA;
    }
  }
  DIF_CUDA_OK(cudaStreamSynchronize(st));   // `info` (a host vector) is consumed, flags are zero
  guard.s = nullptr;
  g->shard = s;
  return DIF_OK;
}

int dif_gallery_search_sharded(dif_gallery_t* g, void* nccl_comm, int rank, int world, const float* queries, int n_queries,
                               int k, float* scores, int64_t* ids, int64_t* grows, void* stream) {
  DIF_REQUIRE(g && queries && scores && ids, DIF_ERR_INVALID, "dif_gallery_search_sharded: null argument");
  dif_shard_state* s = g->shard;
  DIF_REQUIRE(s, DIF_ERR_STATE, "dif_gallery_shard_attach has not been called on this gallery");
  DIF_REQUIRE(s->comm == nccl_comm && s->rank == rank && s->world == world, DIF_ERR_INVALID,
              "the gallery is attached as rank %d of %d on another communicator", s->rank, s->world);
  DIF_REQUIRE(n_queries > 0 && n_queries <= s->max_q && k >= 1 && k <= s->max_k, DIF_ERR_CAPACITY,
              "attached for at most %d queries x top-%d; got %d x %d", s->max_q, s->max_k, n_queries, k);
  DIF_REQUIRE(s->has_ids == g->has_ids && s->id_base == g->id_base, DIF_ERR_STATE,
              "the gallery's ids changed since dif_gallery_shard_attach (explicit ids / id_base): attach again on every rank");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t cb = chunk_bytes(n_queries, k, s->with_ids);
  const int64_t launches0 = dif_launch_count();
  ShardMergeParams mp{nullptr, cb, s->info, world, n_queries, k, g->metric, s->with_ids, scores, ids, grows};
  if (s->transport == DIF_TRANSPORT_NCCL) {
    const NcclApi* api = nccl_api();
    DIF_REQUIRE(api, DIF_ERR_STATE, "NCCL is not available in this process");
    char* mine = s->region + (size_t)rank * cb;
    if (int rc = dif_gallery_search_packed(g, queries, n_queries, k, s->with_ids, mine, st)) return rc;
    if (world > 1) {
      DIF_NCCL_OK(api, api->AllGather(mine, s->region, cb, kNcclUint8, static_cast<ncclComm_t>(nccl_comm), st));
    }
    mp.chunks = s->region;
    shard_merge_kernel<<<std::min(n_queries, 148 * 8), kMergeThreads, merge_smem(world, k), st>>>(mp);
    DIF_LAUNCH_OK();
  } else {
    if (int rc = dif_gallery_search_packed(g, queries, n_queries, k, s->with_ids, s->send, st)) return rc;
    PeerParams pp{};
    for (int w = 0; w < world; ++w) pp.region[w] = s->peer_region[w];
    s->epoch += 1;
    pp.slot_off = kFlagBytes + (size_t)(s->epoch & 1u) * world * s->chunk_cap;
    pp.my_chunk = s->send;
    pp.my_bytes = cb;
    pp.rank = rank;
    pp.epoch = s->epoch;
    pp.done = s->done;
    mp.chunks = s->region + pp.slot_off;
    // co-resident grid: at most 4 blocks of 128 threads per SM
    const int grid = std::max(1, std::min(n_queries, device_sm_count() * 4));
    shard_peer_exchange_merge_kernel<<<grid, kMergeThreads, merge_smem(world, k), st>>>(mp, pp);
    DIF_LAUNCH_OK();
  }
  DIF_CUDA_OK(cudaEventRecord(g->ev_phase[3], st));
  g->phase_sharded = true;
  g->stats[1] = dif_launch_count() - launches0;
  return DIF_OK;
}

int dif_gallery_search_sharded_host(dif_gallery_t* g, void* nccl_comm, int rank, int world, const float* queries_host,
                                    int bcast_root, int n_queries, int k, float* scores_host, int64_t* ids_host,
                                    int64_t* grows_host) {
  DIF_REQUIRE(g, DIF_ERR_INVALID, "dif_gallery_search_sharded_host: null gallery");
  dif_shard_state* s = g->shard;
  DIF_REQUIRE(s, DIF_ERR_STATE, "dif_gallery_shard_attach has not been called on this gallery");
  DIF_REQUIRE(bcast_root < world && bcast_root >= DIF_UPLOAD_SLICED, DIF_ERR_INVALID, "bcast_root %d out of range", bcast_root);
  const bool sliced = bcast_root == DIF_UPLOAD_SLICED && world > 1;
  const bool i_upload = bcast_root < 0 || bcast_root == rank;
  DIF_REQUIRE(!i_upload || queries_host, DIF_ERR_INVALID, "dif_gallery_search_sharded_host: this rank must supply the queries");
  DIF_REQUIRE(n_queries > 0 && n_queries <= s->max_q && k >= 1 && k <= s->max_k, DIF_ERR_CAPACITY,
              "attached for at most %d queries x top-%d; got %d x %d", s->max_q, s->max_k, n_queries, k);
  DIF_REQUIRE((scores_host != nullptr) == (ids_host != nullptr), DIF_ERR_INVALID, "pass both scores_host and ids_host, or neither");
  const size_t row_bytes = (size_t)g->D * 4;
  const int slice_rows = (n_queries + world - 1) / world;          // sliced upload: rows each rank sends over PCIe
  const size_t qb = sliced ? (size_t)slice_rows * world * row_bytes : (size_t)n_queries * row_bytes;
  const size_t nk = (size_t)n_queries * k;
  const size_t grows_off = (nk * 12 + 7) & ~(size_t)7;
  const size_t ob = grows_off + nk * 8;
  const size_t qb_al = (qb + 255) & ~(size_t)255;
  if (int rc = gallery_ensure_stage(g, qb_al + ob)) return rc;
  char* hp = static_cast<char*>(g->h_pin);
  float* dq = static_cast<float*>(g->d_stage);
  cudaStream_t st = g->own_stream;
  if (i_upload) {
    // sliced: this rank moves only rows [rank * slice_rows, ...) over its PCIe link; NVLink assembles the batch below
    const int row0 = sliced ? std::min(n_queries, rank * slice_rows) : 0;
    const int rows = sliced ? std::min(slice_rows, n_queries - row0) : n_queries;
    const size_t off0 = (size_t)row0 * row_bytes, ub = (size_t)rows * row_bytes;
    const char* src = reinterpret_cast<const char*>(queries_host) + off0;
    char* dst = reinterpret_cast<char*>(dq) + off0;
    cudaPointerAttributes attr{};
    const bool pinned = cudaPointerGetAttributes(&attr, queries_host) == cudaSuccess && attr.type == cudaMemoryTypeHost;
    if (!pinned) cudaGetLastError();
    if (pinned) {
      if (ub) DIF_CUDA_OK(cudaMemcpyAsync(dst, src, ub, cudaMemcpyHostToDevice, st));
    } else {
      const size_t piece = (size_t)1 << 20;
      for (size_t off = 0; off < ub; off += piece) {
        const size_t n = std::min(piece, ub - off);
        memcpy(hp + off, src + off, n);
        DIF_CUDA_OK(cudaMemcpyAsync(dst + off, hp + off, n, cudaMemcpyHostToDevice, st));
      }
    }
  }
  if (sliced) {
    const NcclApi* api = nccl_api();
    DIF_REQUIRE(api, DIF_ERR_STATE, "NCCL is not available in this process");
    const size_t sb = (size_t)slice_rows * row_bytes;
    DIF_NCCL_OK(api, api->AllGather(reinterpret_cast<char*>(dq) + (size_t)rank * sb, dq, sb, kNcclUint8,
                                    static_cast<ncclComm_t>(nccl_comm), st));
  }
  if (bcast_root >= 0 && world > 1) {
    const NcclApi* api = nccl_api();
    DIF_REQUIRE(api, DIF_ERR_STATE, "NCCL is not available in this process");
    DIF_NCCL_OK(api, api->Broadcast(dq, dq, qb, kNcclUint8, bcast_root, static_cast<ncclComm_t>(nccl_comm), st));
  }
  if (int rc = dif_gallery_search_sharded(g, nccl_comm, rank, world, dq, n_queries, k, reinterpret_cast<float*>(s->d_out + nk * 8),
                                          reinterpret_cast<int64_t*>(s->d_out),
                                          grows_host ? reinterpret_cast<int64_t*>(s->d_out + grows_off) : nullptr, st))
    return rc;
  if (scores_host)   // one D2H of the whole result block
    DIF_CUDA_OK(cudaMemcpyAsync(hp + qb_al, s->d_out, grows_host ? ob : nk * 12, cudaMemcpyDeviceToHost, st));
  DIF_CUDA_OK(cudaStreamSynchronize(st));
  if (scores_host) {
    memcpy(ids_host, hp + qb_al, nk * 8);
    memcpy(scores_host, hp + qb_al + nk * 8, nk * 4);
    if (grows_host) memcpy(grows_host, hp + qb_al + grows_off, nk * 8);
  }
  return DIF_OK;
}

}  // extern "C"
