// Library-wide state of libdif_b200.so: error string, device selection, launch counter,
// TMA descriptor construction.
#include <cudaTypedefs.h>

#include <atomic>
#include <mutex>
#include <vector>

#include "dif_common.cuh"

namespace dif {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};
static int g_sm_count = 0;
static int g_device = -1;   // the one device this process is bound to (dif_init)

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int count_launch(int n) {
  g_launches.fetch_add(n, std::memory_order_relaxed);
  return 0;
}

int device_sm_count() { return g_sm_count; }

// Outgrown workspace blocks are parked, not freed: a CUDA graph captured earlier may still replay kernels that
// point at them.  dif_release_retired() frees the lot when the caller knows no such graph is alive.
static std::mutex g_retired_mu;
static std::vector<void*>* g_retired = new std::vector<void*>();   // lives until the process exits
void retire_device_block(void* p) {
  if (!p) return;
  std::lock_guard<std::mutex> lock(g_retired_mu);
  g_retired->push_back(p);
}
int64_t release_retired_blocks() {
  std::vector<void*> blocks;
  {
    std::lock_guard<std::mutex> lock(g_retired_mu);
    blocks.swap(*g_retired);
  }
  for (void* p : blocks) cudaFree(p);
  return (int64_t)blocks.size();
}

static PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  });
  return fn;
}

int make_tmap_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t pitch_bytes,
                 uint32_t box_rows, uint32_t box_cols, int dtype, int atom32) {
  auto fn = get_encode_fn();
  DIF_REQUIRE(fn != nullptr, DIF_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  const int esz = dtype == 0 ? 4 : 2;
  DIF_REQUIRE(box_cols * esz == 128, DIF_ERR_INVALID, "tensor-map box must span one 128-byte swizzle row");
  DIF_REQUIRE(box_rows >= 1 && box_rows <= 256, DIF_ERR_INVALID, "tensor-map box rows %u out of range", box_rows);
  DIF_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15u) == 0 && (pitch_bytes & 15u) == 0, DIF_ERR_INVALID,
              "tensor-map base/pitch must be 16-byte aligned");
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {pitch_bytes};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, dtype == 0 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                  const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DIF_REQUIRE(r == CUDA_SUCCESS, DIF_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return DIF_OK;
}

}  // namespace dif

extern "C" {

int dif_init(int device) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    dif::set_error("no CUDA device available (%s); libdif_b200 has no CPU fallback",
                   e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    return DIF_ERR_NO_DEVICE;
  }
  DIF_REQUIRE(device >= 0 && device < n, DIF_ERR_INVALID, "device %d out of range [0,%d)", device, n);
  DIF_REQUIRE(dif::g_device < 0 || dif::g_device == device, DIF_ERR_STATE,
              "this process is bound to device %d (one process per GPU); dif_init(%d) refused", dif::g_device, device);
  DIF_CUDA_OK(cudaSetDevice(device));
  cudaDeviceProp prop;
  DIF_CUDA_OK(cudaGetDeviceProperties(&prop, device));
  DIF_REQUIRE(prop.major == 10, DIF_ERR_NO_DEVICE,
              "device %d is sm_%d%d; libdif_b200 is built for sm_100a (B200) only", device, prop.major, prop.minor);
  dif::g_sm_count = prop.multiProcessorCount;
  DIF_CUDA_OK(cudaFree(0));
  dif::g_device = device;
  return DIF_OK;
}

const char* dif_last_error(void) { return dif::g_err; }

int dif_sync(void* stream) {
  DIF_CUDA_OK(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
  return DIF_OK;
}

const char* dif_version(void) { return "dif_b200 0.1.0 (sm_100a)"; }

int64_t dif_launch_count(void) { return dif::g_launches.load(); }
int64_t dif_release_retired(void) { return dif::release_retired_blocks(); }

}  // extern "C"
