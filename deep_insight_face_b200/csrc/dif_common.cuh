// Host-side plumbing shared by the C-ABI translation units: error string, CUDA checks,
// TMA tensor-map construction through the driver entry point (no link-time libcuda dependency,
// so the library still loads on a box without a GPU driver).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../include/dif_b200.h"

namespace dif {

void set_error(const char* fmt, ...);
int count_launch(int n = 1);   // bumps the library-wide launch counter (dif_launch_count)

#define DIF_CUDA_OK(expr)                                                                       \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess) {                                                                    \
      dif::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return DIF_ERR_CUDA;                                                                      \
    }                                                                                           \
  } while (0)

#define DIF_REQUIRE(cond, code, ...)   \
  do {                                 \
    if (!(cond)) {                     \
      dif::set_error(__VA_ARGS__);     \
      return (code);                   \
    }                                  \
  } while (0)

#define DIF_LAUNCH_OK()                                                                   \
  do {                                                                                    \
    cudaError_t _e = cudaGetLastError();                                                  \
    if (_e != cudaSuccess) {                                                              \
      dif::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
      return DIF_ERR_CUDA;                                                                \
    }                                                                                     \
    dif::count_launch();                                                                  \
  } while (0)

// 2-D row-major [rows][cols] tensor of `elem_bytes` elements, box = [box_rows][box_cols],
// 128-byte swizzle (box_cols * elem_bytes must be 128).  dtype: 0 = fp32 (consumed as tf32), 1 = bf16.
// atom32 = 1: CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B (32-byte chunks swizzled over 4 rows) - the only shared-memory
// layout the tensor core accepts for an MN-major TF32 operand (UMMA layout type SWIZZLE_128B_BASE32B).
int make_tmap_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t pitch_bytes,
                 uint32_t box_rows, uint32_t box_cols, int dtype, int atom32 = 0);

inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

int device_sm_count();
// Loss workspaces only grow.  A block that is outgrown is retired, not freed: CUDA graphs captured earlier
// (BatchHardStep / ArcFaceStep of another shape) hold its address in their kernel nodes and must stay valid.
void retire_device_block(void* p);
int64_t release_retired_blocks();

}  // namespace dif
