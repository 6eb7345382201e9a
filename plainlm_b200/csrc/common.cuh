// Host-side helpers shared by every translation unit of libplainlm_b200.so:
// status codes, last-error slot, launch checks, TMA tensor-map encoding.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/plainlm_b200.h"

namespace plm {

// Records a failure message (thread-local) and returns the status code so call sites can `return fail(...)`.
int fail(int status, const char* fmt, ...);

// Returns PLM_OK or records the CUDA error of the launch that was just enqueued.
int check_launch(const char* what);

int sm_count();

// Makes the CUDA context that owns `device_ptr` current on the calling thread if the thread has none.
int ensure_context(const void* device_ptr);

// Encodes a 2-D bf16 tensor map: tensor is [rows, cols] row-major with leading dimension `ld` (elements);
// a box is [box_rows, box_cols] elements, 128-byte swizzle (box_cols * 2 bytes must be 128).
// Out-of-bounds elements are zero-filled by the hardware.
int make_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                      uint32_t box_rows, uint32_t box_cols);

// Same for an fp32 tensor: box_cols * 4 bytes must be 128 (128-byte swizzle) or 64 (64-byte swizzle).
int make_tmap_f32_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                     uint32_t box_cols);

#define PLM_ENSURE_CONTEXT(ptr)                            \
  do {                                                    \
    if ((ptr) != nullptr) {                               \
      const int rc_ = ::plm::ensure_context(ptr);         \
      if (rc_ != PLM_OK) return rc_;                      \
    }                                                     \
  } while (0)

#define PLM_REQUIRE(cond, ...)                                  \
  do {                                                          \
    if (!(cond)) return ::plm::fail(PLM_ERR_INVALID, __VA_ARGS__); \
  } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace plm
