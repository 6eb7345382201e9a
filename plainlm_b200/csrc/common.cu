#include "common.cuh"

#include <cstdarg>
#include <cstdio>
#include <mutex>

namespace plm {

static thread_local char g_err[512] = "";

int fail(int status, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return status;
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(PLM_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
  return PLM_OK;
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

// ---- driver entry points (resolved once through the runtime; no link-time dependency on libcuda)
typedef CUresult (*CtxGetCurrentFn)(CUcontext*);
typedef CUresult (*CtxSetCurrentFn)(CUcontext);
typedef CUresult (*PointerGetAttributeFn)(void*, CUpointer_attribute, CUdeviceptr);

static void* driver_fn(const char* name) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess)
    return nullptr;
  return p;
}

int ensure_context(const void* device_ptr) {
  // Calls may arrive on a thread that has no current CUDA context yet (the autograd engine's worker thread sets its
  // device lazily).  Bind the context that owns the pointer — never device 0 by default: ranks > 0 own other GPUs.
  static CtxGetCurrentFn get_cur = reinterpret_cast<CtxGetCurrentFn>(driver_fn("cuCtxGetCurrent"));
  static CtxSetCurrentFn set_cur = reinterpret_cast<CtxSetCurrentFn>(driver_fn("cuCtxSetCurrent"));
  static PointerGetAttributeFn ptr_attr = reinterpret_cast<PointerGetAttributeFn>(driver_fn("cuPointerGetAttribute"));
  if (!get_cur || !set_cur || !ptr_attr) return fail(PLM_ERR_CUDA, "CUDA driver entry points unavailable");
  CUcontext cur = nullptr;
  if (get_cur(&cur) == CUDA_SUCCESS && cur != nullptr) return PLM_OK;
  CUcontext owner = nullptr;
  CUresult r = ptr_attr(&owner, CU_POINTER_ATTRIBUTE_CONTEXT, reinterpret_cast<CUdeviceptr>(device_ptr));
  if (r != CUDA_SUCCESS || owner == nullptr)
    return fail(PLM_ERR_INVALID, "pointer %p is not a device pointer of any CUDA context (%d)", device_ptr, (int)r);
  r = set_cur(owner);
  if (r != CUDA_SUCCESS) return fail(PLM_ERR_CUDA, "cuCtxSetCurrent failed (%d)", (int)r);
  return PLM_OK;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<EncodeTiledFn>(p);
    }
  });
  return fn;
}

int make_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                      uint32_t box_rows, uint32_t box_cols) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return fail(PLM_ERR_CUDA, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  if (!aligned16(base) || (ld % 8) != 0) return fail(PLM_ERR_INVALID, "tensor map: base/ld not 16-byte aligned");
  if (box_cols * 2 != 128 || box_rows == 0 || box_rows > 256)
    return fail(PLM_ERR_INVALID, "tensor map: bad box %ux%u", box_rows, box_cols);
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estride[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estride,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(PLM_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) rows=%llu cols=%llu ld=%llu", (int)r,
                (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld);
  return PLM_OK;
}

int make_tmap_f32_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                     uint32_t box_cols) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return fail(PLM_ERR_CUDA, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  if (!aligned16(base) || (ld % 4) != 0) return fail(PLM_ERR_INVALID, "tensor map: base/ld not 16-byte aligned");
  if ((box_cols * 4 != 128 && box_cols * 4 != 64) || box_rows == 0 || box_rows > 256)
    return fail(PLM_ERR_INVALID, "tensor map: bad fp32 box %ux%u", box_rows, box_cols);
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld * 4};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estride[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstride, box, estride,
                  CU_TENSOR_MAP_INTERLEAVE_NONE,
                  box_cols * 4 == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(PLM_ERR_CUDA, "cuTensorMapEncodeTiled(f32) failed (%d) rows=%llu cols=%llu ld=%llu", (int)r,
                (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld);
  return PLM_OK;
}

}  // namespace plm

extern "C" {

int plm_abi_version(void) { return PLM_ABI_VERSION; }

const char* plm_last_error(void) { return plm::g_err; }

int plm_device_check(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return plm::fail(PLM_ERR_CUDA, "cudaGetDevice: %s", cudaGetErrorString(e));
  int major = 0;
  e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e != cudaSuccess) return plm::fail(PLM_ERR_CUDA, "cudaDeviceGetAttribute: %s", cudaGetErrorString(e));
  if (major != 10) return plm::fail(PLM_ERR_UNSUPPORTED, "device compute capability %d.x, need 10.x (B200)", major);
  return PLM_OK;
}

}  // extern "C"
