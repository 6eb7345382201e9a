// gemm_kernel.cuh instantiated for one epilogue kind: plain bf16 store.
#include "gemm_kernel.cuh"

namespace plm {
PLM_DEFINE_GEMM_EPI_GENERAL(PLM_EPI_BF16)
}  // namespace plm
