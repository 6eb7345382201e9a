// gemm_kernel.cuh instantiated for one epilogue kind: fp32 store / fp32 bulk reduce-add (PLM_EPI_ATOMIC_F32: weight gradients, split-K).
#include "gemm_kernel.cuh"

namespace plm {
PLM_DEFINE_GEMM_EPI_GENERAL(PLM_EPI_F32)
}  // namespace plm
