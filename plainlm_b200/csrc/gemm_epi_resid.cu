// gemm_kernel.cuh instantiated for one epilogue kind: fp32 residual add (attention out-projection, fc2).
#include "gemm_kernel.cuh"

namespace plm {
PLM_DEFINE_GEMM_EPI_GENERAL(PLM_EPI_RESID_F32)
}  // namespace plm
