// gemm_kernel.cuh instantiated for one epilogue kind: fc2's input-gradient GEMM fused with the GLU backward.
#include "gemm_kernel.cuh"

namespace plm {
PLM_DEFINE_GEMM_EPI_DGRAD(PLM_EPI_BF16_GLU_BWD)
}  // namespace plm
