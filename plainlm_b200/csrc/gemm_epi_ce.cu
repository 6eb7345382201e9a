// gemm_kernel.cuh instantiated for one epilogue kind: LM head with the cross-entropy statistics.
#include "gemm_kernel.cuh"

namespace plm {
PLM_DEFINE_GEMM_EPI_FORWARD(PLM_EPI_BF16_CE)
}  // namespace plm
