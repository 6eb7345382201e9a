// gemm_kernel.cuh instantiated for one epilogue kind: bf16 store with RoPE on the q|k columns (QKV projection).
#include "gemm_kernel.cuh"

namespace plm {
PLM_DEFINE_GEMM_EPI_FORWARD(PLM_EPI_BF16_ROPE)
}  // namespace plm
