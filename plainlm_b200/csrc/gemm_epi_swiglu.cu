// gemm_kernel.cuh instantiated for one epilogue kind: fc1 with the GLU gate.
#include "gemm_kernel.cuh"

namespace plm {
PLM_DEFINE_GEMM_EPI_FORWARD(PLM_EPI_BF16_SWIGLU)
}  // namespace plm
