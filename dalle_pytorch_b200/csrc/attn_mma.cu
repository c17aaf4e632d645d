// Tensor-core (bf16) fused attention.  Placeholder translation unit: until the tensor-core kernels land, every
// request is served by the fp32-arithmetic CUDA-core kernels of attn_simt.cu (which also handle bf16 storage).
#include "attn_common.cuh"

namespace db200 {

bool attn_mma_supported(const db200_attn_fwd_params&) { return false; }
int attn_fwd_mma_launch(const db200_attn_fwd_params&, cudaStream_t) { return set_error(DB200_ERR_UNSUPPORTED, "attn_mma: not built"); }
int attn_bwd_mma_launch(const db200_attn_bwd_params&, cudaStream_t) { return set_error(DB200_ERR_UNSUPPORTED, "attn_mma: not built"); }

}  // namespace db200
