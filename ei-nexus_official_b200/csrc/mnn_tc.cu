// Tensor-core (tcgen05) similarity tiles for the MNN matcher -- placeholder until the kernel lands.
#include "common.cuh"

size_t einx_mnn_tc_scratch_bytes(int, int, int, int, int) { return 0; }

int einx_mnn_tc(einx_ctx* ctx, const float*, const float*, const int32_t*, const int32_t*, int, int, int, int,
                int precision, unsigned long long*, unsigned long long*, unsigned char*, size_t, cudaStream_t) {
    return einx_fail(ctx, EINX_ERR_UNSUPPORTED, "einx_mnn: precision %d (tensor-core path) is not built yet", precision);
}
