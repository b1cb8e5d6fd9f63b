// Detector head post-processing: channel softmax (or logistic) + dustbin removal + pixel shuffle
// (SURVEY.md section 8 f, row 2).  Semantics: reference core/modules/utils/detector_util.py:18-39 and
// :42-77 (see include/einx.h).
//
// One thread per coarse cell keeps the cell's <= 65 logits in registers: the logits are read once
// (coalesced along the coarse row, one plane per channel) and the cell x cell block of scores leaves
// as 16-byte stores, so the probability tensor of the unfused reference never exists.
#include "common.cuh"

namespace {

constexpr int kMaxC = 65;

template <int MODE>
__global__ void __launch_bounds__(128)
head_kernel(const float* __restrict__ logits, int C, int Hc, int Wc, int cell, float* __restrict__ out) {
    const int b = blockIdx.y;
    const int cellidx = blockIdx.x * blockDim.x + threadIdx.x;
    const int plane = Hc * Wc;
    if (cellidx >= plane) return;
    const float* src = logits + (size_t)b * C * plane + cellidx;
    float v[kMaxC];
#pragma unroll
    for (int c = 0; c < kMaxC; ++c) v[c] = c < C ? __ldg(src + (size_t)c * plane) : -INFINITY;
    if (MODE != EINX_HEAD_SHUFFLE) {
        if (C == 1) {
            v[0] = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-v[0])));  // 1 / (1 + exp(-x)), detector_util.py:36
        } else {
            float mx = v[0];
#pragma unroll
            for (int c = 1; c < kMaxC; ++c) mx = fmaxf(mx, v[c]);
            float sum = 0.0f;
#pragma unroll
            for (int c = 0; c < kMaxC; ++c) {
                v[c] = c < C ? expf(__fsub_rn(v[c], mx)) : 0.0f;
                sum = __fadd_rn(sum, v[c]);
            }
#pragma unroll
            for (int c = 0; c < kMaxC; ++c) v[c] = __fdiv_rn(v[c], sum);
        }
    }
    if (MODE == EINX_HEAD_PROB) {
        float* dst = out + (size_t)b * C * plane + cellidx;
#pragma unroll
        for (int c = 0; c < kMaxC; ++c)
            if (c < C) dst[(size_t)c * plane] = v[c];
        return;
    }
    // pixel shuffle: out[b, 0, hc*cell + i, wc*cell + j] = prob[b, i*cell + j, hc, wc]; the dustbin is dropped
    const int hc = cellidx / Wc, wc = cellidx - hc * Wc;
    const int W = Wc * cell;
    float* dst = out + ((size_t)b * Hc * cell + (size_t)hc * cell) * W + (size_t)wc * cell;
    if (cell == 8 && (reinterpret_cast<uintptr_t>(out) & 15u) == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float4* row = reinterpret_cast<float4*>(dst + (size_t)i * W);
            row[0] = make_float4(v[8 * i], v[8 * i + 1], v[8 * i + 2], v[8 * i + 3]);
            row[1] = make_float4(v[8 * i + 4], v[8 * i + 5], v[8 * i + 6], v[8 * i + 7]);
        }
    } else {
#pragma unroll
        for (int c = 0; c < kMaxC - 1; ++c)
            if (c < cell * cell) dst[(size_t)(c / cell) * W + (c % cell)] = v[c];
    }
}

}  // namespace

extern "C" int einx_logits_to_score(einx_ctx* ctx, const float* logits, int B, int C, int Hc, int Wc, int cell, int mode,
                                    float* out, einx_stream stream_) {
    if (!ctx) return EINX_ERR_INVALID;
    if (B < 0 || C <= 0 || Hc <= 0 || Wc <= 0 || cell <= 0)
        return einx_fail(ctx, EINX_ERR_INVALID, "einx_logits_to_score: bad shape B=%d C=%d Hc=%d Wc=%d cell=%d", B, C, Hc, Wc, cell);
    if (mode != EINX_HEAD_PROB && !((cell > 1 && C == cell * cell + 1) || (cell == 1 && C == 1)))
        return einx_fail(ctx, EINX_ERR_INVALID, "einx_logits_to_score: C=%d does not match cell=%d (detector_util.py:66,74 assert)", C, cell);
    if (C > kMaxC) return einx_fail(ctx, EINX_ERR_UNSUPPORTED, "einx_logits_to_score: C=%d > %d channels", C, kMaxC);
    if (B == 0) return EINX_OK;
    if (!logits || !out) return einx_fail(ctx, EINX_ERR_INVALID, "einx_logits_to_score: NULL pointer argument");
    if (B > 65535) return einx_fail(ctx, EINX_ERR_UNSUPPORTED, "einx_logits_to_score: B=%d > 65535", B);
    DeviceGuard guard(ctx->device);
    cudaStream_t stream = (cudaStream_t)stream_;
    const dim3 grid((Hc * Wc + 127) / 128, B);
    switch (mode) {
        case EINX_HEAD_SCORE: head_kernel<EINX_HEAD_SCORE><<<grid, 128, 0, stream>>>(logits, C, Hc, Wc, cell, out); break;
        case EINX_HEAD_PROB: head_kernel<EINX_HEAD_PROB><<<grid, 128, 0, stream>>>(logits, C, Hc, Wc, cell, out); break;
        case EINX_HEAD_SHUFFLE: head_kernel<EINX_HEAD_SHUFFLE><<<grid, 128, 0, stream>>>(logits, C, Hc, Wc, cell, out); break;
        default: return einx_fail(ctx, EINX_ERR_INVALID, "einx_logits_to_score: mode %d", mode);
    }
    EINX_CHECK_LAUNCH(ctx);
    return EINX_OK;
}
