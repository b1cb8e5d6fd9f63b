// Descriptor sampling at the kept keypoints + L2 normalisation (one warp per keypoint).
// Semantics: reference core/modules/utils/descriptor_util.py:21-28, :50-71 (gather) and :74-128
// (bilinear grid_sample, align_corners=False, zeros padding) -- see include/einx.h.
#include "common.cuh"

namespace {

constexpr int kWarpsPerBlock = 8;
constexpr int kMaxPerLane = 16;  // channels per lane held in registers: C <= 512

template <int MODE>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
sample_kernel(const float* __restrict__ raw, int C, int Hd, int Wd, float Hp, float Wp,
              const float* __restrict__ kpts, const int32_t* __restrict__ counts, int kcap, float scale,
              int normalize, float* __restrict__ desc) {
    const int b = blockIdx.y;
    const int k = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (k >= kcap) return;
    float* out = desc + ((size_t)b * kcap + k) * C;
    int cnt = counts[b];
    if (cnt > kcap) cnt = kcap;
    if (k >= cnt) {  // padding rows are defined (zero) so a batched matcher can ignore them safely
        for (int c = lane; c < C; c += 32) out[c] = 0.0f;
        return;
    }
    const float* kp = kpts + ((size_t)b * kcap + k) * 3;
    const float py = kp[0], px = kp[1];
    const float* img = raw + (size_t)b * C * Hd * Wd;
    const size_t plane = (size_t)Hd * Wd;
    float v[kMaxPerLane];
    float ss = 0.0f;
    if (MODE == EINX_SAMPLE_GATHER) {
        // pos.floor().long() -> raw[i, :, y, x]                          (:57-60)
        int yy = (int)floorf(py), xx = (int)floorf(px);
        yy = min(max(yy, 0), Hd - 1);
        xx = min(max(xx, 0), Wd - 1);
        const float* src = img + (size_t)yy * Wd + xx;
#pragma unroll
        for (int j = 0; j < kMaxPerLane; ++j) {
            const int c = lane + 32 * j;
            v[j] = c < C ? __ldg(src + c * plane) : 0.0f;
        }
    } else {
        // pos - 0.5 -> [-1, 1] on the padded image -> grid_sample un-normalisation  (:105-120)
        const float gy = __fsub_rn(__fmul_rn(2.0f, __fdiv_rn(__fsub_rn(py, 0.5f), __fsub_rn(Hp, 1.0f))), 1.0f);
        const float gx = __fsub_rn(__fmul_rn(2.0f, __fdiv_rn(__fsub_rn(px, 0.5f), __fsub_rn(Wp, 1.0f))), 1.0f);
        const float iy = __fdiv_rn(__fsub_rn(__fmul_rn(__fadd_rn(gy, 1.0f), (float)Hd), 1.0f), 2.0f);
        const float ix = __fdiv_rn(__fsub_rn(__fmul_rn(__fadd_rn(gx, 1.0f), (float)Wd), 1.0f), 2.0f);
        const float fy = floorf(iy), fx = floorf(ix);
        const int y0 = (int)fy, x0 = (int)fx;
        const float wy1 = __fsub_rn(iy, fy), wx1 = __fsub_rn(ix, fx);
        const float wy0 = __fsub_rn(__fadd_rn(fy, 1.0f), iy), wx0 = __fsub_rn(__fadd_rn(fx, 1.0f), ix);
        const float w00 = __fmul_rn(wx0, wy0), w01 = __fmul_rn(wx1, wy0);
        const float w10 = __fmul_rn(wx0, wy1), w11 = __fmul_rn(wx1, wy1);
        const bool oy0 = (y0 >= 0) & (y0 < Hd), oy1 = (y0 + 1 >= 0) & (y0 + 1 < Hd);
        const bool ox0 = (x0 >= 0) & (x0 < Wd), ox1 = (x0 + 1 >= 0) & (x0 + 1 < Wd);
        const float* src = img + (ptrdiff_t)y0 * Wd + x0;
#pragma unroll
        for (int j = 0; j < kMaxPerLane; ++j) {
            const int c = lane + 32 * j;
            float acc = 0.0f;
            if (c < C) {
                const float* s = src + c * plane;
                // tap order nw, ne, sw, se; out-of-map taps contribute zero
                if (oy0 && ox0) acc = __fadd_rn(acc, __fmul_rn(__ldg(s), w00));
                if (oy0 && ox1) acc = __fadd_rn(acc, __fmul_rn(__ldg(s + 1), w01));
                if (oy1 && ox0) acc = __fadd_rn(acc, __fmul_rn(__ldg(s + Wd), w10));
                if (oy1 && ox1) acc = __fadd_rn(acc, __fmul_rn(__ldg(s + Wd + 1), w11));
            }
            v[j] = acc;
        }
    }
#pragma unroll
    for (int j = 0; j < kMaxPerLane; ++j) ss = fmaf(v[j], v[j], ss);
    float mul = scale, den = 1.0f;
    if (normalize) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        den = fmaxf(sqrtf(ss), 1e-12f);  // F.normalize eps
    }
#pragma unroll
    for (int j = 0; j < kMaxPerLane; ++j) {
        const int c = lane + 32 * j;
        if (c < C) out[c] = __fmul_rn(mul, normalize ? __fdiv_rn(v[j], den) : v[j]);
    }
}

}  // namespace

extern "C" int einx_sample(einx_ctx* ctx, const float* raw, int B, int C, int Hd, int Wd, int mode, int Hp, int Wp,
                           const float* kpts, const int32_t* counts, int kcap, float scale, int normalize,
                           float* desc, einx_stream stream_) {
    if (!ctx) return EINX_ERR_INVALID;
    if (B < 0 || C <= 0 || Hd <= 0 || Wd <= 0 || kcap < 0)
        return einx_fail(ctx, EINX_ERR_INVALID, "einx_sample: bad shape B=%d C=%d Hd=%d Wd=%d kcap=%d", B, C, Hd, Wd, kcap);
    if (B == 0 || kcap == 0) return EINX_OK;
    if (!raw || !kpts || !counts || !desc) return einx_fail(ctx, EINX_ERR_INVALID, "einx_sample: NULL pointer argument");
    if (C > 32 * kMaxPerLane) return einx_fail(ctx, EINX_ERR_UNSUPPORTED, "einx_sample: C=%d > %d", C, 32 * kMaxPerLane);
    if (B > 65535) return einx_fail(ctx, EINX_ERR_UNSUPPORTED, "einx_sample: B=%d > 65535", B);
    DeviceGuard guard(ctx->device);
    cudaStream_t stream = (cudaStream_t)stream_;
    dim3 grid((kcap + kWarpsPerBlock - 1) / kWarpsPerBlock, B);
    if (mode == EINX_SAMPLE_GATHER) {
        sample_kernel<EINX_SAMPLE_GATHER><<<grid, kWarpsPerBlock * 32, 0, stream>>>(
            raw, C, Hd, Wd, (float)Hp, (float)Wp, kpts, counts, kcap, scale, normalize, desc);
    } else if (mode == EINX_SAMPLE_BILINEAR) {
        if (Hp <= 1 || Wp <= 1) return einx_fail(ctx, EINX_ERR_INVALID, "einx_sample: bilinear needs Hp, Wp > 1");
        sample_kernel<EINX_SAMPLE_BILINEAR><<<grid, kWarpsPerBlock * 32, 0, stream>>>(
            raw, C, Hd, Wd, (float)Hp, (float)Wp, kpts, counts, kcap, scale, normalize, desc);
    } else {
        return einx_fail(ctx, EINX_ERR_INVALID, "einx_sample: unknown mode %d", mode);
    }
    EINX_CHECK_LAUNCH(ctx);
    return EINX_OK;
}
