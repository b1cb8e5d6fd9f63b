// Event accumulation image and detector event mask (SURVEY.md section 8 f, row 1).
// Semantics: reference datasets/visualize.py:23-49 (dict branch) and
// core/modules/event_extractors/EventExtractors.py:357-363 (see include/einx.h).
//
// The reference builds the image with one Python-level loop iteration per event; here a window is a
// histogram of integer pixel hits (int32 atomics into an L2-resident count plane) followed by one
// CTA per window that reduces min / max and writes the fp64 min-max normalised uint8 image.
#include <limits.h>

#include "common.cuh"

namespace {

// pol == nullptr: += 1 per event (dict branch, visualize.py:37-40); else += 2 * p - 1 (array branch, :42-44)
template <typename T>
__global__ void __launch_bounds__(256)
events_count_kernel(const T* __restrict__ x, const T* __restrict__ y, const T* __restrict__ pol,
                    const int64_t* __restrict__ off, int H, int W, int* __restrict__ cnt) {
    const int b = blockIdx.y;
    const int64_t beg = off[b], end = off[b + 1];
    int* c = cnt + (size_t)b * H * W;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = beg + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < end; i += stride) {
        const int ix = (int)__ldg(x + i), iy = (int)__ldg(y + i);  // int() truncates toward zero
        const int w = pol ? (int)(2.0 * (double)__ldg(pol + i) - 1.0) : 1;  // integer-valued polarities (host-checked)
        if ((unsigned)ix < (unsigned)W && (unsigned)iy < (unsigned)H) atomicAdd(c + (size_t)iy * W + ix, w);
    }
}

constexpr int kNormThreads = 1024;

__global__ void __launch_bounds__(kNormThreads)
events_normalize_kernel(const int* __restrict__ cnt, int npix, uint8_t* __restrict__ image) {
    const int b = blockIdx.x;
    const int* c = cnt + (size_t)b * npix;
    uint8_t* out = image + (size_t)b * npix;
    int mn = INT_MAX, mx = INT_MIN;
    for (int i = threadIdx.x; i < npix; i += kNormThreads) {
        const int v = c[i];
        mn = min(mn, v);
        mx = max(mx, v);
    }
    mn = __reduce_min_sync(0xffffffffu, mn);
    mx = __reduce_max_sync(0xffffffffu, mx);
    __shared__ int smn[kNormThreads / 32], smx[kNormThreads / 32];
    if ((threadIdx.x & 31) == 0) { smn[threadIdx.x >> 5] = mn; smx[threadIdx.x >> 5] = mx; }
    __syncthreads();
    mn = smn[0]; mx = smx[0];
#pragma unroll
    for (int w = 1; w < kNormThreads / 32; ++w) { mn = min(mn, smn[w]); mx = max(mx, smx[w]); }
    const double range = (double)(mx - mn);
    for (int i = threadIdx.x; i < npix; i += kNormThreads) {
        uint8_t o = 0;
        if (mx != mn) {
            double v = (double)(c[i] - mn) / range * 255.0;  // visualize.py:45, numpy fp64
            if (v > 255.0) v = 255.0;                        // :46
            o = (uint8_t)v;                                  // :48 astype(uint8) truncates
        }
        out[i] = o;
    }
}

__global__ void __launch_bounds__(256)
mask_dilate_kernel(const uint8_t* __restrict__ image, int H, int W, int pad_top, int pad_left, int Hp, int Wp,
                   uint8_t* __restrict__ mask) {
    const int b = blockIdx.z;
    const int xp = blockIdx.x * blockDim.x + threadIdx.x, yp = blockIdx.y;
    if (xp >= Wp) return;
    const uint8_t* img = image + (size_t)b * H * W;
    const int y = yp - pad_top, x = xp - pad_left;
    int any = 0;
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
        const int yy = y + dy;
        if ((unsigned)yy >= (unsigned)H) continue;
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) {
            const int xx = x + dx;
            if ((unsigned)xx < (unsigned)W) any |= img[(size_t)yy * W + xx];
        }
    }
    mask[((size_t)b * Hp + yp) * Wp + xp] = any ? 1 : 0;
}

// compact wire format of integer-pixel events -> the fp32 SoA the voxeliser reads
__global__ void __launch_bounds__(256)
unpack_events_kernel(const uint16_t* __restrict__ x, const uint16_t* __restrict__ y, const int8_t* __restrict__ p, size_t n,
                     float* __restrict__ xo, float* __restrict__ yo, float* __restrict__ po) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        xo[i] = (float)x[i];
        yo[i] = (float)y[i];
        po[i] = (float)p[i];
    }
}

}  // namespace

namespace {

int events_image_impl(einx_ctx* ctx, const char* who, const void* x, const void* y, const void* pol, int coord_f64,
                      const int64_t* ev_offsets, int B, int H, int W, uint8_t* image, einx_stream stream_) {
    if (B < 0 || H <= 0 || W <= 0) return einx_fail(ctx, EINX_ERR_INVALID, "%s: bad shape B=%d H=%d W=%d", who, B, H, W);
    if (B == 0) return EINX_OK;
    if (!x || !y || !ev_offsets || !image) return einx_fail(ctx, EINX_ERR_INVALID, "%s: NULL pointer argument", who);
    if (B > 65535) return einx_fail(ctx, EINX_ERR_UNSUPPORTED, "%s: B=%d > 65535", who, B);
    if ((size_t)H * W > (size_t)INT_MAX) return einx_fail(ctx, EINX_ERR_UNSUPPORTED, "%s: image too large", who);
    DeviceGuard guard(ctx->device);
    cudaStream_t stream = (cudaStream_t)stream_;
    const size_t npix = (size_t)H * W;
    int rc = einx_ws_reserve(ctx, sizeof(int) * npix * B, stream);
    if (rc) return rc;
    int* cnt = (int*)ctx->ws;
    EINX_CUDA(ctx, cudaMemsetAsync(cnt, 0, sizeof(int) * npix * B, stream));
    int per_window = (ctx->num_sms * 8 + B - 1) / B;
    if (per_window < 1) per_window = 1;
    if (per_window > 1024) per_window = 1024;
    if (coord_f64)
        events_count_kernel<double><<<dim3(per_window, B), 256, 0, stream>>>((const double*)x, (const double*)y, (const double*)pol, ev_offsets, H, W, cnt);
    else
        events_count_kernel<float><<<dim3(per_window, B), 256, 0, stream>>>((const float*)x, (const float*)y, (const float*)pol, ev_offsets, H, W, cnt);
    EINX_CHECK_LAUNCH(ctx);
    events_normalize_kernel<<<B, kNormThreads, 0, stream>>>(cnt, (int)npix, image);
    EINX_CHECK_LAUNCH(ctx);
    return EINX_OK;
}

}  // namespace

extern "C" int einx_events_image(einx_ctx* ctx, const void* x, const void* y, int coord_f64, const int64_t* ev_offsets,
                                 int B, int H, int W, uint8_t* image, einx_stream stream_) {
    if (!ctx) return EINX_ERR_INVALID;
    return events_image_impl(ctx, "einx_events_image", x, y, nullptr, coord_f64, ev_offsets, B, H, W, image, stream_);
}

extern "C" int einx_events_image_signed(einx_ctx* ctx, const void* x, const void* y, const void* p, int coord_f64,
                                        const int64_t* ev_offsets, int B, int H, int W, uint8_t* image, einx_stream stream_) {
    if (!ctx) return EINX_ERR_INVALID;
    if (B > 0 && !p) return einx_fail(ctx, EINX_ERR_INVALID, "einx_events_image_signed: NULL pointer argument");
    return events_image_impl(ctx, "einx_events_image_signed", x, y, p, coord_f64, ev_offsets, B, H, W, image, stream_);
}

extern "C" int einx_mask_dilate(einx_ctx* ctx, const uint8_t* image, int B, int H, int W, int pad_top, int pad_left, int Hp,
                                int Wp, uint8_t* mask, einx_stream stream_) {
    if (!ctx) return EINX_ERR_INVALID;
    if (B < 0 || H <= 0 || W <= 0 || pad_top < 0 || pad_left < 0 || Hp < H + pad_top || Wp < W + pad_left)
        return einx_fail(ctx, EINX_ERR_INVALID, "einx_mask_dilate: bad shape %dx%d at (%d,%d) in %dx%d", H, W, pad_top, pad_left, Hp, Wp);
    if (B == 0) return EINX_OK;
    if (!image || !mask) return einx_fail(ctx, EINX_ERR_INVALID, "einx_mask_dilate: NULL pointer argument");
    if (B > 65535 || Hp > 65535) return einx_fail(ctx, EINX_ERR_UNSUPPORTED, "einx_mask_dilate: B or Hp > 65535");
    DeviceGuard guard(ctx->device);
    mask_dilate_kernel<<<dim3((Wp + 255) / 256, Hp, B), 256, 0, (cudaStream_t)stream_>>>(image, H, W, pad_top, pad_left, Hp, Wp, mask);
    EINX_CHECK_LAUNCH(ctx);
    return EINX_OK;
}

extern "C" int einx_unpack_events(einx_ctx* ctx, const uint16_t* x, const uint16_t* y, const int8_t* p, int64_t n, float* xo,
                                  float* yo, float* po, einx_stream stream_) {
    if (!ctx) return EINX_ERR_INVALID;
    if (n < 0) return einx_fail(ctx, EINX_ERR_INVALID, "einx_unpack_events: n=%lld", (long long)n);
    if (n == 0) return EINX_OK;
    if (!x || !y || !p || !xo || !yo || !po) return einx_fail(ctx, EINX_ERR_INVALID, "einx_unpack_events: NULL pointer argument");
    DeviceGuard guard(ctx->device);
    size_t blocks = ((size_t)n + 255) / 256;
    if (blocks > (size_t)ctx->num_sms * 16) blocks = (size_t)ctx->num_sms * 16;
    unpack_events_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream_>>>(x, y, p, (size_t)n, xo, yo, po);
    EINX_CHECK_LAUNCH(ctx);
    return EINX_OK;
}
