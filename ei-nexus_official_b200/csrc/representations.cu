// The reference's other two scatter representations (SURVEY.md section 8 f, row 4): event stack and time surface.
// Semantics: reference datasets/representations.py:177-214 (events_to_event_stack) and :25-63
// (events_to_time_surface), both after time_normalization (:8-22); see include/einx.h.
//
// Both bin the time-sorted events of a window with np.searchsorted(t, i*dt, 'left') .. searchsorted(t, i*dt + dt,
// 'right'): an event belongs to bin i iff  i*dt <= t <= i*dt + dt  in fp64 -- both ends inclusive, so an event
// exactly on a boundary lands in two bins.  Each thread re-evaluates those two comparisons for the (at most three)
// candidate bins around floor(t * bins) with the same fp64 expressions, so membership is bit-exact.
#include "common.cuh"

namespace {

struct Norm {
    double t0, denom;  // t <- (t - t[0]) / (t[-1] - t[0] + 1e-8)     representations.py:19-20
};

__device__ __forceinline__ Norm window_norm(const double* __restrict__ t, int64_t beg, int64_t end) {
    Norm n;
    n.t0 = t[beg];
    n.denom = (t[end - 1] - n.t0) + 1e-8;
    return n;
}

// MODE 0: event stack  (int32 sum of 2p-1 per (bin, y, x))          :192, :209-210
// MODE 1: time surface (latest normalised time per (2*bin + p, y, x)) :57; time-sorted => the maximum
template <int MODE>
__global__ void __launch_bounds__(256)
binned_scatter_kernel(const float* __restrict__ x, const float* __restrict__ y, const double* __restrict__ t,
                      const float* __restrict__ p, const int64_t* __restrict__ off, int nbins, int channels, int H, int W,
                      int* __restrict__ out) {
    const int b = blockIdx.y;
    const int64_t beg = off[b], end = off[b + 1];
    if (end - beg <= 0) return;
    const Norm nm = window_norm(t, beg, end);
    const double dt = 1.0 / (double)nbins;
    int* g = out + (size_t)b * channels * H * W;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = beg + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < end; i += stride) {
        const double tn = (__ldg(t + i) - nm.t0) / nm.denom;
        const int ix = (int)__ldg(x + i), iy = (int)__ldg(y + i), ip = (int)__ldg(p + i);  // astype(np.int32) truncates
        if ((unsigned)ix >= (unsigned)W || (unsigned)iy >= (unsigned)H) continue;
        const int c0 = (int)(tn * (double)nbins);
#pragma unroll
        for (int d = -1; d <= 1; ++d) {
            const int k = c0 + d;
            if (k < 0 || k >= nbins) continue;
            const double lo = (double)k * dt, hi = lo + dt;  // t0_bin = i_bin * dt_bin; t1_bin = t0_bin + dt_bin
            if (!(lo <= tn && tn <= hi)) continue;
            if (MODE == 0) {
                atomicAdd(g + ((size_t)k * H + iy) * W + ix, 2 * ip - 1);
            } else {
                const int ch = 2 * k + ip;
                if (ch < 0 || ch >= channels) continue;  // polarities other than 0/1 have no channel of their own
                atomicMax(reinterpret_cast<unsigned int*>(g) + ((size_t)ch * H + iy) * W + ix,
                          __float_as_uint((float)tn));  // non-negative floats order like their bit patterns
            }
        }
    }
}

__global__ void int_to_float_kernel(const int* __restrict__ src, float* __restrict__ dst, size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = (float)src[i];
}

int launch_binned(einx_ctx* ctx, int mode, const float* x, const float* y, const double* t, const float* p,
                  const int64_t* off, int B, int nbins, int channels, int H, int W, float* out, cudaStream_t stream) {
    const size_t n = (size_t)B * channels * H * W;
    int per_window = (ctx->num_sms * 8 + B - 1) / B;
    if (per_window < 1) per_window = 1;
    if (per_window > 1024) per_window = 1024;
    if (mode == 0) {
        int rc = einx_ws_reserve(ctx, sizeof(int) * n, stream);
        if (rc) return rc;
        int* acc = (int*)ctx->ws;
        EINX_CUDA(ctx, cudaMemsetAsync(acc, 0, sizeof(int) * n, stream));
        binned_scatter_kernel<0><<<dim3(per_window, B), 256, 0, stream>>>(x, y, t, p, off, nbins, channels, H, W, acc);
        EINX_CHECK_LAUNCH(ctx);
        size_t blocks = (n + 255) / 256;
        if (blocks > (size_t)ctx->num_sms * 16) blocks = (size_t)ctx->num_sms * 16;
        int_to_float_kernel<<<(unsigned)blocks, 256, 0, stream>>>(acc, out, n);
        EINX_CHECK_LAUNCH(ctx);
    } else {
        EINX_CUDA(ctx, cudaMemsetAsync(out, 0, sizeof(float) * n, stream));
        binned_scatter_kernel<1><<<dim3(per_window, B), 256, 0, stream>>>(x, y, t, p, off, nbins, channels, H, W, (int*)out);
        EINX_CHECK_LAUNCH(ctx);
    }
    return EINX_OK;
}

int check_args(einx_ctx* ctx, const char* who, const void* x, const void* y, const void* t, const void* p, const void* off,
               int B, int bins, int H, int W, const void* out) {
    if (B < 0 || bins <= 0 || H <= 0 || W <= 0) return einx_fail(ctx, EINX_ERR_INVALID, "%s: bad shape B=%d bins=%d H=%d W=%d", who, B, bins, H, W);
    if (B > 0 && (!x || !y || !t || !p || !off || !out)) return einx_fail(ctx, EINX_ERR_INVALID, "%s: NULL pointer argument", who);
    if (B > 65535) return einx_fail(ctx, EINX_ERR_UNSUPPORTED, "%s: B=%d > 65535", who, B);
    return EINX_OK;
}

}  // namespace

extern "C" int einx_event_stack(einx_ctx* ctx, const float* x, const float* y, const double* t, const float* p,
                                const int64_t* ev_offsets, int B, int bins, int H, int W, float* out, einx_stream stream) {
    if (!ctx) return EINX_ERR_INVALID;
    int rc = check_args(ctx, "einx_event_stack", x, y, t, p, ev_offsets, B, bins, H, W, out);
    if (rc || B == 0) return rc;
    DeviceGuard guard(ctx->device);
    return launch_binned(ctx, 0, x, y, t, p, ev_offsets, B, bins, bins, H, W, out, (cudaStream_t)stream);
}

extern "C" int einx_time_surface(einx_ctx* ctx, const float* x, const float* y, const double* t, const float* p,
                                 const int64_t* ev_offsets, int B, int bins, int H, int W, float* out, einx_stream stream) {
    if (!ctx) return EINX_ERR_INVALID;
    int rc = check_args(ctx, "einx_time_surface", x, y, t, p, ev_offsets, B, bins, H, W, out);
    if (rc || B == 0) return rc;
    if (bins < 2) return einx_fail(ctx, EINX_ERR_INVALID, "einx_time_surface: bins=%d (n_bins = bins // 2 would be 0)", bins);
    DeviceGuard guard(ctx->device);
    return launch_binned(ctx, 1, x, y, t, p, ev_offsets, B, bins / 2, bins, H, W, out, (cudaStream_t)stream);
}
