// Event distance map (SURVEY.md section 8 f, row 4).  Semantics: reference datasets/representations.py:215-248
// (events_to_distance_map): per time bin, cv.distanceTransform(1 - event_map, DIST_L2, 3) -- the 3x3 chamfer
// distance (axial weight a = 0.955, diagonal b = 1.3693) of every pixel to the nearest event pixel of the bin.
//
// On an unobstructed grid the two-pass chamfer scan equals the closed form
//     D(p) = min over event pixels q of  b * min(|dx|, |dy|) + a * (max(|dx|, |dy|) - min(|dx|, |dy|)),
// evaluated here in fp64 and rounded once (the test-side restatement is pinned to OpenCV within 1e-6
// relative).  The form is monotone in |dx| for a fixed |dy|, so the nearest event of a ROW is its horizontally
// nearest one: three passes, none of them iterative --
//   1. mark      events -> one bit per (bin, y, x)   (bin membership: i/bins <= t <= (i+1)/bins in fp64, both ends
//                inclusive: np.searchsorted 'left' .. 'right')
//   2. row pass  horizontal distance to the nearest set bit of the same row (uint16, 0xffff: empty row)
//   3. column pass  D(x, y) = min over rows y' of d(rowdist[y'][x], |y - y'|), walking outwards from y and
//                stopping once a * |dy| can no longer beat the best -- a few rows on event-dense maps.
// A bin without events yields FLT_MAX everywhere (OpenCV 4.x).
#include <float.h>

#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256)
dm_mark_kernel(const float* __restrict__ x, const float* __restrict__ y, const double* __restrict__ t,
               const int64_t* __restrict__ off, int nbins, int H, int W, int WW, uint32_t* __restrict__ bits) {
    const int b = blockIdx.y;
    const int64_t beg = off[b], end = off[b + 1];
    if (end - beg <= 0) return;
    const double t0 = t[beg], denom = (t[end - 1] - t0) + 1e-8;   // representations.py:19-20
    const double ct = 1.0 / (double)nbins;                         // channel_t = 1 / bins
    uint32_t* g = bits + (size_t)b * nbins * H * WW;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = beg + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < end; i += stride) {
        const double tn = (__ldg(t + i) - t0) / denom;
        const int ix = (int)__ldg(x + i), iy = (int)__ldg(y + i);  // astype(np.int32) truncates
        if ((unsigned)ix >= (unsigned)W || (unsigned)iy >= (unsigned)H) continue;
        const int c0 = (int)(tn * (double)nbins);
#pragma unroll
        for (int d = -1; d <= 1; ++d) {
            const int k = c0 + d;
            if (k < 0 || k >= nbins) continue;
            if (!((double)k * ct <= tn && tn <= (double)(k + 1) * ct)) continue;  // i * channel_t .. (i + 1) * channel_t
            atomicOr(g + ((size_t)k * H + iy) * WW + (ix >> 5), 1u << (ix & 31));
        }
    }
}

// one thread per (row, 32-pixel word): nearest set bit to the left / right of each of its pixels
__global__ void __launch_bounds__(128)
dm_row_kernel(const uint32_t* __restrict__ bits, int W, int WW, size_t nrows, uint16_t* __restrict__ rowdist) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nrows * WW) return;
    const size_t row = idx / WW;
    const int w = (int)(idx - row * WW);
    const uint32_t* r = bits + row * WW;
    const uint32_t cur = r[w];
    int left = -0x10000, right = 0x20000;  // x of the nearest set bit in the words before / after (none: far away)
    for (int k = w - 1; k >= 0; --k) {
        const uint32_t v = r[k];
        if (v) { left = 32 * k + 31 - __clz(v); break; }
    }
    for (int k = w + 1; k < WW; ++k) {
        const uint32_t v = r[k];
        if (v) { right = 32 * k + __ffs(v) - 1; break; }
    }
    uint16_t* out = rowdist + row * W;
#pragma unroll 4
    for (int bpos = 0; bpos < 32; ++bpos) {
        const int xx = 32 * w + bpos;
        if (xx >= W) break;
        const uint32_t le = cur & (0xffffffffu >> (31 - bpos));   // bits at or left of bpos
        const uint32_t ge = cur & (0xffffffffu << bpos);          // bits at or right of bpos
        const int xl = le ? 32 * w + 31 - __clz(le) : left;
        const int xr = ge ? 32 * w + __ffs(ge) - 1 : right;
        const int d = min(xx - xl, xr - xx);
        out[xx] = (uint16_t)(d > 0xfffe ? 0xffff : d);
    }
}

__device__ __forceinline__ double chamfer(int dx, int dy) {
    const double a = (double)0.955f, b = (double)1.3693f;  // OpenCV's fp32 weights, widened
    const int mn = min(dx, dy), mx = max(dx, dy);
    return __dadd_rn(__dmul_rn(b, (double)mn), __dmul_rn(a, (double)(mx - mn)));  // no contraction: numpy does not fuse
}

__global__ void __launch_bounds__(256)
dm_column_kernel(const uint16_t* __restrict__ rowdist, int H, int W, size_t nplanes, float* __restrict__ out) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nplanes * H * W) return;
    const size_t plane = idx / ((size_t)H * W);
    const int rem = (int)(idx - plane * H * W);
    const int yy = rem / W, xx = rem - yy * W;
    const uint16_t* rd = rowdist + plane * H * W + xx;
    const double a = (double)0.955f;
    double best = DBL_MAX;
    for (int dy = 0; dy < H; ++dy) {
        if (__dmul_rn(a, (double)dy) >= best) break;  // d >= a * max(|dx|, |dy|) >= a * |dy|
        const int yu = yy - dy, yd = yy + dy;
        if (yu < 0 && yd >= H) break;
        if (yu >= 0) {
            const int dx = rd[(size_t)yu * W];
            if (dx != 0xffff) best = fmin(best, chamfer(dx, dy));
        }
        if (dy > 0 && yd < H) {
            const int dx = rd[(size_t)yd * W];
            if (dx != 0xffff) best = fmin(best, chamfer(dx, dy));
        }
    }
    out[idx] = best == DBL_MAX ? FLT_MAX : (float)best;
}

}  // namespace

extern "C" int einx_distance_map(einx_ctx* ctx, const float* x, const float* y, const double* t, const int64_t* ev_offsets,
                                 int B, int bins, int H, int W, float* out, einx_stream stream_) {
    if (!ctx) return EINX_ERR_INVALID;
    if (B < 0 || bins <= 0 || H <= 0 || W <= 0)
        return einx_fail(ctx, EINX_ERR_INVALID, "einx_distance_map: bad shape B=%d bins=%d H=%d W=%d", B, bins, H, W);
    if (B == 0) return EINX_OK;
    if (!x || !y || !t || !ev_offsets || !out) return einx_fail(ctx, EINX_ERR_INVALID, "einx_distance_map: NULL pointer argument");
    if (B > 65535) return einx_fail(ctx, EINX_ERR_UNSUPPORTED, "einx_distance_map: B=%d > 65535", B);
    if (W > 0xfffe || H > 0xfffe) return einx_fail(ctx, EINX_ERR_UNSUPPORTED, "einx_distance_map: %dx%d beyond 16-bit row distances", H, W);
    DeviceGuard guard(ctx->device);
    cudaStream_t stream = (cudaStream_t)stream_;
    const int WW = (W + 31) / 32;
    const size_t nplanes = (size_t)B * bins, nrows = nplanes * H;
    const size_t bits_bytes = align_up(nrows * WW * sizeof(uint32_t), 256);
    const size_t rd_bytes = nrows * W * sizeof(uint16_t);
    int rc = einx_ws_reserve(ctx, bits_bytes + rd_bytes, stream);
    if (rc) return rc;
    uint32_t* bits = (uint32_t*)ctx->ws;
    uint16_t* rowdist = (uint16_t*)((unsigned char*)ctx->ws + bits_bytes);
    EINX_CUDA(ctx, cudaMemsetAsync(bits, 0, nrows * WW * sizeof(uint32_t), stream));
    int per_window = (ctx->num_sms * 8 + B - 1) / B;
    per_window = per_window < 1 ? 1 : (per_window > 1024 ? 1024 : per_window);
    dm_mark_kernel<<<dim3(per_window, B), 256, 0, stream>>>(x, y, t, ev_offsets, bins, H, W, WW, bits);
    EINX_CHECK_LAUNCH(ctx);
    dm_row_kernel<<<(unsigned)((nrows * WW + 127) / 128), 128, 0, stream>>>(bits, W, WW, nrows, rowdist);
    EINX_CHECK_LAUNCH(ctx);
    dm_column_kernel<<<(unsigned)((nplanes * H * W + 255) / 256), 256, 0, stream>>>(rowdist, H, W, nplanes, out);
    EINX_CHECK_LAUNCH(ctx);
    return EINX_OK;
}
