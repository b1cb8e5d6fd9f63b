// Metric-side N x M reductions (SURVEY.md section 8 f, row 4): a pairwise distance matrix between two keypoint
// sets that is only ever reduced along both axes -- the same row / column "best of" shape as the MNN matcher,
// with two coordinates instead of a descriptor.  The matrix is never materialised.
//
//   einx_pairwise_min_dist  core/metrics/keypoints_metrics.py:110-124 (Repeatability.update_one): Euclidean norm of
//                           the (N, M, 2) difference tensor, minimum along both axes.
//   einx_gt_assign          core/geometry/gt_generation.py:96-126 (gt_matches_from_pose_depth): the two reprojection
//                           distance matrices, their maximum masked by visibility, argmin along both axes, the mutual
//                           check against pos_th and the negatives against neg_th.
//
// One thread owns one point of a side and walks the other side through shared memory (256 points per refill);
// blockIdx.z selects which side owns.  First index wins ties (strict '<' while walking upwards), as torch.min.
#include <math.h>

#include "common.cuh"

namespace {

constexpr int kOwners = 128;   // threads per CTA = points owned
constexpr int kChunk = 256;    // points of the other side staged per refill

// ---- Repeatability ---------------------------------------------------------------------------- //
__global__ void __launch_bounds__(kOwners)
pairwise_min_dist_kernel(const float* __restrict__ a, const float* __restrict__ b, const int32_t* __restrict__ na,
                         const int32_t* __restrict__ nb, int N, int M, float* __restrict__ rowmin, float* __restrict__ colmin) {
    __shared__ float2 other[kChunk];
    const int bi = blockIdx.y, side = blockIdx.z;
    const int n = na ? min(na[bi], N) : N, m = nb ? min(nb[bi], M) : M;
    const float2* own_p = reinterpret_cast<const float2*>(side == 0 ? a + (size_t)bi * N * 2 : b + (size_t)bi * M * 2);
    const float2* oth_p = reinterpret_cast<const float2*>(side == 0 ? b + (size_t)bi * M * 2 : a + (size_t)bi * N * 2);
    const int n_own = side == 0 ? n : m, n_oth = side == 0 ? m : n;
    const int cap_own = side == 0 ? N : M;
    if (blockIdx.x * kOwners >= cap_own) return;
    const int i = blockIdx.x * kOwners + threadIdx.x;
    const bool live = i < n_own;
    const float2 me = live ? own_p[i] : make_float2(0.f, 0.f);
    double best = INFINITY;
    for (int j0 = 0; j0 < n_oth; j0 += kChunk) {
        __syncthreads();
        for (int k = threadIdx.x; k < kChunk && j0 + k < n_oth; k += kOwners) other[k] = oth_p[j0 + k];
        __syncthreads();
        const int lim = min(kChunk, n_oth - j0);
        for (int k = 0; k < lim; ++k) {
            // the difference is taken in fp32 (side 0 minus side 1: the sign is squared away), squares and sum in fp64
            const double dx = (double)__fsub_rn(me.x, other[k].x), dy = (double)__fsub_rn(me.y, other[k].y);
            const double v = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
            best = (v < best || v != v) ? v : best;   // NaN sticks, like torch.min
        }
    }
    float* out = side == 0 ? rowmin + (size_t)bi * N : colmin + (size_t)bi * M;
    if (i < cap_own) out[i] = live ? (float)sqrt(best) : INFINITY;
}

// ---- ground-truth assignment ------------------------------------------------------------------ //
struct GtParams {
    const float *kp0, *kp1, *kp0_1, *kp1_0;         // (B, N, 2), (B, M, 2), (B, N, 2), (B, M, 2)
    const uint8_t *visible0, *visible1, *valid0, *valid1;
    int N, M;
    float pos2, neg2;                                // thresholds squared (pos_th**2, neg_th**2)
    int32_t *min0, *min1;                            // (B, N), (B, M): argmin of the masked distance
    float *dmin0, *dmin1;                            // masked distance at the argmin
    uint8_t *neg0, *neg1;                            // negative0 / negative1 of :117-118
};

__device__ __forceinline__ float sqdist(float2 p, float2 q) {
    const float dx = __fsub_rn(p.x, q.x), dy = __fsub_rn(p.y, q.y);
    return __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));   // torch.sum((p - q) ** 2, -1) over two elements
}

__global__ void __launch_bounds__(kOwners) gt_reduce_kernel(const GtParams P) {
    __shared__ float2 oth_pt[kChunk], oth_proj[kChunk];
    __shared__ uint8_t oth_vis[kChunk];
    const int bi = blockIdx.y, side = blockIdx.z;
    const int N = P.N, M = P.M;
    const int n_own = side == 0 ? N : M, n_oth = side == 0 ? M : N;
    if (blockIdx.x * kOwners >= n_own) return;
    const int i = blockIdx.x * kOwners + threadIdx.x;
    const bool live = i < n_own;
    const float2* kp0 = reinterpret_cast<const float2*>(P.kp0) + (size_t)bi * N;
    const float2* kp1 = reinterpret_cast<const float2*>(P.kp1) + (size_t)bi * M;
    const float2* kp0_1 = reinterpret_cast<const float2*>(P.kp0_1) + (size_t)bi * N;
    const float2* kp1_0 = reinterpret_cast<const float2*>(P.kp1_0) + (size_t)bi * M;
    // side 0 owns i: dist0 = |kp0_1[i] - kp1[j]|^2, dist1 = |kp0[i] - kp1_0[j]|^2; side 1 owns j, same two terms
    float2 me_pt = make_float2(0.f, 0.f), me_proj = me_pt;
    bool me_vis = false;
    if (live) {
        me_pt = side == 0 ? kp0[i] : kp1[i];
        me_proj = side == 0 ? kp0_1[i] : kp1_0[i];
        me_vis = (side == 0 ? P.visible0[(size_t)bi * N + i] : P.visible1[(size_t)bi * M + i]) != 0;
    }
    float best = INFINITY, own_term_min = INFINITY;  // masked max(dist0, dist1); unmasked dist0 (side 0) / dist1 (side 1)
    int best_k = 0;
    for (int j0 = 0; j0 < n_oth; j0 += kChunk) {
        __syncthreads();
        for (int k = threadIdx.x; k < kChunk && j0 + k < n_oth; k += kOwners) {
            const int j = j0 + k;
            oth_pt[k] = side == 0 ? kp1[j] : kp0[j];
            oth_proj[k] = side == 0 ? kp1_0[j] : kp0_1[j];
            oth_vis[k] = side == 0 ? P.visible1[(size_t)bi * M + j] : P.visible0[(size_t)bi * N + j];
        }
        __syncthreads();
        const int lim = min(kChunk, n_oth - j0);
        for (int k = 0; k < lim; ++k) {
            // my projection against the other's point, my point against the other's projection
            const float d_mine = sqdist(me_proj, oth_pt[k]);    // side 0: dist0[i, j]; side 1: dist1[i, j]
            const float d_other = sqdist(me_pt, oth_proj[k]);   // side 0: dist1[i, j]; side 1: dist0[i, j]
            own_term_min = (d_mine < own_term_min || d_mine != d_mine) ? d_mine : own_term_min;
            float d = fmaxf(d_mine, d_other);
            if (d_mine != d_mine || d_other != d_other) d = NAN;  // torch.max propagates NaN
            d = (me_vis && oth_vis[k]) ? d : INFINITY;            // torch.where(mask_visible, dist, inf)
            if (d < best || (d != d && best == best)) { best = d; best_k = j0 + k; }
        }
    }
    if (!live) return;
    const size_t o = (size_t)bi * n_own + i;
    const bool valid = (side == 0 ? P.valid0[o] : P.valid1[o]) != 0;
    (side == 0 ? P.min0 : P.min1)[o] = best_k;
    (side == 0 ? P.dmin0 : P.dmin1)[o] = best;
    (side == 0 ? P.neg0 : P.neg1)[o] = (own_term_min > (side == 0 ? P.neg2 : P.neg2)) && valid;  // NaN > x is false
}

__global__ void gt_finalize_kernel(const GtParams P, int B, int64_t* __restrict__ m0, int64_t* __restrict__ m1) {
    const int N = P.N, M = P.M;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t tot0 = (size_t)B * N, tot1 = (size_t)B * M;
    if (idx < tot0) {
        const int bi = (int)(idx / N), i = (int)(idx - (size_t)bi * N);
        const int j = P.min0[idx];
        // positive.any(-1): j = min0[i] is the only candidate; mutual (min1[j] == i) and closer than pos_th
        const bool pos = M > 0 && P.min1[(size_t)bi * M + j] == i && P.dmin0[idx] < P.pos2;
        m0[idx] = P.neg0[idx] ? -1 : (pos ? (int64_t)j : -2);
    }
    if (idx < tot1) {
        const int bi = (int)(idx / M), j = (int)(idx - (size_t)bi * M);
        const int i = P.min1[idx];
        const bool pos = N > 0 && P.min0[(size_t)bi * N + i] == j && P.dmin1[idx] < P.pos2;
        m1[idx] = P.neg1[idx] ? -1 : (pos ? (int64_t)i : -2);
    }
}

}  // namespace

extern "C" int einx_pairwise_min_dist(einx_ctx* ctx, const float* a, const float* b, const int32_t* na, const int32_t* nb,
                                      int B, int N, int M, float* rowmin, float* colmin, einx_stream stream_) {
    if (!ctx) return EINX_ERR_INVALID;
    if (B < 0 || N < 0 || M < 0) return einx_fail(ctx, EINX_ERR_INVALID, "einx_pairwise_min_dist: bad shape B=%d N=%d M=%d", B, N, M);
    if (B == 0 || (N == 0 && M == 0)) return EINX_OK;
    if ((N > 0 && (!a || !rowmin)) || (M > 0 && (!b || !colmin)))
        return einx_fail(ctx, EINX_ERR_INVALID, "einx_pairwise_min_dist: NULL pointer argument");
    if (B > 65535) return einx_fail(ctx, EINX_ERR_UNSUPPORTED, "einx_pairwise_min_dist: B=%d > 65535", B);
    DeviceGuard guard(ctx->device);
    const int big = N > M ? N : M;
    pairwise_min_dist_kernel<<<dim3((big + kOwners - 1) / kOwners, B, 2), kOwners, 0, (cudaStream_t)stream_>>>(a, b, na, nb, N, M, rowmin, colmin);
    EINX_CHECK_LAUNCH(ctx);
    return EINX_OK;
}

extern "C" int einx_gt_assign(einx_ctx* ctx, const float* kp0, const float* kp1, const float* kp0_1, const float* kp1_0,
                              const uint8_t* visible0, const uint8_t* visible1, const uint8_t* valid0, const uint8_t* valid1,
                              int B, int N, int M, float pos_th, float neg_th, int64_t* m0, int64_t* m1, int32_t* min0,
                              int32_t* min1, einx_stream stream_) {
    if (!ctx) return EINX_ERR_INVALID;
    if (B < 0 || N < 0 || M < 0) return einx_fail(ctx, EINX_ERR_INVALID, "einx_gt_assign: bad shape B=%d N=%d M=%d", B, N, M);
    if (B == 0 || (N == 0 && M == 0)) return EINX_OK;
    if (N == 0 || M == 0) return einx_fail(ctx, EINX_ERR_INVALID, "einx_gt_assign: an empty side is the caller's early return (gt_generation.py:63-71)");
    if (!kp0 || !kp1 || !kp0_1 || !kp1_0 || !visible0 || !visible1 || !valid0 || !valid1 || !m0 || !m1)
        return einx_fail(ctx, EINX_ERR_INVALID, "einx_gt_assign: NULL pointer argument");
    if (B > 65535) return einx_fail(ctx, EINX_ERR_UNSUPPORTED, "einx_gt_assign: B=%d > 65535", B);
    DeviceGuard guard(ctx->device);
    cudaStream_t stream = (cudaStream_t)stream_;
    const size_t t0 = (size_t)B * N, t1 = (size_t)B * M;
    // scratch: argmins (unless the caller wants them), distances at the argmin, negative flags
    const size_t bytes = align_up((t0 + t1) * 4, 256) * 2 + align_up(t0 + t1, 256);
    int rc = einx_ws_reserve(ctx, bytes, stream);
    if (rc) return rc;
    unsigned char* ws = (unsigned char*)ctx->ws;
    GtParams P = {};
    P.kp0 = kp0; P.kp1 = kp1; P.kp0_1 = kp0_1; P.kp1_0 = kp1_0;
    P.visible0 = visible0; P.visible1 = visible1; P.valid0 = valid0; P.valid1 = valid1;
    P.N = N; P.M = M;
    P.pos2 = __builtin_powif(pos_th, 2); P.neg2 = __builtin_powif(neg_th, 2);
    P.min0 = min0 ? min0 : (int32_t*)ws;
    P.min1 = min1 ? min1 : (int32_t*)ws + t0;
    P.dmin0 = (float*)(ws + align_up((t0 + t1) * 4, 256));
    P.dmin1 = P.dmin0 + t0;
    P.neg0 = ws + 2 * align_up((t0 + t1) * 4, 256);
    P.neg1 = P.neg0 + t0;
    const int big = N > M ? N : M;
    gt_reduce_kernel<<<dim3((big + kOwners - 1) / kOwners, B, 2), kOwners, 0, stream>>>(P);
    EINX_CHECK_LAUNCH(ctx);
    const size_t tot = t0 > t1 ? t0 : t1;
    gt_finalize_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, stream>>>(P, B, m0, m1);
    EINX_CHECK_LAUNCH(ctx);
    return EINX_OK;
}
