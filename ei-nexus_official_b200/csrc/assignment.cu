// LightGlue match filtering on a log-assignment matrix (SURVEY.md section 8 f, row 3).
// Semantics: reference core/modules/matchers/lightglue.py:402-418 (see include/einx.h).
//
// One pass over scores[:, :-1, :-1]: a CTA owns a strip of 32 rows; each warp walks 32-column blocks
// of the strip with coalesced row segments, reduces its 32x32 block along the columns in registers
// (tournament argmax, lowest row wins) and along the rows through a padded shared-memory transpose
// (lowest column wins).  Row results stay in the CTA, column results merge across strips with one
// 64-bit atomicMax of (orderable value, ~index) -- the key format and tie-breaking of the MNN matcher.
#include "common.cuh"

namespace {

constexpr int kStripThreads = 256;
constexpr int kStripWarps = kStripThreads / 32;
constexpr int kPitch = 33;

__global__ void __launch_bounds__(kStripThreads, 2)
assignment_best_kernel(const float* __restrict__ scores, int M, int N, unsigned long long* __restrict__ rowkey,
                       unsigned long long* __restrict__ colkey) {
    const int b = blockIdx.y;
    const int i0 = blockIdx.x * 32;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t ld = (size_t)N + 1;
    const float* S = scores + (size_t)b * (M + 1) * ld;
    __shared__ float scratch[kStripWarps][32 * kPitch];
    __shared__ unsigned long long rowred[32];
    if (threadIdx.x < 32) rowred[threadIdx.x] = 0ull;
    __syncthreads();
    float* sc = scratch[warp];
    float best = -INFINITY;
    int best_j = 0;
    const int nblk = (N + 31) / 32;
    for (int jb = warp; jb < nblk; jb += kStripWarps) {
        const int j = jb * 32 + lane;
        float v[32];
#pragma unroll
        for (int r = 0; r < 32; ++r) v[r] = (i0 + r < M && j < N) ? __ldg(S + (size_t)(i0 + r) * ld + j) : -INFINITY;
#pragma unroll
        for (int r = 0; r < 32; ++r) sc[r * kPitch + lane] = v[r];
        float cv;
        int cr;
        argmax32(v, cv, cr);  // column j over the strip's rows
        if (j < N && cv > -INFINITY) atomicMax(colkey + (size_t)b * N + j, pack_best(cv, (uint32_t)(i0 + cr)));
        __syncwarp();
        float g[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) g[k] = sc[lane * kPitch + k];
        float rv;
        int rk;
        argmax32(g, rv, rk);  // row i0 + lane over this block's columns
        if (rv > best) { best = rv; best_j = jb * 32 + rk; }
        __syncwarp();
    }
    if (best > -INFINITY) atomicMax(&rowred[lane], pack_best(best, (uint32_t)best_j));
    __syncthreads();
    if (threadIdx.x < 32 && i0 + threadIdx.x < M) rowkey[(size_t)b * M + i0 + threadIdx.x] = rowred[threadIdx.x];
}

__device__ __forceinline__ int key_index(unsigned long long k) { return (int)(0xffffffffu - (uint32_t)(k & 0xffffffffull)); }
__device__ __forceinline__ float key_value(unsigned long long k) { return f32_from_orderable((uint32_t)(k >> 32)); }

__global__ void __launch_bounds__(256)
assignment_filter_kernel(const unsigned long long* __restrict__ rowkey, const unsigned long long* __restrict__ colkey, int M,
                         int N, float th, int64_t* __restrict__ m0, int64_t* __restrict__ m1, float* __restrict__ ms0,
                         float* __restrict__ ms1) {
    const int b = blockIdx.y;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long* rk = rowkey + (size_t)b * M;
    const unsigned long long* ck = colkey + (size_t)b * N;
    // A key of 0 means "no candidate": the carried-keys producer skips rows / columns whose best value is -inf or
    // NaN (a diverged model), and key_index(0) would be -1.
    if (t < M) {
        const unsigned long long k = rk[t];
        const int j = k ? key_index(k) : 0;
        const unsigned long long kc = k ? ck[j] : 0ull;
        const bool mutual = k && kc && key_index(kc) == t;              // indices0 == m1.gather(1, m0)
        const float s = mutual ? expf(key_value(k)) : 0.0f;             // where(mutual0, max0.exp(), 0)
        m0[(size_t)b * M + t] = (mutual && s > th) ? j : -1;
        ms0[(size_t)b * M + t] = s;
    }
    if (t < N) {
        const unsigned long long kc = ck[t];
        const int i = kc ? key_index(kc) : 0;
        const unsigned long long k = kc ? rk[i] : 0ull;
        const bool mutual = kc && k && key_index(k) == t;               // indices1 == m0.gather(1, m1)
        const float s = mutual ? expf(key_value(k)) : 0.0f;             // mscores0.gather(1, m1) under mutual1
        m1[(size_t)b * N + t] = (mutual && s > th) ? i : -1;            // valid0.gather(1, m1)
        ms1[(size_t)b * N + t] = s;
    }
}

}  // namespace

extern "C" int einx_filter_matches(einx_ctx* ctx, const float* scores, int B, int M, int N, float th, int64_t* m0,
                                   int64_t* m1, float* ms0, float* ms1, einx_stream stream_) {
    if (!ctx) return EINX_ERR_INVALID;
    if (B < 0 || M <= 0 || N <= 0)
        return einx_fail(ctx, EINX_ERR_INVALID, "einx_filter_matches: bad shape B=%d M=%d N=%d (torch.max over an empty dimension raises)", B, M, N);
    if (B == 0) return EINX_OK;
    if (!scores || !m0 || !m1 || !ms0 || !ms1) return einx_fail(ctx, EINX_ERR_INVALID, "einx_filter_matches: NULL pointer argument");
    if (B > 65535) return einx_fail(ctx, EINX_ERR_UNSUPPORTED, "einx_filter_matches: B=%d > 65535", B);
    DeviceGuard guard(ctx->device);
    cudaStream_t stream = (cudaStream_t)stream_;
    const size_t nkeys = (size_t)B * ((size_t)M + N);
    int rc = einx_ws_reserve(ctx, sizeof(unsigned long long) * nkeys, stream);
    if (rc) return rc;
    unsigned long long* rowkey = (unsigned long long*)ctx->ws;
    unsigned long long* colkey = rowkey + (size_t)B * M;
    EINX_CUDA(ctx, cudaMemsetAsync(colkey, 0, sizeof(unsigned long long) * (size_t)B * N, stream));
    assignment_best_kernel<<<dim3((M + 31) / 32, B), kStripThreads, 0, stream>>>(scores, M, N, rowkey, colkey);
    EINX_CHECK_LAUNCH(ctx);
    const int mx = M > N ? M : N;
    assignment_filter_kernel<<<dim3((mx + 255) / 256, B), 256, 0, stream>>>(rowkey, colkey, M, N, th, m0, m1, ms0, ms1);
    EINX_CHECK_LAUNCH(ctx);
    return EINX_OK;
}

// The same filter on keys that einx_log_double_softmax left while writing the matrix (no pass over the matrix).
extern "C" int einx_filter_matches_keys(einx_ctx* ctx, const uint64_t* best_keys, int B, int M, int N, float th, int64_t* m0,
                                        int64_t* m1, float* ms0, float* ms1, einx_stream stream_) {
    if (!ctx) return EINX_ERR_INVALID;
    if (B < 0 || M <= 0 || N <= 0) return einx_fail(ctx, EINX_ERR_INVALID, "einx_filter_matches_keys: bad shape B=%d M=%d N=%d", B, M, N);
    if (B == 0) return EINX_OK;
    if (!best_keys || !m0 || !m1 || !ms0 || !ms1) return einx_fail(ctx, EINX_ERR_INVALID, "einx_filter_matches_keys: NULL pointer argument");
    if (B > 65535) return einx_fail(ctx, EINX_ERR_UNSUPPORTED, "einx_filter_matches_keys: B=%d > 65535", B);
    DeviceGuard guard(ctx->device);
    const unsigned long long* rowkey = (const unsigned long long*)best_keys;
    const unsigned long long* colkey = rowkey + (size_t)B * M;
    const int mx = M > N ? M : N;
    assignment_filter_kernel<<<dim3((mx + 255) / 256, B), 256, 0, (cudaStream_t)stream_>>>(rowkey, colkey, M, N, th, m0, m1, ms0, ms1);
    EINX_CHECK_LAUNCH(ctx);
    return EINX_OK;
}
