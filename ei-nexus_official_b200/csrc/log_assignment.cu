// LightGlue log-assignment matrix from similarities and matchability logits (SURVEY.md section 8 f, row 3).
// Semantics: reference core/modules/matchers/lightglue.py:365-377 (sigmoid_log_double_softmax; see include/einx.h).
//
// scores[b, i, j] = log_softmax_j(sim)[i, j] + log_softmax_i(sim)[i, j] + logsigmoid(z0[i]) + logsigmoid(z1[j])
// needs a (max, log-sum-exp) per row AND per column before the first output can be written, so the
// similarity matrix is read twice.  Three launches per batch chunk:
//   1. lds_stats_kernel   : 128 x 256 tiles.  A warp owns 32 columns of the tile and walks the rows in
//                           32-row blocks: the column statistics stay in registers (lane = column, online
//                           max / rescaled sum), the row statistics of each block come out of a padded
//                           shared-memory transpose (lane = row) and are merged across the CTA's 8 warps.
//                           Partials (max, sum) go to the workspace: M/128 per column, N/256 per row.
//   2. lds_merge_kernel   : folds the partials; leaves (max, log sum, logsigmoid(z), logsigmoid(-z)) per row / column.
//   3. lds_write_kernel   : second read of sim, one coalesced write of the (M+1) x (N+1) matrix incl. the
//                           unmatched row / column and the zero corner.
// The host loop issues the three per chunk of batch items whose similarities fit a fraction of L2, so the
// second read is served by L2: HBM sees sim once and scores once (8 B per element instead of 12).
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int kTileRows = 128;
constexpr int kTileCols = 256;
constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kPitch = 33;

// merge (m2, s2) into (m, s): both describe sum_k exp(x_k) as s * exp(m); an empty partial is (-inf, 0)
__device__ __forceinline__ void lse_merge(float& m, float& s, float m2, float s2) {
    const float mn = fmaxf(m, m2);
    if (mn == -INFINITY) return;  // both empty (or all -inf): keep (-inf, 0)
    s = s * expf(m - mn) + s2 * expf(m2 - mn);
    m = mn;
}

// torch: min(x, 0) - log1p(exp(-|x|))
__device__ __forceinline__ float log_sigmoid(float x) { return fminf(x, 0.0f) - log1pf(expf(-fabsf(x))); }

__global__ void __launch_bounds__(kThreads, 2)
lds_stats_kernel(const float* __restrict__ sim, int M, int N, int nrt, int nct, float2* __restrict__ rowpart,
                 float2* __restrict__ colpart) {
    const int b = blockIdx.z;
    const int i0 = blockIdx.y * kTileRows;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int j = blockIdx.x * kTileCols + warp * 32 + lane;
    const float* S = sim + (size_t)b * M * N;
    __shared__ float scratch[kWarps][32 * kPitch];
    __shared__ float2 rowred[kTileRows / 32][kWarps][32];
    float* sc = scratch[warp];
    float cm = -INFINITY, cs = 0.0f;
#pragma unroll 1
    for (int rb = 0; rb < kTileRows / 32; ++rb) {
        const int r0 = i0 + rb * 32;
        float v[32];
#pragma unroll
        for (int r = 0; r < 32; ++r) v[r] = (r0 + r < M && j < N) ? __ldg(S + (size_t)(r0 + r) * N + j) : -INFINITY;
#pragma unroll
        for (int r = 0; r < 32; ++r) sc[r * kPitch + lane] = v[r];
        // column j over these 32 rows
        float bm = v[0];
#pragma unroll
        for (int r = 1; r < 32; ++r) bm = fmaxf(bm, v[r]);
        if (bm > -INFINITY) {
            float bs = 0.0f;
#pragma unroll
            for (int r = 0; r < 32; ++r) bs += expf(v[r] - bm);
            lse_merge(cm, cs, bm, bs);
        }
        __syncwarp();
        // row r0 + lane over this warp's 32 columns
        float g[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) g[k] = sc[lane * kPitch + k];
        float rm = g[0];
#pragma unroll
        for (int k = 1; k < 32; ++k) rm = fmaxf(rm, g[k]);
        float rs = 0.0f;
        if (rm > -INFINITY) {
#pragma unroll
            for (int k = 0; k < 32; ++k) rs += expf(g[k] - rm);
        }
        rowred[rb][warp][lane] = make_float2(rm, rs);
        __syncwarp();
    }
    if (j < N) colpart[((size_t)b * nrt + blockIdx.y) * N + j] = make_float2(cm, cs);
    __syncthreads();
    if (threadIdx.x < kTileRows) {
        const int rb = threadIdx.x >> 5, l = threadIdx.x & 31;
        float m = -INFINITY, s = 0.0f;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) {
            const float2 p = rowred[rb][w][l];
            lse_merge(m, s, p.x, p.y);
        }
        const int i = i0 + threadIdx.x;
        if (i < M) rowpart[((size_t)b * nct + blockIdx.x) * M + i] = make_float2(m, s);
    }
}

// stat = (max, log(sum exp(x - max)), logsigmoid(z), logsigmoid(-z)) per row (t < M) or column (t >= M)
__global__ void __launch_bounds__(256)
lds_merge_kernel(const float2* __restrict__ rowpart, const float2* __restrict__ colpart, const float* __restrict__ z0,
                 const float* __restrict__ z1, int M, int N, int nrt, int nct, float4* __restrict__ rowstat,
                 float4* __restrict__ colstat) {
    const int b = blockIdx.y;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < M) {
        float m = -INFINITY, s = 0.0f;
        for (int c = 0; c < nct; ++c) {
            const float2 p = rowpart[((size_t)b * nct + c) * M + t];
            lse_merge(m, s, p.x, p.y);
        }
        const float z = z0[(size_t)b * M + t];
        rowstat[(size_t)b * M + t] = make_float4(m, logf(s), log_sigmoid(z), log_sigmoid(-z));
    } else if (t < M + N) {
        const int j = t - M;
        float m = -INFINITY, s = 0.0f;
        for (int r = 0; r < nrt; ++r) {
            const float2 p = colpart[((size_t)b * nrt + r) * N + j];
            lse_merge(m, s, p.x, p.y);
        }
        const float z = z1[(size_t)b * N + j];
        colstat[(size_t)b * N + j] = make_float4(m, logf(s), log_sigmoid(z), log_sigmoid(-z));
    }
}

constexpr int kWriteRows = 16;

__global__ void __launch_bounds__(256)
lds_write_kernel(const float* __restrict__ sim, const float4* __restrict__ rowstat, const float4* __restrict__ colstat, int M,
                 int N, float* __restrict__ scores) {
    const int b = blockIdx.z;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i0 = blockIdx.y * kWriteRows;
    if (j > N) return;
    const float* S = sim + (size_t)b * M * N;
    float* O = scores + (size_t)b * (M + 1) * ((size_t)N + 1);
    const float4* rs = rowstat + (size_t)b * M;
    if (j == N) {  // unmatched column: logsigmoid(-z0), corner 0
#pragma unroll 4
        for (int r = 0; r < kWriteRows; ++r) {
            const int i = i0 + r;
            if (i <= M) O[(size_t)i * (N + 1) + N] = i < M ? rs[i].w : 0.0f;
        }
        return;
    }
    const float4 c = colstat[(size_t)b * N + j];
    float v[kWriteRows];
#pragma unroll
    for (int r = 0; r < kWriteRows; ++r) v[r] = (i0 + r < M) ? __ldg(S + (size_t)(i0 + r) * N + j) : 0.0f;
#pragma unroll
    for (int r = 0; r < kWriteRows; ++r) {
        const int i = i0 + r;
        if (i < M) {
            const float4 q = rs[i];
            // (log_softmax over j) + (log_softmax over i) + (logsigmoid(z0) + logsigmoid(z1)), torch's association
            const float s0 = __fsub_rn(__fsub_rn(v[r], q.x), q.y);
            const float s1 = __fsub_rn(__fsub_rn(v[r], c.x), c.y);
            O[(size_t)i * (N + 1) + j] = __fadd_rn(__fadd_rn(s0, s1), __fadd_rn(q.z, c.z));
        } else if (i == M) {
            O[(size_t)i * (N + 1) + j] = c.w;  // unmatched row: logsigmoid(-z1)
        }
    }
}

}  // namespace

extern "C" int einx_log_double_softmax(einx_ctx* ctx, const float* sim, const float* z0, const float* z1, int B, int M, int N,
                                       float* scores, einx_stream stream_) {
    if (!ctx) return EINX_ERR_INVALID;
    if (B < 0 || M <= 0 || N <= 0)
        return einx_fail(ctx, EINX_ERR_INVALID, "einx_log_double_softmax: bad shape B=%d M=%d N=%d", B, M, N);
    if (B == 0) return EINX_OK;
    if (!sim || !z0 || !z1 || !scores) return einx_fail(ctx, EINX_ERR_INVALID, "einx_log_double_softmax: NULL pointer argument");
    DeviceGuard guard(ctx->device);
    cudaStream_t stream = (cudaStream_t)stream_;
    const int nrt = (M + kTileRows - 1) / kTileRows, nct = (N + kTileCols - 1) / kTileCols;
    const int wrt = (M + 1 + kWriteRows - 1) / kWriteRows, wct = (N + 1 + 255) / 256;
    if (nrt > 65535 || wrt > 65535) return einx_fail(ctx, EINX_ERR_UNSUPPORTED, "einx_log_double_softmax: M=%d too large", M);
    // batch items per chunk: similarities of a chunk <= 32 MB, so that the write pass re-reads them from L2
    // (EINX_LDS_CHUNK_MB overrides the 32 MB: a measurement knob, see tools/kbench.py next)
    const size_t item = sizeof(float) * (size_t)M * N;
    size_t budget = (size_t)32 << 20;
    if (const char* e = getenv("EINX_LDS_CHUNK_MB")) {
        const long mb = atol(e);
        if (mb > 0) budget = (size_t)mb << 20;
    }
    int chunk = (int)(budget / item);
    chunk = chunk < 1 ? 1 : (chunk > B ? B : chunk);
    if (chunk > 65535) chunk = 65535;
    const size_t rowpart_b = align_up(sizeof(float2) * (size_t)chunk * nct * M, 256);
    const size_t colpart_b = align_up(sizeof(float2) * (size_t)chunk * nrt * N, 256);
    const size_t rowstat_b = align_up(sizeof(float4) * (size_t)chunk * M, 256);
    const size_t colstat_b = align_up(sizeof(float4) * (size_t)chunk * N, 256);
    int rc = einx_ws_reserve(ctx, rowpart_b + colpart_b + rowstat_b + colstat_b);
    if (rc) return rc;
    char* ws = (char*)ctx->ws;
    float2* rowpart = (float2*)ws;
    float2* colpart = (float2*)(ws + rowpart_b);
    float4* rowstat = (float4*)(ws + rowpart_b + colpart_b);
    float4* colstat = (float4*)(ws + rowpart_b + colpart_b + rowstat_b);
    for (int b0 = 0; b0 < B; b0 += chunk) {
        const int nb = B - b0 < chunk ? B - b0 : chunk;
        const float* S = sim + (size_t)b0 * M * N;
        lds_stats_kernel<<<dim3(nct, nrt, nb), kThreads, 0, stream>>>(S, M, N, nrt, nct, rowpart, colpart);
        EINX_CHECK_LAUNCH(ctx);
        lds_merge_kernel<<<dim3((M + N + 255) / 256, nb), 256, 0, stream>>>(rowpart, colpart, z0 + (size_t)b0 * M, z1 + (size_t)b0 * N,
                                                                          M, N, nrt, nct, rowstat, colstat);
        EINX_CHECK_LAUNCH(ctx);
        lds_write_kernel<<<dim3(wct, wrt, nb), 256, 0, stream>>>(S, rowstat, colstat, M, N, scores + (size_t)b0 * (M + 1) * ((size_t)N + 1));
        EINX_CHECK_LAUNCH(ctx);
    }
    return EINX_OK;
}
