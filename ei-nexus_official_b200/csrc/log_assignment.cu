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
//   2. lds_merge_kernel   : folds the partials; leaves (max, log sum, logsigmoid(z)) per row / column and writes the
//                           unmatched row / column, logsigmoid(-z), and the zero corner of the matrix.
//   3. lds_write_plain_kernel / lds_write_keys_kernel : second read of sim, one coalesced write of the M x N interior
//                           of the matrix; the keys form also reduces what it writes for filter_matches.
// HBM sees sim twice and scores once: 12 B per element (walking the batch in L2-sized chunks so that the second
// read hits L2 was measured and lost to the smaller launches; see the entry point).
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

// exp of a non-positive difference inside the sums: ex2.approx(d * log2 e), 2 + |1.17 d| ulp -- the terms that carry
// the sum have d near 0 (2 ulp), the ones with a large |d| weigh exp(d); expf() here made the pass issue bound
__device__ __forceinline__ float exp_term(float d) { return __expf(d); }

template <bool FULL>
__device__ __forceinline__ void lds_stats_tile(const float* __restrict__ S, int M, int N, int i0, int j, int warp, int lane,
                                               float* sc, float2 (*rowred)[kWarps][32], float& cm, float& cs) {
#pragma unroll 1
    for (int rb = 0; rb < kTileRows / 32; ++rb) {
        const int r0 = i0 + rb * 32;
        const float* p = S + (size_t)r0 * N + j;
        float v[32];
#pragma unroll
        for (int r = 0; r < 32; ++r) v[r] = (FULL || (r0 + r < M && j < N)) ? __ldg(p + (size_t)r * N) : -INFINITY;
#pragma unroll
        for (int r = 0; r < 32; ++r) sc[r * kPitch + lane] = v[r];
        // column j over these 32 rows
        float bm = v[0];
#pragma unroll
        for (int r = 1; r < 32; ++r) bm = fmaxf(bm, v[r]);
        if (FULL || bm > -INFINITY) {
            float bs = 0.0f;
#pragma unroll
            for (int r = 0; r < 32; ++r) bs += exp_term(v[r] - bm);
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
        if (FULL || rm > -INFINITY) {
#pragma unroll
            for (int k = 0; k < 32; ++k) rs += exp_term(g[k] - rm);
        }
        rowred[rb][warp][lane] = make_float2(rm, rs);
        __syncwarp();
    }
}

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
    float cm = -INFINITY, cs = 0.0f;
    if (i0 + kTileRows <= M && (blockIdx.x + 1) * kTileCols <= N)  // CTA-uniform: interior tiles skip the bounds tests
        lds_stats_tile<true>(S, M, N, i0, j, warp, lane, scratch[warp], rowred, cm, cs);
    else
        lds_stats_tile<false>(S, M, N, i0, j, warp, lane, scratch[warp], rowred, cm, cs);
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

// Folds the partials: stat = (max, log(sum exp(x - max)), logsigmoid(z)) per row (t < M) or column (M <= t < M + N);
// also writes the unmatched column / row of the matrix, logsigmoid(-z), and the zero corner (t == M + N).
__global__ void __launch_bounds__(256)
lds_merge_kernel(const float2* __restrict__ rowpart, const float2* __restrict__ colpart, const float* __restrict__ z0,
                 const float* __restrict__ z1, int M, int N, int nrt, int nct, float4* __restrict__ rowstat,
                 float4* __restrict__ colstat, float* __restrict__ scores) {
    const int b = blockIdx.y;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    float* O = scores + (size_t)b * (M + 1) * ((size_t)N + 1);
    if (t < M) {
        float m = -INFINITY, s = 0.0f;
        for (int c = 0; c < nct; ++c) {
            const float2 p = rowpart[((size_t)b * nct + c) * M + t];
            lse_merge(m, s, p.x, p.y);
        }
        const float z = z0[(size_t)b * M + t];
        rowstat[(size_t)b * M + t] = make_float4(m, logf(s), log_sigmoid(z), 0.0f);
        O[(size_t)t * (N + 1) + N] = log_sigmoid(-z);
    } else if (t < M + N) {
        const int j = t - M;
        float m = -INFINITY, s = 0.0f;
        for (int r = 0; r < nrt; ++r) {
            const float2 p = colpart[((size_t)b * nrt + r) * N + j];
            lse_merge(m, s, p.x, p.y);
        }
        const float z = z1[(size_t)b * N + j];
        colstat[(size_t)b * N + j] = make_float4(m, logf(s), log_sigmoid(z), 0.0f);
        O[(size_t)M * (N + 1) + j] = log_sigmoid(-z);
    } else if (t == M + N) {
        O[(size_t)M * (N + 1) + N] = 0.0f;
    }
}

// Rows in flight per thread of the write pass.  Measured on B200, C2 shape, whole entry point: 4 rows 0.180 ms,
// 8 rows 0.170 ms, 16 rows 0.187 ms, 32 rows 0.204 ms; streaming (.cs) loads / stores change nothing.
constexpr int kWriteRows = 8;
// Rows per CTA when the pass also reduces the row / column maxima (KEYS): taller tiles mean fewer 64-bit atomics per
// column (M / tile); measured flat on B200 (C2 shape, entry point: 16 and 32 rows 0.219 ms, 64 rows 0.224, 128 rows
// 0.232, 256 rows 0.244).
constexpr int kKeyTileRows = 32;

// Interior of the matrix: thread = column, kWriteRows rows in flight.  KEYS: the values being written are also reduced
// to (max, first index) per row and per column -- exactly what filter_matches (lightglue.py:402-418) would recompute
// from the stored matrix -- as 64-bit (orderable value, ~index) keys, the format of the MNN matcher.
template <bool KEYS, bool FULL>
__device__ __forceinline__ void lds_write_rows(const float* __restrict__ p, const float4* __restrict__ rs, const float4 c, int rows,
                                               int N, float* __restrict__ o, bool valid, int j, int row0, int ib,
                                               unsigned long long* rowred, float& cb, int& ci) {
    float v[kWriteRows];
#pragma unroll
    for (int r = 0; r < kWriteRows; ++r) v[r] = (FULL || r < rows) ? __ldg(p + (size_t)r * N) : 0.0f;
#pragma unroll
    for (int r = 0; r < kWriteRows; ++r) {
        if (FULL || r < rows) {  // CTA-uniform
            const float4 q = rs[r];
            // (log_softmax over j) + (log_softmax over i) + (logsigmoid(z0) + logsigmoid(z1)), torch's association
            const float s0 = __fsub_rn(__fsub_rn(v[r], q.x), q.y);
            const float s1 = __fsub_rn(__fsub_rn(v[r], c.x), c.y);
            const float out = __fadd_rn(__fadd_rn(s0, s1), __fadd_rn(q.z, c.z));
            if (!KEYS || valid) o[(size_t)r * (N + 1)] = out;
            if (KEYS) {
                const uint32_t ord = valid ? f32_orderable(out + 0.0f) : 0u;
                const uint32_t mx = __reduce_max_sync(0xffffffffu, ord);
                const uint32_t eq = __ballot_sync(0xffffffffu, valid && ord == mx);
                // (max, first column holding it) of this warp's 32 columns, merged over the CTA's warps in shared memory
                // (keeping the keys in registers until the end of the tile was measured slower: 0.247 vs 0.222 ms)
                if ((threadIdx.x & 31) == 0 && eq)
                    atomicMax(&rowred[row0 + r],
                              ((unsigned long long)mx << 32) | (unsigned long long)(0xffffffffu - (uint32_t)(j + __ffs(eq) - 1)));
                if (out > cb) { cb = out; ci = ib + r; }  // rows ascend: strict > keeps the first
            }
        }
    }
}

// interior of the matrix without the reductions: thread = column, one group of kWriteRows rows per CTA
__global__ void __launch_bounds__(256)
lds_write_plain_kernel(const float* __restrict__ sim, const float4* __restrict__ rowstat, const float4* __restrict__ colstat, int M,
                       int N, float* __restrict__ scores) {
    const int b = blockIdx.z;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i0 = blockIdx.y * kWriteRows;
    if (j >= N) return;
    const float* p = sim + ((size_t)b * M + i0) * N + j;
    float* o = scores + ((size_t)b * (M + 1) + i0) * ((size_t)N + 1) + j;
    const float4* rs = rowstat + (size_t)b * M + i0;
    const float4 c = colstat[(size_t)b * N + j];
    float cb = 0.0f;
    int ci = 0;
    if (i0 + kWriteRows <= M)
        lds_write_rows<false, true>(p, rs, c, kWriteRows, N, o, true, j, 0, i0, nullptr, cb, ci);
    else
        lds_write_rows<false, false>(p, rs, c, M - i0, N, o, true, j, 0, i0, nullptr, cb, ci);
}

// interior of the matrix with the reductions: thread = column, kKeyTileRows rows per CTA in groups of kWriteRows
__global__ void __launch_bounds__(256)
lds_write_keys_kernel(const float* __restrict__ sim, const float4* __restrict__ rowstat, const float4* __restrict__ colstat, int M,
                      int N, float* __restrict__ scores, unsigned long long* __restrict__ rowkey,
                      unsigned long long* __restrict__ colkey) {
    const int b = blockIdx.z;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i0 = blockIdx.y * kKeyTileRows;
    const bool valid = j < N;  // invalid lanes stay for the warp collectives and the barriers
    __shared__ unsigned long long rowred[kKeyTileRows];
    if (threadIdx.x < kKeyTileRows) rowred[threadIdx.x] = 0ull;
    __syncthreads();
    const int jc = valid ? j : N - 1;  // invalid lanes shadow the last column (never stored, never reduced)
    const float4 c = colstat[(size_t)b * N + jc];
    float cb = -INFINITY;
    int ci = 0;
#pragma unroll 1
    for (int ib = i0; ib < min(i0 + kKeyTileRows, M); ib += kWriteRows) {
        const float* p = sim + ((size_t)b * M + ib) * N + jc;
        float* o = scores + ((size_t)b * (M + 1) + ib) * ((size_t)N + 1) + jc;
        const float4* rs = rowstat + (size_t)b * M + ib;
        if (ib + kWriteRows <= M)
            lds_write_rows<true, true>(p, rs, c, kWriteRows, N, o, valid, j, ib - i0, ib, rowred, cb, ci);
        else
            lds_write_rows<true, false>(p, rs, c, M - ib, N, o, valid, j, ib - i0, ib, rowred, cb, ci);
    }
    if (valid && cb > -INFINITY) atomicMax(colkey + (size_t)b * N + j, pack_best(cb, (uint32_t)ci));
    __syncthreads();
    if (threadIdx.x < kKeyTileRows && i0 + threadIdx.x < M && rowred[threadIdx.x])
        atomicMax(rowkey + (size_t)b * M + i0 + threadIdx.x, rowred[threadIdx.x]);
}

}  // namespace

extern "C" int einx_log_double_softmax(einx_ctx* ctx, const float* sim, const float* z0, const float* z1, int B, int M, int N,
                                       float* scores, uint64_t* best_keys, einx_stream stream_) {
    if (!ctx) return EINX_ERR_INVALID;
    if (B < 0 || M <= 0 || N <= 0)
        return einx_fail(ctx, EINX_ERR_INVALID, "einx_log_double_softmax: bad shape B=%d M=%d N=%d", B, M, N);
    if (B == 0) return EINX_OK;
    if (!sim || !z0 || !z1 || !scores) return einx_fail(ctx, EINX_ERR_INVALID, "einx_log_double_softmax: NULL pointer argument");
    DeviceGuard guard(ctx->device);
    cudaStream_t stream = (cudaStream_t)stream_;
    const int nrt = (M + kTileRows - 1) / kTileRows, nct = (N + kTileCols - 1) / kTileCols;
    const int wtile = best_keys ? kKeyTileRows : kWriteRows;
    const int wrt = (M + wtile - 1) / wtile, wct = (N + 255) / 256;
    unsigned long long* rowkey = (unsigned long long*)best_keys;
    unsigned long long* colkey = rowkey ? rowkey + (size_t)B * M : nullptr;
    if (best_keys) EINX_CUDA(ctx, cudaMemsetAsync(best_keys, 0, sizeof(uint64_t) * (size_t)B * ((size_t)M + N), stream));
    if (nrt > 65535 || wrt > 65535) return einx_fail(ctx, EINX_ERR_UNSUPPORTED, "einx_log_double_softmax: M=%d too large", M);
    // One chunk = the whole batch by default.  EINX_LDS_CHUNK_MB (a measurement knob, tools/kbench.py next) walks the
    // batch in chunks of that many MB of similarities so that the write pass re-reads them from L2; measured on B200
    // (C2 shape, 64 x 1024 x 1024) the smaller launches cost more than the saved HBM read: 0.34 ms at 32 MB, 0.28 ms at
    // 64 MB against 0.22 ms for the whole batch (0.26 / 0.23 against 0.19 ms with the
    // final kernels).
    const size_t item = sizeof(float) * (size_t)M * N;
    int chunk = B;
    if (const char* e = getenv("EINX_LDS_CHUNK_MB")) {
        const long mb = atol(e);
        if (mb > 0) chunk = (int)(((size_t)mb << 20) / item);
    }
    chunk = chunk < 1 ? 1 : (chunk > B ? B : chunk);
    if (chunk > 65535) chunk = 65535;
    const size_t rowpart_b = align_up(sizeof(float2) * (size_t)chunk * nct * M, 256);
    const size_t colpart_b = align_up(sizeof(float2) * (size_t)chunk * nrt * N, 256);
    const size_t rowstat_b = align_up(sizeof(float4) * (size_t)chunk * M, 256);
    const size_t colstat_b = align_up(sizeof(float4) * (size_t)chunk * N, 256);
    int rc = einx_ws_reserve(ctx, rowpart_b + colpart_b + rowstat_b + colstat_b, stream);
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
        float* O = scores + (size_t)b0 * (M + 1) * ((size_t)N + 1);
        lds_merge_kernel<<<dim3((M + N + 1 + 255) / 256, nb), 256, 0, stream>>>(rowpart, colpart, z0 + (size_t)b0 * M, z1 + (size_t)b0 * N,
                                                                              M, N, nrt, nct, rowstat, colstat, O);
        EINX_CHECK_LAUNCH(ctx);
        if (best_keys)
            lds_write_keys_kernel<<<dim3(wct, wrt, nb), 256, 0, stream>>>(S, rowstat, colstat, M, N, O, rowkey + (size_t)b0 * M,
                                                                         colkey + (size_t)b0 * N);
        else
            lds_write_plain_kernel<<<dim3(wct, wrt, nb), 256, 0, stream>>>(S, rowstat, colstat, M, N, O);
        EINX_CHECK_LAUNCH(ctx);
    }
    return EINX_OK;
}
