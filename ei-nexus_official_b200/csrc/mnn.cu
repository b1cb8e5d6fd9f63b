// Mutual-nearest-neighbour matching, fp32-exact path + the shared finalisation kernels.
// Semantics: reference core/modules/matchers/MNN.py:11-32, :88-129 (see include/einx.h).
//
// The similarity matrix is never materialised: each 128x128 tile of d0 . d1^T is reduced in the
// epilogue to per-row / per-column (value, lowest index) keys that are merged across tiles with one
// 64-bit atomicMax per row/column.  A tiny second kernel applies the thresholds and the mutual
// check straight from the keys.  The tensor-core variants (mnn_tc.cu) reuse the same key format and
// finalisation, so every precision mode has identical tie-breaking.
#include "common.cuh"
#include "mnn_keys.cuh"

namespace {

constexpr int BM = 128, BN = 128, BK = 8;
constexpr int TM = 8, TN = 8;
constexpr int kGemmThreads = 256;

// MODE 0: best keys.  MODE 1: best keys excluding the already-known best (second best, for the
// ratio test).  MODE 2: store the similarity tile (opt-in dense output).
template <int MODE>
__global__ void __launch_bounds__(kGemmThreads)
mnn_fp32_kernel(const float* __restrict__ d0, const float* __restrict__ d1, const int32_t* __restrict__ n0,
                const int32_t* __restrict__ n1, int ncap, int mcap, int D, unsigned long long* __restrict__ rowkey,
                unsigned long long* __restrict__ colkey, const unsigned long long* __restrict__ rowbest,
                const unsigned long long* __restrict__ colbest, float* __restrict__ sim_out) {
    const int b = blockIdx.z;
    const int N = n0 ? min(n0[b], ncap) : ncap;
    const int M = n1 ? min(n1[b], mcap) : mcap;
    const int i0 = blockIdx.y * BM, j0 = blockIdx.x * BN;
    if (i0 >= N || j0 >= M) return;
    const float* A = d0 + (size_t)b * ncap * D;
    const float* Bm = d1 + (size_t)b * mcap * D;

    __shared__ float As[2][BK][BM + 4];
    __shared__ float Bs[2][BK][BN + 4];
    __shared__ unsigned long long colred[kGemmThreads / 32][BN];

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads, each an 8x8 micro-tile (4+4 split)
    // global -> shared staging: 128 rows x 8 k per tile = 1024 floats = 256 threads x 4 (one float4)
    const int lrow = tid >> 1, lk = (tid & 1) * 4;
    const bool vec = (D & 3) == 0;

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.0f;

    auto load_tile = [&](const float* base, int row0, int rows, int k0, float out[4]) {
        const int r = row0 + lrow;
        out[0] = out[1] = out[2] = out[3] = 0.0f;
        if (r < rows) {
            const float* src = base + (size_t)r * D + k0 + lk;
            if (vec && k0 + lk + 3 < D) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(src));
                out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w;
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (k0 + lk + q < D) out[q] = __ldg(src + q);
            }
        }
    };

    float ra[4], rb[4];
    load_tile(A, i0, N, 0, ra);
    load_tile(Bm, j0, M, 0, rb);
#pragma unroll
    for (int q = 0; q < 4; ++q) { As[0][lk + q][lrow] = ra[q]; Bs[0][lk + q][lrow] = rb[q]; }
    __syncthreads();
    const int nk = (D + BK - 1) / BK;
    for (int kt = 0; kt < nk; ++kt) {
        const int cur = kt & 1;
        if (kt + 1 < nk) {
            load_tile(A, i0, N, (kt + 1) * BK, ra);
            load_tile(Bm, j0, M, (kt + 1) * BK, rb);
        }
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[TM], bb[TN];
            const float4 a0 = *reinterpret_cast<const float4*>(&As[cur][k][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[cur][k][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[cur][k][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[cur][k][64 + tx * 4]);
            a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
            bb[0] = b0.x; bb[1] = b0.y; bb[2] = b0.z; bb[3] = b0.w; bb[4] = b1.x; bb[5] = b1.y; bb[6] = b1.z; bb[7] = b1.w;
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
        }
        if (kt + 1 < nk) {
#pragma unroll
            for (int q = 0; q < 4; ++q) { As[cur ^ 1][lk + q][lrow] = ra[q]; Bs[cur ^ 1][lk + q][lrow] = rb[q]; }
        }
        __syncthreads();
    }

    // micro-tile coordinates: rows i0 + {ty*4..+3, 64+ty*4..+3}, cols j0 + {tx*4..+3, 64+tx*4..+3}
    int gi[TM], gj[TN];
#pragma unroll
    for (int i = 0; i < TM; ++i) gi[i] = i0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
#pragma unroll
    for (int j = 0; j < TN; ++j) gj[j] = j0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));

    if (MODE == 2) {
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
            for (int j = 0; j < TN; ++j)
                if (gi[i] < N && gj[j] < M) sim_out[((size_t)b * N + gi[i]) * M + gj[j]] = acc[i][j];
        return;
    }

    unsigned long long* rk = rowkey + (size_t)b * ncap;
    unsigned long long* ck = colkey + (size_t)b * mcap;
    unsigned int excl_col[TM], excl_row[TN];  // MODE 1: index of the known best to leave out
    if (MODE == 1) {
#pragma unroll
        for (int i = 0; i < TM; ++i) excl_col[i] = gi[i] < N ? key_index(rowbest[(size_t)b * ncap + gi[i]]) : 0xffffffffu;
#pragma unroll
        for (int j = 0; j < TN; ++j) excl_row[j] = gj[j] < M ? key_index(colbest[(size_t)b * mcap + gj[j]]) : 0xffffffffu;
    }
    unsigned long long rbest[TM], cbest[TN];
#pragma unroll
    for (int i = 0; i < TM; ++i) rbest[i] = 0ull;
#pragma unroll
    for (int j = 0; j < TN; ++j) cbest[j] = 0ull;
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            if (gi[i] < N && gj[j] < M) {
                const unsigned int ob = f32_orderable(acc[i][j] + 0.0f);
                if (MODE == 0 || excl_col[i] != (unsigned)gj[j])
                    rbest[i] = max(rbest[i], ((unsigned long long)ob << 32) | (0xffffffffu - (unsigned)gj[j]));
                if (MODE == 0 || excl_row[j] != (unsigned)gi[i])
                    cbest[j] = max(cbest[j], ((unsigned long long)ob << 32) | (0xffffffffu - (unsigned)gi[i]));
            }
        }
    // rows: the 16 threads sharing a row are the 16 lanes of a half-warp
#pragma unroll
    for (int i = 0; i < TM; ++i) {
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) rbest[i] = max(rbest[i], __shfl_xor_sync(0xffffffffu, rbest[i], o));
        if (tx == 0 && rbest[i]) atomicMax(rk + gi[i], rbest[i]);
    }
    // columns: combine the two ty of a warp, then the 8 warps through shared memory
    const int warp = tid >> 5;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
        cbest[j] = max(cbest[j], __shfl_xor_sync(0xffffffffu, cbest[j], 16));
        if ((tid & 16) == 0) colred[warp][gj[j] - j0] = cbest[j];
    }
    __syncthreads();
    if (tid < BN) {
        unsigned long long m = colred[0][tid];
#pragma unroll
        for (int w = 1; w < kGemmThreads / 32; ++w) m = max(m, colred[w][tid]);
        if (m && j0 + tid < M) atomicMax(ck + j0 + tid, m);
    }
}

struct FinalizeParams {
    const unsigned long long *rowkey, *colkey, *row2, *col2;
    const int32_t *n0, *n1;
    int ncap, mcap;
    float ratio_sq, dist_sq;
    int use_ratio, use_dist, mutual;
    int64_t *m0, *m1;
    float *s0, *s1;
};

// find_nn of MNN.py:11-22 for one query, straight from the packed keys
__device__ __forceinline__ int nn_from_keys(const unsigned long long* best, const unsigned long long* second,
                                            int idx, const FinalizeParams& P) {
    const unsigned long long k = best[idx];
    if (k == 0ull) return -1;
    const float sim = key_value(k);
    const float dist = __fmul_rn(2.0f, __fsub_rn(1.0f, sim));  // dist_nn = 2 * (1 - sim_nn)
    bool ok = true;
    if (P.use_ratio) {
        const unsigned long long k2 = second[idx];
        const float dist2 = __fmul_rn(2.0f, __fsub_rn(1.0f, key_value(k2)));
        ok = ok && (k2 != 0ull) && (dist <= __fmul_rn(P.ratio_sq, dist2));
    }
    if (P.use_dist) ok = ok && (dist <= P.dist_sq);
    return ok ? (int)key_index(k) : -1;
}

__global__ void __launch_bounds__(256) mnn_finalize_kernel(const FinalizeParams P) {
    const int b = blockIdx.y;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int N = P.n0 ? min(P.n0[b], P.ncap) : P.ncap;
    const int M = P.n1 ? min(P.n1[b], P.mcap) : P.mcap;
    const unsigned long long* rk = P.rowkey + (size_t)b * P.ncap;
    const unsigned long long* ck = P.colkey + (size_t)b * P.mcap;
    const unsigned long long* r2 = P.row2 ? P.row2 + (size_t)b * P.ncap : nullptr;
    const unsigned long long* c2 = P.col2 ? P.col2 + (size_t)b * P.mcap : nullptr;
    if (t < P.ncap) {
        int m = -1;
        if (t < N) {
            m = nn_from_keys(rk, r2, t, P);
            if (P.mutual && m >= 0 && nn_from_keys(ck, c2, m, P) != t) m = -1;  // mutual_check, MNN.py:25-32
        }
        P.m0[(size_t)b * P.ncap + t] = m;
        P.s0[(size_t)b * P.ncap + t] = m >= 0 ? 1.0f : 0.0f;
    }
    if (t < P.mcap) {
        int m = -1;
        if (t < M) {
            m = nn_from_keys(ck, c2, t, P);
            if (P.mutual && m >= 0 && nn_from_keys(rk, r2, m, P) != t) m = -1;
        }
        P.m1[(size_t)b * P.mcap + t] = m;
        P.s1[(size_t)b * P.mcap + t] = m >= 0 ? 1.0f : 0.0f;
    }
}

// matched_kpts0 = kpts0[m0 > -1], matched_kpts1 = kpts1[m0[m0 > -1]] in ascending i (MNN.py:103-129)
__global__ void __launch_bounds__(256)
mnn_gather_kernel(const int64_t* __restrict__ m0, const float* __restrict__ k0, const float* __restrict__ k1,
                  int ncap, int mcap, float* __restrict__ mk0, float* __restrict__ mk1, int32_t* __restrict__ nmatch) {
    const int b = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __shared__ int wsum[8];
    __shared__ int run_s;
    if (tid == 0) run_s = 0;
    __syncthreads();
    for (int base = 0; base < ncap; base += 256) {
        const int i = base + tid;
        const long long m = i < ncap ? m0[(size_t)b * ncap + i] : -1;
        const bool keep = m >= 0;
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) wsum[warp] = __popc(bal);
        __syncthreads();
        int woff = 0, tot = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            if (w < warp) woff += wsum[w];
            tot += wsum[w];
        }
        const int run = run_s;
        if (keep) {
            const int pos = run + woff + __popc(bal & ((1u << lane) - 1u));
            const float* a = k0 + ((size_t)b * ncap + i) * 3;
            const float* c = k1 + ((size_t)b * mcap + m) * 3;
            float* oa = mk0 + ((size_t)b * ncap + pos) * 3;
            float* oc = mk1 + ((size_t)b * ncap + pos) * 3;
            oa[0] = a[0]; oa[1] = a[1]; oa[2] = a[2];
            oc[0] = c[0]; oc[1] = c[1]; oc[2] = c[2];
        }
        __syncthreads();
        if (tid == 0) run_s = run + tot;
        __syncthreads();
    }
    if (tid == 0) nmatch[b] = run_s;
}

// Finalisation and matched-keypoint compaction in one launch: one 1024-thread CTA per pair decides its
// rows and columns chunk by chunk (same arithmetic as mnn_finalize_kernel) and compacts the matched
// keypoint rows of the chunk in ascending i with a block scan -- the two tiny kernels this replaces sat
// back to back on the critical path of a step behind the similarity kernel.
constexpr int kFinThreads = 1024;

__global__ void __launch_bounds__(kFinThreads)
mnn_finalize_gather_kernel(const FinalizeParams P, const float* __restrict__ k0, const float* __restrict__ k1,
                           float* __restrict__ mk0, float* __restrict__ mk1, int32_t* __restrict__ nmatch) {
    const int b = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int N = P.n0 ? min(P.n0[b], P.ncap) : P.ncap;
    const int M = P.n1 ? min(P.n1[b], P.mcap) : P.mcap;
    const unsigned long long* rk = P.rowkey + (size_t)b * P.ncap;
    const unsigned long long* ck = P.colkey + (size_t)b * P.mcap;
    const unsigned long long* r2 = P.row2 ? P.row2 + (size_t)b * P.ncap : nullptr;
    const unsigned long long* c2 = P.col2 ? P.col2 + (size_t)b * P.mcap : nullptr;
    __shared__ int wsum[kFinThreads / 32];
    __shared__ int run_s;
    if (tid == 0) run_s = 0;
    __syncthreads();
    const int mx = max(P.ncap, P.mcap);
    for (int base = 0; base < mx; base += kFinThreads) {
        const int t = base + tid;
        int m = -1;
        if (t < P.ncap) {
            if (t < N) {
                m = nn_from_keys(rk, r2, t, P);
                if (P.mutual && m >= 0 && nn_from_keys(ck, c2, m, P) != t) m = -1;  // mutual_check, MNN.py:25-32
            }
            P.m0[(size_t)b * P.ncap + t] = m;
            P.s0[(size_t)b * P.ncap + t] = m >= 0 ? 1.0f : 0.0f;
        }
        if (t < P.mcap) {
            int mm = -1;
            if (t < M) {
                mm = nn_from_keys(ck, c2, t, P);
                if (P.mutual && mm >= 0 && nn_from_keys(rk, r2, mm, P) != t) mm = -1;
            }
            P.m1[(size_t)b * P.mcap + t] = mm;
            P.s1[(size_t)b * P.mcap + t] = mm >= 0 ? 1.0f : 0.0f;
        }
        // matched_kpts0 = kpts0[m0 > -1], matched_kpts1 = kpts1[m0[m0 > -1]] in ascending i (MNN.py:103-129)
        const bool keep = m >= 0;
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) wsum[warp] = __popc(bal);
        __syncthreads();
        int woff = 0, tot = 0;
#pragma unroll
        for (int w = 0; w < kFinThreads / 32; ++w) {
            const int c = wsum[w];
            if (w < warp) woff += c;
            tot += c;
        }
        const int run = run_s;
        if (keep) {
            const int pos = run + woff + __popc(bal & ((1u << lane) - 1u));
            const float* a = k0 + ((size_t)b * P.ncap + t) * 3;
            const float* c = k1 + ((size_t)b * P.mcap + m) * 3;
            float* oa = mk0 + ((size_t)b * P.ncap + pos) * 3;
            float* oc = mk1 + ((size_t)b * P.ncap + pos) * 3;
            oa[0] = a[0]; oa[1] = a[1]; oa[2] = a[2];
            oc[0] = c[0]; oc[1] = c[1]; oc[2] = c[2];
        }
        __syncthreads();
        if (tid == 0) run_s = run + tot;
        __syncthreads();
    }
    if (tid == 0) nmatch[b] = run_s;
}

// ---- opt-in dense by-products -------------------------------------------------------------- //
__global__ void lse_rows_kernel(const float* __restrict__ sim, int N, int M, float* __restrict__ lse) {
    const int b = blockIdx.y, i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= N) return;
    const float* row = sim + ((size_t)b * N + i) * M;
    float mx = -INFINITY;
    for (int j = lane; j < M; j += 32) mx = fmaxf(mx, row[j]);
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float s = 0.0f;
    for (int j = lane; j < M; j += 32) s += expf(row[j] - mx);
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) lse[(size_t)b * N + i] = mx + logf(s);
}
__global__ void lse_cols_kernel(const float* __restrict__ sim, int N, int M, float* __restrict__ lse) {
    const int b = blockIdx.y, j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= M) return;
    const float* col = sim + (size_t)b * N * M + j;
    float mx = -INFINITY;
    for (int i = 0; i < N; ++i) mx = fmaxf(mx, col[(size_t)i * M]);
    float s = 0.0f;
    for (int i = 0; i < N; ++i) s += expf(col[(size_t)i * M] - mx);
    lse[(size_t)b * M + j] = mx + logf(s);
}
// log_assignment[:N,:M] = log_softmax(sim,-1) + log_softmax(sim,-2); last row/col 0 (MNN.py:96-98)
__global__ void log_assignment_kernel(const float* __restrict__ sim, const float* __restrict__ lr,
                                      const float* __restrict__ lc, int N, int M, float* __restrict__ la) {
    const int b = blockIdx.z, i = blockIdx.y, j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j > M) return;
    float v = 0.0f;
    if (i < N && j < M) {
        const float s = sim[((size_t)b * N + i) * M + j];
        v = (s - lr[(size_t)b * N + i]) + (s - lc[(size_t)b * M + j]);
    }
    la[((size_t)b * (N + 1) + i) * (M + 1) + j] = v;
}

}  // namespace

int einx_mnn_tc(einx_ctx* ctx, const float* d0, const float* d1, const int32_t* n0, const int32_t* n1, int B, int ncap,
                int mcap, int D, int precision, unsigned long long* rowkey, unsigned long long* colkey,
                unsigned char* scratch, size_t scratch_bytes, const uint16_t* split0, const uint16_t* split1,
                cudaStream_t stream);
size_t einx_mnn_tc_scratch_bytes(int B, int ncap, int mcap, int D, int precision, bool have_split);
bool einx_mnn_tc_supported(const float* d0, const float* d1, int D, int precision);

extern "C" int einx_mnn(einx_ctx* ctx, const float* d0, const float* d1, const int32_t* n0, const int32_t* n1, int B,
                        int ncap, int mcap, int D, float ratio_thresh, float distance_thresh, int mutual,
                        int precision, int64_t* m0, int64_t* m1, float* s0, float* s1, const float* kpts0,
                        const float* kpts1, float* mk0, float* mk1, int32_t* nmatch, einx_stream stream_) {
    return einx_mnn_split(ctx, d0, d1, nullptr, nullptr, n0, n1, B, ncap, mcap, D, ratio_thresh, distance_thresh, mutual,
                          precision, m0, m1, s0, s1, kpts0, kpts1, mk0, mk1, nmatch, stream_);
}

extern "C" int einx_mnn_split(einx_ctx* ctx, const float* d0, const float* d1, const uint16_t* split0,
                              const uint16_t* split1, const int32_t* n0, const int32_t* n1, int B, int ncap, int mcap,
                              int D, float ratio_thresh, float distance_thresh, int mutual, int precision, int64_t* m0,
                              int64_t* m1, float* s0, float* s1, const float* kpts0, const float* kpts1, float* mk0,
                              float* mk1, int32_t* nmatch, einx_stream stream_) {
    if (!ctx) return EINX_ERR_INVALID;
    if ((split0 != nullptr) != (split1 != nullptr))
        return einx_fail(ctx, EINX_ERR_INVALID, "einx_mnn_split: split0 and split1 come together");
    if (split0 && (((uintptr_t)split0 | (uintptr_t)split1) & 15))
        return einx_fail(ctx, EINX_ERR_INVALID, "einx_mnn_split: split operands must be 16-byte aligned");
    if (B < 0 || ncap < 0 || mcap < 0 || D <= 0)
        return einx_fail(ctx, EINX_ERR_INVALID, "einx_mnn: bad shape B=%d N=%d M=%d D=%d", B, ncap, mcap, D);
    if (B == 0) return EINX_OK;
    if (B > 65535) return einx_fail(ctx, EINX_ERR_UNSUPPORTED, "einx_mnn: B=%d > 65535", B);
    if (!m0 || !m1 || !s0 || !s1) return einx_fail(ctx, EINX_ERR_INVALID, "einx_mnn: NULL output pointer");
    if ((ncap > 0 && !d0) || (mcap > 0 && !d1)) return einx_fail(ctx, EINX_ERR_INVALID, "einx_mnn: NULL descriptor pointer");
    if (kpts0 && (!kpts1 || !mk0 || !mk1 || !nmatch))
        return einx_fail(ctx, EINX_ERR_INVALID, "einx_mnn: kpts0 given but kpts1/mk0/mk1/nmatch missing");
    if (precision < EINX_MNN_FP32 || precision > EINX_MNN_FP16X3)
        return einx_fail(ctx, EINX_ERR_INVALID, "einx_mnn: unknown precision %d", precision);
    DeviceGuard guard(ctx->device);
    cudaStream_t stream = (cudaStream_t)stream_;
    const bool use_ratio = ratio_thresh > 0.0f;
    // TMA needs 16-byte row pitches; other feature sizes take the (more exact) FFMA path
    if (precision != EINX_MNN_FP32 && !einx_mnn_tc_supported(d0, d1, D, precision)) precision = EINX_MNN_FP32;

    const size_t rk_bytes = align_up((size_t)B * ncap * 8, 256), ck_bytes = align_up((size_t)B * mcap * 8, 256);
    const size_t key_bytes = (rk_bytes + ck_bytes) * (use_ratio ? 2 : 1);
    const size_t tc_bytes = precision == EINX_MNN_FP32 ? 0 : einx_mnn_tc_scratch_bytes(B, ncap, mcap, D, precision, split0 != nullptr);
    int rc = einx_ws_reserve(ctx, key_bytes + tc_bytes + 256, stream);
    if (rc) return rc;
    unsigned char* ws = (unsigned char*)ctx->ws;
    unsigned long long* rowkey = (unsigned long long*)ws;
    unsigned long long* colkey = (unsigned long long*)(ws + rk_bytes);
    unsigned long long* row2 = use_ratio ? (unsigned long long*)(ws + rk_bytes + ck_bytes) : nullptr;
    unsigned long long* col2 = use_ratio ? (unsigned long long*)(ws + 2 * rk_bytes + ck_bytes) : nullptr;
    if (key_bytes) EINX_CUDA(ctx, cudaMemsetAsync(ws, 0, key_bytes, stream));

    if (ncap > 0 && mcap > 0) {
        dim3 grid((mcap + BN - 1) / BN, (ncap + BM - 1) / BM, B);
        if (precision == EINX_MNN_FP32 || use_ratio) {
            einx_prof_begin(ctx, 3, stream);
            // the ratio test needs exact second-best values: it always runs on the fp32 path
            mnn_fp32_kernel<0><<<grid, kGemmThreads, 0, stream>>>(d0, d1, n0, n1, ncap, mcap, D, rowkey, colkey,
                                                                  nullptr, nullptr, nullptr);
            einx_prof_end(ctx, 3, stream);
            EINX_CHECK_LAUNCH(ctx);
            if (use_ratio) {
                mnn_fp32_kernel<1><<<grid, kGemmThreads, 0, stream>>>(d0, d1, n0, n1, ncap, mcap, D, row2, col2,
                                                                      rowkey, colkey, nullptr);
                EINX_CHECK_LAUNCH(ctx);
            }
        } else {
            rc = einx_mnn_tc(ctx, d0, d1, n0, n1, B, ncap, mcap, D, precision, rowkey, colkey, ws + key_bytes,
                             tc_bytes, split0, split1, stream);
            if (rc) return rc;
        }
    }
    FinalizeParams F = {};
    F.rowkey = rowkey; F.colkey = colkey; F.row2 = row2; F.col2 = col2;
    F.n0 = n0; F.n1 = n1; F.ncap = ncap; F.mcap = mcap;
    F.use_ratio = use_ratio; F.use_dist = distance_thresh > 0.0f; F.mutual = mutual != 0;
    F.ratio_sq = (float)((double)ratio_thresh * (double)ratio_thresh);
    F.dist_sq = (float)((double)distance_thresh * (double)distance_thresh);
    F.m0 = m0; F.m1 = m1; F.s0 = s0; F.s1 = s1;
    const int mx = ncap > mcap ? ncap : mcap;
    if (kpts0 && mx > 0) {
        mnn_finalize_gather_kernel<<<B, kFinThreads, 0, stream>>>(F, kpts0, kpts1, mk0, mk1, nmatch);
        EINX_CHECK_LAUNCH(ctx);
        return EINX_OK;
    }
    if (mx > 0) {
        mnn_finalize_kernel<<<dim3((mx + 255) / 256, B), 256, 0, stream>>>(F);
        EINX_CHECK_LAUNCH(ctx);
    }
    if (kpts0) {
        mnn_gather_kernel<<<B, 256, 0, stream>>>(m0, kpts0, kpts1, ncap, mcap, mk0, mk1, nmatch);
        EINX_CHECK_LAUNCH(ctx);
    }
    return EINX_OK;
}

extern "C" int einx_mnn_dense(einx_ctx* ctx, const float* d0, const float* d1, int B, int N, int M, int D,
                              float* similarity, float* log_assignment, einx_stream stream_) {
    if (!ctx) return EINX_ERR_INVALID;
    if (B <= 0 || N <= 0 || M <= 0 || D <= 0 || !d0 || !d1 || !similarity)
        return einx_fail(ctx, EINX_ERR_INVALID, "einx_mnn_dense: bad argument");
    if (B > 65535 || N > 65535) return einx_fail(ctx, EINX_ERR_UNSUPPORTED, "einx_mnn_dense: B or N > 65535");
    DeviceGuard guard(ctx->device);
    cudaStream_t stream = (cudaStream_t)stream_;
    dim3 grid((M + BN - 1) / BN, (N + BM - 1) / BM, B);
    mnn_fp32_kernel<2><<<grid, kGemmThreads, 0, stream>>>(d0, d1, nullptr, nullptr, N, M, D, nullptr, nullptr, nullptr,
                                                          nullptr, similarity);
    EINX_CHECK_LAUNCH(ctx);
    if (log_assignment) {
        int rc = einx_ws_reserve(ctx, sizeof(float) * (size_t)B * (N + M), stream);
        if (rc) return rc;
        float* lr = (float*)ctx->ws;
        float* lc = lr + (size_t)B * N;
        lse_rows_kernel<<<dim3((N + 7) / 8, B), 256, 0, stream>>>(similarity, N, M, lr);
        EINX_CHECK_LAUNCH(ctx);
        lse_cols_kernel<<<dim3((M + 127) / 128, B), 128, 0, stream>>>(similarity, N, M, lc);
        EINX_CHECK_LAUNCH(ctx);
        log_assignment_kernel<<<dim3((M + 1 + 127) / 128, N + 1, B), 128, 0, stream>>>(similarity, lr, lc, N, M,
                                                                                      log_assignment);
        EINX_CHECK_LAUNCH(ctx);
    }
    return EINX_OK;
}
