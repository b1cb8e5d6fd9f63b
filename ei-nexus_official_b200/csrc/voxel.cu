// Event voxelisation: trilinear scatter + non-zero mean/std normalisation.
// Semantics: reference datasets/representations.py:8-22 and :66-124 (see include/einx.h).
//
// Layout in HBM: SoA events (x, y, p fp32; t fp64) for the whole ragged batch, grid
// (B, bins, H, W) fp32.  The grid of a batch is L2-resident on B200 (C2: 64 x 0.86 MB), so the
// zero / scatter / stats / apply passes hit L2, and DRAM sees the events once and the grid once.
#include <cooperative_groups.h>
#include <stdlib.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int kScatterThreads = 256;
constexpr int kEventsPerThread = 4;

__device__ __forceinline__ void red_add(float* addr, float v) {
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void red_add2(float* addr, float a, float b) {
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(a), "f"(b) : "memory");
}

__device__ __forceinline__ void red_add4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

struct TimeBase {
    double t0;     // first timestamp of the window
    double denom;  // (t[-1] - t[0]) + 1e-8                  representations.py:19-20
    float tf0;     // fp32 of the normalised first timestamp  (:76, :81)
    float span;    // t_norm[-1] - t_norm[0] in fp32          (:81)
};

__device__ __forceinline__ TimeBase make_time_base(const double* __restrict__ t, int64_t beg, int64_t end) {
    TimeBase tb;
    tb.t0 = t[beg];
    double tl = t[end - 1];
    tb.denom = (tl - tb.t0) + 1e-8;
    tb.tf0 = (float)((tb.t0 - tb.t0) / tb.denom);
    float tfN = (float)((tl - tb.t0) / tb.denom);
    tb.span = __fsub_rn(tfN, tb.tf0);
    return tb;
}

// One event -> up to 8 corner contributions.  An L2 reduction costs the same whatever its width
// (tools/microbench_red.cu), so x-adjacent corners go out as ONE 16-byte vector reduction on the
// aligned group of four cells that holds both (zeros in the other two lanes: x + 0 = x), unless the
// pair straddles two groups or the group leaves the buffer [lo, hi).  Exact-zero weights
// (integer-pixel EC events) are skipped, which cannot change any cell nor the `!= 0` mask.
__device__ __forceinline__ void splat_event(float* __restrict__ g, float xf, float yf, float tn, float pf,
                                            int bins, int H, int W, const float* lo, const float* hi) {
    if (tn != tn) return;  // 0/0 time span: the reference's NaN bin index is out of range
    const float pol = pf < 1.0f ? -1.0f : pf;  // value[value < 1] = -1   (:88-89)
    const int x0 = (int)xf, y0 = (int)yf, t0 = (int)tn;  // .int() truncates (:83-85)
    if ((float)x0 == xf && (float)y0 == yf) {
        // Integer-pixel event (the EC dataset, raw sensor streams): the x+1 and y+1 corners carry weight pol * 0 and can
        // change neither a cell nor the `!= 0` mask, and the remaining weight pol * 1 * 1 * wt is exactly pol * wt --
        // two scalar reductions, same bits as the general path below.
        if ((unsigned)x0 >= (unsigned)W || (unsigned)y0 >= (unsigned)H) return;
        float* cell = g + ((size_t)t0 * H + y0) * W + x0;
#pragma unroll
        for (int dt = 0; dt < 2; ++dt) {
            const int tl = t0 + dt;
            if (tl < 0 || tl >= bins) continue;
            const float w = __fmul_rn(pol, __fsub_rn(1.0f, fabsf(__fsub_rn((float)tl, tn))));
            if (w != 0.0f) red_add(cell + (size_t)dt * H * W, w);
        }
        return;
    }
    const float wx0 = __fmul_rn(pol, __fsub_rn(1.0f, fabsf(__fsub_rn((float)x0, xf))));
    const float wx1 = __fmul_rn(pol, __fsub_rn(1.0f, fabsf(__fsub_rn((float)(x0 + 1), xf))));
    const bool vx0 = (x0 >= 0) & (x0 < W);
    const bool vx1 = (x0 + 1 >= 0) & (x0 + 1 < W);
#pragma unroll
    for (int dy = 0; dy < 2; ++dy) {
        const int yl = y0 + dy;
        if (yl < 0 || yl >= H) continue;
        const float wy = __fsub_rn(1.0f, fabsf(__fsub_rn((float)yl, yf)));
        const float a0 = __fmul_rn(wx0, wy), a1 = __fmul_rn(wx1, wy);
#pragma unroll
        for (int dt = 0; dt < 2; ++dt) {
            const int tl = t0 + dt;
            if (tl < 0 || tl >= bins) continue;
            const float wt = __fsub_rn(1.0f, fabsf(__fsub_rn((float)tl, tn)));
            const float w0 = vx0 ? __fmul_rn(a0, wt) : 0.0f;
            const float w1 = vx1 ? __fmul_rn(a1, wt) : 0.0f;
            if (w0 == 0.0f && w1 == 0.0f) continue;
            float* cell = g + ((size_t)tl * H + yl) * W + x0;
            const unsigned off = (unsigned)(reinterpret_cast<uintptr_t>(cell) & 15u) >> 2;  // x0's slot in its group
            float* group = cell - off;
            if (vx0 && vx1 && off < 3u && group >= lo && group + 4 <= hi && w0 != 0.0f && w1 != 0.0f) {
                red_add4(group, off == 0 ? w0 : 0.0f, off == 0 ? w1 : (off == 1 ? w0 : 0.0f),
                         off == 1 ? w1 : (off == 2 ? w0 : 0.0f), off == 2 ? w1 : 0.0f);
            } else if (vx0 && vx1 && ((reinterpret_cast<uintptr_t>(cell) & 7u) == 0)) {
                red_add2(cell, w0, w1);
            } else {
                if (w0 != 0.0f) red_add(cell, w0);
                if (w1 != 0.0f) red_add(cell + 1, w1);
            }
        }
    }
}

__global__ void __launch_bounds__(kScatterThreads)
voxel_scatter_kernel(const float* __restrict__ x, const float* __restrict__ y, const double* __restrict__ t,
                     const float* __restrict__ p, const int64_t* __restrict__ off, int bins, int H, int W,
                     float* __restrict__ out) {
    const int b = blockIdx.y;
    const int64_t beg = off[b], end = off[b + 1];
    if (end - beg <= 0) return;
    const TimeBase tb = make_time_base(t, beg, end);
    const float bm1 = (float)(bins - 1);
    float* g = out + (size_t)b * bins * H * W;
    const float* out_end = out + (size_t)gridDim.y * bins * H * W;
    const int64_t stride = (int64_t)gridDim.x * kScatterThreads * kEventsPerThread;
    for (int64_t base = beg + (int64_t)blockIdx.x * kScatterThreads * kEventsPerThread; base < end; base += stride) {
        float xf[kEventsPerThread], yf[kEventsPerThread], pf[kEventsPerThread];
        double td[kEventsPerThread];
        // issue all loads of this thread's events before any dependent math (coalesced per warp)
#pragma unroll
        for (int k = 0; k < kEventsPerThread; ++k) {
            const int64_t i = base + k * kScatterThreads + threadIdx.x;
            const bool ok = i < end;
            xf[k] = ok ? __ldg(x + i) : 0.f;
            yf[k] = ok ? __ldg(y + i) : 0.f;
            pf[k] = ok ? __ldg(p + i) : 0.f;
            td[k] = ok ? __ldg(t + i) : tb.t0;
        }
#pragma unroll
        for (int k = 0; k < kEventsPerThread; ++k) {
            const int64_t i = base + k * kScatterThreads + threadIdx.x;
            if (i >= end) continue;
            const float tf = (float)((td[k] - tb.t0) / tb.denom);                          // :19-20, :76
            const float tn = __fdiv_rn(__fmul_rn(bm1, __fsub_rn(tf, tb.tf0)), tb.span);  // :81
            splat_event(g, xf[k], yf[k], tn, pf[k], bins, H, W, out, out_end);
        }
    }
}

// ---- normalisation: statistics over cells != 0, then apply -------------------------------- //
struct Stats {
    double n, sum, sumsq;
};

__global__ void __launch_bounds__(256)
voxel_stats_kernel(const float* __restrict__ grid, size_t ncell, double* __restrict__ stats) {
    const int b = blockIdx.y;
    const float* g = grid + (size_t)b * ncell;
    double n = 0, s = 0, ss = 0;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t nthr = (size_t)gridDim.x * blockDim.x;
    if (((ncell & 3) == 0) && ((reinterpret_cast<uintptr_t>(g) & 15u) == 0)) {
        const float4* g4 = reinterpret_cast<const float4*>(g);
        for (size_t i = tid; i < ncell / 4; i += nthr) {
            const float4 v = g4[i];
            const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (e[k] != 0.0f) { n += 1.0; s += (double)e[k]; ss += (double)e[k] * (double)e[k]; }
        }
    } else {
        for (size_t i = tid; i < ncell; i += nthr) {
            const float e = g[i];
            if (e != 0.0f) { n += 1.0; s += (double)e; ss += (double)e * (double)e; }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        n += __shfl_xor_sync(0xffffffffu, n, o);
        s += __shfl_xor_sync(0xffffffffu, s, o);
        ss += __shfl_xor_sync(0xffffffffu, ss, o);
    }
    __shared__ double sh[3][8];
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { sh[0][w] = n; sh[1][w] = s; sh[2][w] = ss; }
    __syncthreads();
    if (threadIdx.x < 3) {
        double a = 0;
        for (int k = 0; k < 8; ++k) a += sh[threadIdx.x][k];
        if (a != 0.0) atomicAdd(stats + 3 * b + threadIdx.x, a);
    }
}

__global__ void __launch_bounds__(256)
voxel_apply_kernel(float* __restrict__ grid, size_t ncell, const double* __restrict__ stats) {
    const int b = blockIdx.y;
    const double n = stats[3 * b], s = stats[3 * b + 1], ss = stats[3 * b + 2];
    if (n <= 0.0) return;
    const double mean_d = s / n;
    const float mean = (float)mean_d;
    // unbiased std (torch.Tensor.std); a single cell gives nan, which fails `std > 0` (:118-121)
    float sd = nanf("");
    if (n > 1.0) {
        double var = (ss - n * mean_d * mean_d) / (n - 1.0);
        sd = (float)sqrt(var > 0.0 ? var : 0.0);
    }
    const bool divide = sd > 0.0f;
    float* g = grid + (size_t)b * ncell;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t nthr = (size_t)gridDim.x * blockDim.x;
    if (((ncell & 3) == 0) && ((reinterpret_cast<uintptr_t>(g) & 15u) == 0)) {
        float4* g4 = reinterpret_cast<float4*>(g);
        for (size_t i = tid; i < ncell / 4; i += nthr) {
            float4 v = g4[i];
            float e[4] = {v.x, v.y, v.z, v.w};
            bool any = false;
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (e[k] != 0.0f) {
                    const float c = __fsub_rn(e[k], mean);
                    e[k] = divide ? __fdiv_rn(c, sd) : c;
                    any = true;
                }
            if (any) g4[i] = make_float4(e[0], e[1], e[2], e[3]);
        }
    } else {
        for (size_t i = tid; i < ncell; i += nthr) {
            const float e = g[i];
            if (e != 0.0f) {
                const float c = __fsub_rn(e, mean);
                g[i] = divide ? __fdiv_rn(c, sd) : c;
            }
        }
    }
}

// ---- normalisation, fused: one thread-block cluster per window ------------------------------- //
// Each CTA of the cluster pulls its slice of the window's grid (L2-resident after the scatter) into
// shared memory once, the (count, sum, sum of squares) of the non-zero cells is reduced across the
// cluster through DSMEM, and the slice is normalised out of shared memory: the grid is read once and
// written once, with no statistics buffer, memset or second launch.
constexpr int kNormThreads = 512;

struct NormShared {
    double red[3][kNormThreads / 32];
    double part[3];
};

__global__ void __launch_bounds__(kNormThreads)
voxel_norm_cluster_kernel(float* __restrict__ grid, size_t ncell, int slice) {
    cg::cluster_group cluster = cg::this_cluster();
    const int CS = (int)cluster.num_blocks(), r = (int)cluster.block_rank();
    const int b = blockIdx.x / CS;
    extern __shared__ __align__(16) unsigned char norm_raw[];
    NormShared& sh = *reinterpret_cast<NormShared*>(norm_raw);
    float* buf = reinterpret_cast<float*>(norm_raw + align_up(sizeof(NormShared), 16));
    const size_t start = (size_t)r * slice;
    const int n = start < ncell ? (int)min((size_t)slice, ncell - start) : 0;  // slice % 4 == 0, ncell % 4 == 0
    float4* g4 = reinterpret_cast<float4*>(grid + (size_t)b * ncell + start);
    float4* b4 = reinterpret_cast<float4*>(buf);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double cn = 0.0, cs = 0.0, css = 0.0;
    // the whole slice in flight at once (cp.async, no register staging): one memory latency instead of one per iteration
    {
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(b4);
        for (int i = tid; i < n / 4; i += kNormThreads)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16u * (uint32_t)i), "l"(g4 + i) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    for (int i = tid; i < n / 4; i += kNormThreads) {  // (a thread reads back exactly the chunks it copied)
        const float4 v = b4[i];
        const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (e[k] != 0.0f) { cn += 1.0; cs += (double)e[k]; css += (double)e[k] * (double)e[k]; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        cn += __shfl_xor_sync(0xffffffffu, cn, o);
        cs += __shfl_xor_sync(0xffffffffu, cs, o);
        css += __shfl_xor_sync(0xffffffffu, css, o);
    }
    if (lane == 0) { sh.red[0][warp] = cn; sh.red[1][warp] = cs; sh.red[2][warp] = css; }
    __syncthreads();
    if (tid < 3) {
        double a = 0.0;
        for (int w = 0; w < kNormThreads / 32; ++w) a += sh.red[tid][w];
        sh.part[tid] = a;
    }
    cluster.sync();
    double wn = 0.0, ws = 0.0, wss = 0.0;
    for (int q = 0; q < CS; ++q) {  // same order in every CTA: identical statistics
        const double* pq = cluster.map_shared_rank(sh.part, q);
        wn += pq[0]; ws += pq[1]; wss += pq[2];
    }
    cluster.sync();  // peers have read this CTA's partials: it may exit
    if (wn <= 0.0) return;
    const double mean_d = ws / wn;
    const float mean = (float)mean_d;
    // unbiased std (torch.Tensor.std); a single cell gives nan, which fails `std > 0` (:118-121)
    float sd = nanf("");
    if (wn > 1.0) {
        const double var = (wss - wn * mean_d * mean_d) / (wn - 1.0);
        sd = (float)sqrt(var > 0.0 ? var : 0.0);
    }
    const bool divide = sd > 0.0f;
    for (int i = tid; i < n / 4; i += kNormThreads) {
        const float4 v = b4[i];
        float e[4] = {v.x, v.y, v.z, v.w};
        bool any = false;
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (e[k] != 0.0f) {
                const float c = __fsub_rn(e[k], mean);
                e[k] = divide ? __fdiv_rn(c, sd) : c;
                any = true;
            }
        if (any) g4[i] = make_float4(e[0], e[1], e[2], e[3]);
    }
}


// ---- shared-memory tile path: one thread-block cluster per window ----------------------------- //
// CTA (bin, row band) of the cluster keeps its slice of the grid in shared memory.  Events are
// time-sorted, so the events that touch bin k -- those with t_norm in (k-1, k+1) -- are one contiguous
// range, found by a warp-wide 32-ary search over the same fp32 expression the splat uses.  The CTA
// accumulates that range with shared-memory atomics, the cluster exchanges the (count, sum, sum of
// squares) of its non-zero cells through DSMEM, and every slice is normalised and written to HBM
// exactly once: no memset, no global reductions, no second pass over the grid.  Each event is read by
// the (at most two) bins it touches; the second read comes from L2.  Unsorted windows are detected
// (cluster-wide check of t) and handled by scanning the whole window in every CTA.
constexpr int kTileThreads = 1024;

struct TileShared {
    double red[3][kTileThreads / 32];
    double part[3];
    int unsorted;
    long long bound[2];
};

__device__ __forceinline__ float event_tnorm(double td, const TimeBase& tb, float bm1) {
    const float tf = (float)((td - tb.t0) / tb.denom);                          // :19-20, :76
    return __fdiv_rn(__fmul_rn(bm1, __fsub_rn(tf, tb.tf0)), tb.span);           // :81
}

// first i in [L, R) for which t_norm(i) > bound (STRICT) or >= bound; R if none.  One warp.
template <bool STRICT>
__device__ long long tnorm_search(const double* __restrict__ t, long long L, long long R, const TimeBase& tb, float bm1,
                                  float bound, int lane) {
    while (R > L) {
        const long long n = R - L;
        const long long s = (n + 31) / 32;
        const long long i = L + (long long)lane * s;
        bool pr = false;
        if (i < R) {
            const float tn = event_tnorm(__ldg(t + i), tb, bm1);
            pr = STRICT ? (tn > bound) : (tn >= bound);
        }
        const unsigned valid = __ballot_sync(0xffffffffu, i < R);
        const unsigned hit = __ballot_sync(0xffffffffu, pr);
        if (hit) {
            const int j = __ffs(hit) - 1;
            R = L + (long long)j * s;                 // first probe that satisfies the predicate
            if (j == 0) break;
            L = L + (long long)(j - 1) * s + 1;       // the probe before it does not
        } else {
            const int last = 31 - __clz(valid);
            L = L + (long long)last * s + 1;
        }
        if (s == 1) break;  // probes were consecutive: R is exact
    }
    return R;
}

__device__ __forceinline__ void splat_tile(float* __restrict__ tile, float xf, float yf, float tn, float pf, int bin, int ys,
                                           int nrows, int W) {
    if (tn != tn) return;  // 0/0 time span: the reference's NaN bin index is out of range
    const int t0 = (int)tn;
    if (t0 != bin && t0 + 1 != bin) return;
    const float wt = __fsub_rn(1.0f, fabsf(__fsub_rn((float)bin, tn)));
    const float pol = pf < 1.0f ? -1.0f : pf;
    const int x0 = (int)xf, y0 = (int)yf;
    const float wx0 = __fmul_rn(pol, __fsub_rn(1.0f, fabsf(__fsub_rn((float)x0, xf))));
    const float wx1 = __fmul_rn(pol, __fsub_rn(1.0f, fabsf(__fsub_rn((float)(x0 + 1), xf))));
    const bool vx0 = (x0 >= 0) & (x0 < W);
    const bool vx1 = (x0 + 1 >= 0) & (x0 + 1 < W);
#pragma unroll
    for (int dy = 0; dy < 2; ++dy) {
        const int yl = y0 + dy;
        const int row = yl - ys;
        if ((unsigned)row >= (unsigned)nrows) continue;  // band rows lie inside [0, H)
        const float wy = __fsub_rn(1.0f, fabsf(__fsub_rn((float)yl, yf)));
        const float w0 = __fmul_rn(__fmul_rn(wx0, wy), wt);
        const float w1 = __fmul_rn(__fmul_rn(wx1, wy), wt);
        float* cell = tile + row * W + x0;
        if (vx0 && w0 != 0.0f) atomicAdd(cell, w0);
        if (vx1 && w1 != 0.0f) atomicAdd(cell + 1, w1);
    }
}

__global__ void __launch_bounds__(kTileThreads, 1)
voxel_tile_kernel(const float* __restrict__ x, const float* __restrict__ y, const double* __restrict__ t,
                  const float* __restrict__ p, const int64_t* __restrict__ off, int bins, int H, int W, int bands,
                  int band_rows, int normalize, float* __restrict__ out) {
    cg::cluster_group cluster = cg::this_cluster();
    const int CS = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
    const int b = blockIdx.x / CS;
    const int bin = rank / bands, band = rank - bin * bands;
    const int ys = band * band_rows;
    const int nrows = max(0, min(band_rows, H - ys));
    const int ncell = nrows * W;
    extern __shared__ __align__(16) unsigned char tile_raw[];
    TileShared& sh = *reinterpret_cast<TileShared*>(tile_raw);
    float* tile = reinterpret_cast<float*>(tile_raw + align_up(sizeof(TileShared), 16));
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    for (int i = tid; i < (ncell + 3) / 4; i += kTileThreads) reinterpret_cast<float4*>(tile)[i] = make_float4(0.f, 0.f, 0.f, 0.f);

    const long long beg = off[b], end = off[b + 1];
    const long long nev = end - beg;
    TimeBase tb = {};
    if (nev > 0) tb = make_time_base(t, beg, end);
    const float bm1 = (float)(bins - 1);

    // sortedness: this CTA checks its share of adjacent timestamp pairs
    int bad = 0;
    if (nev > 1) {
        const long long pairs = nev - 1, share = (pairs + CS - 1) / CS;
        const long long s0 = beg + (long long)rank * share, s1 = min(s0 + share, beg + pairs);
        for (long long i = s0 + tid; i < s1; i += kTileThreads) bad |= (__ldg(t + i) > __ldg(t + i + 1));
    }
    bad = __syncthreads_or(bad);  // also orders the tile zeroing before the atomics below
    if (tid == 0) sh.unsorted = bad;
    cluster.sync();
    int unsorted = 0;
    for (int r = 0; r < CS; ++r) unsorted |= *cluster.map_shared_rank(&sh.unsorted, r);

    long long lo = beg, hi = end;
    if (!unsorted && nev > 0) {
        if (warp == 0) {
            const long long v = bin == 0 ? beg : tnorm_search<true>(t, beg, end, tb, bm1, (float)(bin - 1), lane);
            if (lane == 0) sh.bound[0] = v;
        } else if (warp == 1) {
            const long long v = bin == bins - 1 ? end : tnorm_search<false>(t, beg, end, tb, bm1, (float)(bin + 1), lane);
            if (lane == 0) sh.bound[1] = v;
        }
        __syncthreads();
        lo = sh.bound[0];
        hi = sh.bound[1];
    }
    if (ncell > 0) {
        constexpr int U = 4;
        for (long long base = lo + tid; base < hi; base += (long long)U * kTileThreads) {
            float xf[U], yf[U], pf[U];
            double td[U];
#pragma unroll
            for (int k = 0; k < U; ++k) {
                const long long i = base + (long long)k * kTileThreads;
                const bool ok = i < hi;
                xf[k] = ok ? __ldg(x + i) : 0.f;
                yf[k] = ok ? __ldg(y + i) : 0.f;
                pf[k] = ok ? __ldg(p + i) : 0.f;
                td[k] = ok ? __ldg(t + i) : tb.t0;
            }
#pragma unroll
            for (int k = 0; k < U; ++k) {
                if (base + (long long)k * kTileThreads >= hi) continue;
                splat_tile(tile, xf[k], yf[k], event_tnorm(td[k], tb, bm1), pf[k], bin, ys, nrows, W);
            }
        }
    }
    __syncthreads();

    // statistics of the non-zero cells of the whole window, through DSMEM
    double cn = 0.0, cs = 0.0, css = 0.0;
    if (normalize) {
        for (int i = tid; i < ncell; i += kTileThreads) {
            const float e = tile[i];
            if (e != 0.0f) { cn += 1.0; cs += (double)e; css += (double)e * (double)e; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            cn += __shfl_xor_sync(0xffffffffu, cn, o);
            cs += __shfl_xor_sync(0xffffffffu, cs, o);
            css += __shfl_xor_sync(0xffffffffu, css, o);
        }
        if (lane == 0) { sh.red[0][warp] = cn; sh.red[1][warp] = cs; sh.red[2][warp] = css; }
        __syncthreads();
        if (tid < 3) {
            double a = 0.0;
            for (int w = 0; w < kTileThreads / 32; ++w) a += sh.red[tid][w];
            sh.part[tid] = a;
        }
    }
    cluster.sync();
    double wn = 0.0, ws = 0.0, wss = 0.0;
    if (normalize)
        for (int q = 0; q < CS; ++q) {  // same order in every CTA: identical statistics
            const double* pq = cluster.map_shared_rank(sh.part, q);
            wn += pq[0]; ws += pq[1]; wss += pq[2];
        }
    cluster.sync();  // peers have read this CTA's flag and partials: it may exit
    float mean = 0.0f, sd = 1.0f;
    bool shift = false, divide = false;
    if (normalize && wn > 0.0) {
        const double mean_d = ws / wn;
        mean = (float)mean_d;
        shift = true;
        // unbiased std (torch.Tensor.std); a single cell gives nan, which fails `std > 0` (:118-121)
        sd = nanf("");
        if (wn > 1.0) {
            const double var = (wss - wn * mean_d * mean_d) / (wn - 1.0);
            sd = (float)sqrt(var > 0.0 ? var : 0.0);
        }
        divide = sd > 0.0f;
    }
    float* dst = out + (((size_t)b * bins + bin) * H + ys) * W;
    auto fin = [&](float e) {
        if (shift && e != 0.0f) {
            const float c = __fsub_rn(e, mean);
            e = divide ? __fdiv_rn(c, sd) : c;
        }
        return e;
    };
    if ((ncell & 3) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15u) == 0) {
        for (int i = tid; i < ncell / 4; i += kTileThreads) {
            const float4 v = reinterpret_cast<const float4*>(tile)[i];
            __stcs(reinterpret_cast<float4*>(dst) + i, make_float4(fin(v.x), fin(v.y), fin(v.z), fin(v.w)));
        }
    } else {
        for (int i = tid; i < ncell; i += kTileThreads) dst[i] = fin(tile[i]);
    }
}

}  // namespace

static int voxelize_chunk(einx_ctx* ctx, const float* x, const float* y, const double* t, const float* p,
                          const int64_t* ev_offsets, int B, int bins, int H, int W, int normalize, float* out,
                          einx_stream stream_);

extern "C" int einx_voxelize(einx_ctx* ctx, const float* x, const float* y, const double* t, const float* p,
                             const int64_t* ev_offsets, int B, int bins, int H, int W, int normalize,
                             float* out, einx_stream stream_) {
    if (!ctx) return EINX_ERR_INVALID;
    // Development knob (EINX_VOXEL_CHUNK_MB): zero / scatter / normalise the batch in chunks of windows whose grids fit
    // that many MB, so that a chunk's grid stays in L2 between its three kernels.
    static const int chunk_mb = getenv("EINX_VOXEL_CHUNK_MB") ? atoi(getenv("EINX_VOXEL_CHUNK_MB")) : 0;
    if (chunk_mb > 0 && B > 1 && bins > 0 && H > 0 && W > 0 && ev_offsets && out) {
        const size_t per = (size_t)bins * H * W * 4;
        int cw = (int)(((size_t)chunk_mb << 20) / per);
        if (cw < 1) cw = 1;
        if (cw < B) {
            const int nchunks = (B + cw - 1) / cw;
            cw = (B + nchunks - 1) / nchunks;  // balanced
            for (int b0 = 0; b0 < B; b0 += cw) {
                const int nb = B - b0 < cw ? B - b0 : cw;
                int rc = voxelize_chunk(ctx, x, y, t, p, ev_offsets + b0, nb, bins, H, W, normalize, out + (size_t)b0 * per / 4, stream_);
                if (rc) return rc;
            }
            return EINX_OK;
        }
    }
    return voxelize_chunk(ctx, x, y, t, p, ev_offsets, B, bins, H, W, normalize, out, stream_);
}

static int voxelize_chunk(einx_ctx* ctx, const float* x, const float* y, const double* t, const float* p,
                          const int64_t* ev_offsets, int B, int bins, int H, int W, int normalize, float* out,
                          einx_stream stream_) {
    if (!ctx) return EINX_ERR_INVALID;
    if (B < 0 || bins <= 0 || H <= 0 || W <= 0)
        return einx_fail(ctx, EINX_ERR_INVALID, "einx_voxelize: bad shape B=%d bins=%d H=%d W=%d", B, bins, H, W);
    if (B == 0) return EINX_OK;
    if (!x || !y || !t || !p || !ev_offsets || !out)
        return einx_fail(ctx, EINX_ERR_INVALID, "einx_voxelize: NULL pointer argument");
    if (B > 65535) return einx_fail(ctx, EINX_ERR_UNSUPPORTED, "einx_voxelize: B=%d > 65535", B);
    DeviceGuard guard(ctx->device);
    cudaStream_t stream = (cudaStream_t)stream_;
    const size_t ncell = (size_t)bins * H * W;
    // Shared-memory tile path: small sensors whose bin slice fits one CTA (EC 180x240: 169 KB).  With
    // more than one row band per bin every band re-reads the bin's events, and sub-pixel events cost
    // four shared-memory CAS loops each, so larger sensors stay on the L2-reduction path below.
    {
        static const int tile_env = getenv("EINX_VOXEL_TILE") ? atoi(getenv("EINX_VOXEL_TILE")) : -1;
        const size_t fixed = align_up(sizeof(TileShared), 16);
        const size_t budget = (size_t)ctx->max_smem_optin - fixed;
        int bands = (int)(((size_t)H * W * 4 + budget - 1) / budget);
        const int band_rows = (H + bands - 1) / bands;
        bands = (H + band_rows - 1) / band_rows;
        const int CS = bins * bands;
        const size_t smem = fixed + align_up((size_t)band_rows * W * 4, 16);
        // Measured on B200 (C2, B=64; profiles/r1_voxel_tile.txt): 123 us for the whole window set against
        // 105 us for memset + L2-reduction scatter + cluster normalisation -- with one 170 KB CTA per SM
        // the zero / search / scan / statistics / flush phases each expose their latency and the
        // clusters run in 2.5 waves.  The path is therefore opt-in (EINX_VOXEL_TILE=1).
        const bool want = tile_env > 0;
        if (want && CS <= 16 && smem <= (size_t)ctx->max_smem_optin && (size_t)B * CS <= 0x7fffffffu) {
            EINX_CUDA(ctx, cudaFuncSetAttribute(voxel_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            if (CS > 8) EINX_CUDA(ctx, cudaFuncSetAttribute(voxel_tile_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(B * CS);
            cfg.blockDim = dim3(kTileThreads);
            cfg.dynamicSmemBytes = smem;
            cfg.stream = stream;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = CS;
            attr[0].val.clusterDim.y = 1;
            attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            int max_clusters = 0;
            if (cudaOccupancyMaxActiveClusters(&max_clusters, voxel_tile_kernel, &cfg) == cudaSuccess && max_clusters > 0) {
                einx_prof_begin(ctx, 0, stream);
                EINX_CUDA(ctx, cudaLaunchKernelEx(&cfg, voxel_tile_kernel, x, y, t, p, ev_offsets, bins, H, W, bands, band_rows,
                                                  normalize, out));
                einx_prof_end(ctx, 0, stream);
                ctx->launches++;
                return EINX_OK;
            }
            cudaGetLastError();
        }
    }
    EINX_CUDA(ctx, cudaMemsetAsync(out, 0, sizeof(float) * ncell * B, stream));
    // enough CTAs per window to cover the machine a few times over; windows are ragged, so a
    // CTA grid-strides over its own window only
    int per_window = (ctx->num_sms * 8 + B - 1) / B;
    if (per_window < 1) per_window = 1;
    if (per_window > 1024) per_window = 1024;
    einx_prof_begin(ctx, 0, stream);
    voxel_scatter_kernel<<<dim3(per_window, B), kScatterThreads, 0, stream>>>(x, y, t, p, ev_offsets, bins, H, W, out);
    einx_prof_end(ctx, 0, stream);
    EINX_CHECK_LAUNCH(ctx);
    static const bool two_kernel_norm = getenv("EINX_VOXEL_NORM_2K") != nullptr;
    if (normalize && !two_kernel_norm && (ncell & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 15u) == 0) {
        // fused path: smallest cluster whose slices fit shared memory
        const size_t fixed = align_up(sizeof(NormShared), 16);
        int CS = 0;
        // (two CTAs per SM: with one 225 KB CTA per SM -- MVSEC grids -- the two-kernel path below is faster)
        for (int c = 1; c <= 8; c *= 2) {
            const size_t slice = align_up((ncell + c - 1) / c, 4);
            if (slice * 4 <= (size_t)110 * 1024) { CS = c; break; }
        }
        // small batches: more, smaller CTAs (several per SM) hide the load -> reduce -> store latency chain
        while (CS && CS < 8 && (size_t)B * CS < (size_t)ctx->num_sms * 3) CS *= 2;
        if (CS) {
            const int slice = (int)align_up((ncell + CS - 1) / CS, 4);
            const size_t smem = fixed + (size_t)slice * 4;
            EINX_CUDA(ctx, cudaFuncSetAttribute(voxel_norm_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(B * CS);
            cfg.blockDim = dim3(kNormThreads);
            cfg.dynamicSmemBytes = smem;
            cfg.stream = stream;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = CS;
            attr[0].val.clusterDim.y = 1;
            attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            EINX_CUDA(ctx, cudaLaunchKernelEx(&cfg, voxel_norm_cluster_kernel, out, ncell, slice));
            ctx->launches++;
            return EINX_OK;
        }
    }
    if (normalize) {
        int rc = einx_ws_reserve(ctx, sizeof(double) * 3 * B, stream);
        if (rc) return rc;
        double* stats = (double*)ctx->ws;
        EINX_CUDA(ctx, cudaMemsetAsync(stats, 0, sizeof(double) * 3 * B, stream));
        int chunks = (int)((ncell / 4 + 255) / 256);
        int cap = (ctx->num_sms * 8 + B - 1) / B;
        if (chunks > cap) chunks = cap;
        if (chunks < 1) chunks = 1;
        voxel_stats_kernel<<<dim3(chunks, B), 256, 0, stream>>>(out, ncell, stats);
        EINX_CHECK_LAUNCH(ctx);
        voxel_apply_kernel<<<dim3(chunks, B), 256, 0, stream>>>(out, ncell, stats);
        EINX_CHECK_LAUNCH(ctx);
    }
    return EINX_OK;
}
