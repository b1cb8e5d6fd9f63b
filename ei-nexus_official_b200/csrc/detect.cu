// Detection post-processing: border removal -> iterative NMS fixpoint -> top-k threshold ->
// raster-ordered keypoint rows.  Semantics: reference core/modules/utils/detector_util.py:80-135,
// :138-164, :243-337, :451-484 (see include/einx.h); bit-exact for non-negative, NaN-free maps.
//
// One image per CTA (or per thread-block CLUSTER of row bands when the map does not fit one CTA's
// shared memory); the map is read from HBM once, lives in shared memory through all rounds, and the
// keypoints are written once.
//
// The fixpoint of detector_util.py:286-335 is greedy NMS in (value desc,
// raster asc) order, so the rounds only have to respect two facts: a pixel that is the first-occurrence
// maximum of its CURRENT window is selected for good, and everything else in that window is dead for
// good.  State per pixel: V (value; zeroed once dead, in dense rounds), LM (selected), UB (undecided =
// positive, not selected, no selected pixel in its window).
//
//   dense round   one sweep per thread over 4 columns x a run of rows: the horizontal window maxima of a
//                 row come from three float4 loads, the vertical ones from a register ring of 3-row
//                 partial maxima (9 = 3 x 3 rows), ~9 instructions per pixel, no second plane in shared
//                 memory.  A pixel equal to its window maximum (rare) takes the slow path: exact
//                 first-occurrence test, LM bit, entry in the list of new maxima.  Then every new maximum
//                 zeroes its window in V and clears it in UB (a scatter over list x window rows); the
//                 count of undecided pixels is kept exact from the bits those atomics actually cleared.
//   sparse round  once the undecided pixels fit the worklist: per undecided pixel, look only at the
//                 undecided neighbours (UB bits of the 2R+1 window rows; selected pixels can not be in
//                 the window of an undecided one), then clear the windows of the new maxima in UB.  V is
//                 no longer touched.
//
// The loop ends when no pixel is undecided -- the same fixpoint the reference reaches when its
// batch-wide count of maxima stops changing.  Bands of a cluster keep R-row copies of their neighbours'
// edge rows; every write goes to all copies (distributed shared memory), so there is no halo refresh.
#include <cooperative_groups.h>
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"
#include "detect_common.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int kThreads = 512;
constexpr int kWarps = kThreads / 32;
constexpr int kMaxCluster = 8;

struct NmsParams {
    float* score;
    const uint8_t* mask;
    float* nms_map;
    float* kpts;
    int32_t* counts;
    int B, Hp, Wp, border, kcap;
    int T;        // CTAs (row bands) per image
    int W4;       // float4 column groups per row: ceil(Wp / 4)
    int NCW;      // warps across a row: ceil(W4 / 32)
    int WS;       // row pitch of V in floats: 4 * W4 + 2 * PAD
    int SB;       // row pitch of the bitmaps in words: 4 * NCW + 2 (one zero word on either side)
    int RB;       // rows of the largest band
    int NSEG, SR; // a band is swept as NSEG runs of SR rows per column warp
    int LC;       // entries of the list buffer (new maxima of a dense round / two worklist halves)
    int vec;      // pixels per global access: 4, 2 or 1 (row alignment of `score`)
    int tail_smem;  // the survivor lists of the tail fit the shared-memory scratch (single-CTA images)
    float prob_thresh;
    int use_topk;   // 1: threshold from order statistics rank_lo / rank_hi; 2: top_k >= n (thr_k = 0)
    int rank_lo, rank_hi;
    int scap;       // survivor list capacity per image
    float* surv_val;
    int32_t* surv_idx;
    long long* trace;  // developer aid (EINX_DETECT_TRACE=1): clock64() at phase boundaries of CTA 0
};

#define EINX_TRACE(slot)                                                                                  \
    do {                                                                                                  \
        if (P.trace && blockIdx.x == 0 && threadIdx.x == 0 && (slot) < 126) P.trace[(slot)] = clock64(); \
    } while (0)

struct Shared {
    int und;        // undecided pixels in own rows, exact
    int n_new[2];   // new maxima appended by the dense pass of round k -> n_new[k & 1]
    int wl_n[2];    // worklist lengths (two halves of the list buffer)
    int xcnt[2];
    int warp_scan[kWarps + 1];
    unsigned int hist[256];
    unsigned int sel_prefix, sel_rank, sel_min, sel_cnt;
};

__device__ __forceinline__ float fmax3(float a, float b, float c) { return fmaxf(fmaxf(a, b), c); }  // one FMNMX3

// shared-memory loads from a 32-bit shared-window address: one register + immediate per access in the sweep
__device__ __forceinline__ float4 lds128(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ float lds32(uint32_t a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
    return v;
}

template <int I, int N, typename F>
__device__ __forceinline__ void static_for(F&& f) {
    if constexpr (I < N) {
        f(std::integral_constant<int, I>{});
        static_for<I + 1, N>(f);
    }
}

__device__ __forceinline__ int block_excl_scan(int v, int* scratch, int& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += n;
    }
    __syncthreads();  // protect scratch from the previous call
    if (lane == 31) scratch[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int w = lane < kWarps ? scratch[lane] : 0;
        int winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int n = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += n;
        }
        if (lane < kWarps) scratch[lane] = winc - w;
        if (lane == 31) scratch[kWarps] = winc;
    }
    __syncthreads();
    total = scratch[kWarps];
    return inc - v + scratch[warp];
}

// j-th smallest (0-based) of the positive floats in list[0..n) via 4 radix passes on their bit
// patterns, then the next order statistic; every thread returns the same (a, b).
template <bool GLOBAL>
__device__ __forceinline__ float list_ld(const float* list, int i) {
    return GLOBAL ? __ldcg(list + i) : list[i];  // global lists are written by other CTAs of the cluster: L2 only
}

template <bool GLOBAL>
__device__ void select_two(const float* list, int n, int j, bool need_next, Shared& sh, float& a_out, float& b_out) {
    if (threadIdx.x == 0) { sh.sel_prefix = 0; sh.sel_rank = (unsigned)j; }
    constexpr int kHeld = 8;  // the usual list (a few thousand survivors) is read once and kept in registers
    const bool held = n <= kHeld * kThreads;
    unsigned ev[kHeld];
#pragma unroll
    for (int u = 0; u < kHeld; ++u) {
        const int i = threadIdx.x + u * kThreads;
        ev[u] = (held && i < n) ? __float_as_uint(list_ld<GLOBAL>(list, i)) : 0u;
    }
    unsigned mask = 0;
    for (int shift = 24; shift >= 0; shift -= 8) {
        for (int i = threadIdx.x; i < 256; i += kThreads) sh.hist[i] = 0;
        __syncthreads();
        const unsigned prefix = sh.sel_prefix;
        if (held) {
#pragma unroll
            for (int u = 0; u < kHeld; ++u) {
                const bool in = threadIdx.x + u * kThreads < n && (ev[u] & mask) == prefix;
                // warp-aggregate equal digits (score values share their exponent byte): one atomic per distinct digit
                const unsigned digit = (ev[u] >> shift) & 255u;
                const unsigned peers = __match_any_sync(0xffffffffu, in ? digit : 0x100u);
                if (in && (int)(__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicAdd(&sh.hist[digit], (unsigned)__popc(peers));
            }
        } else {
            for (int i = threadIdx.x; i < n; i += kThreads) {
                const unsigned e = __float_as_uint(list_ld<GLOBAL>(list, i));
                if ((e & mask) == prefix) atomicAdd(&sh.hist[(e >> shift) & 255u], 1u);
            }
        }
        __syncthreads();
        if (threadIdx.x < 32) {
            // warp 0 locates the bin holding rank r: 8 bins per lane, warp prefix, then a short scan
            const unsigned r = sh.sel_rank;
            unsigned c[8], mine = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) { c[k] = sh.hist[threadIdx.x * 8 + k]; mine += c[k]; }
            unsigned inc = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned nn = __shfl_up_sync(0xffffffffu, inc, o);
                if ((int)threadIdx.x >= o) inc += nn;
            }
            const unsigned before = inc - mine;
            const bool here = (before <= r) && (r < inc);  // exactly one lane (r < total count)
            if (here) {
                unsigned cum = before;
                int bin = 0;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    if (cum + c[k] <= r) { cum += c[k]; bin = k + 1; }
                    else break;
                }
                sh.sel_rank = r - cum;
                sh.sel_prefix = prefix | ((unsigned)(threadIdx.x * 8 + bin) << shift);
            }
        }
        mask |= 255u << shift;
        __syncthreads();
    }
    const unsigned abits = sh.sel_prefix;
    float a = __uint_as_float(abits), b = a;
    if (need_next) {
        if (threadIdx.x == 0) { sh.sel_min = 0xffffffffu; sh.sel_cnt = 0; }
        __syncthreads();
        unsigned cnt = 0, mn = 0xffffffffu;
        if (held) {
#pragma unroll
            for (int u = 0; u < kHeld; ++u)
                if (threadIdx.x + u * kThreads < n) {
                    if (ev[u] <= abits) cnt++;
                    else mn = min(mn, ev[u]);
                }
        } else {
            for (int i = threadIdx.x; i < n; i += kThreads) {
                const unsigned e = __float_as_uint(list_ld<GLOBAL>(list, i));
                if (e <= abits) cnt++;
                else mn = min(mn, e);
            }
        }
        cnt = __reduce_add_sync(0xffffffffu, cnt);
        mn = __reduce_min_sync(0xffffffffu, mn);
        if ((threadIdx.x & 31) == 0) {
            atomicAdd(&sh.sel_cnt, cnt);
            atomicMin(&sh.sel_min, mn);
        }
        __syncthreads();
        // the (j+1)-th smallest equals a when a is duplicated past position j
        b = (sh.sel_cnt > (unsigned)j + 1u) ? a : __uint_as_float(sh.sel_min);
    }
    __syncthreads();
    a_out = a;
    b_out = b;
}

// PAD columns of zeros on both sides of a band row; a multiple of 4 so that pixel 0 of every row is
// 16-byte aligned and the sweep can move float4.
template <int R>
struct Geo {
    static constexpr int PAD = (R + 3) / 4 * 4;
    static constexpr int P2 = 2 * R + 1;
};

// Horizontal window maxima of 4 neighbouring pixels.  a[] holds the 4 + 2*PAD values starting PAD
// to the left of the first pixel; o[i] = max a[PAD+i-R .. PAD+i+R].  The values shared by all four
// windows are reduced once, then extended left / right.
template <int R>
__device__ __forceinline__ void hmax4(const float* a, float (&o)[4]) {
    constexpr int PAD = Geo<R>::PAD;
    if constexpr (R >= 2) {
        float common = a[PAD + 3 - R];
#pragma unroll
        for (int k = PAD + 4 - R; k <= PAD + R; ++k) common = fmaxf(common, a[k]);
        const float l1 = a[PAD + 2 - R], l2 = fmaxf(a[PAD + 1 - R], l1), l3 = fmaxf(a[PAD - R], l2);
        const float r1 = a[PAD + R + 1], r2 = fmaxf(r1, a[PAD + R + 2]), r3 = fmaxf(r2, a[PAD + R + 3]);
        o[0] = fmaxf(common, l3);
        o[1] = fmax3(common, l2, r1);
        o[2] = fmax3(common, l1, r2);
        o[3] = fmaxf(common, r3);
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float m = a[PAD + i];
#pragma unroll
            for (int d = 1; d <= R; ++d) m = fmax3(m, a[PAD + i - d], a[PAD + i + d]);
            o[i] = m;
        }
    }
}

// The copies of band-local row l (image row ys - R + l) that live in the neighbouring bands of the cluster.
struct Copies {
    int rank, T, nrows, nprev;
    // local row index of the same image row in the band above / below, or -1 when that band holds no copy
    __device__ __forceinline__ int up(int l, int R2) const { return (rank > 0 && l < R2) ? l + nprev : -1; }
    __device__ __forceinline__ int down(int l) const { return (rank < T - 1 && l >= nrows) ? l - nrows : -1; }
};

template <int R, bool MULTI>
__global__ void __launch_bounds__(kThreads, 1) nms_kernel(const NmsParams P) {
    constexpr int PAD = Geo<R>::PAD;
    constexpr int P2 = Geo<R>::P2;
    constexpr unsigned kWinMask = (1u << P2) - 1u;
    cg::cluster_group cluster = cg::this_cluster();
    const int T = MULTI ? P.T : 1;
    const int rank = MULTI ? (int)cluster.block_rank() : 0;
    const int b = blockIdx.x / T;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int WS = P.WS, SB = P.SB, Hp = P.Hp, Wp = P.Wp, W4 = P.W4, NCW = P.NCW;

    // balanced row bands
    const int base_rows = Hp / T, rem = Hp % T;
    const int nrows = base_rows + (rank < rem ? 1 : 0);
    const int ys = rank * base_rows + min(rank, rem);
    Copies cp;
    cp.rank = rank; cp.T = T; cp.nrows = nrows;
    cp.nprev = base_rows + ((rank - 1) < rem ? 1 : 0);  // rows of the band above
    const int L = nrows + 2 * R;                        // local rows: R halo + own + R halo

    extern __shared__ __align__(16) unsigned char smem_raw[];
    Shared& sh = *reinterpret_cast<Shared*>(smem_raw);
    // Local row l of every array is image row ys - R + l: own rows are l in [R, R + nrows).
    const int lrows = P.RB + 2 * R;
    size_t so = align_up(sizeof(Shared), 16);
    float* const V = reinterpret_cast<float*>(smem_raw + so);
    so += sizeof(float) * (size_t)lrows * WS;
    uint32_t* const LM = reinterpret_cast<uint32_t*>(smem_raw + so);
    so += sizeof(uint32_t) * (size_t)lrows * SB;
    uint32_t* const UB = reinterpret_cast<uint32_t*>(smem_raw + so);
    so += sizeof(uint32_t) * (size_t)lrows * SB;
    unsigned int* const list = reinterpret_cast<unsigned int*>(smem_raw + so);
    const uint32_t V_s = (uint32_t)__cvta_generic_to_shared(V);
    // neighbours' arrays (same offsets in their shared memory)
    float *Vup = nullptr, *Vdn = nullptr;
    uint32_t *UBup = nullptr, *UBdn = nullptr;
    int *und_up = nullptr, *und_dn = nullptr;
    if (MULTI) {
        if (rank > 0) {
            Vup = cluster.map_shared_rank(V, rank - 1);
            UBup = cluster.map_shared_rank(UB, rank - 1);
            und_up = cluster.map_shared_rank(&sh.und, rank - 1);
        }
        if (rank < T - 1) {
            Vdn = cluster.map_shared_rank(V, rank + 1);
            UBdn = cluster.map_shared_rank(UB, rank + 1);
            und_dn = cluster.map_shared_rank(&sh.und, rank + 1);
        }
    }

    // ---- load the band: border + mask zeroing (in place on `score`), zero padding, UB bits -------- //
    for (int i = tid; i < lrows * SB; i += kThreads) { LM[i] = 0u; UB[i] = 0u; }
    if (tid == 0) { sh.und = 0; sh.n_new[0] = sh.n_new[1] = 0; sh.wl_n[0] = sh.wl_n[1] = 0; sh.xcnt[0] = sh.xcnt[1] = 0; }
    // zero pads of every row: PAD floats on the left, PAD on the right
    if constexpr (PAD > 0) {
        constexpr int PF = PAD / 2;  // float4 per row
        for (int i = tid; i < L * PF; i += kThreads) {
            const int l = i / PF, k = i - l * PF;
            float* rowp = V + (size_t)l * WS;
            const int off = k < PAD / 4 ? 4 * k : 4 * W4 + PAD + 4 * (k - PAD / 4);
            *reinterpret_cast<float4*>(rowp + off) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    __syncthreads();
    {
        float* simg = P.score + (size_t)b * Hp * Wp;
        const uint8_t* mimg = P.mask ? P.mask + (size_t)b * Hp * Wp : nullptr;
        const int bd = P.border;
        const int vec = P.vec;                       // pixels per lane and access
        const int gl = 32 / vec;                     // lanes per bitmap word
        const int tasks_per_row = (4 * W4 + 32 * vec - 1) / (32 * vec);
        int cnt = 0;
        for (int t = warp; t < L * tasks_per_row; t += kWarps) {
            const int l = t / tasks_per_row, cw = t - l * tasks_per_row;
            const int y = ys - R + l;
            const bool own = l >= R && l < R + nrows;
            const bool rowin = y >= 0 && y < Hp;
            const bool rowkill = (y < bd) | (y >= Hp - bd);
            const int x = (cw * 32 + lane) * vec;
            float* srow = simg + (size_t)(rowin ? y : 0) * Wp;
            const uint8_t* mrow = mimg ? mimg + (size_t)(rowin ? y : 0) * Wp : nullptr;
            float e[4] = {0.f, 0.f, 0.f, 0.f};
            const bool in = rowin && x < Wp;
            if (in) {
                if (vec == 4) {
                    const float4 q = *reinterpret_cast<const float4*>(srow + x);
                    e[0] = q.x; e[1] = q.y; e[2] = q.z; e[3] = q.w;
                } else if (vec == 2) {
                    const float2 q = *reinterpret_cast<const float2*>(srow + x);
                    e[0] = q.x; e[1] = q.y;
                } else {
                    e[0] = srow[x];
                }
                bool changed = false;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (j < vec) {
                        const int xx = x + j;
                        bool kill = rowkill | (xx < bd) | (xx >= Wp - bd);
                        if (mrow) kill |= (mrow[xx] == 0);
                        if (kill) {
                            changed |= (e[j] != 0.0f);
                            e[j] = 0.0f;
                        }
                    }
                }
                if (changed && own) {  // the band that owns the row writes the zeroed frame / mask back
                    if (vec == 4) *reinterpret_cast<float4*>(srow + x) = make_float4(e[0], e[1], e[2], e[3]);
                    else if (vec == 2) *reinterpret_cast<float2*>(srow + x) = make_float2(e[0], e[1]);
                    else srow[x] = e[0];
                }
            }
            unsigned bits = 0;
            if (x < 4 * W4) {
                float* vrow = V + (size_t)l * WS + PAD + x;
                if (vec == 4) *reinterpret_cast<float4*>(vrow) = make_float4(e[0], e[1], e[2], e[3]);
                else if (vec == 2) *reinterpret_cast<float2*>(vrow) = make_float2(e[0], e[1]);
                else vrow[0] = e[0];
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (j < vec && e[j] > 0.0f) bits |= 1u << j;
            }
            if (own) cnt += __popc(bits);
            unsigned word = bits << (vec * (lane & (gl - 1)));
            for (int o = 1; o < gl; o <<= 1) word |= __shfl_xor_sync(0xffffffffu, word, o);
            if ((lane & (gl - 1)) == 0 && x < 4 * W4) UB[(size_t)l * SB + 1 + (x >> 5)] = word;
        }
        cnt = __reduce_add_sync(0xffffffffu, cnt);
        if (lane == 0 && cnt) atomicAdd(&sh.und, cnt);
    }
    EINX_TRACE(0);

    // ---- NMS rounds ------------------------------------------------------------------------ //
    auto sync_all = [&]() {
        if (MULTI) cluster.sync();
        else __syncthreads();
    };
    // clear the window of a (new) maximum at column x in local row l2 of UB, in every copy of that row;
    // returns the undecided bits this call cleared in the owner's copy, by owner (self / up / down)
    auto clear_window = [&](int l2, int x, int& c_self, int& c_up, int& c_dn) {
        const int xl = x - R + 32;
        const int wi = xl >> 5, shf = xl & 31;
        const unsigned long long m64 = (unsigned long long)kWinMask << shf;
        const uint32_t lo = (uint32_t)m64, hi = (uint32_t)(m64 >> 32);
        const bool mine = l2 >= R && l2 < R + nrows;
        uint32_t* u = UB + (size_t)l2 * SB + wi;
        uint32_t o0 = atomicAnd(u, ~lo), o1 = 0;
        if (hi) o1 = atomicAnd(u + 1, ~hi);
        if (mine) c_self += __popc(o0 & lo) + __popc(o1 & hi);
        if (MULTI) {
            const int lu = cp.up(l2, 2 * R), ld = cp.down(l2);
            if (lu >= 0) {
                uint32_t* ur = UBup + (size_t)lu * SB + wi;
                o0 = atomicAnd(ur, ~lo);
                o1 = hi ? atomicAnd(ur + 1, ~hi) : 0u;
                if (l2 < R) c_up += __popc(o0 & lo) + __popc(o1 & hi);  // rows [0, R) belong to the band above
            }
            if (ld >= 0) {
                uint32_t* ur = UBdn + (size_t)ld * SB + wi;
                o0 = atomicAnd(ur, ~lo);
                o1 = hi ? atomicAnd(ur + 1, ~hi) : 0u;
                if (l2 >= R + nrows) c_dn += __popc(o0 & lo) + __popc(o1 & hi);
            }
        }
    };
    auto settle_counts = [&](int c_self, int c_up, int c_dn) {
        c_self = __reduce_add_sync(0xffffffffu, c_self);
        if (lane == 0 && c_self) atomicSub(&sh.und, c_self);
        if (MULTI) {
            c_up = __reduce_add_sync(0xffffffffu, c_up);
            c_dn = __reduce_add_sync(0xffffffffu, c_dn);
            if (lane == 0 && c_up) atomicSub(und_up, c_up);
            if (lane == 0 && c_dn) atomicSub(und_dn, c_dn);
        }
    };

    if constexpr (R > 0) {
        const int half = P.LC / 2;  // worklist capacity (two halves of the list buffer)
        bool sparse = false;
        int wl_cur = 0;
        int trace_slot = 1;
        for (int round = 0;; ++round) {
            sync_all();  // every band's und / V / UB is final for this round
            trace_slot = round < 19 ? 1 + 6 * round : 126;
            EINX_TRACE(trace_slot); ++trace_slot;
            int tot = sh.und, mx = tot;
            if (MULTI) {
                tot = 0; mx = 0;
                for (int r = 0; r < T; ++r) {
                    const int u = *cluster.map_shared_rank(&sh.und, r);
                    tot += u;
                    mx = max(mx, u);
                }
            }
            if (tot == 0 || round > Hp + Wp) break;  // (the round bound only guards against a corrupted count: a
                                                     // round always decides at least one pixel)
            if (!sparse && mx <= half) {
                // build the worklist of undecided pixels of the own rows, raster order
                sparse = true;
                wl_cur = 0;
                const int nwords = nrows * (SB - 2);
                int run = 0;
                for (int base = 0; base < nwords; base += kThreads) {
                    const int wi = base + tid;
                    int lr = 0, s = 0;
                    uint32_t w = 0;
                    if (wi < nwords) {
                        lr = wi / (SB - 2);
                        s = wi - lr * (SB - 2);
                        w = UB[(size_t)(lr + R) * SB + 1 + s];
                    }
                    int totw;
                    int pos = run + block_excl_scan(__popc(w), sh.warp_scan, totw);
                    while (w) {
                        const int bit = __ffs(w) - 1;
                        w &= w - 1;
                        list[pos++] = ((unsigned)(lr + R) << 16) | (unsigned)(32 * s + bit);
                    }
                    run += totw;
                }
                if (tid == 0) { sh.wl_n[0] = run; sh.wl_n[1] = 0; }
                __syncthreads();
            }
            if (!sparse) {
                // ---- dense pass: sweep, collect the new maxima ---------------------------------- //
                int* const n_new = &sh.n_new[round & 1];
                const int nunits = NCW * P.NSEG;
                for (int u = warp; u < nunits; u += kWarps) {
                    const int seg = u / NCW, cw = u - seg * NCW;
                    const int a = R + seg * P.SR;                   // local output rows [a, e)
                    const int e = min(a + P.SR, R + nrows);
                    if (a >= e) continue;
                    const int g = cw * 32 + lane;
                    const bool active = g < W4;
                    // shared-window byte address of the leftmost float4 the thread reads in row a - R
                    uint32_t rp = V_s + 4u * (uint32_t)((a - R) * WS + 4 * (active ? g : W4 - 1));
                    const uint32_t row_bytes = 4u * (uint32_t)WS;
                    float h[P2][4], p3[P2][4];
#pragma unroll
                    for (int j = 0; j < P2; ++j)
#pragma unroll
                        for (int c = 0; c < 4; ++c) { h[j][c] = 0.0f; p3[j][c] = 0.0f; }
                    // ring slot of a row = (row - (a - R)) mod P2.  Ingesting a row: its horizontal window maxima
                    // (h) and the 3-row partial maximum that ends with it (p3 of the row two above).
                    auto ingest = [&](auto slot) {
                        constexpr int j = decltype(slot)::value;
                        constexpr int j1 = (j + P2 - 1) % P2, j2 = (j + P2 - 2) % P2;
                        float av[4 + 2 * PAD];
#pragma unroll
                        for (int k = 0; k < 1 + PAD / 2; ++k) {
                            const float4 q = lds128(rp + 16u * k);
                            av[4 * k] = q.x; av[4 * k + 1] = q.y; av[4 * k + 2] = q.z; av[4 * k + 3] = q.w;
                        }
                        hmax4<R>(av, h[j]);
#pragma unroll
                        for (int c = 0; c < 4; ++c) p3[j2][c] = fmax3(h[j2][c], h[j1][c], h[j][c]);
                    };
                    // Output row y = (row just ingested in slot j) - R: all of its window rows are in the ring.
                    auto emit = [&](auto slot, int y) {
                        constexpr int j = decltype(slot)::value;
                        constexpr int j2 = (j + P2 - 2) % P2;
                        float M[4];
#pragma unroll
                        for (int c = 0; c < 4; ++c) M[c] = p3[j2][c];  // rows y+R-2 .. y+R
#pragma unroll
                        for (int i = 0; 3 * i < P2 - 3; ++i) {
                            const int ji = (j + 3 * i + 1) % P2;       // rows y-R+3i .. y-R+3i+2
#pragma unroll
                            for (int c = 0; c < 4; ++c) M[c] = fmaxf(M[c], p3[ji][c]);
                        }
                        const uint32_t cpa = rp - (uint32_t)R * row_bytes + 4u * PAD;  // the thread's 4 pixels of row y
                        const float4 cv = lds128(cpa);
                        const float vc[4] = {cv.x, cv.y, cv.z, cv.w};
                        bool any;
                        if constexpr (R >= 3) {
                            // a zero pixel equal to its (all-zero) window implies the thread's other pixels are zero too
                            any = (vc[0] == M[0]) | (vc[1] == M[1]) | (vc[2] == M[2]) | (vc[3] == M[3]);
                            any = any && fmaxf(fmaxf(vc[0], vc[1]), fmaxf(vc[2], vc[3])) > 0.0f;
                        } else {
                            any = false;
#pragma unroll
                            for (int c = 0; c < 4; ++c) any |= (vc[c] > 0.0f && vc[c] == M[c]);
                        }
                        if (any && active) {
                            // slow path (about one pixel in (2R+1)^2): exact first-occurrence test of
                            // detector_util.py:298-308 -- no equal value earlier in raster order
#pragma unroll
                            for (int c = 0; c < 4; ++c) {
                                const float v = vc[c];
                                if (!(v > 0.0f && v == M[c])) continue;
                                const int x = 4 * g + c;
                                const uint32_t px = cpa + 4u * c;
                                bool tie = false;
#pragma unroll
                                for (int dy = 1; dy <= R; ++dy)
#pragma unroll
                                    for (int dx = -R; dx <= R; ++dx) tie |= (lds32(px - dy * row_bytes + 4 * dx) == v);
#pragma unroll
                                for (int d = 1; d <= R; ++d) tie |= (lds32(px - 4 * d) == v);
                                if (!tie) {
                                    const uint32_t bit = 1u << (x & 31);
                                    const uint32_t old = atomicOr(&LM[(size_t)y * SB + 1 + (x >> 5)], bit);
                                    if (!(old & bit)) {
                                        const int pos = atomicAdd(n_new, 1);
                                        if (pos < P.LC) list[pos] = ((unsigned)y << 16) | (unsigned)x;
                                    }
                                }
                            }
                        }
                    };
                    // prologue: rows a-R .. a+R-1 fill the ring (slots 0 .. 2R-1), nothing to emit yet
                    static_for<0, P2 - 1>([&](auto slot) {
                        ingest(slot);
                        rp += row_bytes;
                    });
                    // steady state: ingest row y + R (slot (2R + k) mod P2), emit row y
                    for (int y0 = a; y0 < e; y0 += P2) {
                        static_for<0, P2>([&](auto kk) {
                            constexpr int k = decltype(kk)::value;
                            if (y0 + k < e) {
                                using Slot = std::integral_constant<int, (P2 - 1 + k) % P2>;
                                ingest(Slot{});
                                emit(Slot{}, y0 + k);
                                rp += row_bytes;
                            }
                        });
                    }
                }
                EINX_TRACE(trace_slot); ++trace_slot;
                sync_all();
                EINX_TRACE(trace_slot); ++trace_slot;
                // ---- scatter: every new maximum kills its window (V and UB, all copies) ---------- //
                {
                    const int nn = min(*n_new, P.LC);
                    if (tid == 0) sh.n_new[(round + 1) & 1] = 0;
                    int c_self = 0, c_up = 0, c_dn = 0;
                    for (int it = tid; it < nn * P2; it += kThreads) {
                        const int en = it / P2, dyi = it - en * P2;
                        const unsigned ent = list[en];
                        const int l = (int)(ent >> 16), x = (int)(ent & 0xffffu);
                        const int l2 = l + dyi - R;
                        float* row = V + (size_t)l2 * WS + PAD + x;
                        const int lu = MULTI ? cp.up(l2, 2 * R) : -1, ld = MULTI ? cp.down(l2) : -1;
                        float* rowu = lu >= 0 ? Vup + (size_t)lu * WS + PAD + x : nullptr;
                        float* rowd = ld >= 0 ? Vdn + (size_t)ld * WS + PAD + x : nullptr;
#pragma unroll
                        for (int dx = -R; dx <= R; ++dx) {
                            if (dx == 0 && dyi == R) continue;
                            row[dx] = 0.0f;
                            if (MULTI) {
                                if (rowu) rowu[dx] = 0.0f;
                                if (rowd) rowd[dx] = 0.0f;
                            }
                        }
                        clear_window(l2, x, c_self, c_up, c_dn);
                    }
                    settle_counts(c_self, c_up, c_dn);
                }
                EINX_TRACE(trace_slot); ++trace_slot;
            } else {
                // ---- sparse round, phase 1: an undecided pixel looks at its undecided neighbours ---- //
                unsigned int* const wl = list + wl_cur * half;
                const int n = sh.wl_n[wl_cur];
                for (int en = tid; en < n; en += kThreads) {
                    const unsigned ent = wl[en];
                    const int l = (int)(ent >> 16), x = (int)(ent & 0xffffu);
                    const float* cpx = V + (size_t)l * WS + PAD + x;
                    const float v = cpx[0];
                    const int xl = x - R + 32;
                    const int wi = xl >> 5, shf = xl & 31;
                    bool beaten = false;
#pragma unroll
                    for (int dy = -R; dy <= R; ++dy) {
                        const uint32_t* u = UB + (size_t)(l + dy) * SB + wi;
                        const unsigned long long both = (unsigned long long)u[0] | ((unsigned long long)u[1] << 32);
                        uint32_t f = (uint32_t)(both >> shf) & kWinMask;
                        if (dy == 0) f &= ~(1u << R);
                        while (f) {
                            const int bit = __ffs(f) - 1;
                            f &= f - 1;
                            const float w = cpx[dy * WS + bit - R];
                            // raster-earlier neighbours win ties (first-occurrence argmax)
                            const bool earlier = dy < 0 || (dy == 0 && bit < R);
                            if (w > v || (earlier && w == v)) { beaten = true; f = 0; }
                        }
                    }
                    if (!beaten) {
                        atomicOr(&LM[(size_t)l * SB + 1 + (x >> 5)], 1u << (x & 31));
                        wl[en] = ent | 0x80000000u;  // a new maximum (local rows < 32768)
                    }
                }
                EINX_TRACE(trace_slot); ++trace_slot;
                sync_all();
                EINX_TRACE(trace_slot); ++trace_slot;
                // ---- phase 2: the new maxima clear their windows in UB (all copies) ---------------- //
                {
                    if (tid == 0) sh.wl_n[wl_cur ^ 1] = 0;
                    int c_self = 0, c_up = 0, c_dn = 0;
                    for (int en = tid; en < n; en += kThreads) {
                        const unsigned ent = wl[en];
                        if (!(ent & 0x80000000u)) continue;
                        const int l = (int)((ent >> 16) & 0x7fffu), x = (int)(ent & 0xffffu);
#pragma unroll
                        for (int dy = -R; dy <= R; ++dy) clear_window(l + dy, x, c_self, c_up, c_dn);
                    }
                    settle_counts(c_self, c_up, c_dn);
                }
                sync_all();
                EINX_TRACE(trace_slot); ++trace_slot;
                // ---- phase 3: keep the still-undecided entries ---------------------------------- //
                {
                    unsigned int* const wn = list + (wl_cur ^ 1) * half;
                    for (int base = 0; base < n; base += kThreads) {
                        const int en = base + tid;
                        unsigned ent = 0;
                        bool keep = false;
                        if (en < n) {
                            ent = wl[en];
                            if (!(ent & 0x80000000u)) {
                                const int l = (int)(ent >> 16), x = (int)(ent & 0xffffu);
                                keep = (UB[(size_t)l * SB + 1 + (x >> 5)] >> (x & 31)) & 1u;
                            }
                        }
                        const unsigned kb = __ballot_sync(0xffffffffu, keep);
                        int pos = 0;
                        if (lane == 0 && kb) pos = atomicAdd(&sh.wl_n[wl_cur ^ 1], __popc(kb));
                        pos = __shfl_sync(0xffffffffu, pos, 0);
                        if (keep) wn[pos + __popc(kb & ((1u << lane) - 1u))] = ent;
                    }
                    wl_cur ^= 1;
                }
            }
        }
    } else {
        // no NMS: the survivors are simply the positive pixels
        __syncthreads();
        for (int i = tid; i < nrows * SB; i += kThreads) LM[(size_t)R * SB + i] = UB[(size_t)R * SB + i];
        __syncthreads();
    }

    EINX_TRACE(120);
    // ---- survivors -> ordered per-image list ------------------------------------------------- //
    // At the fixpoint the selected pixels (LM bits of the own rows) are exactly the survivors.  The lists
    // live in the shared-memory scratch (UB + list buffer, both dead now) for single-CTA images, in the
    // global workspace for clusters.
    float* slist;
    int32_t* sidx;
    const bool smem_lists = !MULTI && P.tail_smem;
    if (smem_lists) {
        slist = reinterpret_cast<float*>(UB);
        sidx = reinterpret_cast<int32_t*>(UB) + P.scap;
    } else {
        slist = P.surv_val + (size_t)b * P.scap;
        sidx = P.surv_idx + (size_t)b * P.scap;
    }
    const int SW = SB - 2;
    const int nwords = nrows * SW;
    int own = 0, offset = 0, total = 0;
    {
        int c = 0;
        for (int wi = tid; wi < nwords; wi += kThreads) {
            const int lr = wi / SW, s = wi - lr * SW;
            c += __popc(LM[(size_t)(lr + R) * SB + 1 + s]);
        }
        c = __reduce_add_sync(0xffffffffu, c);
        if (lane == 0 && c) atomicAdd(&sh.xcnt[0], c);
    }
    sync_all();
    if (MULTI) {
        for (int r = 0; r < T; ++r) {
            const int c = *cluster.map_shared_rank(&sh.xcnt[0], r);
            if (r < rank) offset += c;
            total += c;
        }
    } else {
        total = sh.xcnt[0];
    }
    own = sh.xcnt[0];
    {
        // the bitmap words are read before the lists (which may alias UB, never LM) are written
        int run = offset;
        for (int base = 0; base < nwords; base += kThreads) {
            const int wi = base + tid;
            int lr = 0, s = 0;
            uint32_t w = 0;
            if (wi < nwords) {
                lr = wi / SW;
                s = wi - lr * SW;
                w = LM[(size_t)(lr + R) * SB + 1 + s];
            }
            int tot;
            int pos = run + block_excl_scan(__popc(w), sh.warp_scan, tot);
            while (w) {
                const int bit = __ffs(w) - 1;
                w &= w - 1;
                const int x = 32 * s + bit;
                if (pos < P.scap) {
                    slist[pos] = V[(size_t)(lr + R) * WS + PAD + x];
                    sidx[pos] = (ys + lr) * Wp + x;
                }
                ++pos;
            }
            run += tot;
        }
    }
    if (MULTI) __threadfence();
    EINX_TRACE(121);
    sync_all();  // the whole image's list is visible
    EINX_TRACE(122);

    // ---- threshold (detector_util.py:108-133), computed redundantly by every CTA ------------ //
    float thr = P.prob_thresh;
    if (P.use_topk == 2) {
        thr = fminf(0.0f, P.prob_thresh);
    } else if (P.use_topk == 1) {
        const int n = Hp * Wp;
        const int zeros = n - total;  // ascending order: the zeros come first
        float a = 0.0f, bq = 0.0f;
        if (P.rank_hi >= zeros) {
            if (P.rank_lo >= zeros) {
                if (smem_lists) select_two<false>(slist, total, P.rank_lo - zeros, P.rank_hi != P.rank_lo, sh, a, bq);
                else select_two<true>(slist, total, P.rank_lo - zeros, P.rank_hi != P.rank_lo, sh, a, bq);
            } else {  // lo falls on a zero, hi on the smallest survivor
                float dummy;
                if (smem_lists) select_two<false>(slist, total, 0, false, sh, bq, dummy);
                else select_two<true>(slist, total, 0, false, sh, bq, dummy);
            }
        }
        // torch.lerp(a, b, 0.5) takes the `b - (b - a) * (1 - w)` branch
        const float thr_k = __fsub_rn(bq, __fmul_rn(__fsub_rn(bq, a), 0.5f));
        thr = fminf(thr_k, P.prob_thresh);
    }

    EINX_TRACE(123);
    // ---- keypoint rows in raster order + optional dense map ---------------------------------- //
    {
        int c = 0;
        for (int i = tid; i < own; i += kThreads) c += ((smem_lists ? slist[offset + i] : __ldcg(slist + offset + i)) > thr) ? 1 : 0;
        c = __reduce_add_sync(0xffffffffu, c);
        if (lane == 0 && c) atomicAdd(&sh.xcnt[1], c);
    }
    sync_all();
    int koff = 0, ktotal = 0;
    if (MULTI) {
        for (int r = 0; r < T; ++r) {
            const int c = *cluster.map_shared_rank(&sh.xcnt[1], r);
            if (r < rank) koff += c;
            ktotal += c;
        }
    } else {
        ktotal = sh.xcnt[1];
    }
    if (rank == 0 && tid == 0) P.counts[b] = ktotal;
    {
        float* krows = P.kpts + (size_t)b * P.kcap * 3;
        int run = koff;
        for (int base = 0; base < own; base += kThreads) {
            const int i = base + tid;
            float v = 0.0f;
            int idx = 0;
            bool keep = false;
            if (i < own) {
                v = smem_lists ? slist[offset + i] : __ldcg(slist + offset + i);
                idx = smem_lists ? sidx[offset + i] : __ldcg(sidx + offset + i);
                keep = v > thr;
            }
            int tot;
            const int pos = run + block_excl_scan(keep ? 1 : 0, sh.warp_scan, tot);
            if (keep && pos < P.kcap) {
                const int y = idx / Wp, x = idx - y * Wp;
                krows[(size_t)pos * 3 + 0] = (float)y + 0.5f;
                krows[(size_t)pos * 3 + 1] = (float)x + 0.5f;
                krows[(size_t)pos * 3 + 2] = v;
            }
            run += tot;
        }
    }
    if (P.nms_map) {
        // selected AND above the threshold; V still holds dead values from the sparse rounds, LM decides
        float* out = P.nms_map + (size_t)b * Hp * Wp;
        if (P.vec == 4) {
            for (int e = tid; e < nrows * W4; e += kThreads) {
                const int lr = e / W4, g = e - lr * W4;
                const int x = 4 * g;
                const uint32_t nib = (LM[(size_t)(lr + R) * SB + 1 + (x >> 5)] >> (x & 31)) & 0xfu;
                float4 v = *reinterpret_cast<const float4*>(V + (size_t)(lr + R) * WS + PAD + x);
                v.x = ((nib & 1u) && v.x > thr) ? v.x : 0.0f;
                v.y = ((nib & 2u) && v.y > thr) ? v.y : 0.0f;
                v.z = ((nib & 4u) && v.z > thr) ? v.z : 0.0f;
                v.w = ((nib & 8u) && v.w > thr) ? v.w : 0.0f;
                *reinterpret_cast<float4*>(out + (size_t)(ys + lr) * Wp + x) = v;
            }
        } else {
            for (int e = tid; e < nrows * Wp; e += kThreads) {
                const int lr = e / Wp, x = e - lr * Wp;
                const bool sel = (LM[(size_t)(lr + R) * SB + 1 + (x >> 5)] >> (x & 31)) & 1u;
                const float v = V[(size_t)(lr + R) * WS + PAD + x];
                out[(size_t)(ys + lr) * Wp + x] = (sel && v > thr) ? v : 0.0f;
            }
        }
    }
    EINX_TRACE(124);
    if (MULTI) cluster.sync();  // nobody leaves while a neighbour may still read its shared memory
    EINX_TRACE(125);
}

template <int R, bool MULTI>
int launch_nms(einx_ctx* ctx, const NmsParams& P, size_t smem, cudaStream_t stream) {
    auto kern = nms_kernel<R, MULTI>;
    EINX_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(P.B * P.T);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = P.T;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    einx_prof_begin(ctx, 1, stream);
    cudaError_t le = cudaLaunchKernelEx(&cfg, kern, P);
    einx_prof_end(ctx, 1, stream);
    EINX_CUDA(ctx, le);
    ctx->launches++;
    if (P.trace) {  // developer aid: print the phase timeline of CTA 0 (synchronises)
        long long h[128];
        cudaStreamSynchronize(stream);
        cudaMemcpy(h, P.trace, sizeof(h), cudaMemcpyDeviceToHost);
        fprintf(stderr, "[einx_detect trace] T=%d smem=%zu:", P.T, smem);
        long long prev = h[0];
        for (int i = 0; i < 128; ++i)
            if (h[i]) { fprintf(stderr, " %d:+%lld", i, h[i] - prev); prev = h[i]; }
        fprintf(stderr, "\n");
        cudaMemset(P.trace, 0, sizeof(h));
    }
    return EINX_OK;
}

template <bool MULTI>
int dispatch_radius(einx_ctx* ctx, int R, const NmsParams& P, size_t smem, cudaStream_t stream) {
    switch (R) {
        case 0: return launch_nms<0, MULTI>(ctx, P, smem, stream);
        case 1: return launch_nms<1, MULTI>(ctx, P, smem, stream);
        case 2: return launch_nms<2, MULTI>(ctx, P, smem, stream);
        case 3: return launch_nms<3, MULTI>(ctx, P, smem, stream);
        case 4: return launch_nms<4, MULTI>(ctx, P, smem, stream);
        case 5: return launch_nms<5, MULTI>(ctx, P, smem, stream);
        case 6: return launch_nms<6, MULTI>(ctx, P, smem, stream);
        case 7: return launch_nms<7, MULTI>(ctx, P, smem, stream);
        case 8: return launch_nms<8, MULTI>(ctx, P, smem, stream);
    }
    return einx_fail(ctx, EINX_ERR_UNSUPPORTED, "einx_detect: nms_radius %d not in [0, 8]", R);
}

}  // namespace

// fp32 emulation of q = (n-k)/n, rank = q*(n-1) (detector_util.py:113-124; torch divides an
// int64 tensor by a Python int in fp32 and quantile scales q in the input dtype)
void einx_topk_ranks(int n, int k, int* lo, int* hi) {
    volatile float q = (float)(n - k) / (float)n;
    volatile float rank = q * (float)(n - 1);
    *lo = (int)floorf(rank);
    *hi = (int)ceilf(rank);
}

extern "C" int einx_detect(einx_ctx* ctx, float* score, const uint8_t* mask, int B, int Hp, int Wp, int nms_radius,
                           int border, float prob_thresh, int top_k, float* nms_map, float* kpts, int kcap,
                           int32_t* counts, einx_stream stream_) {
    if (!ctx) return EINX_ERR_INVALID;
    if (B < 0 || Hp <= 0 || Wp <= 0 || nms_radius < 0 || border < 0 || kcap < 0)
        return einx_fail(ctx, EINX_ERR_INVALID, "einx_detect: bad argument B=%d Hp=%d Wp=%d r=%d border=%d kcap=%d", B,
                         Hp, Wp, nms_radius, border, kcap);
    if (B == 0) return EINX_OK;
    if (!score || !kpts || !counts) return einx_fail(ctx, EINX_ERR_INVALID, "einx_detect: NULL pointer argument");
    if ((long long)Hp * Wp > (1ll << 30)) return einx_fail(ctx, EINX_ERR_UNSUPPORTED, "einx_detect: map too large");
    if (nms_radius > 8) return einx_fail(ctx, EINX_ERR_UNSUPPORTED, "einx_detect: nms_radius %d not in [0, 8]", nms_radius);
    DeviceGuard guard(ctx->device);
    cudaStream_t stream = (cudaStream_t)stream_;
    const int R = nms_radius;

    NmsParams P = {};
    P.score = score; P.mask = mask; P.nms_map = nms_map; P.kpts = kpts; P.counts = counts;
    P.B = B; P.Hp = Hp; P.Wp = Wp; P.border = border; P.kcap = kcap;
    const int PAD = (R + 3) / 4 * 4;
    P.W4 = (Wp + 3) / 4;
    P.NCW = (P.W4 + 31) / 32;
    P.WS = 4 * P.W4 + 2 * PAD;
    P.SB = 4 * P.NCW + 2;
    P.vec = (Wp % 4 == 0 && (uintptr_t)score % 16 == 0 && (!nms_map || (uintptr_t)nms_map % 16 == 0)) ? 4
            : (Wp % 2 == 0 && (uintptr_t)score % 8 == 0) ? 2 : 1;
    P.prob_thresh = prob_thresh;
    const int n = Hp * Wp;
    if (top_k > 0) {
        if (top_k >= n) P.use_topk = 2;
        else { P.use_topk = 1; einx_topk_ranks(n, top_k, &P.rank_lo, &P.rank_hi); }
    }
    P.scap = R == 0 ? n : ((Hp + R) / (R + 1)) * ((Wp + R) / (R + 1));

    // Bands per image: the smallest cluster whose band (values, two bitmaps, list buffer) fits one CTA's shared
    // memory; then wider while even two such launches side by side (the two sides of a pair run on concurrent
    // streams) leave SMs idle.  Bands of a cluster are at least 2R rows, so a row has at most two copies.
    const size_t fixed = align_up(sizeof(Shared), 16);
    const size_t budget = (size_t)ctx->max_smem_optin;
    auto band_cap = [&](int rb) {  // maxima a band can hold (the new-maxima list of a dense round) -- R > 0
        return R == 0 ? 0 : ((rb + R) / (R + 1) + 1) * ((Wp + R) / (R + 1));
    };
    // list buffer: at least the new maxima a dense round can produce; beyond that, a larger buffer means an
    // earlier switch to sparse rounds (half of it is the worklist capacity)
    auto smem_for = [&](int rb, int lc) {
        return fixed + (size_t)(rb + 2 * R) * ((size_t)P.WS * 4 + (size_t)P.SB * 8) + (size_t)lc * 4;
    };
    auto list_entries = [&](int rb) {  // 0: the band does not fit
        for (int want = 4096; want >= 1024; want >>= 1) {
            int lc = band_cap(rb) > want ? band_cap(rb) : want;
            lc += lc & 1;
            if (smem_for(rb, lc) <= budget) return lc;
        }
        return 0;
    };
    const char* force_env = getenv("EINX_DETECT_CLUSTER");  // testing aid: bands per image (read per call)
    const int force_t = force_env ? atoi(force_env) : 0;
    int T = 0;
    for (int t = 1; t <= kMaxCluster; ++t) {
        if (force_t > 0 && t != force_t) continue;
        const int rb = (Hp + t - 1) / t;
        if (t > 1 && Hp / t < 2 * (R > 0 ? R : 1)) break;
        if (rb + 2 * R >= 32768 || Wp >= 65536) break;  // worklist entries pack (row << 16 | x), bit 31 = flag
        if (list_entries(rb) > 0) { T = t; break; }
    }
    if (T == 0)  // no cluster of bands holds the map in shared memory: L2-resident variant
        return einx_detect_large(ctx, score, mask, B, Hp, Wp, nms_radius, border, prob_thresh, top_k, nms_map, kpts, kcap,
                                 counts, stream_);
    if (force_t <= 0)
        while (T * 2 <= kMaxCluster && (long long)B * T * 2 * 2 <= ctx->num_sms && Hp / (T * 2) >= 2 * (R > 0 ? R : 1) + 8) T *= 2;
    P.T = T;
    P.RB = (Hp + T - 1) / T;
    P.LC = list_entries(P.RB);
    if (P.LC == 0) return einx_fail(ctx, EINX_ERR_UNSUPPORTED, "einx_detect: %dx%d band does not fit shared memory", P.RB, Wp);
    // sweep runs: NSEG x NCW units over the warps of a CTA
    P.NSEG = kWarps / P.NCW > 0 ? kWarps / P.NCW : 1;
    if (P.NSEG > P.RB) P.NSEG = P.RB;
    P.SR = (P.RB + P.NSEG - 1) / P.NSEG;
    // survivor lists of the tail: shared-memory scratch (UB + list buffer) when a single CTA owns the image
    const size_t scratch = (size_t)(P.RB + 2 * R) * P.SB * 4 + (size_t)P.LC * 4;
    P.tail_smem = (T == 1 && (size_t)P.scap * 8 <= scratch) ? 1 : 0;

    const size_t list_bytes = align_up((size_t)B * P.scap * 4, 256);
    int rc = einx_ws_reserve(ctx, P.tail_smem ? 256 : 2 * list_bytes);
    if (rc) return rc;
    unsigned char* ws = (unsigned char*)ctx->ws;
    P.surv_val = (float*)ws;
    P.surv_idx = (int32_t*)(ws + list_bytes);
    static const bool want_trace = getenv("EINX_DETECT_TRACE") != nullptr;
    if (want_trace) {
        static long long* trace_buf = nullptr;
        if (!trace_buf && cudaMalloc(&trace_buf, 128 * sizeof(long long)) == cudaSuccess) cudaMemset(trace_buf, 0, 128 * sizeof(long long));
        P.trace = trace_buf;
    }
    const size_t smem = smem_for(P.RB, P.LC);
    if (T == 1) return dispatch_radius<false>(ctx, R, P, smem, stream);
    return dispatch_radius<true>(ctx, R, P, smem, stream);
}
