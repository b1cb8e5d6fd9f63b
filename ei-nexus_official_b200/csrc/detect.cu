// Detection post-processing: border removal -> iterative NMS fixpoint -> top-k threshold ->
// raster-ordered keypoint rows.  Semantics: reference core/modules/utils/detector_util.py:80-135,
// :138-164, :243-337, :451-484 (see include/einx.h); bit-exact for non-negative, NaN-free maps.
//
// One image per CTA (or per thread-block CLUSTER of row bands when the map does not fit one CTA's
// shared memory; or, beyond one cluster, per row TILE of the map, each tile a cluster of its own -- see NmsParams);
// the map is read from HBM once, lives in shared memory through all rounds, and the keypoints are written once.
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
//                 memory.  A positive pixel equal to its window maximum is a candidate (a bit per pixel, 8 rows x
//                 4 columns per register); pass 2 settles the first-occurrence rule for the candidates only (no
//                 equal value earlier in raster order inside the window), sets the LM bit of the new maxima; these
//                 are dilated on 32-bit words, and the undecided pixels under the dilation are cleared in UB and
//                 zeroed in V; the count of undecided pixels is kept exact.
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

struct NmsSide {   // one batch of maps and its outputs
    float* score;
    const uint8_t* mask;
    float* nms_map;
    float* kpts;
    int32_t* counts;
};

struct NmsParams {
    NmsSide side[2];  // images [0, Bsplit) belong to side[0], [Bsplit, B) to side[1] (the two sides of a batch of pairs)
    int B, Bsplit, Hp, Wp, border, kcap;
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
    // Tiled form for maps that no single cluster holds (1280 x 720): an image is cut into NT row tiles, each run by its
    // own cluster on the sub-image [own rows +- apron).  After k rounds a pixel's state depends on the initial values
    // within 2 R k rows of it, so the own rows of a tile are exact whenever its run took at most apron / (2 R) rounds;
    // a tile that needed more raises its image's flag and the exact large-map kernel redoes that image.
    int NT;          // tiles per image (0: untiled)
    int tile_rows;   // own rows per tile
    int apron;       // rows of context on either side of the own rows
    int seg_cap;     // survivor slots per (tile, band) segment
    int32_t* seg_cnt;   // [B * NT * T] survivors written per segment
    int32_t* redo;      // [B] set to 1 when a tile of the image ran more rounds than its apron covers
};

#define EINX_TRACE(slot)                                                                                  \
    do {                                                                                                  \
        if (P.trace && blockIdx.x == 0 && threadIdx.x == 0 && (slot) < 126) P.trace[(slot)] = clock64(); \
    } while (0)

struct Shared {
    int und;        // undecided pixels in own rows, exact
    int wl_n[2];    // worklist length after the compaction of a sparse round
    int xcnt[2];
    int warp_scan[kWarps + 1];
    unsigned int hist[4][256];  // one per radix pass of the threshold selection
    unsigned int sel_min, sel_cnt;
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

// j-th smallest (0-based) of the positive floats in list[0..n) via 4 radix passes on their bit patterns, then the
// next order statistic; every thread returns the same (a, b).  One histogram per pass (zeroed by the caller long
// before) and every warp locating the bin for itself: a pass costs one barrier.
// One histogram increment per lane.  Score values share their leading bytes (probabilities in [0.5, 1) have one
// exponent), so in the first pass -- and in every pass of a tie-heavy map -- most lanes of a warp hit the same bin: the
// lanes that agree with the first active lane go out as ONE atomic, the rest individually.  (match.any would aggregate
// every group, but its cost grows with the number of distinct digits in the warp: ~30 in the later passes.)
__device__ __forceinline__ void hist_add(unsigned int* hist, bool in, unsigned digit) {
    const unsigned act = __ballot_sync(0xffffffffu, in);
    if (act == 0u) return;  // warp-uniform
    const int leader = __ffs(act) - 1;
    const unsigned dl = __shfl_sync(0xffffffffu, digit, leader);
    const unsigned same = __ballot_sync(0xffffffffu, in && digit == dl);
    if ((int)(threadIdx.x & 31) == leader) atomicAdd(&hist[dl], (unsigned)__popc(same));
    else if (in && digit != dl) atomicAdd(&hist[digit], 1u);
}

template <bool GLOBAL>
__device__ __forceinline__ float list_ld(const float* list, int i) {
    return GLOBAL ? __ldcg(list + i) : list[i];  // global lists are written by other CTAs of the cluster: L2 only
}

template <bool GLOBAL>
__device__ void select_two(const float* list, int n, int j, bool need_next, Shared& sh, float& a_out, float& b_out) {
    const int lane = threadIdx.x & 31;
    constexpr int kHeld = 8;  // the usual list (a few thousand survivors) is read once and kept in registers
    const bool held = n <= kHeld * kThreads;
    unsigned ev[kHeld];
#pragma unroll
    for (int u = 0; u < kHeld; ++u) {
        const int i = threadIdx.x + u * kThreads;
        ev[u] = (held && i < n) ? __float_as_uint(list_ld<GLOBAL>(list, i)) : 0u;
    }
    unsigned mask = 0, prefix = 0, rank = (unsigned)j;
#pragma unroll 1
    for (int p = 0; p < 4; ++p) {
        const int shift = 24 - 8 * p;
        unsigned int* hist = sh.hist[p];
        if (held) {
#pragma unroll
            for (int u = 0; u < kHeld; ++u) {
                if (u * kThreads >= n) break;  // uniform
                const bool in = threadIdx.x + u * kThreads < n && (ev[u] & mask) == prefix;
                hist_add(hist, in, (ev[u] >> shift) & 255u);
            }
        } else {
            for (int i = threadIdx.x; i < n; i += kThreads) {
                const unsigned e = __float_as_uint(list_ld<GLOBAL>(list, i));
                if ((e & mask) == prefix) atomicAdd(&hist[(e >> shift) & 255u], 1u);
            }
        }
        __syncthreads();
        // the bin holding `rank`: 8 bins per lane, warp prefix, then a short scan (every warp, same result)
        unsigned c[8], mine = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) { c[k] = hist[lane * 8 + k]; mine += c[k]; }
        unsigned inc = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned nn = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += nn;
        }
        const unsigned before = inc - mine;
        const bool here = (before <= rank) && (rank < inc);  // exactly one lane (rank < total count)
        unsigned cum = before;
        int bin = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (cum + c[k] <= rank && bin == k) { cum += c[k]; bin = k + 1; }
        }
        const unsigned src = __ffs(__ballot_sync(0xffffffffu, here)) - 1;
        rank = __shfl_sync(0xffffffffu, rank - cum, src);
        prefix |= __shfl_sync(0xffffffffu, (unsigned)(lane * 8 + bin), src) << shift;
        mask |= 255u << shift;
    }
    const unsigned abits = prefix;
    float a = __uint_as_float(abits), b = a;
    if (need_next) {
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
        if (lane == 0) {
            if (cnt) atomicAdd(&sh.sel_cnt, cnt);
            atomicMin(&sh.sel_min, mn);
        }
        __syncthreads();
        // the (j+1)-th smallest equals a when a is duplicated past position j
        b = (sh.sel_cnt > (unsigned)j + 1u) ? a : __uint_as_float(sh.sel_min);
    }
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
    const int unit = blockIdx.x / T;                       // image (or image tile) of the launch
    const bool tiled = MULTI && P.NT > 0;
    const int bg = tiled ? unit / P.NT : unit;             // image of the launch
    const int tile = tiled ? unit - bg * P.NT : 0;
    const NmsSide& S = P.side[bg >= P.Bsplit ? 1 : 0];
    const int b = bg >= P.Bsplit ? bg - P.Bsplit : bg;    // image of its side
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int WS = P.WS, SB = P.SB, SW = P.SB - 2, Wp = P.Wp, W4 = P.W4, NCW = P.NCW;
    const int Hfull = P.Hp;                                // rows of the whole map
    // the (sub-)image this cluster works on: rows [y_off, y_off + Hp) of the map; own rows [own_lo, own_hi) (map rows)
    const int own_lo = tiled ? tile * P.tile_rows : 0;
    const int own_hi = tiled ? min(Hfull, own_lo + P.tile_rows) : Hfull;
    const int y_off = tiled ? max(0, own_lo - P.apron) : 0;
    const int Hp = tiled ? min(Hfull, own_hi + P.apron) - y_off : Hfull;
    EINX_TRACE(0);

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
    // Local row l of every array is image row ys - R + l: own rows are l in [R, R + nrows).  Bitmap rows have one
    // zero word on either side (pixel x is bit x & 31 of word 1 + (x >> 5)), so windows never index out of a row.
    const int lrows = P.RB + 2 * R;
    size_t so = align_up(sizeof(Shared), 16);
    float* const V = reinterpret_cast<float*>(smem_raw + so);
    so += sizeof(float) * (size_t)lrows * WS;
    uint32_t* const LM = reinterpret_cast<uint32_t*>(smem_raw + so);
    so += sizeof(uint32_t) * (size_t)lrows * SB;
    uint32_t* const UB = reinterpret_cast<uint32_t*>(smem_raw + so);
    so += sizeof(uint32_t) * (size_t)lrows * SB;
    // dense rounds: one more bitmap plane (candidates of the sweep, then the horizontally dilated maxima);
    // sparse rounds: the worklist; tail: the survivor lists (together with UB, dead by then)
    unsigned int* const list = reinterpret_cast<unsigned int*>(smem_raw + so);
    uint32_t* const XB = reinterpret_cast<uint32_t*>(list);   // dilated maxima, [local row][word]
    uint32_t* const CB = XB;                                   // candidates, [8-row block of own rows][column group]
    const int CBW = 32 * NCW;
    const uint32_t V_s = (uint32_t)__cvta_generic_to_shared(V);
    // neighbours' arrays (same offsets in their shared memory)
    float *Vup = nullptr, *Vdn = nullptr;
    uint32_t *UBup = nullptr, *UBdn = nullptr, *LMup = nullptr, *LMdn = nullptr;
    if (MULTI) {
        if (rank > 0) {
            Vup = cluster.map_shared_rank(V, rank - 1);
            UBup = cluster.map_shared_rank(UB, rank - 1);
            LMup = cluster.map_shared_rank(LM, rank - 1);
        }
        if (rank < T - 1) {
            Vdn = cluster.map_shared_rank(V, rank + 1);
            UBdn = cluster.map_shared_rank(UB, rank + 1);
            LMdn = cluster.map_shared_rank(LM, rank + 1);
        }
    }

    // ---- load the band: border + mask zeroing (in place on `score`), zero padding, UB bits -------- //
    {
        // bitmaps (LM, UB and the list plane are contiguous): 16-byte stores
        uint4* z = reinterpret_cast<uint4*>(LM);
        const int nz4 = (int)((sizeof(uint32_t) * ((size_t)2 * lrows * SB + (size_t)lrows * SB)) / 16);
        for (int i = tid; i < nz4; i += kThreads) z[i] = make_uint4(0u, 0u, 0u, 0u);
        for (int i = 4 * nz4 + tid; i < 3 * lrows * SB; i += kThreads) LM[i] = 0u;
    }
    if (tid == 0) { sh.und = 0; sh.wl_n[0] = sh.wl_n[1] = 0; sh.xcnt[0] = sh.xcnt[1] = 0; sh.sel_min = 0xffffffffu; sh.sel_cnt = 0; }
    for (int i = tid; i < 4 * 256; i += kThreads) (&sh.hist[0][0])[i] = 0u;
    {
        // The whole band goes from global to shared memory as asynchronous copies (cp.async, 4 * vec bytes each, no
        // register staging), all in flight at once -- one memory latency for the band.  A warp owns rows
        // (warp, warp + 16, ...), a lane the chunks (lane, lane + 32, ...) of a row: no division per chunk.  Rows
        // outside the image, the pad columns and the columns between Wp and the float4 boundary are zero-filled by
        // plain stores to disjoint addresses.
        const float* simg = S.score + ((size_t)b * Hfull + y_off) * Wp;
        const int vec = P.vec;
        const int cpr = Wp / vec;                 // copies per row (Wp % vec == 0)
        const int y_lo = max(0, R - ys), y_hi = min(L, Hp - ys + R);  // local rows inside the image: [y_lo, y_hi)
        const int tailc = 4 * W4 - Wp;            // columns between Wp and the float4 boundary (0..3)
        for (int l = warp; l < L; l += kWarps) {
            float* vrow = V + (size_t)l * WS;
            const uint32_t drow = V_s + 4u * (uint32_t)(l * WS + PAD);
            if (l >= y_lo && l < y_hi) {
                const float* srow = simg + (size_t)(ys - R + l) * Wp;
                if (vec == 4) {
                    for (int c = lane; c < cpr; c += 32)
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(drow + 16u * c), "l"(srow + 4 * c) : "memory");
                } else if (vec == 2) {
                    for (int c = lane; c < cpr; c += 32)
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(drow + 8u * c), "l"(srow + 2 * c) : "memory");
                } else {
                    for (int c = lane; c < cpr; c += 32)
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(drow + 4u * c), "l"(srow + c) : "memory");
                }
                if (lane < tailc) vrow[PAD + Wp + lane] = 0.0f;
            } else {
                for (int g = lane; g < W4; g += 32) *reinterpret_cast<float4*>(vrow + PAD + 4 * g) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            if constexpr (PAD > 0) {  // PAD floats of zeros on either side of the row
                if (lane < PAD / 2) {
                    const int off = lane < PAD / 4 ? 4 * lane : 4 * W4 + PAD + 4 * (lane - PAD / 4);
                    *reinterpret_cast<float4*>(vrow + off) = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    EINX_TRACE(1);
    {
        // Phase 2, over shared memory: the bitmap of positive pixels, one 32-pixel word per thread and step (eight
        // float4 loads, rotated by the lane so that a quarter-warp touches eight different bank groups), and -- on the
        // few words that hold border-frame or masked pixels -- the zeroing of those pixels in the band and, by the
        // band that owns the row, in `score` itself (detector_util.py:138-164 works in place).
        float* simg = S.score + ((size_t)b * Hfull + y_off) * Wp;
        const uint8_t* mimg = S.mask ? S.mask + ((size_t)b * Hfull + y_off) * Wp : nullptr;
        const int bd = P.border;
        const int xe = Wp - bd;  // columns >= xe belong to the frame
        int cnt = 0;
        for (int wi = tid; wi < L * SW; wi += kThreads) {
            const int l = wi / SW, sidx = wi - l * SW;
            const int y = ys - R + l, x0 = 32 * sidx;
            const bool rowin = y >= 0 && y < Hp;
            const int gy = y + y_off;  // row of the whole map
            const bool band_own = l >= R && l < R + nrows;               // rows this band counts as undecided
            const bool own = band_own && gy >= own_lo && gy < own_hi;      // rows this band writes back to `score`
            uint32_t pos = 0;
            if (rowin && x0 < Wp) {
                const uint32_t base = V_s + 4u * (uint32_t)(l * WS + PAD + x0);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int kk = (k + lane) & 7;
                    if (x0 + 4 * kk < 4 * W4) {
                        const float4 q = lds128(base + 16u * kk);
                        const uint32_t nib = (q.x > 0.0f ? 1u : 0u) | (q.y > 0.0f ? 2u : 0u) | (q.z > 0.0f ? 4u : 0u) | (q.w > 0.0f ? 8u : 0u);
                        pos |= nib << (4 * kk);
                    }
                }
                uint32_t kill = 0;
                if (gy < bd || gy >= Hfull - bd) {
                    kill = 0xffffffffu;
                } else {
                    if (x0 < bd) kill |= (bd - x0 >= 32) ? 0xffffffffu : ((1u << (bd - x0)) - 1u);
                    if (x0 + 32 > xe) kill |= (xe <= x0) ? 0xffffffffu : (0xffffffffu << (xe - x0));
                }
                const uint32_t inimg = (Wp - x0 >= 32) ? 0xffffffffu : ((1u << (Wp - x0)) - 1u);
                if (mimg) {
                    const uint8_t* mrow = mimg + (size_t)y * Wp + x0;
                    for (int j = 0; j < 32; ++j)
                        if (((inimg >> j) & 1u) && mrow[j] == 0) kill |= 1u << j;
                }
                kill &= inimg;
                if (kill) {
                    uint32_t kk = kill;
                    float* vrow = V + (size_t)l * WS + PAD + x0;
                    float* srow = simg + (size_t)y * Wp + x0;
                    while (kk) {
                        const int j = __ffs(kk) - 1;
                        kk &= kk - 1;
                        if (vrow[j] != 0.0f) {
                            vrow[j] = 0.0f;
                            if (own) srow[j] = 0.0f;  // the band that owns the row writes the zeroed frame / mask back
                        }
                    }
                    pos &= ~kill;
                }
            }
            UB[(size_t)l * SB + 1 + sidx] = pos;
            if (band_own) cnt += __popc(pos);
        }
        cnt = __reduce_add_sync(0xffffffffu, cnt);
        if (lane == 0 && cnt) atomicAdd(&sh.und, cnt);
    }
    EINX_TRACE(126);

    // ---- NMS rounds ------------------------------------------------------------------------ //
    auto sync_all = [&]() {
        if (MULTI) cluster.sync();
        else __syncthreads();
    };

    int rounds_done = 0;  // rounds that changed the state (uniform over the cluster)
    if constexpr (R > 0) {
        const int cap = min(P.LC, 8 * kThreads);   // worklist capacity (phase 3 holds 8 entries per thread)
        bool sparse = false;
        int n_wl = 0;           // worklist length (sparse rounds; uniform in the CTA)
        int trace_slot = 2;
        for (int round = 0;; ++round) {
            rounds_done = round;
            sync_all();  // every band's V / UB / LM / count is final for this round
            trace_slot = round < 16 ? 2 + 7 * round : 126;
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
            if (!sparse && mx <= cap) {
                // build the worklist of undecided pixels of the own rows, raster order; a thread takes a run of
                // consecutive words so that one block scan places everything
                sparse = true;
                const int nwords = nrows * SW;
                const int per = (nwords + kThreads - 1) / kThreads;
                const int w0 = tid * per, w1 = min(w0 + per, nwords);
                int mine = 0;
                for (int wi = w0; wi < w1; ++wi) {
                    const int lr = wi / SW, s = wi - lr * SW;
                    mine += __popc(UB[(size_t)(lr + R) * SB + 1 + s]);
                }
                int totw;
                int pos = block_excl_scan(mine, sh.warp_scan, totw);
                for (int wi = w0; wi < w1; ++wi) {
                    const int lr = wi / SW, s = wi - lr * SW;
                    uint32_t w = UB[(size_t)(lr + R) * SB + 1 + s];
                    while (w) {
                        const int bit = __ffs(w) - 1;
                        w &= w - 1;
                        list[pos++] = ((unsigned)(lr + R) << 16) | (unsigned)(32 * s + bit);
                    }
                }
                n_wl = totw;
                __syncthreads();
            }
            if (!sparse) {
                // ---- dense pass 1: sweep; pixels equal to their window maximum are candidates ------- //
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
                    // ring slot of a row = (row - (a - R)) mod P2: horizontal window maxima (h), the 3-row partial
                    // maximum that ends with the row (p3 of the row two above) and the thread's own 4 pixels (cv)
                    float h[P2][4], p3[P2][4], cv[P2][4];
#pragma unroll
                    for (int j = 0; j < P2; ++j)
#pragma unroll
                        for (int c = 0; c < 4; ++c) { h[j][c] = 0.0f; p3[j][c] = 0.0f; cv[j][c] = 0.0f; }
                    auto ingest = [&](auto slot) {
                        constexpr int j = decltype(slot)::value;
                        constexpr int j1 = (j + P2 - 1) % P2, j2 = (j + P2 - 2) % P2;
                        float av[4 + 2 * PAD];
#pragma unroll
                        for (int k = 0; k < 1 + PAD / 2; ++k) {
                            const float4 q = lds128(rp + 16u * k);
                            av[4 * k] = q.x; av[4 * k + 1] = q.y; av[4 * k + 2] = q.z; av[4 * k + 3] = q.w;
                        }
#pragma unroll
                        for (int c = 0; c < 4; ++c) cv[j][c] = av[PAD + c];
                        hmax4<R>(av, h[j]);
#pragma unroll
                        for (int c = 0; c < 4; ++c) p3[j2][c] = fmax3(h[j2][c], h[j1][c], h[j][c]);
                    };
                    // Output row y = (row just ingested in slot j) - R: all of its window rows are in the ring.  A positive
                    // pixel equal to its window maximum is a CANDIDATE; the first-occurrence rule (no equal value earlier in
                    // raster order inside the window) is settled for the few candidates in pass 2, not here for every pixel.
                    // Candidates of 8 rows x 4 columns collect in one register.
                    unsigned acc = 0;
                    auto emit = [&](auto slot, int y) {
                        constexpr int j = decltype(slot)::value;
                        constexpr int j2 = (j + P2 - 2) % P2;
                        constexpr int jc = (j + P2 - R) % P2;          // slot of row y itself
                        float M[4];
#pragma unroll
                        for (int c = 0; c < 4; ++c) M[c] = p3[j2][c];  // rows y+R-2 .. y+R
#pragma unroll
                        for (int i = 0; 3 * i < P2 - 3; ++i) {
                            const int ji = (j + 3 * i + 1) % P2;       // rows y-R+3i .. y-R+3i+2
#pragma unroll
                            for (int c = 0; c < 4; ++c) M[c] = fmaxf(M[c], p3[ji][c]);
                        }
                        unsigned nib = 0;
#pragma unroll
                        for (int c = 0; c < 4; ++c)
                            if (cv[jc][c] >= M[c] && cv[jc][c] > 0.0f) nib |= 1u << c;
                        if (active) acc |= nib << (4 * ((y - R) & 7));
                        if (((y - R) & 7) == 7 || y == e - 1) {
                            // runs start on multiples of 8 rows, so this (8 rows x 4 columns) word has one writer
                            CB[(size_t)((y - R) >> 3) * CBW + g] = acc;
                            acc = 0;
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
                __syncthreads();
                EINX_TRACE(trace_slot); ++trace_slot;
                // ---- dense pass 2: candidates -> maxima.  The first-occurrence test of detector_util.py:298-308 (the
                // pooling index is the first maximum of the window in raster order): no equal value in the window rows
                // above, none to the left in the same row.  Only candidates pay for it (one in ~80 pixels of a round).
                for (int wi = tid; wi < ((nrows + 7) >> 3) * CBW; wi += kThreads) {
                    uint32_t c = CB[wi];
                    if (!c) continue;
                    const int blk = wi / CBW, g = wi - blk * CBW;
                    while (c) {
                        const int bit = __ffs(c) - 1;
                        c &= c - 1;
                        const int l = R + 8 * blk + (bit >> 2), x = 4 * g + (bit & 3);
                        const size_t widx = (size_t)l * SB + 1 + (x >> 5);
                        const uint32_t m = 1u << (x & 31);
                        if (LM[widx] & m) continue;  // a maximum of an earlier round stays one
                        const uint32_t px = V_s + 4u * (uint32_t)(l * WS + PAD + x);
                        const float v = lds32(px);
                        bool tie = false;
#pragma unroll
                        for (int d = 1; d <= R; ++d) tie |= (lds32(px - 4 * d) == v);
#pragma unroll
                        for (int dy = 1; dy <= R; ++dy) {
                            const uint32_t rowp = px - 4u * (uint32_t)(dy * WS);
#pragma unroll
                            for (int d = -R; d <= R; ++d) tie |= (lds32(rowp + 4 * d) == v);
                        }
                        if (tie) continue;
                        atomicOr(&LM[widx], m);
                        if (MULTI) {  // the neighbours dilate their copies of my edge rows
                            const int lu = cp.up(l, 2 * R), ld = cp.down(l);
                            if (lu >= 0) atomicOr(&LMup[(size_t)lu * SB + 1 + (x >> 5)], m);
                            if (ld >= 0) atomicOr(&LMdn[(size_t)ld * SB + 1 + (x >> 5)], m);
                        }
                    }
                }
                EINX_TRACE(trace_slot); ++trace_slot;
                sync_all();
                // ---- dense pass 3: horizontal dilation of the maxima, all local rows, on words ------ //
                if (tid == 0) sh.und = 0;
                for (int wi = tid; wi < L * SW; wi += kThreads) {
                    const int l = wi / SW, s = wi - l * SW;
                    const size_t widx = (size_t)l * SB + 1 + s;
                    const uint32_t w = LM[widx], wl = LM[widx - 1], wr = LM[widx + 1];
                    uint32_t acc = w;
#pragma unroll
                    for (int d = 1; d <= R; ++d) acc |= __funnelshift_l(wl, w, d) | __funnelshift_r(w, wr, d);
                    XB[widx] = acc;
                }
                __syncthreads();
                EINX_TRACE(trace_slot); ++trace_slot;
                // ---- dense pass 4: vertical dilation; undecided pixels under it are dead: clear them in UB and
                // zero them in V (the owner writes every copy of its rows); count what stays undecided
                {
                    int left = 0;
                    const int nwords = nrows * SW;
                    for (int base = warp * 32; base < nwords; base += kThreads) {
                        const int wi = base + lane;
                        uint32_t sup = 0;
                        int l = 0, s = 0;
                        if (wi < nwords) {
                            const int lr = wi / SW;
                            s = wi - lr * SW;
                            l = lr + R;
                            const size_t widx = (size_t)l * SB + 1 + s;
                            uint32_t d = 0;
#pragma unroll
                            for (int dy = -R; dy <= R; ++dy) d |= XB[widx + dy * SB];
                            const uint32_t ub = UB[widx];
                            const uint32_t keep = ub & ~d;
                            sup = ub & d & ~LM[widx];
                            if (keep != ub) {
                                UB[widx] = keep;
                                if (MULTI) {
                                    const int lu = cp.up(l, 2 * R), ld = cp.down(l);
                                    if (lu >= 0) UBup[(size_t)lu * SB + 1 + s] = keep;
                                    if (ld >= 0) UBdn[(size_t)ld * SB + 1 + s] = keep;
                                }
                            }
                            left += __popc(keep);
                        }
                        // zero the dead pixels: the warp walks its 32 words as 256 float4 chunks, lanes on consecutive chunks
                        if (__any_sync(0xffffffffu, sup != 0)) {
#pragma unroll
                            for (int q = 0; q < 8; ++q) {
                                const int src = 4 * q + (lane >> 3);            // word (lane of the batch) owning this chunk
                                const uint32_t sw = __shfl_sync(0xffffffffu, sup, src);
                                const int sls = __shfl_sync(0xffffffffu, (l << 8) | s, src);  // s < 256 words
                                const uint32_t nib = (sw >> (4 * (lane & 7))) & 0xfu;
                                if (nib) {
                                    const int sl = sls >> 8, ss = sls & 255;
                                    const int x = 32 * ss + 4 * (lane & 7);
                                    float4* cell = reinterpret_cast<float4*>(V + (size_t)sl * WS + PAD + x);
                                    float4 v = *cell;
                                    if (nib & 1u) v.x = 0.0f;
                                    if (nib & 2u) v.y = 0.0f;
                                    if (nib & 4u) v.z = 0.0f;
                                    if (nib & 8u) v.w = 0.0f;
                                    *cell = v;
                                    if (MULTI) {
                                        const int lu = cp.up(sl, 2 * R), ld = cp.down(sl);
                                        if (lu >= 0) *reinterpret_cast<float4*>(Vup + (size_t)lu * WS + PAD + x) = v;
                                        if (ld >= 0) *reinterpret_cast<float4*>(Vdn + (size_t)ld * WS + PAD + x) = v;
                                    }
                                }
                            }
                        }
                    }
                    left = __reduce_add_sync(0xffffffffu, left);
                    if (lane == 0 && left) atomicAdd(&sh.und, left);
                }
                EINX_TRACE(trace_slot); ++trace_slot;
            } else {
                // ---- sparse round, phase 1: an undecided pixel looks at its undecided neighbours ---- //
                const int n = n_wl;
                for (int en = tid; en < n; en += kThreads) {
                    const unsigned ent = list[en];
                    const int l = (int)(ent >> 16), x = (int)(ent & 0xffffu);
                    const float* cpx = V + (size_t)l * WS + PAD + x;
                    const float v = cpx[0];
                    const int xl = x - R + 32;
                    const int wi = xl >> 5, shf = xl & 31;
                    // Undecided pixels cluster, so a data-dependent loop over the set bits runs as long as the fullest
                    // window of the warp: every position is visited instead, its load predicated on the bit.
                    float early = 0.0f, late = 0.0f;  // raster-earlier / raster-later undecided neighbours (values > 0)
#pragma unroll
                    for (int dy = -R; dy <= R; ++dy) {
                        const uint32_t* u = UB + (size_t)(l + dy) * SB + wi;
                        const unsigned long long both = (unsigned long long)u[0] | ((unsigned long long)u[1] << 32);
                        uint32_t f = (uint32_t)(both >> shf) & kWinMask;
                        if (dy == 0) f &= ~(1u << R);
                        if (f) {
#pragma unroll
                            for (int k = 0; k < P2; ++k) {
                                const float w = (f & (1u << k)) ? cpx[dy * WS + k - R] : 0.0f;
                                if (dy < 0 || (dy == 0 && k < R)) early = fmaxf(early, w);
                                else late = fmaxf(late, w);
                            }
                        }
                    }
                    // raster-earlier neighbours win ties (first-occurrence argmax)
                    const bool beaten = !(v > early && v >= late);
                    if (!beaten) {
                        atomicOr(&LM[(size_t)l * SB + 1 + (x >> 5)], 1u << (x & 31));
                        list[en] = ent | 0x80000000u;  // a new maximum (local rows < 32768)
                    }
                }
                EINX_TRACE(trace_slot); ++trace_slot;
                sync_all();
                EINX_TRACE(trace_slot); ++trace_slot;
                // ---- phase 2: the new maxima clear their windows in UB (every copy of a row; the copies are
                // identical, so the local one tells whether there is anything to clear) --------------------- //
                if (tid == 0) sh.wl_n[0] = 0;
                for (int en = tid; en < n; en += kThreads) {
                    const unsigned ent = list[en];
                    if (!(ent & 0x80000000u)) continue;
                    const int l = (int)((ent >> 16) & 0x7fffu), x = (int)(ent & 0xffffu);
                    const int xl = x - R + 32;
                    const int wi = xl >> 5, shf = xl & 31;
                    const unsigned long long m64 = (unsigned long long)kWinMask << shf;
                    const uint32_t lo = (uint32_t)m64, hi = (uint32_t)(m64 >> 32);
#pragma unroll
                    for (int dy = -R; dy <= R; ++dy) {
                        const int l2 = l + dy;
                        uint32_t* u = UB + (size_t)l2 * SB + wi;
                        const bool c0 = (u[0] & lo) != 0, c1 = hi && (u[1] & hi) != 0;
                        if (!(c0 | c1)) continue;
                        if (c0) atomicAnd(u, ~lo);
                        if (c1) atomicAnd(u + 1, ~hi);
                        if (MULTI) {
                            const int lu = cp.up(l2, 2 * R), ld = cp.down(l2);
                            if (lu >= 0) {
                                uint32_t* ur = UBup + (size_t)lu * SB + wi;
                                if (c0) atomicAnd(ur, ~lo);
                                if (c1) atomicAnd(ur + 1, ~hi);
                            }
                            if (ld >= 0) {
                                uint32_t* ur = UBdn + (size_t)ld * SB + wi;
                                if (c0) atomicAnd(ur, ~lo);
                                if (c1) atomicAnd(ur + 1, ~hi);
                            }
                        }
                    }
                }
                sync_all();
                EINX_TRACE(trace_slot); ++trace_slot;
                // ---- phase 3: keep the still-undecided entries (read all, then write in place) ------ //
                {
                    constexpr int kPer = 8;  // cap <= kPer * kThreads is enforced by the host
                    unsigned held[kPer];
                    bool keep[kPer];
#pragma unroll
                    for (int u = 0; u < kPer; ++u) {
                        const int en = tid + u * kThreads;
                        held[u] = 0;
                        keep[u] = false;
                        if (en < n) {
                            const unsigned ent = list[en];
                            held[u] = ent;
                            if (!(ent & 0x80000000u)) {
                                const int l = (int)(ent >> 16), x = (int)(ent & 0xffffu);
                                keep[u] = (UB[(size_t)l * SB + 1 + (x >> 5)] >> (x & 31)) & 1u;
                            }
                        }
                    }
                    __syncthreads();
#pragma unroll
                    for (int u = 0; u < kPer; ++u) {
                        if (u * kThreads >= n) break;  // uniform
                        const unsigned kb = __ballot_sync(0xffffffffu, keep[u]);
                        int pos = 0;
                        if (lane == 0 && kb) pos = atomicAdd(&sh.wl_n[0], __popc(kb));
                        pos = __shfl_sync(0xffffffffu, pos, 0);
                        if (keep[u]) list[pos + __popc(kb & ((1u << lane) - 1u))] = held[u];
                    }
                    __syncthreads();
                    n_wl = sh.wl_n[0];
                    if (tid == 0) sh.und = n_wl;
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
    if (tiled) {
        // ---- tiled form: the survivors of this band's rows that the tile owns, in raster order, into the band's
        // segment; nms_tile_tail_kernel strings the segments of an image together, selects the threshold and writes
        // the keypoints.  A tile whose run outlasted its apron flags the image for the exact redo.
        const size_t seg = (size_t)unit * T + rank;
        float* sl = P.surv_val + seg * P.seg_cap;
        int32_t* si = P.surv_idx + seg * P.seg_cap;
        const int nwords = nrows * SW;
        const int per = (nwords + kThreads - 1) / kThreads;
        const int w0 = min(tid * per, nwords), w1 = min(w0 + per, nwords);
        int mine = 0, own_n = 0;
        for (int wi = w0; wi < w1; ++wi) {
            const int lr = wi / SW, sx = wi - lr * SW;
            const int gy = y_off + ys + lr;
            if (gy >= own_lo && gy < own_hi) mine += __popc(LM[(size_t)(lr + R) * SB + 1 + sx]);
        }
        int pos = block_excl_scan(mine, sh.warp_scan, own_n);
        for (int wi = w0; wi < w1; ++wi) {
            const int lr = wi / SW, sx = wi - lr * SW;
            const int gy = y_off + ys + lr;
            if (gy < own_lo || gy >= own_hi) continue;
            uint32_t w = LM[(size_t)(lr + R) * SB + 1 + sx];
            while (w) {
                const int bit = __ffs(w) - 1;
                w &= w - 1;
                const int x = 32 * sx + bit;
                if (pos < P.seg_cap) {
                    sl[pos] = V[(size_t)(lr + R) * WS + PAD + x];
                    si[pos] = gy * Wp + x;
                }
                ++pos;
            }
        }
        if (tid == 0) {
            P.seg_cnt[seg] = min(own_n, P.seg_cap);
            if (rank == 0) {
                // rows of context actually present on either side (a side that ends at the map's edge needs none)
                const int top = y_off > 0 ? own_lo - y_off : (1 << 28);
                const int bottom = y_off + Hp < Hfull ? (y_off + Hp) - own_hi : (1 << 28);
                if (2 * R * rounds_done > min(top, bottom)) P.redo[bg] = 1;
            }
        }
        cluster.sync();  // nobody leaves while a neighbour may still read its shared memory
        return;
    }
    // ---- survivors -> ordered per-image list ------------------------------------------------- //
    // At the fixpoint the selected pixels (LM bits of the own rows) are exactly the survivors.  The lists
    // live in the shared-memory scratch (UB + list buffer, both dead now) for single-CTA images, in the
    // global workspace for clusters.  A thread takes a run of consecutive bitmap words, so one block scan
    // orders the whole band.
    float* slist;
    int32_t* sidx;
    const bool smem_lists = !MULTI && P.tail_smem;
    if (smem_lists) {
        slist = reinterpret_cast<float*>(UB);
        sidx = reinterpret_cast<int32_t*>(UB) + P.scap;
    } else {
        slist = P.surv_val + (size_t)bg * P.scap;
        sidx = P.surv_idx + (size_t)bg * P.scap;
    }
    const int nwords = nrows * SW;
    const int per = (nwords + kThreads - 1) / kThreads;
    const int w0 = min(tid * per, nwords), w1 = min(w0 + per, nwords);
    int own = 0, offset = 0, total = 0, mypos = 0;
    {
        int mine = 0;
        for (int wi = w0; wi < w1; ++wi) {
            const int lr = wi / SW, s = wi - lr * SW;
            mine += __popc(LM[(size_t)(lr + R) * SB + 1 + s]);
        }
        mypos = block_excl_scan(mine, sh.warp_scan, own);
        if (tid == 0) sh.xcnt[0] = own;
    }
    sync_all();
    if (MULTI) {
        for (int r = 0; r < T; ++r) {
            const int c = *cluster.map_shared_rank(&sh.xcnt[0], r);
            if (r < rank) offset += c;
            total += c;
        }
    } else {
        total = own;
    }
    {
        // (the lists may alias UB, never LM or V)
        int pos = offset + mypos;
        for (int wi = w0; wi < w1; ++wi) {
            const int lr = wi / SW, s = wi - lr * SW;
            uint32_t w = LM[(size_t)(lr + R) * SB + 1 + s];
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
    // a thread takes a run of consecutive survivors of the band: one block scan places the rows
    const int sper = (own + kThreads - 1) / kThreads;
    const int s0 = min(tid * sper, own), s1 = min(s0 + sper, own);
    int kmine = 0;
    for (int i = s0; i < s1; ++i) kmine += ((smem_lists ? slist[offset + i] : __ldcg(slist + offset + i)) > thr) ? 1 : 0;
    int kown;
    int kpos = block_excl_scan(kmine, sh.warp_scan, kown);
    if (tid == 0) sh.xcnt[1] = kown;
    sync_all();
    int koff = 0, ktotal = 0;
    if (MULTI) {
        for (int r = 0; r < T; ++r) {
            const int c = *cluster.map_shared_rank(&sh.xcnt[1], r);
            if (r < rank) koff += c;
            ktotal += c;
        }
    } else {
        ktotal = kown;
    }
    if (rank == 0 && tid == 0) S.counts[b] = ktotal;
    {
        float* krows = S.kpts + (size_t)b * P.kcap * 3;
        int pos = koff + kpos;
        for (int i = s0; i < s1; ++i) {
            const float v = smem_lists ? slist[offset + i] : __ldcg(slist + offset + i);
            if (v > thr) {
                if (pos < P.kcap) {
                    const int idx = smem_lists ? sidx[offset + i] : __ldcg(sidx + offset + i);
                    const int y = idx / Wp, x = idx - y * Wp;
                    krows[(size_t)pos * 3 + 0] = (float)y + 0.5f;
                    krows[(size_t)pos * 3 + 1] = (float)x + 0.5f;
                    krows[(size_t)pos * 3 + 2] = v;
                }
                ++pos;
            }
        }
    }
    if (S.nms_map) {
        // selected AND above the threshold; V still holds dead values from the sparse rounds, LM decides
        float* out = S.nms_map + (size_t)b * Hp * Wp;
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

// ---- tiled form, second kernel: one CTA per image ------------------------------------------------ //
// Selects the top-k threshold over the survivors of all (tile, band) segments of an image
// (detector_util.py:108-133) and writes the keypoint rows in raster order (= segment order) and, optionally, the
// dense map (zeroed by the host beforehand).  The selection only needs the multiset of values: a warp reads its
// segments once, the values stay in registers through the four radix passes.
constexpr int kTailThreads = 1024;
constexpr int kTailWarps = kTailThreads / 32;
constexpr int kTailHeld = 40;

struct TailShared {
    unsigned int hist[4][256];
    unsigned int sel_min, sel_cnt;
    int warp_scan[kTailWarps + 1];
};

__device__ __forceinline__ int tail_block_scan(int v, int* scratch, int& total) {  // exclusive scan over 1024 threads
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += n;
    }
    __syncthreads();
    if (lane == 31) scratch[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        const int w = scratch[lane];
        int winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int n = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += n;
        }
        scratch[lane] = winc - w;
        if (lane == 31) scratch[kTailWarps] = winc;
    }
    __syncthreads();
    total = scratch[kTailWarps];
    return inc - v + scratch[warp];
}

__global__ void __launch_bounds__(kTailThreads, 1) nms_tile_tail_kernel(const NmsParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TailShared& sh = *reinterpret_cast<TailShared*>(smem_raw);
    // work items: 32 consecutive entries of a segment, in segment order (= raster order)
    int* item_off = reinterpret_cast<int*>(smem_raw + align_up(sizeof(TailShared), 16));  // [nseg + 1] first item of a segment
    int* item_pos = item_off + (P.NT * P.T + 1);                                          // [kTailHeld * 32] kept rows before an item
    int* seg_n = item_pos + kTailHeld * kTailWarps;                                       // [nseg] entries of a segment
    unsigned short* item_seg = reinterpret_cast<unsigned short*>(seg_n + P.NT * P.T);     // [kTailHeld * 32] segment of an item
    const int bg = blockIdx.x;
    const NmsSide& S = P.side[bg >= P.Bsplit ? 1 : 0];
    const int b = bg >= P.Bsplit ? bg - P.Bsplit : bg;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nseg = P.NT * P.T, Hp = P.Hp, Wp = P.Wp;
    const int32_t* cnt = P.seg_cnt + (size_t)bg * nseg;
    const float* seg_val = P.surv_val + (size_t)bg * nseg * P.seg_cap;
    const int32_t* seg_idx = P.surv_idx + (size_t)bg * nseg * P.seg_cap;
    if (tid == 0) { sh.sel_min = 0xffffffffu; sh.sel_cnt = 0; }
    for (int i = tid; i < 4 * 256; i += kTailThreads) (&sh.hist[0][0])[i] = 0u;
    int nitems, total;
    {
        const int c = tid < nseg ? __ldcg(cnt + tid) : 0;
        const int ex = tail_block_scan((c + 31) >> 5, sh.warp_scan, nitems);
        if (tid < nseg) { item_off[tid] = ex; seg_n[tid] = c; }
        if (tid == 0) item_off[nseg] = nitems;
        (void)tail_block_scan(c, sh.warp_scan, total);
    }
    __syncthreads();
    for (int sgm = warp; sgm < nseg; sgm += kTailWarps)   // segment of every item (a warp per segment)
        for (int it = item_off[sgm] + lane; it < item_off[sgm + 1]; it += 32) item_seg[it] = (unsigned short)sgm;
    __syncthreads();
    // values into registers: warp w takes items w, w + 32, ...: slot u of a lane is entry `lane` of item w + 32 u
    // (the host guarantees nitems <= kTailHeld * 32)
    unsigned ev[kTailHeld];
    auto item_segment = [&](int it) { return (int)item_seg[it]; };
#pragma unroll
    for (int u = 0; u < kTailHeld; ++u) {
        const int it = warp + kTailWarps * u;
        ev[u] = 0u;
        if (it < nitems) {
            const int sgm = item_segment(it);
            const int i = 32 * (it - item_off[sgm]) + lane;
            if (i < seg_n[sgm]) ev[u] = __float_as_uint(__ldcg(seg_val + (size_t)sgm * P.seg_cap + i));
        }
    }
    float thr = P.prob_thresh;
    if (P.use_topk == 2) {
        thr = fminf(0.0f, P.prob_thresh);
    } else if (P.use_topk == 1) {
        const int n = Hp * Wp;
        const int zeros = n - total;
        float a = 0.0f, bq = 0.0f;
        if (P.rank_hi >= zeros) {
            const bool lo_on_zero = P.rank_lo < zeros;   // lo falls on a zero, hi on the smallest survivor
            const unsigned j = lo_on_zero ? 0u : (unsigned)(P.rank_lo - zeros);
            const bool need_next = !lo_on_zero && P.rank_hi != P.rank_lo;
            // j-th smallest of the positive floats (bit patterns; empty slots hold 0 and are skipped) via 4 radix passes
            unsigned mask = 0, prefix = 0, rank = j;
#pragma unroll 1
            for (int p = 0; p < 4; ++p) {
                const int shift = 24 - 8 * p;
                unsigned int* hist = sh.hist[p];
#pragma unroll
                for (int u = 0; u < kTailHeld; ++u) {
                    if (warp + kTailWarps * u >= nitems) break;  // warp-uniform
                    const bool in = ev[u] != 0u && (ev[u] & mask) == prefix;
                    hist_add(hist, in, (ev[u] >> shift) & 255u);
                }
                __syncthreads();
                unsigned c[8], mine = 0;
#pragma unroll
                for (int k = 0; k < 8; ++k) { c[k] = hist[lane * 8 + k]; mine += c[k]; }
                unsigned inc = mine;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const unsigned nn = __shfl_up_sync(0xffffffffu, inc, o);
                    if (lane >= o) inc += nn;
                }
                const unsigned before = inc - mine;
                const bool here = (before <= rank) && (rank < inc);
                unsigned cum = before;
                int bin = 0;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    if (cum + c[k] <= rank && bin == k) { cum += c[k]; bin = k + 1; }
                }
                const unsigned src = __ffs(__ballot_sync(0xffffffffu, here)) - 1;
                rank = __shfl_sync(0xffffffffu, rank - cum, src);
                prefix |= __shfl_sync(0xffffffffu, (unsigned)(lane * 8 + bin), src) << shift;
                mask |= 255u << shift;
            }
            const unsigned abits = prefix;
            float va = __uint_as_float(abits), vb = va;
            if (need_next) {
                unsigned cn = 0, mn = 0xffffffffu;
#pragma unroll
                for (int u = 0; u < kTailHeld; ++u)
                    if (ev[u] != 0u) {
                        if (ev[u] <= abits) cn++;
                        else mn = min(mn, ev[u]);
                    }
                cn = __reduce_add_sync(0xffffffffu, cn);
                mn = __reduce_min_sync(0xffffffffu, mn);
                if (lane == 0) {
                    if (cn) atomicAdd(&sh.sel_cnt, cn);
                    atomicMin(&sh.sel_min, mn);
                }
                __syncthreads();
                vb = (sh.sel_cnt > j + 1u) ? va : __uint_as_float(sh.sel_min);
            }
            if (lo_on_zero) { bq = va; a = 0.0f; }
            else { a = va; bq = vb; }
        }
        const float thr_k = __fsub_rn(bq, __fmul_rn(__fsub_rn(bq, a), 0.5f));
        thr = fminf(thr_k, P.prob_thresh);
    }
    // kept rows per item, prefix in item order (= raster order), then the rows -- values still from the registers
#pragma unroll
    for (int u = 0; u < kTailHeld; ++u) {
        const bool keep = ev[u] != 0u && __uint_as_float(ev[u]) > thr;
        const unsigned kb = __ballot_sync(0xffffffffu, keep);
        const int it = warp + kTailWarps * u;
        if (lane == 0 && it < nitems) item_pos[it] = __popc(kb);
    }
    __syncthreads();
    int ktotal = 0;
    for (int base = 0; base < nitems; base += kTailThreads) {  // (nitems <= 1280: at most two rounds)
        const int i = base + tid;
        const int c = i < nitems ? item_pos[i] : 0;
        int chunk_total;
        const int ex = tail_block_scan(c, sh.warp_scan, chunk_total);
        __syncthreads();
        if (i < nitems) item_pos[i] = ktotal + ex;
        ktotal += chunk_total;
    }
    __syncthreads();
    if (tid == 0) S.counts[b] = ktotal;
    float* krows = S.kpts + (size_t)b * P.kcap * 3;
    float* out = S.nms_map ? S.nms_map + (size_t)b * Hp * Wp : nullptr;
#pragma unroll
    for (int u = 0; u < kTailHeld; ++u) {
        const int it = warp + kTailWarps * u;
        if (it >= nitems) break;  // warp-uniform
        const bool keep = ev[u] != 0u && __uint_as_float(ev[u]) > thr;
        const unsigned kb = __ballot_sync(0xffffffffu, keep);
        if (kb == 0u) continue;   // warp-uniform
        const int sgm = item_segment(it);
        if (keep) {
            const int my = item_pos[it] + __popc(kb & ((1u << lane) - 1u));
            const int i = 32 * (it - item_off[sgm]) + lane;
            const int idx = __ldcg(seg_idx + (size_t)sgm * P.seg_cap + i);
            const float v = __uint_as_float(ev[u]);
            if (my < P.kcap) {
                const int y = idx / Wp, x = idx - y * Wp;
                krows[(size_t)my * 3 + 0] = (float)y + 0.5f;
                krows[(size_t)my * 3 + 1] = (float)x + 0.5f;
                krows[(size_t)my * 3 + 2] = v;
            }
            if (out) out[idx] = v;
        }
    }
}

template <int R, bool MULTI>
int launch_nms(einx_ctx* ctx, const NmsParams& P, size_t smem, cudaStream_t stream) {
    auto kern = nms_kernel<R, MULTI>;
    EINX_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(P.B * P.T * (P.NT > 0 ? P.NT : 1));
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

namespace {

int detect_impl(einx_ctx* ctx, const NmsSide* sides, int nsides, int Bside, int Hp, int Wp, int nms_radius, int border,
                float prob_thresh, int top_k, int kcap, einx_stream stream_) {
    const int B = Bside * nsides;
    DeviceGuard guard(ctx->device);
    cudaStream_t stream = (cudaStream_t)stream_;
    const int R = nms_radius;

    NmsParams P = {};
    for (int i = 0; i < nsides; ++i) P.side[i] = sides[i];
    P.B = B; P.Bsplit = Bside; P.Hp = Hp; P.Wp = Wp; P.border = border; P.kcap = kcap;
    const int PAD = (R + 3) / 4 * 4;
    P.W4 = (Wp + 3) / 4;
    P.NCW = (P.W4 + 31) / 32;
    P.WS = 4 * P.W4 + 2 * PAD;
    P.SB = 4 * P.NCW + 2;
    P.vec = 4;
    for (int i = 0; i < nsides; ++i) {
        const uintptr_t a = (uintptr_t)sides[i].score, m = (uintptr_t)sides[i].nms_map;
        const int v = (Wp % 4 == 0 && a % 16 == 0 && m % 16 == 0) ? 4 : (Wp % 2 == 0 && a % 8 == 0) ? 2 : 1;
        if (v < P.vec) P.vec = v;
    }
    P.prob_thresh = prob_thresh;
    const int n = Hp * Wp;
    if (top_k > 0) {
        if (top_k >= n) P.use_topk = 2;
        else { P.use_topk = 1; einx_topk_ranks(n, top_k, &P.rank_lo, &P.rank_hi); }
    }
    P.scap = R == 0 ? n : ((Hp + R) / (R + 1)) * ((Wp + R) / (R + 1));

    // Bands per image: the smallest cluster whose band (values, two bitmaps, list buffer) fits one CTA's shared
    // memory; then wider while the launch leaves more than half of the SMs idle.  Bands of a cluster are at least
    // 2R rows, so a row has at most two copies.
    const size_t fixed = align_up(sizeof(Shared), 16);
    const size_t budget = (size_t)ctx->max_smem_optin;
    // list buffer: one bitmap plane in dense rounds (candidates / dilated maxima), the worklist of undecided pixels
    // in sparse rounds -- the larger it is, the earlier the switch to sparse rounds
    auto smem_for = [&](int rb, int lc) {
        return fixed + (size_t)(rb + 2 * R) * ((size_t)P.WS * 4 + (size_t)P.SB * 8) + (size_t)lc * 4;
    };
    auto list_entries = [&](int rb) {  // 0: the band does not fit
        // the plane holds the dilated maxima ([local row][word]) or the sweep's candidates ([8-row block][column group])
        const int hd = (rb + 2 * R) * P.SB, cb = ((rb + 7) / 8 + 1) * 32 * P.NCW;
        const int plane = hd > cb ? hd : cb;
        for (int want = 4096; want >= 1024; want >>= 1) {
            int lc = plane > want ? plane : want;
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
    if (T == 0) {
        // No cluster of bands holds the map in shared memory.  Tiled form: row tiles of `own` rows, each run by its own
        // 8-band cluster on the sub-image [own rows +- apron); exact whenever a tile's run takes at most
        // apron / (2 R) rounds (8 here), checked on the device; the L2-resident kernel of detect_large.cu then redoes
        // only the images that raised their flag.  EINX_DETECT_TILED=0 keeps the old path.
        static const bool tiled_off = getenv("EINX_DETECT_TILED") && atoi(getenv("EINX_DETECT_TILED")) == 0;
        const int Tt = kMaxCluster;
        int rb_max = 0;  // largest band (rows) that fits one CTA
        for (int rb = 2 * (R > 0 ? R : 1); rb + 2 * R < 32768 && Wp < 65536; ++rb) {
            if (list_entries(rb) > 0) rb_max = rb; else break;
        }
        const int apron = 16 * (R > 0 ? R : 1);          // 8 rounds of 2 R rows
        const int own = Tt * rb_max - 2 * apron;         // own rows of an interior tile
        if (!tiled_off && R > 0 && rb_max >= 2 * R + 8 && own >= 4 * R && B <= 65536) {
            P.NT = (Hp + own - 1) / own;
            P.tile_rows = own;
            P.apron = apron;
            P.T = Tt;
            int hs_max = 0;  // tallest sub-image over the tiles
            for (int t = 0; t < P.NT; ++t) {
                const int lo = t * own, hi = lo + own < Hp ? lo + own : Hp;
                const int y0 = lo - apron > 0 ? lo - apron : 0, y1 = hi + apron < Hp ? hi + apron : Hp;
                if (y1 - y0 > hs_max) hs_max = y1 - y0;
            }
            P.RB = (hs_max + Tt - 1) / Tt;
            P.LC = list_entries(P.RB);
            P.NSEG = kWarps / P.NCW > 0 ? kWarps / P.NCW : 1;
            if (P.NSEG > P.RB) P.NSEG = P.RB;
            P.SR = ((P.RB + P.NSEG - 1) / P.NSEG + 7) / 8 * 8;
            P.tail_smem = 0;
            P.seg_cap = ((P.RB + R) / (R + 1)) * ((Wp + R) / (R + 1));
            const size_t nseg = (size_t)B * P.NT * Tt;
            const size_t seg_elems = nseg * P.seg_cap;
            const size_t list_bytes = align_up(seg_elems * 4, 256);
            const size_t cnt_bytes = align_up(nseg * 4, 256);
            int rc = einx_ws_reserve(ctx, 2 * list_bytes + cnt_bytes, stream);
            if (rc) return rc;
            unsigned char* ws = (unsigned char*)ctx->ws;
            P.surv_val = (float*)ws;
            P.surv_idx = (int32_t*)(ws + list_bytes);
            P.seg_cnt = (int32_t*)(ws + 2 * list_bytes);
            P.redo = ctx->redo_flags;
            EINX_CUDA(ctx, cudaMemsetAsync(P.redo, 0, (size_t)B * sizeof(int32_t), stream));
            const size_t smem = smem_for(P.RB, P.LC);
            // one profiling bracket around the three launches (tiles, tail, conditional redo)
            const int prof = ctx->profile;
            einx_prof_begin(ctx, 1, stream);
            ctx->profile = 0;
            struct Restore { einx_ctx* c; int p; cudaStream_t s; ~Restore() { c->profile = p; einx_prof_end(c, 1, s); } } restore{ctx, prof, stream};
            rc = dispatch_radius<true>(ctx, R, P, smem, stream);
            if (rc) return rc;
            if (P.NT * Tt > kTailThreads) return einx_fail(ctx, EINX_ERR_UNSUPPORTED, "einx_detect: %d tiles per image", P.NT);
            for (int i = 0; i < nsides; ++i)   // the dense map (when asked for) is zeros plus the kept survivors
                if (sides[i].nms_map) EINX_CUDA(ctx, cudaMemsetAsync(sides[i].nms_map, 0, sizeof(float) * (size_t)Bside * Hp * Wp, stream));
            // (every survivor of an image sits in a register of the tail CTA: items of 32 entries, kTailHeld per warp)
            if ((size_t)P.scap / 32 + (size_t)P.NT * Tt > (size_t)kTailHeld * kTailWarps)
                return einx_fail(ctx, EINX_ERR_UNSUPPORTED, "einx_detect: %dx%d map has more survivors than the tiled tail holds", Hp, Wp);
            const size_t tail_smem = align_up(sizeof(TailShared), 16) + (size_t)(2 * P.NT * Tt + 1 + kTailHeld * kTailWarps) * sizeof(int) + (size_t)kTailHeld * kTailWarps * sizeof(unsigned short);
            nms_tile_tail_kernel<<<B, kTailThreads, tail_smem, stream>>>(P);
            EINX_CHECK_LAUNCH(ctx);
            // the exact path for the images (if any) whose tiles needed more rounds than their apron covers
            for (int i = 0; i < nsides; ++i) {
                rc = einx_detect_large(ctx, sides[i].score, sides[i].mask, Bside, Hp, Wp, nms_radius, border, prob_thresh, top_k,
                                       sides[i].nms_map, sides[i].kpts, kcap, sides[i].counts, stream_, P.redo + (size_t)i * Bside);
                if (rc) return rc;
            }
            return EINX_OK;
        }
        for (int i = 0; i < nsides; ++i) {
            const int rc = einx_detect_large(ctx, sides[i].score, sides[i].mask, Bside, Hp, Wp, nms_radius, border, prob_thresh,
                                             top_k, sides[i].nms_map, sides[i].kpts, kcap, sides[i].counts, stream_);
            if (rc) return rc;
        }
        return EINX_OK;
    }
    if (force_t <= 0)
        while (T * 2 <= kMaxCluster && (long long)B * T * 2 <= ctx->num_sms && Hp / (T * 2) >= 2 * (R > 0 ? R : 1) + 8) T *= 2;
    P.T = T;
    P.RB = (Hp + T - 1) / T;
    P.LC = list_entries(P.RB);
    if (P.LC == 0) return einx_fail(ctx, EINX_ERR_UNSUPPORTED, "einx_detect: %dx%d band does not fit shared memory", P.RB, Wp);
    // sweep runs: NSEG x NCW units over the warps of a CTA
    P.NSEG = kWarps / P.NCW > 0 ? kWarps / P.NCW : 1;
    if (P.NSEG > P.RB) P.NSEG = P.RB;
    P.SR = ((P.RB + P.NSEG - 1) / P.NSEG + 7) / 8 * 8;  // runs start on multiples of 8 rows (candidate words have one writer)
    // survivor lists of the tail: shared-memory scratch (UB + list buffer) when a single CTA owns the image
    const size_t scratch = (size_t)(P.RB + 2 * R) * P.SB * 4 + (size_t)P.LC * 4;
    P.tail_smem = (T == 1 && (size_t)P.scap * 8 <= scratch) ? 1 : 0;

    const size_t list_bytes = align_up((size_t)B * P.scap * 4, 256);
    int rc = einx_ws_reserve(ctx, P.tail_smem ? 256 : 2 * list_bytes, stream);
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

}  // namespace

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
    const NmsSide side = {score, mask, nms_map, kpts, counts};
    return detect_impl(ctx, &side, 1, B, Hp, Wp, nms_radius, border, prob_thresh, top_k, kcap, stream_);
}

extern "C" int einx_detect_pair(einx_ctx* ctx, float* score0, float* score1, const uint8_t* mask0, const uint8_t* mask1,
                                int B, int Hp, int Wp, int nms_radius, int border, float prob_thresh, int top_k,
                                float* nms_map0, float* nms_map1, float* kpts0, float* kpts1, int kcap, int32_t* counts0,
                                int32_t* counts1, einx_stream stream_) {
    if (!ctx) return EINX_ERR_INVALID;
    if (B < 0 || Hp <= 0 || Wp <= 0 || nms_radius < 0 || border < 0 || kcap < 0)
        return einx_fail(ctx, EINX_ERR_INVALID, "einx_detect_pair: bad argument B=%d Hp=%d Wp=%d r=%d border=%d kcap=%d", B,
                         Hp, Wp, nms_radius, border, kcap);
    if (B == 0) return EINX_OK;
    if (!score0 || !score1 || !kpts0 || !kpts1 || !counts0 || !counts1)
        return einx_fail(ctx, EINX_ERR_INVALID, "einx_detect_pair: NULL pointer argument");
    if ((long long)Hp * Wp > (1ll << 30)) return einx_fail(ctx, EINX_ERR_UNSUPPORTED, "einx_detect_pair: map too large");
    if (nms_radius > 8) return einx_fail(ctx, EINX_ERR_UNSUPPORTED, "einx_detect_pair: nms_radius %d not in [0, 8]", nms_radius);
    const NmsSide sides[2] = {{score0, mask0, nms_map0, kpts0, counts0}, {score1, mask1, nms_map1, kpts1, counts1}};
    return detect_impl(ctx, sides, 2, B, Hp, Wp, nms_radius, border, prob_thresh, top_k, kcap, stream_);
}
