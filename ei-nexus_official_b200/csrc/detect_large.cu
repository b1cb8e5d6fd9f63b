// Detection post-processing for maps TOO LARGE for the shared-memory bands of detect.cu (e.g. 1280x720):
// border removal -> iterative NMS fixpoint -> top-k threshold -> raster-ordered keypoint rows, with the
// value plane and the bitmaps in an L2-resident global scratch.  Semantics: reference core/modules/utils/detector_util.py:80-135,
// :138-164, :243-337, :451-484 (see include/einx.h); bit-exact for non-negative maps.
//
// One thread-block CLUSTER per image.  Each CTA owns a band of rows of the score map in shared
// memory (with an R-row halo refreshed from the neighbouring CTAs' shared memory over DSMEM every
// round), so an NMS round never touches HBM: the map is read once and the keypoints written once.
//
// A round (detector_util.py:286-335 restated):
//   lm(p)  = v(p) > 0  and  v(p) >= every window value  and  no equal value earlier in raster order
//   v(p)   = 0 for every p that has a local maximum in its window and is not one itself
// Local maxima are monotone (values only decrease), and an undecided pixel (positive, not a maximum,
// not suppressed) has no maximum in its window, so only the undecided pixels matter after a round.
// Dense rounds are separable: pass A takes the horizontal window maximum of 4 pixels per thread from
// float4 loads, pass B the vertical one of 8 rows per thread and decides the pixel; suppression is a
// dilation of the maxima bitmap on 32-bit words.  As soon as the undecided pixels of a band fit the
// worklist (i.i.d. maps: 21 % undecided after round 0, 4 % after round 1, 0.4 %, ...), rounds visit those
// pixels only.  The worklists live in the shared-memory rows of the horizontal maxima, which are dead
// outside the dense passes, so a round writes nothing to global memory and the cluster barriers have
// no stores to drain.  (Measured and rejected: following only the still-positive neighbours through a
// bitmap in worklist rounds -- the dependent bit-scan/load chains cost 4x the 81 independent loads.)
// The loop ends when no pixel of the image is undecided -- the same fixpoint the reference reaches
// when its batch-wide count of maxima stops changing.
#include <cooperative_groups.h>
#include <stdlib.h>

#include "common.cuh"
#include "detect_common.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int kThreads = 1024;
constexpr int kWarps = kThreads / 32;
constexpr int kMaxCluster = 8;
constexpr int kWorklistCap = 4096;  // late NMS rounds visit only the still-undecided pixels (per CTA)

struct DetectParams {
    float* score;
    const uint8_t* mask;
    float* nms_map;
    float* kpts;
    int32_t* counts;
    int B, Hp, Wp, border, kcap;
    int CS;      // CTAs per image (cluster size)
    int S;       // 32-column strips per row
    int WS;      // padded row stride of V in floats: 32*S + 2*PAD
    int RBmax;   // max own rows of a band
    int vec4;    // rows of `score` (and `mask`) can be moved as float4 (uchar4)
    int wl_smem; // worklists alias the shared-memory row maxima (bands large enough to host them)
    unsigned magic_s, magic_ch;  // ceil(2^32 / S), ceil(2^32 / (8*S)): t / S == umulhi(t, magic_s) for t < 2^20
    float prob_thresh;
    int use_topk;  // 1: threshold from order statistics rank_lo / rank_hi; 2: top_k >= n (thr_k = 0)
    int rank_lo, rank_hi;
    int scap;  // survivor list capacity per image
    float* surv_val;
    int32_t* surv_idx;
    unsigned int* worklists;  // [B * CS][2][kWorklistCap] entries (local row << 16 | x), L2-resident
    // global-memory variant (maps too large for a cluster's shared memory)
    float* gV;
    float* gH;
    uint32_t* gLM;
    uint32_t* gRD;
    uint32_t* gPS;
    long long* trace;  // developer aid (EINX_DETECT_TRACE=1): clock64() at phase boundaries of CTA 0
    const int32_t* only_if;  // optional [B]: images whose flag is 0 are left alone (the tiled kernel already did them)
};

#define EINX_TRACE(slot)                                                          \
    do {                                                                          \
        if (P.trace && blockIdx.x == 0 && threadIdx.x == 0 && (slot) < 128) P.trace[(slot)] = clock64(); \
    } while (0)

struct Shared {
    int flags[2];
    int xcnt[2];
    int warp_scan[kWarps + 1];
    unsigned int hist[256];
    unsigned int sel_prefix, sel_rank, sel_min, sel_cnt;
    float thr;
    int wl_n[2];  // undecided pixels found by the last suppression pass (list valid while <= kWorklistCap)
};

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
        int w = scratch[lane];
        int winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int n = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += n;
        }
        scratch[lane] = winc - w;
        if (lane == 31) scratch[kWarps] = winc;
    }
    __syncthreads();
    total = scratch[kWarps];
    return inc - v + scratch[warp];
}

// j-th smallest (0-based) of the positive floats in list[0..n) via 4 radix passes on their bit
// patterns, then the next order statistic; every thread returns the same (a, b).
__device__ void select_two(const float* __restrict__ list, int n, int j, bool need_next, Shared& sh, float& a_out,
                           float& b_out) {
    if (threadIdx.x == 0) { sh.sel_prefix = 0; sh.sel_rank = (unsigned)j; }
    // the usual list (a few thousand survivors) is read from L2 once and kept in registers
    constexpr int kHeld = 4;
    const bool held = n <= kHeld * kThreads;
    unsigned ev[kHeld];
#pragma unroll
    for (int u = 0; u < kHeld; ++u) {
        const int i = threadIdx.x + u * kThreads;
        ev[u] = (held && i < n) ? __float_as_uint(__ldcg(list + i)) : 0u;
    }
    unsigned mask = 0;
    for (int shift = 24; shift >= 0; shift -= 8) {
        for (int i = threadIdx.x; i < 256; i += kThreads) sh.hist[i] = 0;
        __syncthreads();
        const unsigned prefix = sh.sel_prefix;
        if (held) {
#pragma unroll
            for (int u = 0; u < kHeld; ++u)
                if (threadIdx.x + u * kThreads < n && (ev[u] & mask) == prefix) atomicAdd(&sh.hist[(ev[u] >> shift) & 255u], 1u);
        } else {
            for (int i = threadIdx.x; i < n; i += kThreads) {
                const unsigned e = __float_as_uint(__ldcg(list + i));
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
                const unsigned n = __shfl_up_sync(0xffffffffu, inc, o);
                if ((int)threadIdx.x >= o) inc += n;
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
                const unsigned e = __float_as_uint(__ldcg(list + i));
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
// 16-byte aligned and the passes below can move float4.
template <int R>
struct Geo {
    static constexpr int PAD = (R + 3) / 4 * 4;
};

// Horizontal window maxima of 4 neighbouring pixels.  a[] holds the 4 + 2*PAD values starting PAD
// to the left of the first pixel; o[i] = max a[PAD+i-R .. PAD+i+R].  The values shared by all four
// windows are reduced once, then extended left / right: 2R+7 max operations for 4 outputs.
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
        o[1] = fmaxf(fmaxf(common, l2), r1);
        o[2] = fmaxf(fmaxf(common, l1), r2);
        o[3] = fmaxf(common, r3);
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float m = a[PAD + i];
#pragma unroll
            for (int d = 1; d <= R; ++d) m = fmaxf(m, fmaxf(a[PAD + i - d], a[PAD + i + d]));
            o[i] = m;
        }
    }
}

// Vertical window maxima of 8 consecutive rows from the 8 + 2R row maxima a[] above/below them:
// o[i] = max a[i .. i+2R].  For 2R >= 8 every window straddles the 7|8 boundary, so a suffix scan of
// a[0..7] and a prefix scan of a[8..] give all eight with 2R+14 operations.
template <int R>
__device__ __forceinline__ void vmax8(const float* a, float (&o)[8]) {
    if constexpr (R >= 4) {
        float suf[8];
        suf[7] = a[7];
#pragma unroll
        for (int i = 6; i >= 0; --i) suf[i] = fmaxf(a[i], suf[i + 1]);
        float pre = a[8];
#pragma unroll
        for (int k = 9; k <= 2 * R; ++k) pre = fmaxf(pre, a[k]);
        o[0] = fmaxf(suf[0], pre);
#pragma unroll
        for (int i = 1; i < 8; ++i) {
            pre = fmaxf(pre, a[i + 2 * R]);
            o[i] = fmaxf(suf[i], pre);
        }
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float m = a[i];
#pragma unroll
            for (int k = 1; k <= 2 * R; ++k) m = fmaxf(m, a[i + k]);
            o[i] = m;
        }
    }
}

template <int R, bool SMEM>
// 48 registers (no spills) instead of the 64 a 1024-thread launch bound allows: the CTA then leaves a quarter of the
// register file free, so CTAs of the concurrent voxel kernels can share its SM and use the issue slots the NMS
// rounds leave idle (the step runs voxelisation and both detect chains on three streams)
__global__ void __maxnreg__(48) detect_kernel(const DetectParams P) {
    constexpr int PAD = Geo<R>::PAD;
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int CS = P.CS;
    const int b = blockIdx.x / CS;
    if (P.only_if && P.only_if[b] == 0) return;  // (uniform over the cluster: nobody waits for anybody)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int S = P.S, WS = P.WS, HS = 32 * P.S, Hp = P.Hp, Wp = P.Wp;

    // balanced row bands
    const int base_rows = Hp / CS, rem = Hp % CS;
    const int nrows = base_rows + (rank < rem ? 1 : 0);
    const int ys = rank * base_rows + min(rank, rem);
    const int nprev = base_rows + ((rank - 1) < rem ? 1 : 0);  // rows of the band above

    extern __shared__ __align__(16) unsigned char smem_raw[];
    Shared& sh = *reinterpret_cast<Shared*>(smem_raw);
    // Local row lr of every array is image row ys - R + lr: own rows are lr in [R, R + nrows), the R
    // rows on either side are the halo (copies of the neighbouring bands in the shared-memory
    // variant; simply the neighbours' rows of the same padded image in the global variant).
    float *V, *Hm;
    uint32_t *LM, *RD, *PS;
    const int lrows = P.RBmax + 2 * R;
    if (SMEM) {
        size_t o = align_up(sizeof(Shared), 16);
        V = reinterpret_cast<float*>(smem_raw + o);
        o += sizeof(float) * (size_t)lrows * WS;
        Hm = reinterpret_cast<float*>(smem_raw + o);
        o += sizeof(float) * (size_t)lrows * HS;
        LM = reinterpret_cast<uint32_t*>(smem_raw + o);
        o += sizeof(uint32_t) * (size_t)lrows * S;
        RD = reinterpret_cast<uint32_t*>(smem_raw + o);
        o += sizeof(uint32_t) * (size_t)lrows * S;
        PS = reinterpret_cast<uint32_t*>(smem_raw + o);
    } else {
        const size_t img_rows = (size_t)Hp + 2 * R;
        V = P.gV + ((size_t)b * img_rows + ys) * WS;
        Hm = P.gH + ((size_t)b * img_rows + ys) * HS;
        LM = P.gLM + ((size_t)b * img_rows + ys) * S;
        RD = P.gRD + ((size_t)b * img_rows + ys) * S;
        PS = P.gPS + ((size_t)b * img_rows + ys) * S;
    }
    // ---- load the band: border + mask zeroing (in place on `score`), zero padding ---------- //
    if (SMEM) {
        float4* v4 = reinterpret_cast<float4*>(V);
        for (int i = tid; i < lrows * WS / 4; i += kThreads) v4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int i = tid; i < lrows * S; i += kThreads) { LM[i] = 0u; RD[i] = 0u; PS[i] = 0u; }
    }
    if (tid == 0) { sh.flags[0] = sh.flags[1] = 0; sh.xcnt[0] = sh.xcnt[1] = 0; sh.wl_n[0] = sh.wl_n[1] = 0; }
    __syncthreads();
    {
        float* simg = P.score + (size_t)b * Hp * Wp;
        const uint8_t* mimg = P.mask ? P.mask + (size_t)b * Hp * Wp : nullptr;
        const int bd = P.border;
        for (int lr = warp; lr < nrows; lr += kWarps) {
            const int y = ys + lr;
            const bool rowkill = (y < bd) | (y >= Hp - bd);
            float* srow = simg + (size_t)y * Wp;
            const uint8_t* mrow = mimg ? mimg + (size_t)y * Wp : nullptr;
            float* vrow = V + (size_t)(lr + R) * WS + PAD;
            if (P.vec4) {
                for (int c = lane; c < (Wp >> 2); c += 32) {
                    float4 v = *reinterpret_cast<const float4*>(srow + 4 * c);
                    float e[4] = {v.x, v.y, v.z, v.w};
                    uchar4 m4 = make_uchar4(1, 1, 1, 1);
                    if (mrow) m4 = *reinterpret_cast<const uchar4*>(mrow + 4 * c);
                    const unsigned char mm[4] = {m4.x, m4.y, m4.z, m4.w};
                    bool changed = false;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int x = 4 * c + j;
                        const bool kill = rowkill | (x < bd) | (x >= Wp - bd) | (mm[j] == 0);
                        if (kill) {
                            changed |= (e[j] != 0.0f);
                            e[j] = 0.0f;
                        }
                    }
                    v = make_float4(e[0], e[1], e[2], e[3]);
                    if (changed) *reinterpret_cast<float4*>(srow + 4 * c) = v;
                    *reinterpret_cast<float4*>(vrow + 4 * c) = v;
                }
            } else {
                for (int x = lane; x < Wp; x += 32) {
                    float v = srow[x];
                    bool kill = rowkill | (x < bd) | (x >= Wp - bd);
                    if (mrow) kill |= (mrow[x] == 0);
                    if (kill) {
                        if (v != 0.0f) srow[x] = 0.0f;
                        v = 0.0f;
                    }
                    vrow[x] = v;
                }
            }
        }
    }
    if (!SMEM) __threadfence();
    EINX_TRACE(0);
    int trace_slot = 1;

    // ---- NMS rounds ------------------------------------------------------------------------ //
    if constexpr (R > 0) {
        constexpr int P2 = 2 * R + 1;
        unsigned int* const wl0 = (SMEM && P.wl_smem) ? reinterpret_cast<unsigned int*>(Hm) : P.worklists + (size_t)blockIdx.x * 2 * kWorklistCap;
        int wl_cur = 0;         // list buffer holding the current undecided set
        bool wl_mode = false;   // CTA-uniform: this round runs on the worklist instead of dense passes
        const int CH = 8 * S;   // float4 chunks per row
        for (int round = 0;; ++round) {
            cluster.sync();  // S1: every band's V (and the previous round's flag) is final
            trace_slot = 1 + 8 * round;
            EINX_TRACE(trace_slot); ++trace_slot;
            if (round > 0) {
                int any = 0;
                for (int r = 0; r < CS; ++r) any |= *cluster.map_shared_rank(&sh.flags[(round - 1) & 1], r);
                if (!any) break;
            }
            if (SMEM) {
                if (rank > 0) {
                    const float4* src = reinterpret_cast<const float4*>(cluster.map_shared_rank(V, rank - 1) + (size_t)nprev * WS);
                    float4* dst = reinterpret_cast<float4*>(V);
                    for (int i = tid; i < R * WS / 4; i += kThreads) dst[i] = src[i];
                }
                if (rank < CS - 1) {
                    const float4* src = reinterpret_cast<const float4*>(cluster.map_shared_rank(V, rank + 1) + (size_t)R * WS);
                    float4* dst = reinterpret_cast<float4*>(V + (size_t)(R + nrows) * WS);
                    for (int i = tid; i < R * WS / 4; i += kThreads) dst[i] = src[i];
                }
                __syncthreads();
            }
            if (!wl_mode) {
                // pass A: Hm = horizontal window maximum of every local row, halo included (in the
                // global variant the halo rows are the neighbours' own rows: both CTAs then store
                // identical values, so no cross-CTA ordering is needed inside a round)
                {
                    for (int t = tid; t < (nrows + 2 * R) * CH; t += kThreads) {
                        const int row = (int)__umulhi((unsigned)t, P.magic_ch);
                        const int ch = t - row * CH;
                        const float4* src = reinterpret_cast<const float4*>(V + (size_t)row * WS) + ch;
                        float a[4 + 2 * PAD];
#pragma unroll
                        for (int k = 0; k < 1 + PAD / 2; ++k) {
                            const float4 q = src[k];
                            a[4 * k] = q.x; a[4 * k + 1] = q.y; a[4 * k + 2] = q.z; a[4 * k + 3] = q.w;
                        }
                        float o[4];
                        hmax4<R>(a, o);
                        *(reinterpret_cast<float4*>(Hm + (size_t)row * HS) + ch) = make_float4(o[0], o[1], o[2], o[3]);
                    }
                }
                __syncthreads();
                EINX_TRACE(trace_slot); ++trace_slot;
                // pass B: a warp decides 8 rows x 32 columns.  lm = v > 0, v == window max, and no
                // equal value earlier in raster order (rows above: their row maxima; same row: the R
                // values to the left) -- the first-occurrence argmax of detector_util.py:298-308.
                {
                    const int nblk = (nrows + 7) >> 3;
                    const int avail_all = nrows + 2 * R;
                    for (int t = warp; t < nblk * S; t += kWarps) {
                        const int rb = (int)__umulhi((unsigned)t, P.magic_s);
                        const int s = t - rb * S;
                        const int r0 = rb * 8;
                        const int x = 32 * s + lane;
                        const int avail = avail_all - r0;
                        float a[8 + 2 * R];
#pragma unroll
                        for (int k = 0; k < 8 + 2 * R; ++k) a[k] = (k < avail) ? Hm[(size_t)(r0 + k) * HS + x] : 0.0f;
                        float m[8];
                        vmax8<R>(a, m);
                        uint32_t lm_mine = 0, pos_mine = 0;
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int lr = r0 + i;
                            const float* crow = V + (size_t)(lr + R) * WS + PAD + x;
                            const float vc = (lr < nrows) ? crow[0] : 0.0f;
                            const bool pos = vc > 0.0f;
                            bool lm = pos && (vc == m[i]);
                            if (lm) {
                                float early = crow[-1];
#pragma unroll
                                for (int d = 2; d <= R; ++d) early = fmaxf(early, crow[-d]);
#pragma unroll
                                for (int k = 0; k < R; ++k) early = fmaxf(early, a[i + k]);
                                lm = early < vc;
                            }
                            const uint32_t lb = __ballot_sync(0xffffffffu, lm);
                            const uint32_t pb = __ballot_sync(0xffffffffu, pos);
                            if (lane == i) { lm_mine = lb; pos_mine = pb; }
                        }
                        if (lane < 8 && r0 + lane < nrows) {
                            LM[(size_t)(r0 + lane + R) * S + s] = lm_mine;
                            PS[(size_t)(r0 + lane + R) * S + s] = pos_mine;
                        }
                    }
                }
            } else {
                // worklist round, phase 1: each undecided pixel scans its own window
                const int n = sh.wl_n[wl_cur];
                for (int e = tid; e < n; e += kThreads) {
                    const unsigned ent = wl0[wl_cur * kWorklistCap + e];
                    const int lr = (int)(ent >> 16), x = (int)(ent & 0xffffu);
                    const float* crow = V + (size_t)(lr + R) * WS + x + PAD;
                    const float vc = crow[0];
                    float emax = 0.0f, lmax = 0.0f;  // raster-earlier / raster-later halves of the window
#pragma unroll
                    for (int dy = 1; dy <= R; ++dy) {
#pragma unroll
                        for (int dx = -R; dx <= R; ++dx) {
                            emax = fmaxf(emax, crow[-dy * WS + dx]);
                            lmax = fmaxf(lmax, crow[dy * WS + dx]);
                        }
                    }
#pragma unroll
                    for (int d = 1; d <= R; ++d) { emax = fmaxf(emax, crow[-d]); lmax = fmaxf(lmax, crow[d]); }
                    if (vc > emax && vc >= lmax) {
                        atomicOr(&LM[(size_t)(lr + R) * S + (x >> 5)], 1u << (x & 31));
                        wl0[wl_cur * kWorklistCap + e] = ent | 0x80000000u;  // decided: a local maximum (rows < 32768)
                    }
                }
            }
            if (!SMEM) __threadfence();
            EINX_TRACE(trace_slot); ++trace_slot;
            cluster.sync();  // S2: own-row maxima bits are ready in every band
            if (SMEM) {
                if (rank > 0) {
                    const uint32_t* src = cluster.map_shared_rank(LM, rank - 1) + (size_t)nprev * S;
                    for (int i = tid; i < R * S; i += kThreads) LM[i] = src[i];
                }
                if (rank < CS - 1) {
                    const uint32_t* src = cluster.map_shared_rank(LM, rank + 1) + (size_t)R * S;
                    uint32_t* dst = LM + (size_t)(R + nrows) * S;
                    for (int i = tid; i < R * S; i += kThreads) dst[i] = src[i];
                }
            }
            const int wl_next = wl_cur ^ 1;
            if (tid == 0) sh.wl_n[wl_next] = 0;
            __syncthreads();
            EINX_TRACE(trace_slot); ++trace_slot;
            int und = 0;
            if (!wl_mode) {
                // horizontal dilation of the maxima bits, on words, for every local row
                for (int t = tid; t < (nrows + 2 * R) * S; t += kThreads) {
                    const int rr = (int)__umulhi((unsigned)t, P.magic_s);
                    const int s = t - rr * S;
                    const size_t i = (size_t)rr * S + s;
                    const uint32_t w = LM[i];
                    const uint32_t wl = s > 0 ? LM[i - 1] : 0u;
                    const uint32_t wr = s < S - 1 ? LM[i + 1] : 0u;
                    uint32_t acc = w;
#pragma unroll
                    for (int d = 1; d <= R; ++d) acc |= (w >> d) | (wr << (32 - d)) | (w << d) | (wl >> (32 - d));
                    RD[i] = acc;
                }
            }
            if (!wl_mode) {
                __syncthreads();
                EINX_TRACE(trace_slot); ++trace_slot;
                // suppression set + undecided census per own word; the undecided pixels become the
                // next round's worklist
                for (int t = tid; t < nrows * S; t += kThreads) {
                    const int lr = (int)__umulhi((unsigned)t, P.magic_s);
                    const int s = t - lr * S;
                    uint32_t dil = 0;
#pragma unroll
                    for (int dy = 0; dy < P2; ++dy) dil |= RD[(size_t)(lr + dy) * S + s];
                    const size_t i = (size_t)(lr + R) * S + s;
                    const uint32_t posw = PS[i];
                    const uint32_t sup = dil & ~LM[i] & posw;  // positive pixels a neighbouring maximum suppresses
                    uint32_t u = posw & ~dil;                   // positive, not a maximum, not suppressed
                    PS[i] = sup;
                    if (u) {
                        und = 1;
                        int pos = atomicAdd(&sh.wl_n[wl_next], __popc(u));
                        while (u) {
                            const int bit = __ffs(u) - 1;
                            u &= u - 1;
                            if (pos < kWorklistCap) wl0[wl_next * kWorklistCap + pos] = ((unsigned)lr << 16) | (unsigned)(32 * s + bit);
                            ++pos;
                        }
                    }
                }
                __syncthreads();
                EINX_TRACE(trace_slot); ++trace_slot;
                // apply: zero the suppressed pixels, 4 at a time
                for (int t = tid; t < nrows * CH; t += kThreads) {
                    const int lr = (int)__umulhi((unsigned)t, P.magic_ch);
                    const int ch = t - lr * CH;
                    const uint32_t bits = (PS[(size_t)(lr + R) * S + (ch >> 3)] >> ((ch & 7) * 4)) & 0xfu;
                    if (bits) {
                        float4* cell = reinterpret_cast<float4*>(V + (size_t)(lr + R) * WS + PAD) + ch;
                        float4 v = *cell;
                        if (bits & 1u) v.x = 0.0f;
                        if (bits & 2u) v.y = 0.0f;
                        if (bits & 4u) v.z = 0.0f;
                        if (bits & 8u) v.w = 0.0f;
                        *cell = v;
                    }
                }
            } else {
                // worklist round, phase 2: drop pixels that now have a local maximum in their window
                const int n = sh.wl_n[wl_cur];
                for (int e = tid; e < n; e += kThreads) {
                    const unsigned ent = wl0[wl_cur * kWorklistCap + e];
                    if (ent & 0x80000000u) continue;  // became a local maximum in phase 1
                    const int lr = (int)(ent >> 16), x = (int)(ent & 0xffffu);
                    const int xl = x - R;
                    const int wi = xl >> 5;  // arithmetic: -1 for the left border
                    const int sh_ = xl - 32 * wi;
                    uint32_t any = 0;
#pragma unroll
                    for (int dy = -R; dy <= R; ++dy) {
                        const uint32_t* rowp = LM + (size_t)(lr + R + dy) * S;
                        const uint32_t w0 = (wi >= 0 && wi < S) ? rowp[wi] : 0u;
                        const uint32_t w1 = (wi + 1 >= 0 && wi + 1 < S) ? rowp[wi + 1] : 0u;
                        const unsigned long long both = (unsigned long long)w0 | ((unsigned long long)w1 << 32);
                        any |= (uint32_t)(both >> sh_) & ((1u << P2) - 1u);
                    }
                    if (any) {
                        V[(size_t)(lr + R) * WS + x + PAD] = 0.0f;
                    } else {
                        const int pos = atomicAdd(&sh.wl_n[wl_next], 1);
                        wl0[wl_next * kWorklistCap + pos] = ent;  // pos < n <= kWorklistCap
                        und = 1;
                    }
                }
            }
            und = __syncthreads_or(und);
            EINX_TRACE(trace_slot); ++trace_slot;
            wl_mode = sh.wl_n[wl_next] <= kWorklistCap;
            wl_cur = wl_next;
            if (tid == 0) sh.flags[round & 1] = und;
            if (!SMEM) __threadfence();
        }
    } else {
        // no NMS: survivors are simply the positive pixels
        __syncthreads();
        for (int wi = warp; wi < nrows * S; wi += kWarps) {
            const int lr = wi / S, s = wi - lr * S;
            const float v = V[(size_t)(lr + R) * WS + 32 * s + lane + PAD];
            const unsigned bits = __ballot_sync(0xffffffffu, v > 0.0f);
            if (lane == 0) LM[(size_t)(lr + R) * S + s] = bits;
        }
        __syncthreads();
    }

    EINX_TRACE(120);
    // ---- survivors -> ordered per-image list (global workspace) ------------------------------ //
    // At the fixpoint every positive pixel is a local maximum, so the maxima bits of the last
    // round are exactly the survivors.
    float* slist = P.surv_val + (size_t)b * P.scap;
    int32_t* sidx = P.surv_idx + (size_t)b * P.scap;
    const int nwords = nrows * S;
    int own = 0;
    {
        int c = 0;
        for (int wi = tid; wi < nwords; wi += kThreads) c += __popc(LM[(size_t)R * S + wi]);
        c = __reduce_add_sync(0xffffffffu, c);
        if (lane == 0 && c) atomicAdd(&sh.xcnt[0], c);
    }
    cluster.sync();
    int offset = 0, total = 0;
    for (int r = 0; r < CS; ++r) {
        const int c = *cluster.map_shared_rank(&sh.xcnt[0], r);
        if (r < rank) offset += c;
        total += c;
    }
    own = sh.xcnt[0];
    {
        int run = offset;
        for (int base = 0; base < nwords; base += kThreads) {
            const int wi = base + tid;
            const uint32_t w = wi < nwords ? LM[(size_t)R * S + wi] : 0u;
            int tot;
            int pos = run + block_excl_scan(__popc(w), sh.warp_scan, tot);
            if (w) {
                const int lr = wi / S, s = wi - lr * S;
                uint32_t bits = w;
                while (bits) {
                    const int bit = __ffs(bits) - 1;
                    bits &= bits - 1;
                    const int x = 32 * s + bit;
                    if (pos < P.scap) {
                        slist[pos] = V[(size_t)(lr + R) * WS + x + PAD];
                        sidx[pos] = (ys + lr) * Wp + x;
                    }
                    ++pos;
                }
            }
            run += tot;
        }
    }
    __threadfence();
    EINX_TRACE(121);
    cluster.sync();  // the whole image's list is visible
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
                select_two(slist, total, P.rank_lo - zeros, P.rank_hi != P.rank_lo, sh, a, bq);
            } else {  // lo falls on a zero, hi on the smallest survivor
                float dummy;
                select_two(slist, total, 0, false, sh, bq, dummy);
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
        for (int i = tid; i < own; i += kThreads) c += (__ldcg(slist + offset + i) > thr) ? 1 : 0;
        c = __reduce_add_sync(0xffffffffu, c);
        if (lane == 0 && c) atomicAdd(&sh.xcnt[1], c);
    }
    cluster.sync();
    int koff = 0, ktotal = 0;
    for (int r = 0; r < CS; ++r) {
        const int c = *cluster.map_shared_rank(&sh.xcnt[1], r);
        if (r < rank) koff += c;
        ktotal += c;
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
                v = __ldcg(slist + offset + i);
                idx = __ldcg(sidx + offset + i);
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
        float* out = P.nms_map + (size_t)b * Hp * Wp;
        for (int e = tid; e < nrows * Wp; e += kThreads) {
            const int lr = e / Wp, x = e - lr * Wp;
            const float v = V[(size_t)(lr + R) * WS + x + PAD];
            out[(size_t)(ys + lr) * Wp + x] = v > thr ? v : 0.0f;
        }
    }
    EINX_TRACE(124);
    cluster.sync();  // nobody leaves while a neighbour may still read its shared memory
    EINX_TRACE(125);
}

template <int R, bool SMEM>
int launch_detect(einx_ctx* ctx, const DetectParams& P, size_t smem, cudaStream_t stream) {
    auto kern = detect_kernel<R, SMEM>;
    EINX_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(P.B * P.CS);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = P.CS;
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
        fprintf(stderr, "[einx_detect trace] CS=%d smem=%zu:", P.CS, smem);
        long long prev = h[0];
        for (int i = 0; i < 128; ++i)
            if (h[i]) { fprintf(stderr, " %d:+%lld", i, h[i] - prev); prev = h[i]; }
        fprintf(stderr, "\n");
        cudaMemset(P.trace, 0, sizeof(h));
    }
    return EINX_OK;
}

template <bool SMEM>
int dispatch_radius(einx_ctx* ctx, int R, const DetectParams& P, size_t smem, cudaStream_t stream) {
    switch (R) {
        case 0: return launch_detect<0, SMEM>(ctx, P, smem, stream);
        case 1: return launch_detect<1, SMEM>(ctx, P, smem, stream);
        case 2: return launch_detect<2, SMEM>(ctx, P, smem, stream);
        case 3: return launch_detect<3, SMEM>(ctx, P, smem, stream);
        case 4: return launch_detect<4, SMEM>(ctx, P, smem, stream);
        case 5: return launch_detect<5, SMEM>(ctx, P, smem, stream);
        case 6: return launch_detect<6, SMEM>(ctx, P, smem, stream);
        case 7: return launch_detect<7, SMEM>(ctx, P, smem, stream);
        case 8: return launch_detect<8, SMEM>(ctx, P, smem, stream);
    }
    return einx_fail(ctx, EINX_ERR_UNSUPPORTED, "einx_detect: nms_radius %d not in [0, 8]", R);
}

}  // namespace

int einx_detect_large(einx_ctx* ctx, float* score, const uint8_t* mask, int B, int Hp, int Wp, int nms_radius,
                           int border, float prob_thresh, int top_k, float* nms_map, float* kpts, int kcap,
                           int32_t* counts, einx_stream stream_, const int32_t* only_if) {
    if (!ctx) return EINX_ERR_INVALID;
    if (B < 0 || Hp <= 0 || Wp <= 0 || nms_radius < 0 || border < 0 || kcap < 0)
        return einx_fail(ctx, EINX_ERR_INVALID, "einx_detect: bad argument B=%d Hp=%d Wp=%d r=%d border=%d kcap=%d", B,
                         Hp, Wp, nms_radius, border, kcap);
    if (B == 0) return EINX_OK;
    if (!score || !kpts || !counts) return einx_fail(ctx, EINX_ERR_INVALID, "einx_detect: NULL pointer argument");
    if ((long long)Hp * Wp > (1ll << 30)) return einx_fail(ctx, EINX_ERR_UNSUPPORTED, "einx_detect: map too large");
    DeviceGuard guard(ctx->device);
    cudaStream_t stream = (cudaStream_t)stream_;
    const int R = nms_radius;

    DetectParams P = {};
    P.score = score; P.mask = mask; P.nms_map = nms_map; P.kpts = kpts; P.counts = counts;
    P.B = B; P.Hp = Hp; P.Wp = Wp; P.border = border; P.kcap = kcap;
    P.only_if = only_if;
    P.S = (Wp + 31) / 32;
    const int PAD = (R + 3) / 4 * 4;
    P.WS = 32 * P.S + 2 * PAD;
    P.magic_s = (unsigned)((0x100000000ull + P.S - 1) / P.S);
    P.magic_ch = (unsigned)((0x100000000ull + 8 * P.S - 1) / (8 * P.S));
    P.vec4 = (Wp % 4 == 0) && ((uintptr_t)score % 16 == 0) && (!mask || (uintptr_t)mask % 4 == 0);
    P.prob_thresh = prob_thresh;
    const int n = Hp * Wp;
    if (top_k > 0) {
        if (top_k >= n) P.use_topk = 2;
        else { P.use_topk = 1; einx_topk_ranks(n, top_k, &P.rank_lo, &P.rank_hi); }
    }
    P.scap = R == 0 ? n : ((Hp + R) / (R + 1)) * ((Wp + R) / (R + 1));

    // the value plane and the bitmaps live in the global scratch; one cluster of row bands per image
    const size_t fixed = align_up(sizeof(Shared), 16);
    const bool use_smem = false;
    int CS = kMaxCluster;
    while (CS > 1 && Hp / CS < (R > 0 ? R : 1)) CS /= 2;
    P.CS = CS;
    P.RBmax = (Hp + CS - 1) / CS;
    if ((long long)(P.RBmax + 2 * R) * 8 * P.S >= (1 << 20))
        return einx_fail(ctx, EINX_ERR_UNSUPPORTED, "einx_detect: %dx%d map too large for one cluster", Hp, Wp);

    // workspace: survivor lists (+ padded global image, row maxima and bitmaps for the large-map variant)
    const size_t list_bytes = align_up((size_t)B * P.scap * 4, 256);
    P.wl_smem = use_smem && (size_t)(P.RBmax + 2 * R) * P.S * 32 * 4 >= (size_t)2 * kWorklistCap * 4;
    const size_t wl_bytes = P.wl_smem ? 0 : align_up((size_t)B * CS * 2 * kWorklistCap * 4, 256);
    size_t ws_bytes = 2 * list_bytes + wl_bytes;
    const size_t img_rows = (size_t)Hp + 2 * R;
    const size_t gv_bytes = align_up((size_t)B * img_rows * P.WS * 4, 256);
    const size_t gh_bytes = align_up((size_t)B * img_rows * P.S * 32 * 4, 256);
    const size_t gw_bytes = align_up((size_t)B * img_rows * P.S * 4, 256);
    if (!use_smem) ws_bytes += gv_bytes + gh_bytes + 3 * gw_bytes;
    int rc = einx_ws_reserve(ctx, ws_bytes, stream);
    if (rc) return rc;
    unsigned char* ws = (unsigned char*)ctx->ws;
    P.surv_val = (float*)ws;
    P.surv_idx = (int32_t*)(ws + list_bytes);
    P.worklists = (unsigned int*)(ws + 2 * list_bytes);
    static const bool want_trace = getenv("EINX_DETECT_TRACE") != nullptr;
    if (want_trace) {
        static long long* trace_buf = nullptr;
        if (!trace_buf && cudaMalloc(&trace_buf, 128 * sizeof(long long)) == cudaSuccess) cudaMemset(trace_buf, 0, 128 * sizeof(long long));
        P.trace = trace_buf;
    }
    unsigned char* g = ws + 2 * list_bytes + wl_bytes;
    P.gV = (float*)g;
    P.gLM = (uint32_t*)(g + gv_bytes);
    P.gRD = (uint32_t*)(g + gv_bytes + gw_bytes);
    P.gPS = (uint32_t*)(g + gv_bytes + 2 * gw_bytes);
    P.gH = (float*)(g + gv_bytes + 3 * gw_bytes);
    EINX_CUDA(ctx, cudaMemsetAsync(P.gV, 0, gv_bytes + 3 * gw_bytes, stream));  // zero padding, empty bitmaps
    return dispatch_radius<false>(ctx, R, P, fixed, stream);
}
