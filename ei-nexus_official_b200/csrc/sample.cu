// Descriptor sampling at the kept keypoints + L2 normalisation (one warp per keypoint).
// Semantics: reference core/modules/utils/descriptor_util.py:21-28, :50-71 (gather) and :74-128
// (bilinear grid_sample, align_corners=False, zeros padding) -- see include/einx.h.
#include <cuda_fp16.h>

#include "common.cuh"

namespace {

// Optional second output: the fp16 hi / lo operands of the matcher's FP16X3 tensor-core mode (einx_mnn_split),
// hi = fp16(2^10 d), lo = fp16(2^10 d - hi), written while the descriptor is still in registers -- the matcher's
// tile pipeline then has nothing to convert.  `hi` == nullptr skips it.
struct SplitOut {
    __half* hi;
    __half* lo;
};
__device__ __forceinline__ void split_store(const SplitOut& so, size_t idx, float d) {
    const float x = d * 1024.0f;
    const __half h = __float2half_rn(x);
    so.hi[idx] = h;
    so.lo[idx] = __float2half_rn(x - __half2float(h));
}
// two values of one lane (channels 32 apart): packed conversions, half the conversion instructions of two split_store calls
__device__ __forceinline__ void split_store2(const SplitOut& so, size_t idx0, size_t idx1, float d0, float d1) {
    const float x0 = d0 * 1024.0f, x1 = d1 * 1024.0f;
    const __half2 h = __floats2half2_rn(x0, x1);
    const float2 f = __half22float2(h);
    const __half2 l = __floats2half2_rn(x0 - f.x, x1 - f.y);
    so.hi[idx0] = __low2half(h);
    so.hi[idx1] = __high2half(h);
    so.lo[idx0] = __low2half(l);
    so.lo[idx1] = __high2half(l);
}
__device__ __forceinline__ void split_store4(const SplitOut& so, size_t idx, float4 d) {  // idx % 4 == 0
    const float x0 = d.x * 1024.0f, x1 = d.y * 1024.0f, x2 = d.z * 1024.0f, x3 = d.w * 1024.0f;
    const __half2 h01 = __floats2half2_rn(x0, x1), h23 = __floats2half2_rn(x2, x3);
    const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
    const __half2 l01 = __floats2half2_rn(x0 - f01.x, x1 - f01.y), l23 = __floats2half2_rn(x2 - f23.x, x3 - f23.y);
    uint2 hv, lv;
    hv.x = *reinterpret_cast<const uint32_t*>(&h01); hv.y = *reinterpret_cast<const uint32_t*>(&h23);
    lv.x = *reinterpret_cast<const uint32_t*>(&l01); lv.y = *reinterpret_cast<const uint32_t*>(&l23);
    *reinterpret_cast<uint2*>(so.hi + idx) = hv;
    *reinterpret_cast<uint2*>(so.lo + idx) = lv;
}

constexpr int kWarpsPerBlock = 8;
constexpr int kMaxPerLane = 16;  // channels per lane held in registers: C <= 512

template <int MODE>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
sample_kernel(const float* __restrict__ raw, int C, int Hd, int Wd, float Hp, float Wp,
              const float* __restrict__ kpts, const int32_t* __restrict__ counts, int kcap, float scale,
              int normalize, float* __restrict__ desc, SplitOut so) {
    const int b = blockIdx.y;
    const int k = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (k >= kcap) return;
    float* out = desc + ((size_t)b * kcap + k) * C;
    int cnt = counts[b];
    if (cnt > kcap) cnt = kcap;
    if (k >= cnt) {  // padding rows are defined (zero) so a batched matcher can ignore them safely
        for (int c = lane; c < C; c += 32) {
            out[c] = 0.0f;
            if (so.hi) split_store(so, ((size_t)b * kcap + k) * C + c, 0.0f);
        }
        return;
    }
    const float* kp = kpts + ((size_t)b * kcap + k) * 3;
    const float py = kp[0], px = kp[1];
    const float* img = raw + (size_t)b * C * Hd * Wd;
    const size_t plane = (size_t)Hd * Wd;
    float v[kMaxPerLane];
    float ss = 0.0f;
    if (MODE == EINX_SAMPLE_GATHER) {
        // pos.floor().long() -> raw[i, :, y, x]                          (:57-60)
        int yy = (int)floorf(py), xx = (int)floorf(px);
        yy = min(max(yy, 0), Hd - 1);
        xx = min(max(xx, 0), Wd - 1);
        const float* src = img + (size_t)yy * Wd + xx;
#pragma unroll
        for (int j = 0; j < kMaxPerLane; ++j) {
            const int c = lane + 32 * j;
            v[j] = c < C ? __ldg(src + c * plane) : 0.0f;
        }
    } else {
        // pos - 0.5 -> [-1, 1] on the padded image -> grid_sample un-normalisation  (:105-120)
        const float gy = __fsub_rn(__fmul_rn(2.0f, __fdiv_rn(__fsub_rn(py, 0.5f), __fsub_rn(Hp, 1.0f))), 1.0f);
        const float gx = __fsub_rn(__fmul_rn(2.0f, __fdiv_rn(__fsub_rn(px, 0.5f), __fsub_rn(Wp, 1.0f))), 1.0f);
        const float iy = __fdiv_rn(__fsub_rn(__fmul_rn(__fadd_rn(gy, 1.0f), (float)Hd), 1.0f), 2.0f);
        const float ix = __fdiv_rn(__fsub_rn(__fmul_rn(__fadd_rn(gx, 1.0f), (float)Wd), 1.0f), 2.0f);
        const float fy = floorf(iy), fx = floorf(ix);
        const int y0 = (int)fy, x0 = (int)fx;
        const float wy1 = __fsub_rn(iy, fy), wx1 = __fsub_rn(ix, fx);
        const float wy0 = __fsub_rn(__fadd_rn(fy, 1.0f), iy), wx0 = __fsub_rn(__fadd_rn(fx, 1.0f), ix);
        const float w00 = __fmul_rn(wx0, wy0), w01 = __fmul_rn(wx1, wy0);
        const float w10 = __fmul_rn(wx0, wy1), w11 = __fmul_rn(wx1, wy1);
        const bool oy0 = (y0 >= 0) & (y0 < Hd), oy1 = (y0 + 1 >= 0) & (y0 + 1 < Hd);
        const bool ox0 = (x0 >= 0) & (x0 < Wd), ox1 = (x0 + 1 >= 0) & (x0 + 1 < Wd);
        const float* src = img + (ptrdiff_t)y0 * Wd + x0;
#pragma unroll
        for (int j = 0; j < kMaxPerLane; ++j) {
            const int c = lane + 32 * j;
            float acc = 0.0f;
            if (c < C) {
                const float* s = src + c * plane;
                // tap order nw, ne, sw, se; out-of-map taps contribute zero
                if (oy0 && ox0) acc = __fadd_rn(acc, __fmul_rn(__ldg(s), w00));
                if (oy0 && ox1) acc = __fadd_rn(acc, __fmul_rn(__ldg(s + 1), w01));
                if (oy1 && ox0) acc = __fadd_rn(acc, __fmul_rn(__ldg(s + Wd), w10));
                if (oy1 && ox1) acc = __fadd_rn(acc, __fmul_rn(__ldg(s + Wd + 1), w11));
            }
            v[j] = acc;
        }
    }
#pragma unroll
    for (int j = 0; j < kMaxPerLane; ++j) ss = fmaf(v[j], v[j], ss);
    float mul = scale;
    if (normalize) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        mul = __fdiv_rn(scale, fmaxf(sqrtf(ss), 1e-12f));  // scale / max(||v||, eps): one division per keypoint
    }
#pragma unroll
    for (int j = 0; j < kMaxPerLane; ++j) {
        const int c = lane + 32 * j;
        if (c < C) {
            const float d = v[j] * mul;
            out[c] = d;
            if (so.hi) split_store(so, ((size_t)b * kcap + k) * C + c, d);
        }
    }
}


// ---- gather mode on a channels-last map ------------------------------------------------------- //
// The SiLK-type gather reads raw[b, :, y, x]: in NCHW that is one 32-byte sector per 4-byte channel value
// (channel stride = Hd*Wd floats).  cuDNN's native layout on Blackwell is channels-last, where the C channels
// of a pixel are contiguous: a warp reads a keypoint's descriptor as float4 per lane (C = 128: one 512-byte
// request), normalises and writes it the same way.  Two keypoints per warp and iteration keep two loads and
// two shuffle reductions in flight.
template <int NV>  // float4 per lane: C == 128 * NV (NV = 0: any C % 4 == 0 up to 512, guarded)
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
sample_gather_nhwc_kernel(const float* __restrict__ raw, int C, int Hd, int Wd, const float* __restrict__ kpts,
                          const int32_t* __restrict__ counts, int kcap, float scale, int normalize,
                          float* __restrict__ desc, SplitOut so) {
    constexpr int MV = NV ? NV : 4;
    const int b = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const int k0 = (blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5)) * 2;
    if (k0 >= kcap) return;
    int cnt = counts[b];
    if (cnt > kcap) cnt = kcap;
    const float* img = raw + (size_t)b * Hd * Wd * C;
    const int c4 = C >> 2;  // float4 per descriptor
    float4 v[2][MV];
    float ss[2] = {0.0f, 0.0f};
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const int k = k0 + u;
        const bool live = k < cnt;
        int yy = 0, xx = 0;
        if (live) {
            const float* kp = kpts + ((size_t)b * kcap + k) * 3;
            yy = min(max((int)floorf(__ldg(kp)), 0), Hd - 1);      // pos.floor().long()   (:57-60)
            xx = min(max((int)floorf(__ldg(kp + 1)), 0), Wd - 1);
        }
        const float4* src = reinterpret_cast<const float4*>(img + ((size_t)yy * Wd + xx) * C);
#pragma unroll
        for (int j = 0; j < MV; ++j) {
            const int i = lane + 32 * j;
            v[u][j] = (live && (NV || i < c4)) ? __ldg(src + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int j = 0; j < MV; ++j) {
            // same accumulation order as the NCHW kernel is not required: the reference's own norm is a
            // library reduction; the result is checked to 2e-6 absolute
            ss[u] = fmaf(v[u][j].x, v[u][j].x, ss[u]);
            ss[u] = fmaf(v[u][j].y, v[u][j].y, ss[u]);
            ss[u] = fmaf(v[u][j].z, v[u][j].z, ss[u]);
            ss[u] = fmaf(v[u][j].w, v[u][j].w, ss[u]);
        }
    float mul[2] = {scale, scale};
    if (normalize) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            ss[0] += __shfl_xor_sync(0xffffffffu, ss[0], o);
            ss[1] += __shfl_xor_sync(0xffffffffu, ss[1], o);
        }
        mul[0] = __fdiv_rn(scale, fmaxf(sqrtf(ss[0]), 1e-12f));
        mul[1] = __fdiv_rn(scale, fmaxf(sqrtf(ss[1]), 1e-12f));
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const int k = k0 + u;
        if (k >= kcap) continue;
        float4* out = reinterpret_cast<float4*>(desc + ((size_t)b * kcap + k) * C);
#pragma unroll
        for (int j = 0; j < MV; ++j) {
            const int i = lane + 32 * j;
            if (NV || i < c4) {
                const float4 q = v[u][j];  // padding rows (k >= cnt) were loaded as zeros
                const float4 d = make_float4(q.x * mul[u], q.y * mul[u], q.z * mul[u], q.w * mul[u]);
                out[i] = d;
                if (so.hi) split_store4(so, ((size_t)b * kcap + k) * C + 4 * i, d);
            }
        }
    }
}

// ---- bilinear mode, coarse map staged in shared memory --------------------------------------- //
// The generic kernel above reads one 32-byte sector per 4-byte tap (channel stride = Hd*Wd floats).
// For SuperPoint-type maps (C=256, 23x30 cells) the two coarse rows a keypoint touches are only
// C * 2 * Wd floats, so one CTA per (image, coarse row pair) stages them with coalesced loads and
// serves every keypoint whose upper tap row is that pair's first row: the descriptor map is read
// about twice from L2/HBM instead of ~8x in sectors, and every tap becomes a conflict-free LDS.
#ifndef EINX_SLAB_THREADS
#define EINX_SLAB_THREADS 256
#endif
#ifndef EINX_SLAB_MINB
#define EINX_SLAB_MINB 3
#endif
constexpr int kSlabThreads = EINX_SLAB_THREADS;
constexpr int kSlabWarps = kSlabThreads / 32;

struct Tap {  // one keypoint's bilinear footprint, computed once and shared by the 32 channel lanes
    int k, xa, xb;
    float w00, w01, w10, w11;
};

__device__ __forceinline__ float bilinear_unnormalize(float p, float size_padded, int size_in) {
    // pos - 0.5 -> [-1, 1] on the padded image -> grid_sample index ((g + 1) * size_in - 1) / 2
    const float g = __fsub_rn(__fmul_rn(2.0f, __fdiv_rn(__fsub_rn(p, 0.5f), __fsub_rn(size_padded, 1.0f))), 1.0f);
    return __fdiv_rn(__fsub_rn(__fmul_rn(__fadd_rn(g, 1.0f), (float)size_in), 1.0f), 2.0f);
}

// one keypoint's descriptor row from a lane's NJ values (channels lane + 32 j): fp32 row and, when asked for, the fp16
// operand pair -- two channel groups per packed conversion
template <int NJ>
__device__ __forceinline__ void store_row(float* __restrict__ out, const SplitOut& so, size_t row, const float (&v)[NJ], float mul,
                                          int lane, int C) {
    if (NJ != kMaxPerLane && NJ % 2 == 0) {
#pragma unroll
        for (int j = 0; j < NJ; j += 2) {
            const int c0 = lane + 32 * j, c1 = c0 + 32;
            const float d0 = v[j] * mul, d1 = v[j + 1] * mul;
            out[c0] = d0;
            out[c1] = d1;
            if (so.hi) split_store2(so, row + c0, row + c1, d0, d1);
        }
    } else {
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const int c = lane + 32 * j;
            if (NJ != kMaxPerLane || c < C) {
                const float d = v[j] * mul;
                out[c] = d;
                if (so.hi) split_store(so, row + c, d);
            }
        }
    }
}

template <int NJ>  // channel groups of 32 held per lane: C == 32 * NJ exactly, or NJ == kMaxPerLane with guards
__global__ void __launch_bounds__(kSlabThreads, EINX_SLAB_MINB)
sample_bilinear_slab_kernel(const float* __restrict__ raw, int C, int Hd, int Wd, float Hp, float Wp,
                            const float* __restrict__ kpts, const int32_t* __restrict__ counts, int kcap, float scale,
                            int normalize, float* __restrict__ desc, int SP, SplitOut so) {
    extern __shared__ __align__(16) float slab[];  // [C][SP]: rows y0, y0+1 of every channel
    __shared__ Tap taps[kSlabThreads];
    __shared__ int s_bound[2];
    const int b = blockIdx.y;
    const int y0 = (int)blockIdx.x - 1;  // upper tap row of this CTA's keypoints, -1 .. Hd-1
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int cnt = counts[b];
    if (cnt > kcap) cnt = kcap;

    // ---- stage rows y0, y0+1 of every channel first: one warp per channel, lanes along the 2*Wd
    // floats, everything in flight at once (cp.async, no register staging); the keypoint scan below
    // runs while the copies land.  Rows outside the map (y0 = -1, y0+1 = Hd) are zero.
    const float* img = raw + (size_t)b * C * Hd * Wd;
    const int row_elems = 2 * Wd;  // <= 128: a lane owns at most 4 elements of a channel
    const bool oy0 = (y0 >= 0) & (y0 < Hd), oy1 = (y0 + 1 >= 0) & (y0 + 1 < Hd);
    if (!(oy0 && oy1)) {
        const int lo = oy0 ? Wd : 0, hi = oy1 ? Wd : row_elems;  // the out-of-map row's span
        for (int c = warp; c < C; c += kSlabWarps)
            for (int e = lo + lane; e < hi; e += 32) slab[c * SP + e] = 0.0f;
    }
    {
        bool ok[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int e = lane + 32 * u;
            ok[u] = e < row_elems && (e < Wd ? oy0 : oy1);
        }
        const float* src = img + ((ptrdiff_t)warp * Hd + y0) * Wd + lane;  // rows y0 and y0+1 are adjacent in memory
        const size_t cstep = (size_t)kSlabWarps * Hd * Wd;
        uint32_t dst = (uint32_t)__cvta_generic_to_shared(slab + warp * SP + lane);
        const uint32_t dstep = (uint32_t)(kSlabWarps * SP * sizeof(float));
#pragma unroll 4
        for (int c = warp; c < C; c += kSlabWarps, src += cstep, dst += dstep) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (ok[u]) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst + 128u * u), "l"(src + 32 * u) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    if (blockIdx.x == 0) {  // padding rows are defined (zero)
        float4* z = reinterpret_cast<float4*>(desc + ((size_t)b * kcap + cnt) * C);  // C % 4 == 0 on this path
        const size_t nz = (size_t)(kcap - cnt) * C / 4;
        for (size_t i = tid; i < nz; i += kSlabThreads) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (so.hi) {  // (fp16 zeros; C % 8 == 0 whenever the split output is requested)
            uint2* zh = reinterpret_cast<uint2*>(so.hi + ((size_t)b * kcap + cnt) * C);
            uint2* zl = reinterpret_cast<uint2*>(so.lo + ((size_t)b * kcap + cnt) * C);
            for (size_t i = tid; i < nz; i += kSlabThreads) { zh[i] = make_uint2(0u, 0u); zl[i] = make_uint2(0u, 0u); }
        }
    }

    // ---- keypoints are in raster order, so floor(iy) never decreases along the list and this CTA's
    // keypoints form one contiguous run [first, last]: two warps find its ends with a 32-ary search over
    // the exact row expression (instead of every CTA evaluating all keypoints of the image)
    const float* kp = kpts + (size_t)b * kcap * 3;
    if (warp < 2) {
        const int target = y0 + warp;  // warp 0: first k with row >= y0; warp 1: first k with row >= y0 + 1
        int L = 0, R = cnt;
        while (R > L) {
            const int n = R - L, st = (n + 31) >> 5, i = L + lane * st;
            const bool in = i < R;
            const bool pr = in && (int)floorf(bilinear_unnormalize(__ldg(kp + 3 * i), Hp, Hd)) >= target;
            const unsigned valid = __ballot_sync(0xffffffffu, in), hit = __ballot_sync(0xffffffffu, pr);
            if (hit) {
                const int j = __ffs(hit) - 1;
                R = L + j * st;
                if (j == 0) break;
                L = L + (j - 1) * st + 1;
            } else {
                L = L + (31 - __clz(valid)) * st + 1;
            }
            if (st == 1) break;
        }
        if (lane == 0) s_bound[warp] = R;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    const int first = s_bound[0], last = s_bound[1] - 1;
    if (last < first) return;

    for (int base = first; base <= last; base += kSlabThreads) {
        if (base != first) __syncthreads();  // taps consumed by the previous pass
        const int k = base + tid;
        if (k <= last) {
            const float iy = bilinear_unnormalize(__ldg(kp + 3 * k), Hp, Hd);
            const float ix = bilinear_unnormalize(__ldg(kp + 3 * k + 1), Wp, Wd);
            const float fy = floorf(iy), fx = floorf(ix);
            const int x0 = (int)fx;
            const float wy1 = __fsub_rn(iy, fy), wx1 = __fsub_rn(ix, fx);
            const float wy0 = __fsub_rn(__fadd_rn(fy, 1.0f), iy), wx0 = __fsub_rn(__fadd_rn(fx, 1.0f), ix);
            const bool ox0 = (x0 >= 0) & (x0 < Wd), ox1 = (x0 + 1 >= 0) & (x0 + 1 < Wd);
            Tap t;
            t.k = ((int)fy == y0) ? k : -1;
            t.xa = ox0 ? x0 : 0;
            t.xb = ox1 ? x0 + 1 : 0;
            // out-of-map taps get weight 0 (grid_sample zeros padding); out-of-map rows are zero in the slab
            t.w00 = ox0 ? __fmul_rn(wx0, wy0) : 0.0f;
            t.w01 = ox1 ? __fmul_rn(wx1, wy0) : 0.0f;
            t.w10 = ox0 ? __fmul_rn(wx0, wy1) : 0.0f;
            t.w11 = ox1 ? __fmul_rn(wx1, wy1) : 0.0f;
            taps[tid] = t;
        }
        __syncthreads();
        const int n = min(kSlabThreads, last - base + 1);
        // two keypoints per warp and iteration: their tap loads, norm reductions (five dependent shuffles
        // each) and stores interleave instead of running back to back
        for (int e = warp; e < n; e += 2 * kSlabWarps) {
            const Tap t0 = taps[e];
            const bool has1 = e + kSlabWarps < n;
            const Tap t1 = taps[has1 ? e + kSlabWarps : e];
            const bool live0 = t0.k >= 0, live1 = has1 && t1.k >= 0;
            if (!live0 && !live1) continue;
            float v0[NJ], v1[NJ];
            float ss0 = 0.0f, ss1 = 0.0f;
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                const int c = lane + 32 * j;
                float a0 = 0.0f, a1 = 0.0f;
                if (NJ != kMaxPerLane || c < C) {
                    const float* sp = slab + c * SP;  // odd pitch: the 32 channel lanes hit 32 banks
                    a0 = __fmul_rn(sp[t0.xa], t0.w00);
                    a1 = __fmul_rn(sp[t1.xa], t1.w00);
                    a0 = fmaf(sp[t0.xb], t0.w01, a0);
                    a1 = fmaf(sp[t1.xb], t1.w01, a1);
                    a0 = fmaf(sp[Wd + t0.xa], t0.w10, a0);
                    a1 = fmaf(sp[Wd + t1.xa], t1.w10, a1);
                    a0 = fmaf(sp[Wd + t0.xb], t0.w11, a0);
                    a1 = fmaf(sp[Wd + t1.xb], t1.w11, a1);
                }
                v0[j] = a0; v1[j] = a1;
                ss0 = fmaf(a0, a0, ss0);
                ss1 = fmaf(a1, a1, ss1);
            }
            float mul0 = scale, mul1 = scale;
            if (normalize) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    ss0 += __shfl_xor_sync(0xffffffffu, ss0, o);
                    ss1 += __shfl_xor_sync(0xffffffffu, ss1, o);
                }
                mul0 = __fdiv_rn(scale, fmaxf(sqrtf(ss0), 1e-12f));
                mul1 = __fdiv_rn(scale, fmaxf(sqrtf(ss1), 1e-12f));
            }
            if (live0) {
                float* out = desc + ((size_t)b * kcap + t0.k) * C;
                store_row<NJ>(out, so, ((size_t)b * kcap + t0.k) * C, v0, mul0, lane, C);
            }
            if (live1) {
                float* out = desc + ((size_t)b * kcap + t1.k) * C;
                store_row<NJ>(out, so, ((size_t)b * kcap + t1.k) * C, v1, mul1, lane, C);
            }
        }
    }
}

}  // namespace

extern "C" int einx_sample(einx_ctx* ctx, const float* raw, int B, int C, int Hd, int Wd, int mode, int Hp, int Wp,
                           const float* kpts, const int32_t* counts, int kcap, float scale, int normalize,
                           float* desc, einx_stream stream_) {
    return einx_sample_split(ctx, raw, B, C, Hd, Wd, mode, Hp, Wp, kpts, counts, kcap, scale, normalize, desc, nullptr, stream_);
}

extern "C" int einx_sample_split(einx_ctx* ctx, const float* raw, int B, int C, int Hd, int Wd, int mode, int Hp, int Wp,
                                 const float* kpts, const int32_t* counts, int kcap, float scale, int normalize,
                                 float* desc, uint16_t* split, einx_stream stream_) {
    if (!ctx) return EINX_ERR_INVALID;
    if (split && (C % 8 != 0 || ((uintptr_t)split & 15)))
        return einx_fail(ctx, EINX_ERR_UNSUPPORTED, "einx_sample_split: the fp16 operands need C %% 8 == 0 and a 16-byte aligned buffer (C=%d)", C);
    SplitOut so = {nullptr, nullptr};
    if (split && B > 0 && kcap > 0) {
        so.hi = reinterpret_cast<__half*>(split);
        so.lo = so.hi + (size_t)B * kcap * C;
    }
    if (B < 0 || C <= 0 || Hd <= 0 || Wd <= 0 || kcap < 0)
        return einx_fail(ctx, EINX_ERR_INVALID, "einx_sample: bad shape B=%d C=%d Hd=%d Wd=%d kcap=%d", B, C, Hd, Wd, kcap);
    if (B == 0 || kcap == 0) return EINX_OK;
    if (!raw || !kpts || !counts || !desc) return einx_fail(ctx, EINX_ERR_INVALID, "einx_sample: NULL pointer argument");
    if (C > 32 * kMaxPerLane) return einx_fail(ctx, EINX_ERR_UNSUPPORTED, "einx_sample: C=%d > %d", C, 32 * kMaxPerLane);
    if (B > 65535) return einx_fail(ctx, EINX_ERR_UNSUPPORTED, "einx_sample: B=%d > 65535", B);
    DeviceGuard guard(ctx->device);
    cudaStream_t stream = (cudaStream_t)stream_;
    dim3 grid((kcap + kWarpsPerBlock - 1) / kWarpsPerBlock, B);
    struct ProfScope {  // brackets whichever sampling kernel runs below
        einx_ctx* c; cudaStream_t s;
        ProfScope(einx_ctx* c_, cudaStream_t s_) : c(c_), s(s_) { einx_prof_begin(c, 2, s); }
        ~ProfScope() { einx_prof_end(c, 2, s); }
    } prof_scope(ctx, stream);
    if (mode == EINX_SAMPLE_GATHER) {
        sample_kernel<EINX_SAMPLE_GATHER><<<grid, kWarpsPerBlock * 32, 0, stream>>>(
            raw, C, Hd, Wd, (float)Hp, (float)Wp, kpts, counts, kcap, scale, normalize, desc, so);
    } else if (mode == EINX_SAMPLE_GATHER_NHWC) {
        if (C % 4 != 0 || (uintptr_t)raw % 16 != 0 || (uintptr_t)desc % 16 != 0)
            return einx_fail(ctx, EINX_ERR_UNSUPPORTED, "einx_sample: channels-last gather needs C %% 4 == 0 and 16-byte aligned buffers (C=%d)", C);
        dim3 g2((kcap + 2 * kWarpsPerBlock - 1) / (2 * kWarpsPerBlock), B);
        const int T = kWarpsPerBlock * 32;
        if (C == 128) sample_gather_nhwc_kernel<1><<<g2, T, 0, stream>>>(raw, C, Hd, Wd, kpts, counts, kcap, scale, normalize, desc, so);
        else if (C == 256) sample_gather_nhwc_kernel<2><<<g2, T, 0, stream>>>(raw, C, Hd, Wd, kpts, counts, kcap, scale, normalize, desc, so);
        else sample_gather_nhwc_kernel<0><<<g2, T, 0, stream>>>(raw, C, Hd, Wd, kpts, counts, kcap, scale, normalize, desc, so);
    } else if (mode == EINX_SAMPLE_BILINEAR) {
        if (Hp <= 1 || Wp <= 1) return einx_fail(ctx, EINX_ERR_INVALID, "einx_sample: bilinear needs Hp, Wp > 1");
        // shared-memory slab variant when two coarse rows of every channel fit (odd pitch: lanes are
        // channels, so an odd channel pitch makes every tap conflict-free)
        const int SP = (2 * Wd) | 1;
        const size_t slab_bytes = (size_t)C * SP * sizeof(float);
        if (slab_bytes <= 100 * 1024 && Hd + 1 <= 65535 && 2 * Wd <= 128 && C % 4 == 0) {
            auto launch = [&](auto kern) -> cudaError_t {
                cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)slab_bytes);
                if (e != cudaSuccess) return e;
                kern<<<dim3(Hd + 1, B), kSlabThreads, slab_bytes, stream>>>(raw, C, Hd, Wd, (float)Hp, (float)Wp, kpts,
                                                                          counts, kcap, scale, normalize, desc, SP, so);
                return cudaSuccess;
            };
            switch (C) {
                case 32: EINX_CUDA(ctx, launch(sample_bilinear_slab_kernel<1>)); break;
                case 64: EINX_CUDA(ctx, launch(sample_bilinear_slab_kernel<2>)); break;
                case 128: EINX_CUDA(ctx, launch(sample_bilinear_slab_kernel<4>)); break;
                case 256: EINX_CUDA(ctx, launch(sample_bilinear_slab_kernel<8>)); break;
                default: EINX_CUDA(ctx, launch(sample_bilinear_slab_kernel<kMaxPerLane>)); break;
            }
            EINX_CHECK_LAUNCH(ctx);
            return EINX_OK;
        }
        sample_kernel<EINX_SAMPLE_BILINEAR><<<grid, kWarpsPerBlock * 32, 0, stream>>>(
            raw, C, Hd, Wd, (float)Hp, (float)Wp, kpts, counts, kcap, scale, normalize, desc, so);
    } else {
        return einx_fail(ctx, EINX_ERR_INVALID, "einx_sample: unknown mode %d", mode);
    }
    EINX_CHECK_LAUNCH(ctx);
    return EINX_OK;
}
