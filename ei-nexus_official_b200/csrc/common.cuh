// Shared declarations of libeinx (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "einx.h"

struct einx_ctx {
    int device;
    int num_sms;
    int max_smem_optin;   // bytes of dynamic shared memory a CTA may opt in to
    void* ws;             // grow-only device workspace
    size_t ws_bytes;
    cudaStream_t ws_stream;  // stream the workspace was allocated on (cudaMallocAsync)
    int32_t* redo_flags;     // [65536] per-image flags of the tiled detect kernel (outside the workspace: the redo pass reuses that)
    int64_t launches;
    int profile;                    // einx_profile_enable
    cudaEvent_t prof_ev[4][2];      // [slot][begin/end], created lazily
    int prof_set[4];
    char err[512];
};

// Bracket the dominant kernel of an entry point with events on the caller's stream (no-op unless
// profiling is enabled).
void einx_prof_begin(einx_ctx* ctx, int slot, cudaStream_t stream);
void einx_prof_end(einx_ctx* ctx, int slot, cudaStream_t stream);

// Grow the workspace to at least `bytes`.  Growth is stream-ordered (cudaMallocAsync / cudaFreeAsync on the caller's
// stream): nothing synchronises with the host and other streams keep running; a context serves one stream at a time,
// so work queued earlier on that stream finishes with the old block before its memory can be reused.
int einx_ws_reserve(einx_ctx* ctx, size_t bytes, cudaStream_t stream);
int einx_fail(einx_ctx* ctx, int code, const char* fmt, ...);

#define EINX_CUDA(ctx, call)                                                                   \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess)                                                                 \
            return einx_fail((ctx), EINX_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call, \
                             cudaGetErrorString(e_));                                          \
    } while (0)

#define EINX_CHECK_LAUNCH(ctx)                                                                   \
    do {                                                                                         \
        (ctx)->launches++;                                                                       \
        cudaError_t e_ = cudaGetLastError();                                                     \
        if (e_ != cudaSuccess)                                                                   \
            return einx_fail((ctx), EINX_ERR_CUDA, "%s:%d kernel launch -> %s", __FILE__, __LINE__, \
                             cudaGetErrorString(e_));                                            \
    } while (0)

struct DeviceGuard {
    int prev;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

__host__ __device__ static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// fp32 -> monotone uint32 (larger float <=> larger unsigned); -0 must be canonicalised first.
__device__ __forceinline__ uint32_t f32_orderable(float f) {
    uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float f32_from_orderable(uint32_t u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}
// (value, lowest index wins) packed so that one unsigned 64-bit max does an argmax
__device__ __forceinline__ unsigned long long pack_best(float v, uint32_t idx) {
    return ((unsigned long long)f32_orderable(v + 0.0f) << 32) | (unsigned long long)(0xffffffffu - idx);
}

// First-occurrence argmax of 32 register values as a tournament tree: 31 (compare, select value,
// select index) triples at depth 5 instead of a 32-long dependent chain; the left operand wins ties.
__device__ __forceinline__ void argmax32(const float (&x)[32], float& val, int& idx) {
    float a[16];
    int ia[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const bool p = x[2 * k + 1] > x[2 * k];
        a[k] = p ? x[2 * k + 1] : x[2 * k];
        ia[k] = p ? 2 * k + 1 : 2 * k;
    }
#pragma unroll
    for (int n = 8; n >= 1; n >>= 1) {
#pragma unroll
        for (int k = 0; k < n; ++k) {
            const bool p = a[2 * k + 1] > a[2 * k];
            a[k] = p ? a[2 * k + 1] : a[2 * k];
            ia[k] = p ? ia[2 * k + 1] : ia[2 * k];
        }
    }
    val = a[0];
    idx = ia[0];
}
