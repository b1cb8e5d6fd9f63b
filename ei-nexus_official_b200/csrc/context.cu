// Context lifetime, workspace and error plumbing of libeinx.
#include <stdarg.h>
#include <stdlib.h>

#include "common.cuh"

static char g_create_err[512] = "";

int einx_fail(einx_ctx* ctx, int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(ctx ? ctx->err : g_create_err, 512, fmt, ap);
    va_end(ap);
    return code;
}

int einx_ws_reserve(einx_ctx* ctx, size_t bytes, cudaStream_t stream) {
    if (bytes <= ctx->ws_bytes) return EINX_OK;
    // Growth only, ordered on the caller's stream: the new block is usable by everything queued after this point,
    // the old one returns to the pool once the kernels queued before it have run.  No host synchronisation.
    size_t want = align_up(bytes + bytes / 4, 1 << 20);
    void* fresh = nullptr;
    cudaError_t e = cudaMallocAsync(&fresh, want, stream);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return einx_fail(ctx, EINX_ERR_NOMEM, "workspace of %zu bytes: %s", want, cudaGetErrorString(e));
    }
    if (ctx->ws) {
        // (a caller that moved the context to another stream: the old stream's work must be done first)
        if (ctx->ws_stream != stream) cudaStreamSynchronize(ctx->ws_stream);
        e = cudaFreeAsync(ctx->ws, stream);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return einx_fail(ctx, EINX_ERR_CUDA, "workspace release: %s", cudaGetErrorString(e));
        }
    }
    ctx->ws = fresh;
    ctx->ws_bytes = want;
    ctx->ws_stream = stream;
    return EINX_OK;
}

void einx_prof_begin(einx_ctx* ctx, int slot, cudaStream_t stream) {
    if (!ctx->profile) return;
    for (int k = 0; k < 2; ++k)
        if (!ctx->prof_ev[slot][k]) cudaEventCreate(&ctx->prof_ev[slot][k]);
    cudaEventRecord(ctx->prof_ev[slot][0], stream);
}
void einx_prof_end(einx_ctx* ctx, int slot, cudaStream_t stream) {
    if (!ctx->profile) return;
    cudaEventRecord(ctx->prof_ev[slot][1], stream);
    ctx->prof_set[slot] = 1;
}

extern "C" {

int einx_profile_enable(einx_ctx* ctx, int on) {
    if (!ctx) return EINX_ERR_INVALID;
    ctx->profile = on != 0;
    for (int s = 0; s < EINX_PROFILE_SLOTS; ++s) ctx->prof_set[s] = 0;
    return EINX_OK;
}

int einx_profile_read(einx_ctx* ctx, float* ms_out) {
    if (!ctx || !ms_out) return EINX_ERR_INVALID;
    DeviceGuard g(ctx->device);
    for (int s = 0; s < EINX_PROFILE_SLOTS; ++s) {
        ms_out[s] = -1.0f;
        if (!ctx->prof_set[s]) continue;
        EINX_CUDA(ctx, cudaEventSynchronize(ctx->prof_ev[s][1]));
        EINX_CUDA(ctx, cudaEventElapsedTime(&ms_out[s], ctx->prof_ev[s][0], ctx->prof_ev[s][1]));
    }
    return EINX_OK;
}

int einx_version(void) { return 100; }

int einx_create(int device, einx_ctx** out) {
    if (!out) return einx_fail(nullptr, EINX_ERR_INVALID, "einx_create: out is NULL");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return einx_fail(nullptr, EINX_ERR_CUDA, "einx_create: no CUDA device (%s); this library has no CPU fallback",
                         cudaGetErrorString(e));
    }
    if (device < 0 || device >= ndev)
        return einx_fail(nullptr, EINX_ERR_INVALID, "einx_create: device %d out of range [0,%d)", device, ndev);
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return einx_fail(nullptr, EINX_ERR_CUDA, "einx_create: %s", cudaGetErrorString(e));
    if (prop.major != 10)
        return einx_fail(nullptr, EINX_ERR_ARCH,
                         "einx_create: device %d is sm_%d%d; libeinx is built for sm_100a (B200) only", device,
                         prop.major, prop.minor);
    einx_ctx* ctx = (einx_ctx*)calloc(1, sizeof(einx_ctx));
    if (!ctx) return einx_fail(nullptr, EINX_ERR_NOMEM, "einx_create: out of host memory");
    ctx->device = device;
    ctx->num_sms = prop.multiProcessorCount;
    ctx->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
    {
        // per-image flags of the tiled detect kernel: allocated here (not on first use) so that no entry point ever
        // calls cudaMalloc -- which is illegal while the caller's stream is being captured into a CUDA graph
        DeviceGuard g(device);
        if (cudaMalloc(&ctx->redo_flags, 65536 * sizeof(int32_t)) != cudaSuccess) {
            cudaGetLastError();
            free(ctx);
            return einx_fail(nullptr, EINX_ERR_NOMEM, "einx_create: out of device memory");
        }
    }
    *out = ctx;
    return EINX_OK;
}

void einx_destroy(einx_ctx* ctx) {
    if (!ctx) return;
    DeviceGuard g(ctx->device);
    if (ctx->ws) {
        cudaDeviceSynchronize();
        cudaFree(ctx->ws);  // (valid for stream-ordered allocations too; the stream may be gone by now)
    }
    if (ctx->redo_flags) cudaFree(ctx->redo_flags);
    for (int s = 0; s < EINX_PROFILE_SLOTS; ++s)
        for (int k = 0; k < 2; ++k)
            if (ctx->prof_ev[s][k]) cudaEventDestroy(ctx->prof_ev[s][k]);
    free(ctx);
}

const char* einx_last_error(const einx_ctx* ctx) { return ctx ? ctx->err : g_create_err; }

int64_t einx_launch_count(const einx_ctx* ctx) { return ctx ? ctx->launches : 0; }

}  // extern "C"
