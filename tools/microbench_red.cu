// Micro-benchmark: throughput of float reductions to random cells of an L2-resident grid
// (decides the design of the voxel scatter).  nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>

__device__ __forceinline__ uint32_t hash(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

template <int MODE>
__global__ void k_red(float* g, uint32_t cells, uint32_t n, int reps) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t stride = gridDim.x * blockDim.x;
    for (int r = 0; r < reps; ++r)
        for (uint32_t e = i; e < n; e += stride) {
            uint32_t c = hash(e * 2654435761u + r) % cells;
            float w = 1e-3f * (c & 7);
            if (MODE == 0) asm volatile("red.global.add.f32 [%0], %1;" ::"l"(g + c), "f"(w) : "memory");
            if (MODE == 1) { c &= ~1u; asm volatile("red.global.add.v2.f32 [%0], {%1,%2};" ::"l"(g + c), "f"(w), "f"(w) : "memory"); }
            if (MODE == 2) { c &= ~3u; asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(g + c), "f"(w), "f"(w), "f"(w), "f"(w) : "memory"); }
            if (MODE == 3) { atomicAdd(g + c, w); }
            if (MODE == 4) { // two scalar reds to adjacent cells (what an unaligned x-pair costs)
                asm volatile("red.global.add.f32 [%0], %1;" ::"l"(g + c), "f"(w) : "memory");
                asm volatile("red.global.add.f32 [%0], %1;" ::"l"(g + c + 1), "f"(w) : "memory"); }
        }
}

__global__ void k_smem_int(int* out, uint32_t n, int reps) {
    extern __shared__ int s[];
    for (int i = threadIdx.x; i < 12288; i += blockDim.x) s[i] = 0;
    __syncthreads();
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
    for (int r = 0; r < reps; ++r)
        for (uint32_t e = i; e < n; e += stride) atomicAdd(&s[hash(e + r) % 12288], 1);
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = s[5];
}
__global__ void k_smem_float(float* out, uint32_t n, int reps) {
    extern __shared__ float sf[];
    for (int i = threadIdx.x; i < 12288; i += blockDim.x) sf[i] = 0;
    __syncthreads();
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
    for (int r = 0; r < reps; ++r)
        for (uint32_t e = i; e < n; e += stride) atomicAdd(&sf[hash(e + r) % 12288], 1.0f);
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = sf[5];
}

template <typename F> float timeit(F f) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}

int main() {
    const uint32_t n = 1u << 24;
    float* g; cudaMalloc(&g, 256u << 20); cudaMemset(g, 0, 256u << 20);
    int sizes[] = {450000 /*MVSEC 5 bins*/, 216000 * 64 /*C2 batch*/, 9216000 /*720p 10 bins*/};
    const char* names[] = {"red.f32", "red.v2.f32", "red.v4.f32", "atomicAdd(ret unused)", "2x red.f32 adjacent"};
    for (int cells : sizes) {
        printf("grid of %d cells (%.1f MB), %u ops\n", cells, cells * 4e-6, n);
        float ms;
        ms = timeit([&] { k_red<0><<<148 * 8, 256>>>(g, cells, n, 1); }); printf("  %-24s %8.3f ms  %7.1f Gop/s\n", names[0], ms, n / ms * 1e-6);
        ms = timeit([&] { k_red<1><<<148 * 8, 256>>>(g, cells, n, 1); }); printf("  %-24s %8.3f ms  %7.1f Gop/s (x2 floats)\n", names[1], ms, n / ms * 1e-6);
        ms = timeit([&] { k_red<2><<<148 * 8, 256>>>(g, cells, n, 1); }); printf("  %-24s %8.3f ms  %7.1f Gop/s (x4 floats)\n", names[2], ms, n / ms * 1e-6);
        ms = timeit([&] { k_red<3><<<148 * 8, 256>>>(g, cells, n, 1); }); printf("  %-24s %8.3f ms  %7.1f Gop/s\n", names[3], ms, n / ms * 1e-6);
        ms = timeit([&] { k_red<4><<<148 * 8, 256>>>(g, cells, n, 1); }); printf("  %-24s %8.3f ms  %7.1f Gpair/s\n", names[4], ms, n / ms * 1e-6);
    }
    int* o; cudaMalloc(&o, 4096 * 4);
    cudaFuncSetAttribute(k_smem_int, cudaFuncAttributeMaxDynamicSharedMemorySize, 49152);
    float ms = timeit([&] { k_smem_int<<<148 * 4, 256, 49152>>>(o, n, 1); }); printf("smem int atomicAdd (12288 cells/CTA)   %8.3f ms %7.1f Gop/s\n", ms, n / ms * 1e-6);
    ms = timeit([&] { k_smem_float<<<148 * 4, 256, 49152>>>((float*)o, n, 1); }); printf("smem float atomicAdd (CAS loop)        %8.3f ms %7.1f Gop/s\n", ms, n / ms * 1e-6);
    return 0;
}
