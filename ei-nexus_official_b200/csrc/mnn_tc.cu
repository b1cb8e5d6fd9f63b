// Tensor-core similarity tiles for the MNN matcher: tcgen05.mma with TMEM accumulators, operands
// staged by TMA (128-byte swizzle), row/column argmax fused into the TMEM epilogue.
// Semantics: reference core/modules/matchers/MNN.py:88-92 (einsum + 2x topk(1)); the thresholds and
// the mutual check run in mnn.cu's finalisation kernel on the keys this kernel produces.
//
// Persistent, warp-specialised CTA (one per SM):
//   warp 0   TMA producer   -- A tile 128 rows x 128 B, B tile 256 rows x 128 B per k-block
//   warp 1   MMA issuer     -- one lane issues tcgen05.mma (M=128, N=256, K=32 bytes) x4 per k-block
//   warp 2   TMEM allocator -- 512 columns = two 128x256 fp32 accumulators (double buffered)
//   warps 4-11 epilogue     -- tcgen05.ld 32 columns at a time; per-row running max in registers,
//                              per-column max through a padded 32x32 shared-memory transpose, then
//                              one 64-bit atomicMax per row / column into the global key arrays
// The similarity matrix therefore never leaves the SM.
//
// Precision modes (include/einx.h): BF16 rounds the descriptors to bf16 (kind::f16); TF32X3 splits
// every fp32 descriptor into hi = tf32(x) and lo = x - hi and accumulates hi*hi + hi*lo + lo*hi
// (kind::tf32) by running the k-loop three times over different operand pairs -- fp32-accurate to
// a few 1e-7 while staying on the tensor pipe.
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"
#include "mnn_keys.cuh"

namespace {

constexpr int TILE_M = 128;          // rows of d0 per tile  (UMMA M, TMEM lanes)
constexpr int TILE_N = 256;          // rows of d1 per tile  (UMMA N, TMEM columns)
constexpr int KBLOCK_BYTES = 128;    // one swizzle-128B row per k-block
constexpr int A_BYTES = TILE_M * KBLOCK_BYTES;  // 16 KB
constexpr int B_BYTES = TILE_N * KBLOCK_BYTES;  // 32 KB
constexpr int STAGES = 3;
constexpr int kEpilogueWarp0 = 4;
constexpr int kEpilogueWarps = 8;   // two warps per TMEM lane quarter, each takes half of the columns
constexpr int kThreads = 32 * (kEpilogueWarp0 + kEpilogueWarps);
constexpr int kScratchPitch = 33;   // floats; 32x32 transpose tile per epilogue warp, conflict-free both ways
constexpr uint32_t kTmemCols = 512;

struct TcParams {
    const int32_t* n0;
    const int32_t* n1;
    int B, ncap, mcap;
    int tiles_m, tiles_n;  // per pair
    int nkb;               // k-blocks per operand pass
    int passes;            // 1 (bf16) or 3 (tf32x3)
    uint32_t idesc;
    unsigned long long* rowkey;
    unsigned long long* colkey;
};

struct __align__(8) Barriers {
    unsigned long long full[STAGES];
    unsigned long long empty[STAGES];
    unsigned long long tmem_full[2];
    unsigned long long tmem_empty[2];
    uint32_t tmem_base;
};

// ---- PTX wrappers ---------------------------------------------------------------------------- //
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, unsigned long long* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(unsigned long long* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
template <int KIND>  // 0: kind::f16 (bf16 inputs), 1: kind::tf32
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    if (KIND == 0) {
        asm volatile(
            "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
            : "memory");
    } else {
        asm volatile(
            "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
            : "memory");
    }
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major operand tile, 128-byte rows, SWIZZLE_128B: 8-row groups 1024 B apart (SBO), descriptor
// version 1 (Blackwell), layout type 2.  The start address advances by 32 B per UMMA_K step.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3ffff) >> 4);   // start address, bits [0,14)
    d |= (uint64_t)1 << 16;                        // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;              // stride byte offset, bits [32,46)
    d |= (uint64_t)1 << 46;                        // version
    d |= (uint64_t)2 << 61;                        // SWIZZLE_128B
    return d;
}

// signed-int key: larger float <=> larger int (for redux.sync.max.s32)
__device__ __forceinline__ int f32_skey(uint32_t bits) { return (int)(bits ^ (((int)bits >> 31) & 0x7fffffff)); }

template <int KIND>
__global__ void __launch_bounds__(kThreads, 1)
mnn_tc_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
              const __grid_constant__ CUtensorMap mapB0, const __grid_constant__ CUtensorMap mapB1, const TcParams P) {
    extern __shared__ __align__(1024) unsigned char smem[];
    // carve: [stages x (A | B)] 1024-aligned, then barriers, then the column-merge buffer
    unsigned char* tiles = (unsigned char*)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
    Barriers* bars = reinterpret_cast<Barriers*>(tiles + (size_t)STAGES * (A_BYTES + B_BYTES));
    unsigned long long* colpart = reinterpret_cast<unsigned long long*>(reinterpret_cast<unsigned char*>(bars) + 128);
    float* scratch_all = reinterpret_cast<float*>(colpart + 4 * TILE_N);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_per_pair = P.tiles_m * P.tiles_n;
    const int total_tiles = P.B * tiles_per_pair;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA0) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB0) : "memory");
        if (P.passes > 1) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA1) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB1) : "memory");
        }
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&bars->full[s], 1); mbar_init(&bars->empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&bars->tmem_full[a], 1); mbar_init(&bars->tmem_empty[a], kEpilogueWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)),
                     "r"(kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;

    auto tile_coords = [&](int t, int& b, int& i0, int& j0, bool& live) {
        b = t / tiles_per_pair;
        const int r = t - b * tiles_per_pair;
        i0 = (r / P.tiles_n) * TILE_M;
        j0 = (r % P.tiles_n) * TILE_N;
        const int N = P.n0 ? min(P.n0[b], P.ncap) : P.ncap;
        const int M = P.n1 ? min(P.n1[b], P.mcap) : P.mcap;
        live = (i0 < N) && (j0 < M);
    };

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
                int b, i0, j0;
                bool live;
                tile_coords(t, b, i0, j0, live);
                if (!live) continue;
                const int rowA = b * P.ncap + i0, rowB = b * P.mcap + j0;
                for (int pass = 0; pass < P.passes; ++pass) {
                    // pass 0: hi*hi   pass 1: hi*lo   pass 2: lo*hi
                    const CUtensorMap* ma = (pass == 2) ? &mapA1 : &mapA0;
                    const CUtensorMap* mb = (pass == 1) ? &mapB1 : &mapB0;
                    for (int kb = 0; kb < P.nkb; ++kb) {
                        mbar_wait(&bars->empty[stage], phase ^ 1);
                        unsigned char* sa = tiles + (size_t)stage * (A_BYTES + B_BYTES);
                        unsigned char* sb = sa + A_BYTES;
                        mbar_expect_tx(&bars->full[stage], A_BYTES + B_BYTES);
                        const int kcoord = kb * (KBLOCK_BYTES / (KIND == 0 ? 2 : 4));
                        tma_load_2d(sa, ma, &bars->full[stage], kcoord, rowA);
                        tma_load_2d(sb, mb, &bars->full[stage], kcoord, rowB);
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
                int b, i0, j0;
                bool live;
                tile_coords(t, b, i0, j0, live);
                if (!live) continue;
                mbar_wait(&bars->tmem_empty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)acc * TILE_N;
                const int nsteps = P.passes * P.nkb;
                for (int step = 0; step < nsteps; ++step) {
                    mbar_wait(&bars->full[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(tiles + (size_t)stage * (A_BYTES + B_BYTES));
                    const uint64_t adesc = make_smem_desc(sa), bdesc = make_smem_desc(sa + A_BYTES);
#pragma unroll
                    for (int k = 0; k < KBLOCK_BYTES / 32; ++k) {
                        // +32 B along K inside the swizzle atom = +2 in the 16-byte-unit address field
                        tc_mma<KIND>(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), P.idesc,
                                     (step > 0 || k > 0) ? 1u : 0u);
                    }
                    tc_commit(&bars->empty[stage]);  // smem slot free once these MMAs have read it
                    if (step == nsteps - 1) tc_commit(&bars->tmem_full[acc]);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            }
        }
    } else if (warp >= kEpilogueWarp0) {
        // ===== epilogue: TMEM -> registers -> row / column best keys =====
        // Rows: the thread that owns TMEM lane r scans its 128 columns with a strict '>' (lowest
        // column wins ties).  Columns: the 32x32 chunk goes through a padded shared-memory tile so
        // that lane c then owns column c and scans the 32 rows the same way (lowest row wins).
        const int q = warp & 3;                          // TMEM lane quarter this warp may access
        const int half = (warp - kEpilogueWarp0) >> 2;   // which 128 columns of the tile
        float* scratch = scratch_all + (size_t)(warp - kEpilogueWarp0) * 32 * kScratchPitch;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
            int b, i0, j0;
            bool live;
            tile_coords(t, b, i0, j0, live);
            if (!live) continue;
            const int N = P.n0 ? min(P.n0[b], P.ncap) : P.ncap;
            const int M = P.n1 ? min(P.n1[b], P.mcap) : P.mcap;
            const int row = i0 + 32 * q + lane;  // this thread's row of d0
            const bool row_ok = row < N;
            const int rows_here = min(max(N - (i0 + 32 * q), 0), 32);  // valid rows of this warp's quarter
            const bool full_tile = (i0 + TILE_M <= N) && (j0 + TILE_N <= M);
            mbar_wait(&bars->tmem_full[acc], acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(acc * TILE_N + 128 * half);
            float best = -INFINITY;
            int best_j = 0;
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                uint32_t v[32];
                tc_ld32(taddr + 32 * c, v);
                const int jc = j0 + 128 * half + 32 * c;
                if (jc < M) {  // warp-uniform; beyond M the tile is padding
#pragma unroll
                    for (int k = 0; k < 32; ++k) {
                        const float f = __uint_as_float(v[k]);
                        const bool ok = full_tile || (jc + k < M);
                        if (ok && f > best) { best = f; best_j = jc + k; }
                        scratch[lane * kScratchPitch + k] = f;
                    }
                    __syncwarp();
                    // two independent scan chains (rows 0-15, 16-31) keep the compare latency hidden
                    float c0 = -INFINITY, c1 = -INFINITY;
                    int r0 = 0, r1 = 16;
#pragma unroll
                    for (int r = 0; r < 16; ++r) {
                        const float x0 = scratch[r * kScratchPitch + lane];
                        const float x1 = scratch[(r + 16) * kScratchPitch + lane];
                        if ((full_tile || r < rows_here) && x0 > c0) { c0 = x0; r0 = r; }
                        if ((full_tile || r + 16 < rows_here) && x1 > c1) { c1 = x1; r1 = r + 16; }
                    }
                    if (c1 > c0) { c0 = c1; r0 = r1; }
                    const bool col_ok = (jc + lane < M) && (c0 > -INFINITY);
                    colpart[q * TILE_N + 128 * half + 32 * c + lane] =
                        col_ok ? (((unsigned long long)f32_orderable(c0 + 0.0f) << 32) |
                                  (0xffffffffu - (uint32_t)(i0 + 32 * q + r0)))
                               : 0ull;
                    __syncwarp();  // the tile is rewritten by the next chunk
                }
            }
            // TMEM accumulator fully read by this warp: hand it back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->tmem_empty[acc]);
            if (row_ok && best > -INFINITY)
                atomicMax(P.rowkey + (size_t)b * P.ncap + row,
                          ((unsigned long long)f32_orderable(best + 0.0f) << 32) | (0xffffffffu - (uint32_t)best_j));
            // merge the 4 lane quarters' column keys: 256 epilogue threads, one column each
            asm volatile("bar.sync 1, 256;" ::: "memory");
            {
                const int cidx = threadIdx.x - kEpilogueWarp0 * 32;
                const int j = j0 + cidx;
                if (j < M) {
                    unsigned long long m = colpart[cidx];
#pragma unroll
                    for (int w = 1; w < 4; ++w) m = max(m, colpart[w * TILE_N + cidx]);
                    if (m) atomicMax(P.colkey + (size_t)b * P.mcap + j, m);
                }
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");  // colpart is reused by the next tile
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

// ---- operand preparation --------------------------------------------------------------------- //
__global__ void to_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, size_t n) {
    const size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i + 3 < n) {
        const float4 v = *reinterpret_cast<const float4*>(src + i);
        __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
        uint2 o;
        o.x = *reinterpret_cast<uint32_t*>(&a);
        o.y = *reinterpret_cast<uint32_t*>(&b);
        *reinterpret_cast<uint2*>(dst + i) = o;
    } else {
        for (size_t k = i; k < n; ++k) dst[k] = __float2bfloat16_rn(src[k]);
    }
}
__global__ void split_tf32_kernel(const float* __restrict__ src, float* __restrict__ hi, float* __restrict__ lo, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const float x = src[i];
        uint32_t h;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));
        const float hf = __uint_as_float(h);
        hi[i] = hf;
        lo[i] = __fsub_rn(x, hf);  // exact; the tensor core truncates it to tf32 (error ~2^-22 |x|)
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// 2-D row-major matrix (rows x D) -> tensor map with a (128 B x box_rows) swizzled box
int make_map(einx_ctx* ctx, CUtensorMap* map, void* base, CUtensorMapDataType dt, int elt, size_t rows, int D, int box_rows) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return einx_fail(ctx, EINX_ERR_CUDA, "einx_mnn: cuTensorMapEncodeTiled entry point unavailable");
    cuuint64_t dims[2] = {(cuuint64_t)D, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)D * elt};
    cuuint32_t box[2] = {(cuuint32_t)(KBLOCK_BYTES / elt), (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, dt, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return einx_fail(ctx, EINX_ERR_CUDA, "einx_mnn: cuTensorMapEncodeTiled failed (%d)", (int)r);
    return EINX_OK;
}

uint32_t make_idesc(int kind) {
    // cute::UMMA::InstrDescriptor: c_format F32 (1) @4, a/b format @7/@10 (BF16 = 1, TF32 = 2),
    // a/b K-major (0) @15/@16, N>>3 @17, M>>4 @24
    const uint32_t fmt = kind == 0 ? 1u : 2u;
    return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(TILE_N >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
}

}  // namespace

size_t einx_mnn_tc_scratch_bytes(int B, int ncap, int mcap, int D, int precision) {
    const size_t elems = (size_t)B * ((size_t)ncap + mcap) * D;
    return precision == EINX_MNN_BF16 ? align_up(elems * 2, 1024) + 2048 : 2 * align_up(elems * 4, 1024) + 4096;
}

bool einx_mnn_tc_supported(int D, int precision) {
    // TMA needs 16-byte aligned row pitches
    return precision == EINX_MNN_BF16 ? (D % 8 == 0) : (D % 4 == 0);
}

int einx_mnn_tc(einx_ctx* ctx, const float* d0, const float* d1, const int32_t* n0, const int32_t* n1, int B, int ncap,
                int mcap, int D, int precision, unsigned long long* rowkey, unsigned long long* colkey,
                unsigned char* scratch, size_t scratch_bytes, cudaStream_t stream) {
    const size_t e0 = (size_t)B * ncap * D, e1 = (size_t)B * mcap * D;
    unsigned char* base = (unsigned char*)(((uintptr_t)scratch + 1023) & ~(uintptr_t)1023);
    CUtensorMap maps[4];
    memset(maps, 0, sizeof(maps));
    TcParams P = {};
    P.n0 = n0; P.n1 = n1; P.B = B; P.ncap = ncap; P.mcap = mcap;
    P.tiles_m = (ncap + TILE_M - 1) / TILE_M;
    P.tiles_n = (mcap + TILE_N - 1) / TILE_N;
    P.rowkey = rowkey; P.colkey = colkey;
    int rc;
    if (precision == EINX_MNN_BF16) {
        __nv_bfloat16* a = (__nv_bfloat16*)base;
        __nv_bfloat16* b = a + e0;  // e0 * 2 bytes: keeps 16-byte alignment when D % 8 == 0
        to_bf16_kernel<<<(unsigned)((e0 / 4 + 255) / 256 + 1), 256, 0, stream>>>(d0, a, e0);
        EINX_CHECK_LAUNCH(ctx);
        to_bf16_kernel<<<(unsigned)((e1 / 4 + 255) / 256 + 1), 256, 0, stream>>>(d1, b, e1);
        EINX_CHECK_LAUNCH(ctx);
        if ((rc = make_map(ctx, &maps[0], a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (size_t)B * ncap, D, TILE_M))) return rc;
        if ((rc = make_map(ctx, &maps[2], b, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (size_t)B * mcap, D, TILE_N))) return rc;
        maps[1] = maps[0];
        maps[3] = maps[2];
        P.nkb = (D * 2 + KBLOCK_BYTES - 1) / KBLOCK_BYTES;
        P.passes = 1;
        P.idesc = make_idesc(0);
    } else {
        float* ahi = (float*)base;
        float* alo = ahi + e0;
        float* bhi = alo + e0;
        float* blo = bhi + e1;
        split_tf32_kernel<<<(unsigned)((e0 + 255) / 256), 256, 0, stream>>>(d0, ahi, alo, e0);
        EINX_CHECK_LAUNCH(ctx);
        split_tf32_kernel<<<(unsigned)((e1 + 255) / 256), 256, 0, stream>>>(d1, bhi, blo, e1);
        EINX_CHECK_LAUNCH(ctx);
        if ((rc = make_map(ctx, &maps[0], ahi, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (size_t)B * ncap, D, TILE_M))) return rc;
        if ((rc = make_map(ctx, &maps[1], alo, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (size_t)B * ncap, D, TILE_M))) return rc;
        if ((rc = make_map(ctx, &maps[2], bhi, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (size_t)B * mcap, D, TILE_N))) return rc;
        if ((rc = make_map(ctx, &maps[3], blo, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (size_t)B * mcap, D, TILE_N))) return rc;
        P.nkb = (D * 4 + KBLOCK_BYTES - 1) / KBLOCK_BYTES;
        P.passes = 3;
        P.idesc = make_idesc(1);
    }
    const size_t smem = 1024 + (size_t)STAGES * (A_BYTES + B_BYTES) + 128 + 4 * TILE_N * sizeof(unsigned long long) +
                        (size_t)kEpilogueWarps * 32 * kScratchPitch * sizeof(float);
    const int total_tiles = B * P.tiles_m * P.tiles_n;
    int grid = ctx->num_sms < total_tiles ? ctx->num_sms : total_tiles;
    if (grid < 1) grid = 1;
    if (precision == EINX_MNN_BF16) {
        EINX_CUDA(ctx, cudaFuncSetAttribute(mnn_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        einx_prof_begin(ctx, 3, stream);
        mnn_tc_kernel<0><<<grid, kThreads, smem, stream>>>(maps[0], maps[1], maps[2], maps[3], P);
    } else {
        EINX_CUDA(ctx, cudaFuncSetAttribute(mnn_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        einx_prof_begin(ctx, 3, stream);
        mnn_tc_kernel<1><<<grid, kThreads, smem, stream>>>(maps[0], maps[1], maps[2], maps[3], P);
    }
    einx_prof_end(ctx, 3, stream);
    EINX_CHECK_LAUNCH(ctx);
    (void)scratch_bytes;
    return EINX_OK;
}
