// Tensor-core similarity tiles for the MNN matcher: tcgen05.mma with TMEM accumulators, operands
// staged by TMA (128-byte swizzle), row/column argmax fused into the TMEM epilogue.
// Semantics: reference core/modules/matchers/MNN.py:88-92 (einsum + 2x topk(1)); the thresholds and
// the mutual check run in mnn.cu's finalisation kernel on the keys this kernel produces.
//
// Persistent, warp-specialised CTA (one per SM; the two CTAs of a cluster share a 256-row tile in the pair form):
//   warp 0      TMA producer   -- per k-block an A tile of 128 rows and a B tile of 256 (pair form: 128) rows
//   warp 1      MMA issuer     -- one lane issues tcgen05.mma (M = 128 or 256 across the pair, N = 256, K = 32 bytes)
//   warp 2      TMEM allocator -- 512 columns = two 128x256 fp32 accumulators (double buffered)
//   warps 4..   epilogue       -- tcgen05.ld 32 columns at a time, one chunk ahead; row argmax as a tournament over
//                                 the chunk's 32 registers, column argmax the same way after a padded 32x32
//                                 shared-memory transpose; one 64-bit atomicMax per row / column into the key arrays
//   then        converters     -- split modes only: derive the low-order operand tiles in shared memory
// The similarity matrix therefore never leaves the SM.
//
// Precision modes (include/einx.h): BF16 rounds the descriptors to bf16 (kind::f16).  TF32X3 splits every fp32
// descriptor into hi = tf32(x) and lo = x - hi and accumulates hi*hi + hi*lo + lo*hi (kind::tf32) by running the
// k-loop three times over different operand pairs.  FP16X3 does the same with hi = fp16(2^10 x), lo = fp16(2^10 x - hi)
// (kind::f16): the same 22 significant bits at the fp16 tensor rate and half the operand bytes.  Both splits are
// fp32-accurate to a few 1e-7 while staying on the tensor pipe.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include "common.cuh"
#include "mnn_keys.cuh"

namespace {

constexpr int TILE_M = 128;          // rows of d0 per tile  (UMMA M, TMEM lanes)
constexpr int TILE_N = 256;          // rows of d1 per tile  (UMMA N, TMEM columns)
constexpr int kEpilogueWarp0 = 4;
// warp budgets (build knobs; the defaults are the measured optimum on B200)
#ifndef EINX_SPLIT_EPI_WARPS
#define EINX_SPLIT_EPI_WARPS 8
#endif
#ifndef EINX_BF16_EPI_WARPS
#define EINX_BF16_EPI_WARPS 8
#endif
#ifndef EINX_FP16_EPI_WARPS
#define EINX_FP16_EPI_WARPS 8
#endif
#ifndef EINX_SPLIT_CONV_WARPS
#define EINX_SPLIT_CONV_WARPS 4
#endif
constexpr int kScratchPitch = 33;    // floats; 32x32 transpose tile per epilogue warp, conflict-free both ways
constexpr uint32_t kTmemCols = 512;

// Per-precision shape of the pipeline.
//   BF16   : stages of (A 16 KB | B 32 KB / CG), 8 epilogue warps.
//   TF32X3 : every stage also holds the low-order tiles (A lo | B lo) that the converter warps derive
//            from the raw fp32 tiles in shared memory, so each descriptor byte crosses L2 -> SM once
//            per tile and no split copy of the descriptors ever exists in HBM.
// CG = 2 is the CTA-pair form (tcgen05 cta_group::2): the two CTAs of a cluster own the two 128-row
// halves of a 256 x 256 tile.  Each loads its own A half and HALF of the B tile, the leader issues one
// M=256 MMA that reads both CTAs' shared memory, and each CTA's accumulator half lands in its own
// TMEM.  Per CTA that is 1/3 less operand traffic from L2 and out of shared memory than two
// independent 128 x 256 tiles -- the two limits the single-CTA kernel runs into at D = 128..256.
template <int KIND, int KB, int CG, int EW = 0>  // EW: epilogue warps of the FP16X3 pair form (0 = default)
struct Cfg {
    // KIND 0: bf16 operands prepared in HBM.  KIND 1: fp32 operands, tf32 MMAs, lo tiles derived in shared
    // memory.  KIND 2: fp32 operands, fp16 MMAs: converter warps write hi / lo fp16 tiles (half the row
    // width of the raw tile, 64-byte swizzle) and the MMAs read only those.
    // KIND 3: fp16 hi / lo operands already split in HBM (by the sampler that produced the descriptors, or by
    // split_fp16_kernel): four TMA loads per stage, no converter warps -- the pipeline is pure TMA -> MMA.
    static constexpr int kElt = (KIND == 0 || KIND == 3) ? 2 : 4;   // bytes per element of the TMA'd (raw) tiles
    static constexpr int kKB = KB;                       // bytes of K per raw stage row (= swizzle span)
    static constexpr int kABytes = TILE_M * KB;
    static constexpr int kBRows = TILE_N / CG;           // rows of d1 this CTA stages per k-block
    static constexpr int kBBytes = kBRows * KB;
    static constexpr bool kConvert = KIND == 1 || KIND == 2;
    static constexpr bool kPresplit = KIND == 3;
    static constexpr int kRawBytes = kABytes + kBBytes;
    static constexpr int kStageBytes = kRawBytes * ((kConvert || kPresplit) ? 2 : 1);   // + lo tiles
    // operand tiles the MMAs read: row width in bytes and tile sizes
    static constexpr int kOpKB = KIND == 2 ? KB / 2 : KB;
    static constexpr int kOpABytes = TILE_M * kOpKB;
    static constexpr int kOpBBytes = kBRows * kOpKB;
    // (single-CTA split kernels keep 4 + 8: their stages are 1.5x larger and two must fit beside the scratch)
    static constexpr int kEpiWarps = KIND == 3 ? (EW ? EW : 8) : KIND == 0 ? EINX_BF16_EPI_WARPS : (CG == 1 ? 4 : (KIND == 1 ? EINX_SPLIT_EPI_WARPS : (EW ? EW : EINX_FP16_EPI_WARPS)));
    static constexpr int kConvWarps = !kConvert ? 0 : (CG == 1 ? 8 : (KIND == 1 ? EINX_SPLIT_CONV_WARPS : 8));
    static constexpr int kConvWarp0 = kEpilogueWarp0 + kEpiWarps;
    static constexpr int kColsPerWarp = TILE_N / (kEpiWarps / 4);
    // warps: 0 TMA, 1 MMA, 2 TMEM alloc, 3 idle, 4.. epilogue, then converters
    static constexpr int kThreads = 32 * (kEpilogueWarp0 + kEpiWarps + kConvWarps);
    // alignment slack + barriers + one 32x32 transpose tile per epilogue warp (reused for the column-key merge)
    static constexpr size_t kFixedSmem = 1024 + 256 + (size_t)kEpiWarps * 32 * kScratchPitch * sizeof(float);
    static constexpr int kStagesFit = (int)((227 * 1024 - kFixedSmem) / kStageBytes);
    static constexpr int kStages = kStagesFit > 8 ? 8 : kStagesFit;
    static constexpr size_t kSmem = kFixedSmem + (size_t)kStages * kStageBytes;
    static_assert(kStages >= 2 && kStages <= 8, "Barriers holds 8 stages");
};

struct TcParams {
    const int32_t* n0;
    const int32_t* n1;
    int B, ncap, mcap;
    int tiles_m, tiles_n;  // per pair
    int nkb;               // k-blocks per tile
    uint32_t idesc;
    float in_scale, out_scale;  // FP16X3: operands are scaled by in_scale before the split, similarities by out_scale after
    unsigned long long* rowkey;
    unsigned long long* colkey;
};

struct __align__(8) Barriers {   // 240 bytes; in the pair form the LEADER's full / conv / tmem_empty are the live ones
    unsigned long long full[8];    // TMA landed the raw tiles of a stage (both CTAs' tiles in the pair form)
    unsigned long long conv[8];    // converter warps wrote the lo tiles of a stage (TF32X3)
    unsigned long long empty[8];   // the MMAs reading a stage have completed
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
// Spin on the phase parity.  A pipeline bug must not hang the GPU: after ~2^31 cycles the kernel
// traps, which surfaces as a CUDA error on the host instead of a dead device.
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done;
    long long t0 = 0;
    for (uint32_t spins = 0;; ++spins) {
        asm volatile(
            "{\n.reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (done) return;
        if ((spins & 0xfffu) == 0xfffu) {
            const long long now = clock64();
            if (t0 == 0) t0 = now;
            else if (now - t0 > (1ll << 31)) __trap();
        }
    }
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
// ---- CTA-pair (cta_group::2) forms ---- //
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;  // clears the CTA-rank bit of a shared::cluster address: "the leader's copy"
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Arrive on the barrier at the same shared-memory offset in CTA `cta` of the cluster.  Plain (CTA-scope
// release) form, as CUTLASS's ClusterBarrier::arrive(cta_id): a `.release.cluster` arrive costs a
// cluster-wide memory fence (~0.8 us per call measured under ncu, stall_membar) and none of the
// hand-offs below publishes generic-proxy global data -- TMEM reads are ordered by the tcgen05
// fences, shared-memory tiles by fence.proxy.async before the arrive.
__device__ __forceinline__ void mbar_arrive_cta(unsigned long long* bar, uint32_t cta) {
    asm volatile(
        "{\n.reg .b32 ra;\n"
        "mapa.shared::cluster.u32 ra, %0, %1;\n"
        "mbarrier.arrive.shared::cluster.b64 _, [ra];\n}\n" ::"r"(smem_u32(bar)),
        "r"(cta)
        : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(unsigned long long* bar, uint32_t parity) {
    // like mbar_wait, acquiring at cluster scope (the arrivals come from the peer CTA as well)
    const uint32_t addr = smem_u32(bar);
    uint32_t done;
    long long t0 = 0;
    for (uint32_t spins = 0;; ++spins) {
        asm volatile(
            "{\n.reg .pred p;\n"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (done) return;
        if ((spins & 0xfffu) == 0xfffu) {
            const long long now = clock64();
            if (t0 == 0) t0 = now;
            else if (now - t0 > (1ll << 31)) __trap();
        }
    }
}
// TMA load whose completion bytes are credited to the LEADER CTA's barrier
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* map, unsigned long long* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1)
        : "memory");
}
// MMA completion -> the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void tc_commit_pair(unsigned long long* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "h"((uint16_t)3)
                 : "memory");
}
template <int KIND, int CG>  // KIND 0: kind::f16 (bf16 inputs), 1: kind::tf32; CG: cta_group
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    if (KIND != 1 && CG == 1) {
        asm volatile(
            "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
            : "memory");
    } else if (KIND != 1) {
        asm volatile(
            "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
            : "memory");
    } else if (CG == 1) {
        asm volatile(
            "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
            : "memory");
    } else {
        asm volatile(
            "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
            "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
            : "memory");
    }
}
__device__ __forceinline__ void tc_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
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
}
// The registers of an issued load are written asynchronously until wait::ld; passing them through
// the wait as in/out operands makes every later use depend on it (the compiler cannot hoist one).
__device__ __forceinline__ void tc_ld_wait(uint32_t (&r)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                   "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]),
                   "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]),
                   "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
                 :
                 : "memory");
}

// K-major operand tile whose rows are one swizzle span (KB = 128 or 64 bytes) wide: 8-row groups are
// 8*KB bytes apart (SBO), descriptor version 1 (Blackwell), layout type 2 (SWIZZLE_128B) or 4
// (SWIZZLE_64B).  The start address advances by 32 B per UMMA_K step inside the span.
template <int KB>
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3ffff) >> 4);   // start address, bits [0,14)
    d |= (uint64_t)1 << 16;                        // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)((8 * KB) >> 4) << 32;          // stride byte offset, bits [32,46)
    d |= (uint64_t)1 << 46;                        // version
    d |= (uint64_t)(KB == 128 ? 2 : 4) << 61;      // swizzle mode
    return d;
}

// low-order part of an fp32 value w.r.t. its tf32 truncation: kind::tf32 reads the top 19 bits of
// each operand word, so hi = x & 0xffffe000 is what the tensor core sees of the raw tile and
// lo = x - hi is exact in fp32 (13 significant bits, of which the tensor core keeps 11)
__device__ __forceinline__ float tf32_lo(float x) {
    return __fsub_rn(x, __uint_as_float(__float_as_uint(x) & 0xffffe000u));
}

template <int KIND, int KB, int CG, int EW = 0>
__global__ void __launch_bounds__(Cfg<KIND, KB, CG, EW>::kThreads, 1)
mnn_tc_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
              const __grid_constant__ CUtensorMap mapA2, const __grid_constant__ CUtensorMap mapB2, const TcParams P) {
    // (mapA2 / mapB2: the low-order operand matrices of the pre-split form, KIND 3; unused otherwise)
    using C = Cfg<KIND, KB, CG, EW>;
    constexpr int A_BYTES = C::kABytes, B_BYTES = C::kBBytes, RAW_BYTES = C::kRawBytes, STAGES = C::kStages,
                  STAGE_BYTES = C::kStageBytes;
    extern __shared__ __align__(1024) unsigned char smem[];
    // carve: [stages x (A | B [| A lo | B lo])] 1024-aligned, then barriers, then the epilogue warps' transpose tiles
    // (offset arithmetic on the shared array, not an integer round trip: the compiler must keep the
    // shared address space, or every access below becomes a generic LD.E/ST.E)
    unsigned char* tiles = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);
    Barriers* bars = reinterpret_cast<Barriers*>(tiles + (size_t)STAGES * STAGE_BYTES);
    float* scratch_all = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(bars) + 256);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;  // which 128-row half of the pair's tile
    const bool leader = rank == 0;
    const int tiles_per_pair = P.tiles_m * P.tiles_n;
    const int total_tiles = P.B * tiles_per_pair;
    const int tile0 = blockIdx.x / CG, tile_step = gridDim.x / CG;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB) : "memory");
        if (C::kPresplit) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA2) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB2) : "memory");
        }
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&bars->full[s], (CG == 2 && !C::kConvert) ? 2 : 1);  // pair form: one arrival per producer
            mbar_init(&bars->conv[s], CG * C::kConvWarps);     // one arrival per converter warp of the pair
            mbar_init(&bars->empty[s], 1);
        }
        for (int a = 0; a < 2; ++a) { mbar_init(&bars->tmem_full[a], 1); mbar_init(&bars->tmem_empty[a], CG * C::kEpiWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        if (CG == 1) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)),
                         "r"(kTmemCols)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)),
                         "r"(kTmemCols)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    if (CG == 2) cluster_sync_all(); else __syncthreads();  // barriers initialised in both CTAs before any remote arrive
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;

    // tile t of the schedule: pair b, rows [i0, i0 + 128*CG) of d0 (this CTA: the 128 from my_i0), columns [j0, j0+256)
    auto tile_coords = [&](int t, int& b, int& i0, int& j0, bool& live) {
        b = t / tiles_per_pair;
        const int r = t - b * tiles_per_pair;
        i0 = (r / P.tiles_n) * (TILE_M * CG);
        j0 = (r % P.tiles_n) * TILE_N;
        const int N = P.n0 ? min(P.n0[b], P.ncap) : P.ncap;
        const int M = P.n1 ? min(P.n1[b], P.mcap) : P.mcap;
        live = (i0 < N) && (j0 < M);  // identical in both CTAs of a pair
    };

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int t = tile0; t < total_tiles; t += tile_step) {
                int b, i0, j0;
                bool live;
                tile_coords(t, b, i0, j0, live);
                if (!live) continue;
                const int rowA = b * P.ncap + i0 + (int)rank * TILE_M;
                const int rowB = b * P.mcap + j0 + (int)rank * C::kBRows;
                for (int kb = 0; kb < P.nkb; ++kb) {
                    mbar_wait(&bars->empty[stage], phase ^ 1);
                    unsigned char* sa = tiles + (size_t)stage * STAGE_BYTES;
                    const int kcoord = kb * (KB / C::kElt);
                    if (CG == 1 || C::kConvert) {
                        // (3xTF32 pair form: each CTA's converters wait for their own tiles, so the bytes are
                        // counted locally; the leader's MMA waits for the converters of both CTAs instead)
                        mbar_expect_tx(&bars->full[stage], RAW_BYTES * (C::kPresplit ? 2 : 1));
                        tma_load_2d(sa, &mapA, &bars->full[stage], kcoord, rowA);
                        tma_load_2d(sa + A_BYTES, &mapB, &bars->full[stage], kcoord, rowB);
                        if (C::kPresplit) {
                            tma_load_2d(sa + RAW_BYTES, &mapA2, &bars->full[stage], kcoord, rowA);
                            tma_load_2d(sa + RAW_BYTES + A_BYTES, &mapB2, &bars->full[stage], kcoord, rowB);
                        }
                    } else {
                        // both CTAs' bytes complete on the leader's barrier; the peer adds its arrival remotely
                        if (leader) mbar_expect_tx(&bars->full[stage], 2 * RAW_BYTES * (C::kPresplit ? 2 : 1));
                        else mbar_arrive_cta(&bars->full[stage], 0);
                        tma_load_2d_pair(sa, &mapA, &bars->full[stage], kcoord, rowA);
                        tma_load_2d_pair(sa + A_BYTES, &mapB, &bars->full[stage], kcoord, rowB);
                        if (C::kPresplit) {
                            tma_load_2d_pair(sa + RAW_BYTES, &mapA2, &bars->full[stage], kcoord, rowA);
                            tma_load_2d_pair(sa + RAW_BYTES + A_BYTES, &mapB2, &bars->full[stage], kcoord, rowB);
                        }
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (the leader CTA's, in the pair form) =====
        if (lane == 0 && leader) {
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int t = tile0; t < total_tiles; t += tile_step) {
                int b, i0, j0;
                bool live;
                tile_coords(t, b, i0, j0, live);
                if (!live) continue;
                if (CG == 2) mbar_wait_cluster(&bars->tmem_empty[acc], acc_phase ^ 1);
                else mbar_wait(&bars->tmem_empty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)acc * TILE_N;
                for (int kb = 0; kb < P.nkb; ++kb) {
                    if (!(C::kConvert && CG == 2)) {
                        if (CG == 2) mbar_wait_cluster(&bars->full[stage], phase);
                        else mbar_wait(&bars->full[stage], phase);
                    }
                    if (C::kConvert) {  // every converter warp (of both CTAs) saw its tiles land and wrote the lo tiles
                        if (CG == 2) mbar_wait_cluster(&bars->conv[stage], phase);
                        else mbar_wait(&bars->conv[stage], phase);
                    }
                    tc_fence_after();
                    const uint32_t sa = smem_u32(tiles + (size_t)stage * STAGE_BYTES);
                    if (KIND == 2) {
                        // fp16 hi / lo tiles behind the raw ones: [A hi | B hi | A lo | B lo], 64-byte-wide rows
                        constexpr int OKB = C::kOpKB, OA = C::kOpABytes, OB = C::kOpBBytes;
                        const uint32_t so = sa + RAW_BYTES;
                        const uint64_t a_hi = make_smem_desc<OKB>(so), b_hi = make_smem_desc<OKB>(so + OA);
                        const uint64_t a_lo = make_smem_desc<OKB>(so + OA + OB), b_lo = make_smem_desc<OKB>(so + 2 * OA + OB);
#pragma unroll
                        for (int k = 0; k < OKB / 32; ++k)
                            tc_mma<KIND, CG>(tmem_d, a_hi + (uint64_t)(2 * k), b_hi + (uint64_t)(2 * k), P.idesc, (kb > 0 || k > 0) ? 1u : 0u);
#pragma unroll
                        for (int k = 0; k < OKB / 32; ++k)
                            tc_mma<KIND, CG>(tmem_d, a_hi + (uint64_t)(2 * k), b_lo + (uint64_t)(2 * k), P.idesc, 1u);
#pragma unroll
                        for (int k = 0; k < OKB / 32; ++k)
                            tc_mma<KIND, CG>(tmem_d, a_lo + (uint64_t)(2 * k), b_hi + (uint64_t)(2 * k), P.idesc, 1u);
                    } else if (KIND == 1 || KIND == 3) {
                        // x.y ~= hi.hi + hi.lo + lo.hi  (3xTF32: the raw tile is its own hi part -- the tensor
                        // core drops the 13 low mantissa bits of a tf32 operand; pre-split fp16: the hi and lo
                        // tiles arrive by TMA at the same offsets)
                        const uint64_t a_hi = make_smem_desc<KB>(sa), b_hi = make_smem_desc<KB>(sa + A_BYTES);
                        const uint64_t a_lo = make_smem_desc<KB>(sa + RAW_BYTES);
                        const uint64_t b_lo = make_smem_desc<KB>(sa + RAW_BYTES + A_BYTES);
#pragma unroll
                        for (int k = 0; k < KB / 32; ++k)
                            tc_mma<KIND, CG>(tmem_d, a_hi + (uint64_t)(2 * k), b_hi + (uint64_t)(2 * k), P.idesc, (kb > 0 || k > 0) ? 1u : 0u);
#pragma unroll
                        for (int k = 0; k < KB / 32; ++k)
                            tc_mma<KIND, CG>(tmem_d, a_hi + (uint64_t)(2 * k), b_lo + (uint64_t)(2 * k), P.idesc, 1u);
#pragma unroll
                        for (int k = 0; k < KB / 32; ++k)
                            tc_mma<KIND, CG>(tmem_d, a_lo + (uint64_t)(2 * k), b_hi + (uint64_t)(2 * k), P.idesc, 1u);
                    } else {
                        const uint64_t a_hi = make_smem_desc<KB>(sa), b_hi = make_smem_desc<KB>(sa + A_BYTES);
#pragma unroll
                        for (int k = 0; k < KB / 32; ++k) {
                            // +32 B along K inside the swizzle span = +2 in the 16-byte-unit address field
                            tc_mma<KIND, CG>(tmem_d, a_hi + (uint64_t)(2 * k), b_hi + (uint64_t)(2 * k), P.idesc, (kb > 0 || k > 0) ? 1u : 0u);
                        }
                    }
                    // smem slot free (in both CTAs) once these MMAs have read it
                    if (CG == 2) tc_commit_pair(&bars->empty[stage]); else tc_commit(&bars->empty[stage]);
                    if (kb == P.nkb - 1) {
                        if (CG == 2) tc_commit_pair(&bars->tmem_full[acc]); else tc_commit(&bars->tmem_full[acc]);
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            }
        }
    } else if (C::kConvert && warp >= C::kConvWarp0 && warp < C::kConvWarp0 + C::kConvWarps) {
        // ===== converter: lo tiles = x - tf32_trunc(x), element-wise (so the swizzle is irrelevant) =====
        const int ct = threadIdx.x - C::kConvWarp0 * 32;
        constexpr int kChunks = RAW_BYTES / 16;
        int stage = 0;
        uint32_t phase = 0;
        for (int t = tile0; t < total_tiles; t += tile_step) {
            int b, i0, j0;
            bool live;
            tile_coords(t, b, i0, j0, live);
            if (!live) continue;
            for (int kb = 0; kb < P.nkb; ++kb) {
                mbar_wait(&bars->full[stage], phase);
                const float4* raw = reinterpret_cast<const float4*>(tiles + (size_t)stage * STAGE_BYTES);
                if (KIND == 1) {
                    float4* lo = reinterpret_cast<float4*>(tiles + (size_t)stage * STAGE_BYTES + RAW_BYTES);
#pragma unroll 8
                    for (int i = ct; i < kChunks; i += C::kConvWarps * 32) {
                        const float4 v = raw[i];
                        lo[i] = make_float4(tf32_lo(v.x), tf32_lo(v.y), tf32_lo(v.z), tf32_lo(v.w));
                    }
                } else {
                    // fp16 split: x' = in_scale * x, hi = fp16(x'), lo = fp16(x' - hi).  The raw tile is
                    // 128-byte swizzled (16-byte chunk c of row r sits at chunk c ^ (r & 7)); the fp16 tiles
                    // have 64-byte rows, 64-byte swizzled (chunk c' of row r at c' ^ ((r >> 1) & 3)), so the
                    // four floats of raw chunk c become the (c & 1) half of fp16 chunk c >> 1.
                    static_assert(KIND != 2 || KB == 128, "FP16X3 stages 32-element k-blocks");
                    unsigned char* ops = tiles + (size_t)stage * STAGE_BYTES + RAW_BYTES;
                    constexpr int kConvThreads = KIND == 2 ? C::kConvWarps * 32 : 256;  // (the branch is dead for the other kinds)
                    constexpr int kAIters = A_BYTES / 16 / kConvThreads, kBIters = B_BYTES / 16 / kConvThreads;
                    static_assert(KIND != 2 || (kConvThreads == 256 && kAIters * kConvThreads * 16 == A_BYTES &&
                                                kBIters * kConvThreads * 16 == B_BYTES),
                                  "a thread's chunks are 32 rows apart: its swizzle terms are loop invariant");
                    const float sc = P.in_scale;
                    // thread ct owns raw chunk ct of every 32-row slab: row r0 + 32 n, physical chunk ct & 7, so the
                    // logical chunk c and both swizzle terms depend on ct only and the destination advances 2 KB per slab
                    const int r0 = ct >> 3, c = (ct & 7) ^ (r0 & 7);
                    const uint32_t off0 = (uint32_t)r0 * 64u + (uint32_t)(((c >> 1) ^ ((r0 >> 1) & 3)) << 4) + (uint32_t)((c & 1) << 3);
                    auto split4 = [&](const float4 v, unsigned char* hi_t, unsigned char* lo_t, uint32_t off) {
                        const float x0 = v.x * sc, x1 = v.y * sc, x2 = v.z * sc, x3 = v.w * sc;
                        const __half2 h01 = __floats2half2_rn(x0, x1), h23 = __floats2half2_rn(x2, x3);
                        const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
                        const __half2 l01 = __floats2half2_rn(x0 - f01.x, x1 - f01.y), l23 = __floats2half2_rn(x2 - f23.x, x3 - f23.y);
                        uint2 hv, lv;
                        hv.x = *reinterpret_cast<const uint32_t*>(&h01); hv.y = *reinterpret_cast<const uint32_t*>(&h23);
                        lv.x = *reinterpret_cast<const uint32_t*>(&l01); lv.y = *reinterpret_cast<const uint32_t*>(&l23);
                        *reinterpret_cast<uint2*>(hi_t + off) = hv;
                        *reinterpret_cast<uint2*>(lo_t + off) = lv;
                    };
                    unsigned char* a_hi = ops;
                    unsigned char* b_hi = ops + C::kOpABytes;
                    constexpr int kLoOff = C::kOpABytes + C::kOpBBytes;
#pragma unroll
                    for (int n = 0; n < kAIters; ++n) split4(raw[ct + n * kConvThreads], a_hi, a_hi + kLoOff, off0 + 2048u * n);
#pragma unroll 4
                    for (int n = 0; n < kBIters; ++n)
                        split4(raw[A_BYTES / 16 + ct + n * kConvThreads], b_hi, b_hi + kLoOff, off0 + 2048u * n);
                }
                // generic-proxy writes must be visible to the tensor core's async-proxy reads
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) {
                    if (CG == 2) mbar_arrive_cta(&bars->conv[stage], 0);
                    else mbar_arrive(&bars->conv[stage]);
                }
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp >= kEpilogueWarp0 && warp < kEpilogueWarp0 + C::kEpiWarps) {
        // ===== epilogue: TMEM -> registers -> row / column best keys =====
        // Rows: the thread that owns TMEM lane r finds its row's best column with a tournament over
        // each 32-column chunk (lowest column wins ties).  Columns: the 32x32 chunk goes through a
        // padded shared-memory transpose so that lane c then owns column c (lowest row wins).
        const int q = warp & 3;                          // TMEM lane quarter this warp may access
        const int part = (warp - kEpilogueWarp0) >> 2;   // which kColsPerWarp columns of the tile
        float* scratch = scratch_all + (size_t)(warp - kEpilogueWarp0) * 32 * kScratchPitch;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int t = tile0; t < total_tiles; t += tile_step) {
            int b, i0, j0;
            bool live;
            tile_coords(t, b, i0, j0, live);
            if (!live) continue;
            i0 += (int)rank * TILE_M;  // this CTA's half of the pair's rows
            const int N = P.n0 ? min(P.n0[b], P.ncap) : P.ncap;
            const int M = P.n1 ? min(P.n1[b], P.mcap) : P.mcap;
            const int row = i0 + 32 * q + lane;  // this thread's row of d0
            const bool row_ok = row < N;
            const bool full_tile = (i0 + TILE_M <= N) && (j0 + TILE_N <= M);
            mbar_wait(&bars->tmem_full[acc], acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(acc * TILE_N + C::kColsPerWarp * part);
            float best = -INFINITY;
            int best_j = 0;
            constexpr int kChunksPerWarp = C::kColsPerWarp / 32;
            // software pipeline: the TMEM load of chunk c+1 is in flight while chunk c is reduced
            uint32_t v[2][32];
            unsigned long long ckey[kChunksPerWarp];  // this lane's column of every chunk: best (value, row) so far
            tc_ld32_issue(taddr, v[0]);
#pragma unroll
            for (int c = 0; c < kChunksPerWarp; ++c) {
                tc_ld_wait(v[c & 1]);
                if (c + 1 < kChunksPerWarp) tc_ld32_issue(taddr + 32 * (c + 1), v[(c + 1) & 1]);
                const int jc = j0 + C::kColsPerWarp * part + 32 * c;
                if (jc < M) {  // warp-uniform; beyond M the tile is padding
                    float f[32];
#pragma unroll
                    for (int k = 0; k < 32; ++k) f[k] = __uint_as_float(v[c & 1][k]);
                    if (!full_tile) {
                        // ragged edge: padding columns and rows can never win a strict '>'
#pragma unroll
                        for (int k = 0; k < 32; ++k)
                            if (jc + k >= M || !row_ok) f[k] = -INFINITY;
                    }
#pragma unroll
                    for (int k = 0; k < 32; ++k) scratch[lane * kScratchPitch + k] = f[k];
                    // row: tournament over this thread's 32 columns (left wins ties = lowest column)
                    float rv;
                    int rk;
                    argmax32(f, rv, rk);
                    if (rv > best) { best = rv; best_j = jc + rk; }
                    __syncwarp();
                    // column: lane c owns column c of the transposed tile (upper wins ties = lowest row)
                    float g[32];
#pragma unroll
                    for (int r = 0; r < 32; ++r) g[r] = scratch[r * kScratchPitch + lane];
                    float cv;
                    int cr;
                    argmax32(g, cv, cr);
                    const bool col_ok = (jc + lane < M) && (cv > -INFINITY);
                    ckey[c] = col_ok ? (((unsigned long long)f32_orderable(((KIND == 2 || KIND == 3) ? cv * P.out_scale : cv) + 0.0f) << 32) |
                                        (0xffffffffu - (uint32_t)(i0 + 32 * q + cr)))
                                     : 0ull;
                    __syncwarp();  // the tile is rewritten by the next chunk
                } else {
                    ckey[c] = 0ull;
                }
            }
            // TMEM accumulator fully read by this warp: hand it back to the MMA warp (the leader's)
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (CG == 2) mbar_arrive_cta(&bars->tmem_empty[acc], 0);
                else mbar_arrive(&bars->tmem_empty[acc]);
            }
            if (row_ok && best > -INFINITY)
                atomicMax(P.rowkey + (size_t)b * P.ncap + row,
                          ((unsigned long long)f32_orderable(((KIND == 2 || KIND == 3) ? best * P.out_scale : best) + 0.0f) << 32) |
                              (0xffffffffu - (uint32_t)best_j));
            // merge the 4 lane quarters' column keys: every warp parks its keys in its own (now idle) transpose
            // tile, then the threads of the epilogue group take the maximum over the four quarters of a column
            {
                unsigned long long* mine = reinterpret_cast<unsigned long long*>(scratch);
#pragma unroll
                for (int c = 0; c < kChunksPerWarp; ++c) mine[32 * c + lane] = ckey[c];
            }
            asm volatile("bar.sync 1, %0;" ::"n"(C::kEpiWarps * 32) : "memory");
            for (int cidx = threadIdx.x - kEpilogueWarp0 * 32; cidx < TILE_N; cidx += C::kEpiWarps * 32) {
                const int j = j0 + cidx;
                if (j < M) {
                    const int pp = cidx / C::kColsPerWarp, within = cidx - pp * C::kColsPerWarp;  // owning part, slot
                    unsigned long long m = 0ull;
#pragma unroll
                    for (int w = 0; w < 4; ++w) {  // epilogue warp (4 * pp + w) holds lane quarter w of that part
                        const unsigned long long* src = reinterpret_cast<const unsigned long long*>(
                            scratch_all + (size_t)(4 * pp + w) * 32 * kScratchPitch);
                        m = max(m, src[within]);
                    }
                    if (m) atomicMax(P.colkey + (size_t)b * P.mcap + j, m);
                }
            }
            asm volatile("bar.sync 1, %0;" ::"n"(C::kEpiWarps * 32) : "memory");  // the tiles are rewritten by the next tile
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
        }
    }
    tc_fence_before();
    if (CG == 2) cluster_sync_all(); else __syncthreads();  // neither CTA leaves while the other may still touch it
    if (warp == 2) {
        tc_fence_after();
        if (CG == 1)
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
        else
            asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

// ---- operand preparation --------------------------------------------------------------------- //
__global__ void to_bf16_kernel(const float* __restrict__ src0, size_t n0, const float* __restrict__ src1, size_t n1,
                               __nv_bfloat16* __restrict__ dst) {
    // dst = [bf16(src0) | bf16(src1)], both sides in one launch; n0 and n1 are multiples of 8
    const size_t stride = (size_t)gridDim.x * blockDim.x * 4;
    const size_t n = n0 + n1;
    for (size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
        const float4 v = i < n0 ? __ldg(reinterpret_cast<const float4*>(src0 + i))
                                : __ldg(reinterpret_cast<const float4*>(src1 + (i - n0)));
        __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
        uint2 o;
        o.x = *reinterpret_cast<uint32_t*>(&a);
        o.y = *reinterpret_cast<uint32_t*>(&b);
        *reinterpret_cast<uint2*>(dst + i) = o;
    }
}

// fp32 descriptors -> the fp16 hi / lo operand matrices of the FP16X3 split: x' = in_scale * x, hi = fp16(x'),
// lo = fp16(x' - hi).  Used when the caller has no pre-split operands (the sampler writes them itself on the
// pipeline path, so there this pass does not run).  n0, n1 multiples of 4.
__global__ void split_fp16_kernel(const float* __restrict__ src0, size_t n0, const float* __restrict__ src1, size_t n1,
                                  __half* __restrict__ hi0, __half* __restrict__ lo0, __half* __restrict__ hi1,
                                  __half* __restrict__ lo1, float sc) {
    const size_t stride = (size_t)gridDim.x * blockDim.x * 4;
    const size_t n = n0 + n1;
    for (size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
        const bool first = i < n0;
        const size_t j = first ? i : i - n0;
        const float4 v = __ldg(reinterpret_cast<const float4*>((first ? src0 : src1) + j));
        const float x0 = v.x * sc, x1 = v.y * sc, x2 = v.z * sc, x3 = v.w * sc;
        const __half2 h01 = __floats2half2_rn(x0, x1), h23 = __floats2half2_rn(x2, x3);
        const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
        const __half2 l01 = __floats2half2_rn(x0 - f01.x, x1 - f01.y), l23 = __floats2half2_rn(x2 - f23.x, x3 - f23.y);
        uint2 hv, lv;
        hv.x = *reinterpret_cast<const uint32_t*>(&h01); hv.y = *reinterpret_cast<const uint32_t*>(&h23);
        lv.x = *reinterpret_cast<const uint32_t*>(&l01); lv.y = *reinterpret_cast<const uint32_t*>(&l23);
        *reinterpret_cast<uint2*>((first ? hi0 : hi1) + j) = hv;
        *reinterpret_cast<uint2*>((first ? lo0 : lo1) + j) = lv;
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

// 2-D row-major matrix (rows x D) -> tensor map with a (kb bytes x box_rows) swizzled box
int make_map(einx_ctx* ctx, CUtensorMap* map, const void* base, CUtensorMapDataType dt, int elt, size_t rows, int D,
             int box_rows, int kb) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return einx_fail(ctx, EINX_ERR_CUDA, "einx_mnn: cuTensorMapEncodeTiled entry point unavailable");
    cuuint64_t dims[2] = {(cuuint64_t)D, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)D * elt};
    cuuint32_t box[2] = {(cuuint32_t)(kb / elt), (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    kb == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return einx_fail(ctx, EINX_ERR_CUDA, "einx_mnn: cuTensorMapEncodeTiled failed (%d)", (int)r);
    return EINX_OK;
}

uint32_t make_idesc(int kind, int cg) {
    // cute::UMMA::InstrDescriptor: c_format F32 (1) @4, a/b format @7/@10 (BF16 = 1, TF32 = 2),
    // a/b K-major (0) @15/@16, N>>3 @17, M>>4 @24 (M = 256 across the CTA pair)
    const uint32_t fmt = kind == 0 ? 1u : (kind == 1 ? 2u : 0u);  // BF16 / TF32 / F16
    return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(TILE_N >> 3) << 17) | ((uint32_t)((TILE_M * cg) >> 4) << 24);
}

template <int KIND, int KB, int CG, int EW = 0>
int launch_tc(einx_ctx* ctx, const CUtensorMap& ma, const CUtensorMap& mb, const TcParams& P, int grid, cudaStream_t stream,
              const CUtensorMap* ma2 = nullptr, const CUtensorMap* mb2 = nullptr) {
    auto kern = mnn_tc_kernel<KIND, KB, CG, EW>;
    using C = Cfg<KIND, KB, CG, EW>;
    const size_t smem = C::kSmem;
    EINX_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(C::kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    einx_prof_begin(ctx, 3, stream);
    EINX_CUDA(ctx, cudaLaunchKernelEx(&cfg, kern, ma, mb, ma2 ? *ma2 : ma, mb2 ? *mb2 : mb, P));
    einx_prof_end(ctx, 3, stream);
    ctx->launches++;
    return EINX_OK;
}

}  // namespace

static bool presplit_wanted() {
    // EINX_MNN_FP16_INKERNEL=1: the previous FP16X3 form (converter warps split the fp32 tiles in shared memory)
    static const bool inkernel = getenv("EINX_MNN_FP16_INKERNEL") && atoi(getenv("EINX_MNN_FP16_INKERNEL")) != 0;
    return !inkernel;
}

size_t einx_mnn_tc_scratch_bytes(int B, int ncap, int mcap, int D, int precision, bool have_split) {
    const size_t elems = (size_t)B * ((size_t)ncap + mcap) * D;
    if (precision == EINX_MNN_BF16) return align_up(elems * 2, 1024) + 2048;
    if (precision == EINX_MNN_FP16X3 && presplit_wanted() && !have_split && D % 8 == 0) return align_up(elems * 4, 1024) + 2048;
    return 0;  // TF32X3 reads the descriptors in place
}

bool einx_mnn_tc_supported(const float* d0, const float* d1, int D, int precision) {
    // TMA needs 16-byte aligned bases and row pitches
    if (precision == EINX_MNN_BF16) return D % 8 == 0 && ((uintptr_t)d0 % 16 == 0) && ((uintptr_t)d1 % 16 == 0);
    return D % 4 == 0 && ((uintptr_t)d0 % 16 == 0) && ((uintptr_t)d1 % 16 == 0);
}

int einx_mnn_tc(einx_ctx* ctx, const float* d0, const float* d1, const int32_t* n0, const int32_t* n1, int B, int ncap,
                int mcap, int D, int precision, unsigned long long* rowkey, unsigned long long* colkey,
                unsigned char* scratch, size_t scratch_bytes, const uint16_t* split0, const uint16_t* split1,
                cudaStream_t stream) {
    const size_t e0 = (size_t)B * ncap * D, e1 = (size_t)B * mcap * D;
    CUtensorMap maps[2];
    memset(maps, 0, sizeof(maps));
    TcParams P = {};
    P.n0 = n0; P.n1 = n1; P.B = B; P.ncap = ncap; P.mcap = mcap;
    // CTA pairs (cta_group::2, 256-row tiles) where the operand path is the limit: measured on B200 at
    // 64x1024x1024x256 / 32x2048x2048x128 / 1x8192x8192x128 the pair form takes 3xTF32 from 150 / 185 /
    // 102 us to 136 / 163 / 92 us (tensor pipe 60 % -> 76 % active), while the single-pass bf16 kernel is
    // bound by its argmax epilogue and is 5 % faster unpaired (49 vs 52 us).  EINX_MNN_CTA_PAIR=0/1 forces.
    static const int pair_env = getenv("EINX_MNN_CTA_PAIR") ? atoi(getenv("EINX_MNN_CTA_PAIR")) : -1;
    const bool want_pair = pair_env < 0 ? precision != EINX_MNN_BF16 : pair_env != 0;
    const int CG = (want_pair && ncap > TILE_M) ? 2 : 1;
    P.tiles_m = (ncap + TILE_M * CG - 1) / (TILE_M * CG);
    P.tiles_n = (mcap + TILE_N - 1) / TILE_N;
    P.rowkey = rowkey; P.colkey = colkey;
    const int total_tiles = B * P.tiles_m * P.tiles_n;
    int grid = (ctx->num_sms / CG) < total_tiles ? (ctx->num_sms / CG) : total_tiles;
    if (grid < 1) grid = 1;
    grid *= CG;
    int rc;
    (void)scratch_bytes;
    if (precision == EINX_MNN_BF16) {
        __nv_bfloat16* a = (__nv_bfloat16*)(((uintptr_t)scratch + 1023) & ~(uintptr_t)1023);
        __nv_bfloat16* b = a + e0;  // e0 * 2 bytes: keeps 16-byte alignment when D % 8 == 0
        const size_t quads = (e0 + e1) / 4;
        unsigned blocks = (unsigned)((quads + 255) / 256);
        if (blocks > (unsigned)ctx->num_sms * 16) blocks = (unsigned)ctx->num_sms * 16;
        to_bf16_kernel<<<blocks ? blocks : 1, 256, 0, stream>>>(d0, e0, d1, e1, a);
        EINX_CHECK_LAUNCH(ctx);
        if ((rc = make_map(ctx, &maps[0], a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (size_t)B * ncap, D, TILE_M, 128))) return rc;
        if ((rc = make_map(ctx, &maps[1], b, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (size_t)B * mcap, D, TILE_N / CG, 128))) return rc;
        P.nkb = (D * 2 + 127) / 128;
        P.idesc = make_idesc(0, CG);
        return CG == 2 ? launch_tc<0, 128, 2>(ctx, maps[0], maps[1], P, grid, stream)
                       : launch_tc<0, 128, 1>(ctx, maps[0], maps[1], P, grid, stream);
    }
    if (precision == EINX_MNN_FP16X3 && presplit_wanted() && D % 8 == 0) {
        // Pre-split fp16 operands: hi = fp16(2^10 d), lo = fp16(2^10 d - hi), each a (rows, D) fp16 matrix -- written
        // by the sampler next to the fp32 descriptors (split0 / split1 = [hi | lo]) or derived here in one pass.
        // The tile pipeline is then TMA -> MMA only (no converter warps): 64-element k-blocks, 128-byte swizzle.
        const __half *hi0, *lo0, *hi1, *lo1;
        if (split0 && split1) {
            hi0 = (const __half*)split0; lo0 = hi0 + e0;
            hi1 = (const __half*)split1; lo1 = hi1 + e1;
        } else {
            __half* base = (__half*)(((uintptr_t)scratch + 1023) & ~(uintptr_t)1023);
            __half *h0 = base, *l0 = h0 + e0, *h1 = l0 + e0, *l1 = h1 + e1;  // e0, e1 multiples of 8: 16-byte aligned
            const size_t quads = (e0 + e1) / 4;
            unsigned blocks = (unsigned)((quads + 255) / 256);
            if (blocks > (unsigned)ctx->num_sms * 16) blocks = (unsigned)ctx->num_sms * 16;
            split_fp16_kernel<<<blocks ? blocks : 1, 256, 0, stream>>>(d0, e0, d1, e1, h0, l0, h1, l1, 1024.0f);
            EINX_CHECK_LAUNCH(ctx);
            hi0 = h0; lo0 = l0; hi1 = h1; lo1 = l1;
        }
        CUtensorMap lom[2];
        memset(lom, 0, sizeof(lom));
        // 64-element k-blocks (128-byte rows) and 8 epilogue warps: measured against 32-element k-blocks with 8 or 16
        // epilogue warps on B200 (C2 shape: 71.7 vs 81.9 / 86.0 us)
        const int kb = 128;
        if ((rc = make_map(ctx, &maps[0], hi0, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (size_t)B * ncap, D, TILE_M, kb))) return rc;
        if ((rc = make_map(ctx, &maps[1], hi1, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (size_t)B * mcap, D, TILE_N / CG, kb))) return rc;
        if ((rc = make_map(ctx, &lom[0], lo0, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (size_t)B * ncap, D, TILE_M, kb))) return rc;
        if ((rc = make_map(ctx, &lom[1], lo1, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (size_t)B * mcap, D, TILE_N / CG, kb))) return rc;
        P.nkb = (D * 2 + kb - 1) / kb;
        P.idesc = make_idesc(2, CG);
        P.in_scale = 1024.0f;
        P.out_scale = 1.0f / (1024.0f * 1024.0f);
        // D <= 128: a tile's MMAs take half as long as at D = 256 and the argmax epilogue is the bound -- 16 epilogue
        // warps over two stages measure 2-4 % faster there (94.2 vs 96.3 us at 32x2048^2x128) and 18 % slower at D = 256
        static const int epi16_env = getenv("EINX_MNN_EPI16") ? atoi(getenv("EINX_MNN_EPI16")) : -1;
        const bool epi16 = epi16_env < 0 ? D <= 128 : epi16_env != 0;
        if (CG == 2 && epi16) return launch_tc<3, 128, 2, 16>(ctx, maps[0], maps[1], P, grid, stream, &lom[0], &lom[1]);
        if (CG == 2) return launch_tc<3, 128, 2>(ctx, maps[0], maps[1], P, grid, stream, &lom[0], &lom[1]);
        return launch_tc<3, 128, 1>(ctx, maps[0], maps[1], P, grid, stream, &lom[0], &lom[1]);
    }
    if (precision == EINX_MNN_FP16X3) {
        // fp32 descriptors in place, 32-element k-blocks (128-byte raw rows); |d| * 2^10 must stay below 65504
        if ((rc = make_map(ctx, &maps[0], d0, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (size_t)B * ncap, D, TILE_M, 128))) return rc;
        if ((rc = make_map(ctx, &maps[1], d1, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (size_t)B * mcap, D, TILE_N / CG, 128))) return rc;
        P.nkb = (D * 4 + 127) / 128;
        P.idesc = make_idesc(2, CG);
        P.in_scale = 1024.0f;
        P.out_scale = 1.0f / (1024.0f * 1024.0f);
        // pair form: 8 epilogue + 8 converter warps (640 threads, 94 registers) and three 64 KB stages -- the column
        // keys live in registers and merge through the idle transpose tiles, which is what lets the third stage fit
        if (CG == 2) return launch_tc<2, 128, 2, 8>(ctx, maps[0], maps[1], P, grid, stream);
        return launch_tc<2, 128, 1>(ctx, maps[0], maps[1], P, grid, stream);
    }
    // TF32X3: 16-element k-blocks (64-byte rows)
    constexpr int KB = 64;
    if ((rc = make_map(ctx, &maps[0], d0, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (size_t)B * ncap, D, TILE_M, KB))) return rc;
    if ((rc = make_map(ctx, &maps[1], d1, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (size_t)B * mcap, D, TILE_N / CG, KB))) return rc;
    P.nkb = (D * 4 + KB - 1) / KB;
    P.idesc = make_idesc(1, CG);
    return CG == 2 ? launch_tc<1, KB, 2>(ctx, maps[0], maps[1], P, grid, stream)
                   : launch_tc<1, KB, 1>(ctx, maps[0], maps[1], P, grid, stream);
}
