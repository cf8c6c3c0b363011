// Blackwell (sm_100a) primitives used by the tensor-core kernels: mbarrier, TMA (cp.async.bulk.tensor),
// tcgen05 (alloc / mma / commit / ld) and the UMMA shared-memory / instruction descriptors.
// Hand-written inline PTX; bit layouts follow the PTX ISA "tcgen05 matrix/instruction descriptor" tables.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// 32-bit shared-memory loads / fire-and-forget global reductions for the metadata-driven epilogues. A `const T*` into
// dynamic shared memory that crossed a function boundary compiles to generic LD.E with 64-bit address arithmetic
// (4 integer instructions per load); these take the 32-bit shared address, so constant offsets fold into the LDS.
// byte permute; a selector nibble with bit 3 set replicates the sign bit of the selected byte over the result byte
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void red_add_f32(float* p, float v) {
    asm volatile("red.relaxed.gpu.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

// ---- mbarrier ------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// try_wait, optionally with a suspend-time hint: with the hint the thread may stay suspended in hardware (no issue slots used)
// until the phase completes or the hint (ns) expires, instead of returning to the software polling loop after the default,
// short limit -- ncu (profiles/r02/r_bwd_builders.ncu-rep) showed 36 .. 47 % of the builder kernels' executed instructions in
// these loops. Measured on the training step: only the weight-gradient builder kernel gains (701 -> 624 us per step); the split
// forward GEMM, whose splitter warps wait for TMA data on the critical path, loses (817 -> 855 us: the wake-up is slower than
// a poll); the other kernels do not move. Opt-in per wait (mbar_wait<SLEEP_NS, true>).
#ifndef GNB_MBAR_HINT_NS
#define GNB_MBAR_HINT_NS 20000
#endif
template <bool HINT = false>
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t addr, uint32_t parity) {
    uint32_t done;
    if (HINT)
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity), "r"((uint32_t)GNB_MBAR_HINT_NS)
            : "memory");
    else
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
    return done;
}
// Bounded wait: a lost arrival traps (reported as a launch error) instead of hanging the GPU.
// SLEEP_NS > 0 backs off between polls so that long waits (epilogue / metadata warps) do not steal issue slots
// from the warps doing the work.
template <int SLEEP_NS = 0, bool HINT = false>
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    if (mbar_try_wait<HINT>(addr, parity)) return;
    long long t0 = 0;
    for (uint32_t it = 1;; ++it) {
        if (SLEEP_NS > 0) __nanosleep(SLEEP_NS);
        if (mbar_try_wait<HINT>(addr, parity)) return;
        if ((it & 4095u) == 0u) {
            const long long now = clock64();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 4000000000LL) {           // ~2 s at 2 GHz
                printf("gnb mbar_wait timeout: block %d thread %d barrier smem 0x%x parity %u\n", (int)blockIdx.x,
                       (int)threadIdx.x, addr, parity);
                __trap();
            }
        }
    }
}

// One lane of a CONVERGED warp. Single-thread roles (TMA producer, MMA issuer) run their loops with the whole warp
// and elect only around the issuing instructions: inside an `if (lane == 0)` region the compiler cannot prove that
// descriptors / addresses are warp-uniform and wraps every tcgen05.mma in an ELECT + R2UR.BROADCAST + BRA.U.ANY
// loop (~110 cycles per MMA measured, against a 64-cycle tensor-pipe floor for 128x128x8 tf32).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

// Wait executed by a CONVERGED warp with a warp-uniform exit (vote): unlike the per-thread spin of mbar_wait, the
// compiler can prove that control flow stays convergent, so loop counters / descriptors computed afterwards live in
// uniform registers and the tcgen05.mma / TMA operands need no R2UR broadcasts.
// Same test with acquire semantics at CLUSTER scope: for barriers whose arrival comes from the peer CTA and orders that
// CTA's shared-memory writes (the operand splitter of the tf32x3 GEMM) before this warp's tcgen05.mma.
__device__ __forceinline__ uint32_t mbar_try_wait_cluster(uint32_t addr, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    return done;
}
template <bool CLUSTER_ACQUIRE = false>
__device__ __forceinline__ void mbar_wait_warp(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    long long t0 = 0;
    for (uint32_t it = 1;; ++it) {
        if (__all_sync(0xffffffffu, CLUSTER_ACQUIRE ? mbar_try_wait_cluster(addr, parity) : mbar_try_wait(addr, parity))) return;
        if ((it & 4095u) == 0u) {
            const long long now = clock64();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 4000000000LL) {
                if ((threadIdx.x & 31) == 0)
                    printf("gnb mbar_wait_warp timeout: block %d warp %d barrier smem 0x%x parity %u\n", (int)blockIdx.x,
                           (int)(threadIdx.x >> 5), addr, parity);
                __trap();
            }
        }
    }
}

// ---- TMA -----------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// 2-D tiled load: coordinates {c0 = innermost (element), c1 = row}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}

// 1-D bulk copy global -> shared (size % 16 == 0, both addresses 16-byte aligned), completion on an mbarrier
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---- tcgen05 -------------------------------------------------------------------------------
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <uint32_t NCOLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst) {   // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "n"(NCOLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <uint32_t NCOLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {     // same warp that allocated
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}

// K-major operand tile in shared memory, 128-byte swizzle, rows of 128 B, 8-row groups 1024 B apart.
__device__ __forceinline__ uint64_t umma_desc_sw128_kmajor(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);        // start address        bits [0,14)
    d |= static_cast<uint64_t>(1) << 16;                            // leading byte offset  bits [16,30) (unused)
    d |= static_cast<uint64_t>(1024 >> 4) << 32;                    // stride byte offset   bits [32,46)
    d |= static_cast<uint64_t>(1) << 46;                            // descriptor version   bits [46,48)
    d |= static_cast<uint64_t>(2) << 61;                            // SWIZZLE_128B         bits [61,64)
    return d;
}

// kind::tf32, fp32 accumulate, both operands K-major.
__host__ __device__ constexpr uint32_t umma_idesc_tf32(uint32_t m, uint32_t n) {
    return (1u << 4)            // D format  = F32
           | (2u << 7)          // A format  = TF32
           | (2u << 10)         // B format  = TF32
           | ((n >> 3) << 17)   // N >> 3
           | ((m >> 4) << 24);  // M >> 4
}

// D[tmem] (+)= A[smem] * B[smem]^T ; issued by ONE thread
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// arrive on an mbarrier when all previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

// each thread of the warp reads its TMEM lane, 32 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
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
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- CTA pair (cta_group::2) ---------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {      // every thread of every CTA of the cluster
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the mbarrier at the same shared-memory offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t cta) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}"
        ::"r"(smem_u32(bar)), "r"(cta)
        : "memory");
}
// Relaxed variant for pure notifications: the data hand-off is already ordered by fence.proxy.async + a named
// barrier inside the producing CTA, so the (expensive: MEMBAR + ERRBAR, ~1300 cycles measured) cluster-scope
// release of the plain arrive is not needed on the signalling warp.
__device__ __forceinline__ void mbar_arrive_cluster_relaxed(uint64_t* bar, uint32_t cta) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [ra];\n\t}"
        ::"r"(smem_u32(bar)), "r"(cta)
        : "memory");
}
template <uint32_t NCOLS>
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t* smem_dst) {   // the same warp in BOTH CTAs
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "n"(NCOLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <uint32_t NCOLS>
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}
// D[tmem of both CTAs] (+)= A * B^T with M = 256 (128 rows from each CTA's smem), B: N/2 rows from each CTA
__device__ __forceinline__ void umma_tf32_2cta(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                               uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// kind::f16 with bf16 operands (UMMA_K = 16 = 32 bytes per instruction), fp32 accumulate, both operands K-major: the
// correction products of the split-operand GEMM. a / b format code 1 = BF16 (0 = F16, 2 = TF32 in the same field).
__host__ __device__ constexpr uint32_t umma_idesc_bf16(uint32_t m, uint32_t n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((m >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16_2cta(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                               uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive (once all prior MMAs of this thread completed) on the barrier at this offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_2cta(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask)
                 : "memory");
}

// 2-D tiled load issued by either CTA of a pair; the transaction bytes are credited to the barrier at the same offset
// in the EVEN (leader) CTA: the CTA rank of a shared::cluster address is bit 24, cleared here.
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1)
        : "memory");
}

__device__ __forceinline__ float round_tf32(float v) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return __uint_as_float(r);
}

}  // namespace tc

// ---- host: tensor-map encoding through the driver entry point (no link-time libcuda dependency) ----
typedef CUresult (*gnb_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                        CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline gnb_encode_tiled_fn gnb_get_encode_tiled() {
    static gnb_encode_tiled_fn fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<gnb_encode_tiled_fn>(p);
    }
    return fn;
}

// fp32 row-major matrix [rows, cols] with row pitch ld (elements): box = [32 cols (=128 B), box_rows],
// 128-byte swizzle, out-of-bounds elements read as zero.
static inline int gnb_make_tmap_f32(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int64_t ld,
                                    uint32_t box_rows, CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
    gnb_encode_tiled_fn enc = gnb_get_encode_tiled();
    if (enc == nullptr) return -2;
    if ((reinterpret_cast<uintptr_t>(base) & 15u) || ((ld * 4) & 15)) return -1;
    cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
    cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld) * 4};
    cuuint32_t box[2] = {32u, box_rows};
    cuuint32_t estr[2] = {1u, 1u};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : -1;
}
// 16-bit (bf16 / fp16) row-major matrix [rows, cols] with row pitch ld_bytes: box = [64 cols (= 128 B), box_rows], 128-byte swizzle.
static inline int gnb_make_tmap_16(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int64_t ld_bytes,
                                   uint32_t box_rows, CUtensorMapDataType dtype) {
    gnb_encode_tiled_fn enc = gnb_get_encode_tiled();
    if (enc == nullptr) return -2;
    if ((reinterpret_cast<uintptr_t>(base) & 15u) || (ld_bytes & 15)) return -1;
    cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
    cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld_bytes)};
    cuuint32_t box[2] = {64u, box_rows};
    cuuint32_t estr[2] = {1u, 1u};
    CUresult r = enc(map, dtype, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : -1;
}
static inline int gnb_make_tmap_bf16(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int64_t ld_bytes,
                                     uint32_t box_rows) {
    return gnb_make_tmap_16(map, base, rows, cols, ld_bytes, box_rows, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
}
