// tcgen05 weight-gradient GEMM of the "tf32" mode:  dW[n_out, k_in] += dz[rows, n_out]^T x[rows, k_in]
// (backward of torch.nn.Linear at src/graphnet/models/gnn/dynedge.py:200-247; reduction over the rows, i.e. over
// all edges / nodes of the batch).
//
// Both operands are row-major with the reduction index (row) as the SLOW dimension, i.e. "MN-major" in UMMA
// terms. For 32-bit MN-major operands the only UMMA layout is "128-byte swizzle with 32-byte atomicity"
// (Swizzle<2,5,2>: atom = 4 rows x 128 B, 32-byte chunks XOR-ed with row & 3; descriptor layout type 1), which
// TMA produces with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B. TMA drops [32 rows x 32 columns] boxes straight into
// that layout (LBO = distance between 32-column blocks = 4096 B, SBO = distance between 4-row groups = 512 B),
// so no thread ever touches the operands:
//     A = x tile   : M = 128 k_in columns  x K = 32 rows     (4 boxes,  16 KiB)
//     B = dz tile  : N <= 256 n_out columns x K = 32 rows    (<= 8 boxes, 32 KiB)
//     D[k_in, n_out] in TMEM (lane = k_in, column = n_out)
// Lane = k_in makes the epilogue's reductions into dW[n_out, k_in] coalesced: for a fixed n_out the 32 lanes of a
// warp hit 32 consecutive floats. The row range is split over gridDim.y CTAs (split-K); partial sums are
// combined with fp32 `red.global.add`. 4-stage TMA/mbarrier pipeline, 1 CTA per SM.
#include "common.cuh"
#include "tc_common.cuh"
#include <cuda_fp16.h>

namespace {

constexpr int WG_BM = 128, WG_BN = 256, WG_BK = 32, WG_STAGES = 4, WG_THREADS = 192;
constexpr uint32_t WG_BOX_BYTES = 32 * 32 * 4;                       // one [32 x 32] fp32 box
constexpr uint32_t WG_A_BYTES = (WG_BM / 32) * WG_BOX_BYTES;         // 16 KiB
constexpr uint32_t WG_B_BYTES = (WG_BN / 32) * WG_BOX_BYTES;         // 32 KiB
constexpr uint32_t WG_STAGE_BYTES = WG_A_BYTES + WG_B_BYTES;
constexpr uint32_t WG_SMEM_BYTES = WG_STAGES * WG_STAGE_BYTES + 1024 + 256;

// MN-major 32-bit operand, 128-byte swizzle with 32-byte atomicity (layout type 1): lbo/sbo in bytes
__device__ __forceinline__ uint64_t umma_desc_sw128_mnmajor(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>((lbo >> 4) & 0x3FFFu) << 16;
    d |= static_cast<uint64_t>((sbo >> 4) & 0x3FFFu) << 32;
    d |= static_cast<uint64_t>(1) << 46;          // descriptor version (sm_100)
    d |= static_cast<uint64_t>(1) << 61;          // SWIZZLE_128B_BASE32B
    return d;
}

__global__ void __launch_bounds__(WG_THREADS, 1)
gemm_tc_wgrad_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_z,
                     float* __restrict__ dw, int64_t lddw, int64_t rows, int n_out, int k_in, int n_tiles_n,
                     int64_t rows_per_split, int swap_lbo_sbo) {
    gnb_pdl_begin();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + WG_STAGES * WG_STAGE_BYTES);
    uint64_t* empty = full + WG_STAGES;
    uint64_t* tmem_full = empty + WG_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile_m = blockIdx.x / n_tiles_n, tile_n = blockIdx.x % n_tiles_n;
    const int kin0 = tile_m * WG_BM;
    const int out0 = tile_n * WG_BN;
    int n_cols = n_out - out0;                       // columns of this N tile, rounded up to 16 for the MMA
    if (n_cols > WG_BN) n_cols = WG_BN;
    const int n_mma = (n_cols + 15) & ~15;
    const int n_boxes_b = (n_cols + 31) >> 5;
    const int64_t r_lo = (int64_t)blockIdx.y * rows_per_split;
    int64_t r_hi = r_lo + rows_per_split;
    if (r_hi > rows) r_hi = rows;
    const int num_kb = r_hi > r_lo ? (int)((r_hi - r_lo + WG_BK - 1) / WG_BK) : 0;

    if (warp == 0 && lane == 0) { tc::tma_prefetch_desc(&tm_x); tc::tma_prefetch_desc(&tm_z); }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < WG_STAGES; ++s) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], 1); }
            tc::mbar_init(tmem_full, 1);
            tc::fence_barrier_init();
            tc::fence_proxy_async();
        }
        __syncwarp();
        tc::tmem_alloc<WG_BN>(tmem_slot);
    }
    tc::tcgen05_fence_before();
    __syncthreads();
    tc::tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (num_kb > 0) {
        if (warp == 0) {
            {   // whole warp walks the ring (uniform control flow, see tc::elect_one); one elected lane issues
                for (int it = 0; it < num_kb; ++it) {
                    const int s = it % WG_STAGES;
                    const uint32_t ph = (it / WG_STAGES) & 1;
                    tc::mbar_wait(&empty[s], ph ^ 1);
                    uint8_t* sa = smem + s * WG_STAGE_BYTES;
                    uint8_t* sb = sa + WG_A_BYTES;
                    const int r = (int)(r_lo + (int64_t)it * WG_BK);
                    if (tc::elect_one()) {
                        tc::mbar_arrive_expect_tx(&full[s], (WG_BM / 32 + n_boxes_b) * WG_BOX_BYTES);
#pragma unroll
                        for (int b = 0; b < WG_BM / 32; ++b)
                            tc::tma_load_2d(sa + b * WG_BOX_BYTES, &tm_x, &full[s], kin0 + 32 * b, r);
                        for (int b = 0; b < n_boxes_b; ++b)
                            tc::tma_load_2d(sb + b * WG_BOX_BYTES, &tm_z, &full[s], out0 + 32 * b, r);
                    }
                    __syncwarp();
                }
            }
        } else if (warp == 1) {
            {   // whole warp, uniform control flow; one elected lane issues the MMAs and commits
                // kind::tf32, fp32 accumulate, A and B MN-major (bits 15, 16)
                const uint32_t idesc = tc::umma_idesc_tf32(WG_BM, (uint32_t)n_mma) | (1u << 15) | (1u << 16);
                // bring-up variants (production: 0): bit0 swaps LBO/SBO, bit1 uses an 8-row K group stride
                const uint32_t kgrp = (swap_lbo_sbo & 2) ? 1024u : 512u;
                const uint32_t lbo = (swap_lbo_sbo & 1) ? kgrp : WG_BOX_BYTES;
                const uint32_t sbo = (swap_lbo_sbo & 1) ? WG_BOX_BYTES : kgrp;
                for (int it = 0; it < num_kb; ++it) {
                    const int s = it % WG_STAGES;
                    const uint32_t ph = (it / WG_STAGES) & 1;
                    tc::mbar_wait(&full[s], ph);
                    tc::tcgen05_fence_after();
                    const uint32_t sa = tc::smem_u32(smem + s * WG_STAGE_BYTES);
                    const uint64_t adesc = umma_desc_sw128_mnmajor(sa, lbo, sbo);
                    const uint64_t bdesc = umma_desc_sw128_mnmajor(sa + WG_A_BYTES, lbo, sbo);
                    if (tc::elect_one()) {                // one election per K block: MMAs + commit issued back to back
#pragma unroll
                        for (int k = 0; k < WG_BK / 8; ++k)   // 8 rows = one 1024-byte swizzle atom per 32-column block
                            tc::umma_tf32(tmem_base, adesc + (uint64_t)(k * (1024 >> 4)), bdesc + (uint64_t)(k * (1024 >> 4)),
                                          idesc, (it | k) != 0 ? 1u : 0u);
                        tc::umma_commit(&empty[s]);
                    }
                    __syncwarp();
                }
                if (tc::elect_one()) tc::umma_commit(tmem_full);
                __syncwarp();
            }
        } else {
            tc::mbar_wait<200>(tmem_full, 0);
            tc::tcgen05_fence_after();
            const int q = warp & 3;
            const int kin = kin0 + q * 32 + lane;
            const bool ok = kin < k_in;
            for (int c = 0; c * 32 < n_cols; ++c) {
                uint32_t r[32];
                tc::tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), r);
                tc::tmem_ld_wait();
                if (ok) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int o = out0 + c * 32 + j;
                        if (o < n_out) atomicAdd(dw + (int64_t)o * lddw + kin, __uint_as_float(r[j]));
                    }
                }
            }
        }
    }
    tc::tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc<WG_BN>(tmem_base);
}


// ---- 16-bit plane variant (precision modes bf16 / bf16x3 / mixed16) ------------------------------------------------------------
// Operands are 16-bit planes (v ~ v0 + v1): NPA planes of x (the A operand), NPB planes of dz (B). 16-bit MN-major operands
// use the plain 128-byte swizzle: atom = 8 rows x 128 B (64 elements of the MN index), SBO = 1024 B between 8-row groups of
// the K (row) index, LBO = one TMA box between 64-column blocks; a kind::f16 MMA consumes 16 rows = 2 atoms. A stage holds BK
// rows of every plane. Products: x0 dz0 (1, 1); x1 dz0 + x0 dz0 (2, 1: dz as ONE scaled fp16 plane, mode mixed16);
// x1 dz0 + x0 dz1 + x0 dz0 (2, 2). Element formats (bf16 / fp16, independently for A and B) come with the instruction
// descriptor; out_scale_bits (optional) = device word holding the fp32 bits of max|g| the dz producer scaled by (see
// gnb_absmax_bits): the epilogue multiplies by the inverse power of two.
constexpr uint32_t WB_SMEM_BYTES = 192 * 1024 + 1024 + 256;

__device__ __forceinline__ uint64_t umma_desc_sw128_mnmajor16(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>((lbo >> 4) & 0x3FFFu) << 16;
    d |= static_cast<uint64_t>((sbo >> 4) & 0x3FFFu) << 32;
    d |= static_cast<uint64_t>(1) << 46;          // descriptor version (sm_100)
    d |= static_cast<uint64_t>(2) << 61;          // SWIZZLE_128B
    return d;
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

template <int NPA, int NPB, int BK>
__global__ void __launch_bounds__(WG_THREADS, 1)
gemm_bf_wgrad_kernel(const __grid_constant__ CUtensorMap tm_x0, const __grid_constant__ CUtensorMap tm_x1,
                     const __grid_constant__ CUtensorMap tm_z0, const __grid_constant__ CUtensorMap tm_z1,
                     float* __restrict__ dw, int64_t lddw, int64_t rows, int n_out, int k_in, int n_tiles_n,
                     int64_t rows_per_split, int dbg, uint32_t fmt_a, uint32_t fmt_b, const unsigned* __restrict__ out_scale_bits,
                     const unsigned* __restrict__ out_scale_bits2) {
    gnb_pdl_begin();
    constexpr uint32_t BOX = BK * 128;                           // one [BK rows x 64 columns] 16-bit box
    constexpr uint32_t A_BYTES = 2 * BOX, B_BYTES = 4 * BOX;     // per plane: 128 k_in columns, 256 n_out columns
    constexpr uint32_t STAGE = NPA * A_BYTES + NPB * B_BYTES;
    constexpr int STAGES = (192 * 1024) / STAGE;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE);
    uint64_t* empty = full + STAGES;
    uint64_t* tmem_full = empty + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile_m = blockIdx.x / n_tiles_n, tile_n = blockIdx.x % n_tiles_n;
    const int kin0 = tile_m * WG_BM;
    const int out0 = tile_n * WG_BN;
    int n_cols = n_out - out0;
    if (n_cols > WG_BN) n_cols = WG_BN;
    const int n_mma = (n_cols + 15) & ~15;
    const int n_boxes_b = (n_cols + 63) >> 6;
    const int64_t r_lo = (int64_t)blockIdx.y * rows_per_split;
    int64_t r_hi = r_lo + rows_per_split;
    if (r_hi > rows) r_hi = rows;
    const int num_kb = r_hi > r_lo ? (int)((r_hi - r_lo + BK - 1) / BK) : 0;

    if (warp == 0 && lane == 0) {
        tc::tma_prefetch_desc(&tm_x0); tc::tma_prefetch_desc(&tm_z0);
        if (NPA == 2) tc::tma_prefetch_desc(&tm_x1);
        if (NPB == 2) tc::tma_prefetch_desc(&tm_z1);
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < STAGES; ++s) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], 1); }
            tc::mbar_init(tmem_full, 1);
            tc::fence_barrier_init();
            tc::fence_proxy_async();
        }
        __syncwarp();
        tc::tmem_alloc<WG_BN>(tmem_slot);
    }
    tc::tcgen05_fence_before();
    __syncthreads();
    tc::tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (num_kb > 0) {
        if (warp == 0) {
            for (int it = 0; it < num_kb; ++it) {
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                tc::mbar_wait(&empty[s], ph ^ 1);
                uint8_t* st = smem + s * STAGE;          // [A plane 0 | A plane 1 | B plane 0 | B plane 1]
                const int r = (int)(r_lo + (int64_t)it * BK);
                if (tc::elect_one()) {
                    tc::mbar_arrive_expect_tx(&full[s], (uint32_t)(NPA * 2 + NPB * n_boxes_b) * BOX);
#pragma unroll
                    for (int pl = 0; pl < NPA; ++pl)
#pragma unroll
                        for (int b = 0; b < 2; ++b)
                            tc::tma_load_2d(st + pl * A_BYTES + b * BOX, pl == 0 ? &tm_x0 : &tm_x1, &full[s], kin0 + 64 * b, r);
#pragma unroll
                    for (int pl = 0; pl < NPB; ++pl)
                        for (int b = 0; b < n_boxes_b; ++b)
                            tc::tma_load_2d(st + NPA * A_BYTES + pl * B_BYTES + b * BOX, pl == 0 ? &tm_z0 : &tm_z1, &full[s], out0 + 64 * b, r);
                }
                __syncwarp();
            }
        } else if (warp == 1) {
            // kind::f16, fp32 accumulate, A and B MN-major (bits 15, 16); formats: 0 = fp16, 1 = bf16
            const uint32_t idesc = (1u << 4) | (fmt_a << 7) | (fmt_b << 10) | (((uint32_t)n_mma >> 3) << 17) | ((WG_BM >> 4) << 24) |
                                   (1u << 15) | (1u << 16);
            // bring-up variants (production: 0): bit0 swaps LBO / SBO
            const uint32_t lbo = (dbg & 1) ? 1024u : BOX;
            const uint32_t sbo = (dbg & 1) ? BOX : 1024u;
            for (int it = 0; it < num_kb; ++it) {
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                tc::mbar_wait(&full[s], ph);
                tc::tcgen05_fence_after();
                const uint32_t sa = tc::smem_u32(smem + s * STAGE);
                const uint64_t a0 = umma_desc_sw128_mnmajor16(sa, lbo, sbo), a1 = umma_desc_sw128_mnmajor16(sa + A_BYTES, lbo, sbo);
                const uint64_t b0 = umma_desc_sw128_mnmajor16(sa + NPA * A_BYTES, lbo, sbo);
                const uint64_t b1 = umma_desc_sw128_mnmajor16(sa + NPA * A_BYTES + B_BYTES, lbo, sbo);
                if (tc::elect_one()) {
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {          // 16 rows = two 1024-byte atoms per 64-column block
                        const uint64_t adv = (uint64_t)(k * (2048 >> 4));
                        if (NPA == 2) umma_f16(tmem_base, a1 + adv, b0 + adv, idesc, (it | k) != 0 ? 1u : 0u);
                        if (NPB == 2) umma_f16(tmem_base, a0 + adv, b1 + adv, idesc, 1u);
                        umma_f16(tmem_base, a0 + adv, b0 + adv, idesc, (NPA == 2 || (it | k) != 0) ? 1u : 0u);
                    }
                    tc::umma_commit(&empty[s]);
                }
                __syncwarp();
            }
            if (tc::elect_one()) tc::umma_commit(tmem_full);
            __syncwarp();
        } else {
            float inv = 1.f;
            if (out_scale_bits != nullptr) inv = gnb_pow2_scale(*out_scale_bits).y;
            if (out_scale_bits2 != nullptr) inv *= gnb_pow2_scale(*out_scale_bits2).y;
            tc::mbar_wait<200>(tmem_full, 0);
            tc::tcgen05_fence_after();
            const int q = warp & 3;
            const int kin = kin0 + q * 32 + lane;
            const bool ok = kin < k_in;
            for (int c = 0; c * 32 < n_cols; ++c) {
                uint32_t r[32];
                tc::tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), r);
                tc::tmem_ld_wait();
                if (ok) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int o = out0 + c * 32 + j;
                        if (o < n_out) atomicAdd(dw + (int64_t)o * lddw + kin, __uint_as_float(r[j]) * inv);
                    }
                }
            }
        }
    }
    tc::tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc<WG_BN>(tmem_base);
}


// ---- weight gradient of the aggregating Linear WITHOUT a stored dz (fp16-plane modes) ------------------------------------------
// dz[(i, s), c] = g[i, c] * bit(i, s, c) is a 9-fold redundant function of the node-level gradient and the ReLU bits of the
// aggregating epilogue. Inputs as prepared by gnb_edge_dz_prep: g16 [n, n_out] = fp16(g * 2^s) and the row-major bits
// rowmask[(i * 9 + s) * (n_out / 32) + c / 32]. Per 64-row stage the TMA warp also stages the <= 9 g16 rows the stage's nodes
// need (one bulk copy); eight builder warps (8 rows each; lane = 8 consecutive channels) read the row's node values with one
// 16-byte shared load, expand the row's byte of channel bits into two half-word masks per register and write the MN-major
// 128-byte-swizzled B tile (row rr at rr * 128 B of its 64-channel box, chunk j at j ^ (rr & 7)) with one 16-byte store: the
// kernel reads h (one fp16 plane) + 0.06 GB instead of h + dz, i.e. 0.54 instead of 0.85 GB per 713 k-row launch.
// Warps: 0 TMA (A = h, g16 rows), 1 MMA, 2..5 epilogue, 6..13 builders. full[s] collects the TMA transaction + the 8 builder warps.
// Two rings: the TMA ring {A = h box pair 16 KiB | the stage's g16 rows <= 4.5 KiB} is WGB_NA = 7 stages deep -- with nothing
// but 20 KiB per stage arriving from memory, the 4-stage ring of the stored-dz kernel left too few bytes in flight (155 us
// floor measured with builders and MMAs switched off) --, the locally built B tile (32 KiB) is double-buffered.
// A CTA owns up to TWO 128-wide k_in tiles (two 256-column accumulators = all of TMEM): the dz tile is the expensive operand
// here (built, not loaded), and it is the same for every k_in tile -- with one tile per CTA a 336-wide layer expanded dz three
// times, with pairs twice (625 -> ... us per step). CTA x = k_in tiles {2 x, 2 x + 1}.
constexpr int WGB_THREADS = 448, WGB_BK = 64, WGB_NA = 4, WGB_NBUF = 2, WGB_NB = 8;
constexpr uint32_t WGB_BOX = WGB_BK * 128, WGB_A_BYTES = 2 * WGB_BOX, WGB_B_BYTES = 4 * WGB_BOX;
constexpr uint32_t WGB_ASTAGE = 2 * WGB_A_BYTES + 5 * 1024;             // {A tile 0 | A tile 1 | g16 rows (<= 9 x 512 B)}; 1 KiB-aligned
constexpr uint32_t WGB_DATA_BYTES = WGB_NA * WGB_ASTAGE + WGB_NBUF * WGB_B_BYTES;
constexpr uint32_t WGB_SMEM_BYTES = WGB_DATA_BYTES + 1024 + 256;

// W = 9 slots per node (rows = 9 n), or W = 8 for graphs without k + 1-neighbour nodes (rows = 8 n: a builder warp's 8 rows are one node).
template <int W>
__device__ __forceinline__ void
wgrad_build_body(const CUtensorMap* tm_xp, const __half* __restrict__ g16,
                 const unsigned* __restrict__ rowmask, float* __restrict__ dw, int64_t lddw, int64_t rows, int n_out,
                 int k_in, int64_t rows_per_split, const unsigned* __restrict__ dz_scale_bits,
                 const unsigned* __restrict__ x_scale_bits, int dbg) {
    const CUtensorMap& tm_x = *tm_xp;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* bring = smem + WGB_NA * WGB_ASTAGE;                       // [WGB_NBUF] x 32 KiB
    uint64_t* afull = reinterpret_cast<uint64_t*>(smem + WGB_DATA_BYTES);
    uint64_t* aempty = afull + WGB_NA;
    uint64_t* bfull = aempty + WGB_NA;                // the 8 builder warps wrote the B tile
    uint64_t* bempty = bfull + WGB_NBUF;              // the MMAs that read it completed
    uint64_t* tmem_full = bempty + WGB_NBUF;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kin0 = blockIdx.x * 2 * WG_BM;
    const int nt = kin0 + WG_BM < k_in ? 2 : 1;                         // k_in tiles of this CTA
    const int n_mma = (n_out + 15) & ~15;
    const int64_t n_nodes = rows / W;
    const int64_t r_lo = (int64_t)blockIdx.y * rows_per_split;
    int64_t r_hi = r_lo + rows_per_split;
    if (r_hi > rows) r_hi = rows;
    const int num_kb = r_hi > r_lo ? (int)((r_hi - r_lo + WGB_BK - 1) / WGB_BK) : 0;

    if (warp == 0 && lane == 0) tc::tma_prefetch_desc(&tm_x);
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < WGB_NA; ++s) { tc::mbar_init(&afull[s], 1); tc::mbar_init(&aempty[s], 1); }
            for (int s = 0; s < WGB_NBUF; ++s) { tc::mbar_init(&bfull[s], WGB_NB); tc::mbar_init(&bempty[s], 1); }
            tc::mbar_init(tmem_full, 1);
            tc::fence_barrier_init();
            tc::fence_proxy_async();
        }
        __syncwarp();
        tc::tmem_alloc<512>(tmem_slot);
    }
    tc::tcgen05_fence_before();
    __syncthreads();
    tc::tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (num_kb > 0) {
        if (warp == 0) {
            for (int it = 0; it < num_kb; ++it) {
                const int s = it % WGB_NA;
                const uint32_t ph = (it / WGB_NA) & 1;
                tc::mbar_wait<0, true>(&aempty[s], ph ^ 1);
                uint8_t* st = smem + s * WGB_ASTAGE;
                const int64_t r = r_lo + (int64_t)it * WGB_BK;
                const int64_t nd0 = r / W;
                int64_t nd1 = (r + WGB_BK - 1) / W + 1;                 // one past the last node of the stage
                nd1 = nd1 < n_nodes ? nd1 : n_nodes;
                const uint32_t gb = nd1 > nd0 ? (uint32_t)(nd1 - nd0) * (uint32_t)n_out * 2u : 0u;
                if (tc::elect_one()) {
                    tc::mbar_arrive_expect_tx(&afull[s], (uint32_t)nt * WGB_A_BYTES + gb);
                    if (gb) tc::bulk_load(st + 2 * WGB_A_BYTES, g16 + nd0 * n_out, gb, &afull[s]);
                    for (int tt = 0; tt < nt; ++tt) {
                        tc::tma_load_2d(st + tt * WGB_A_BYTES, &tm_x, &afull[s], kin0 + tt * WG_BM, (int)r);
                        tc::tma_load_2d(st + tt * WGB_A_BYTES + WGB_BOX, &tm_x, &afull[s], kin0 + tt * WG_BM + 64, (int)r);
                    }
                }
                __syncwarp();
            }
        } else if (warp == 1) {
            const uint32_t idesc = (1u << 4) | (((uint32_t)n_mma >> 3) << 17) | ((WG_BM >> 4) << 24) | (1u << 15) | (1u << 16);   // fp16 x fp16
            for (int it = 0; it < num_kb; ++it) {
                const int s = it % WGB_NA, sb = it % WGB_NBUF;
                tc::mbar_wait<0, true>(&afull[s], (it / WGB_NA) & 1);
                tc::mbar_wait<0, true>(&bfull[sb], (it / WGB_NBUF) & 1);
                tc::tcgen05_fence_after();
                const uint64_t a0 = umma_desc_sw128_mnmajor16(tc::smem_u32(smem + s * WGB_ASTAGE), WGB_BOX, 1024u);
                const uint64_t a1 = umma_desc_sw128_mnmajor16(tc::smem_u32(smem + s * WGB_ASTAGE + WGB_A_BYTES), WGB_BOX, 1024u);
                const uint64_t b0 = umma_desc_sw128_mnmajor16(tc::smem_u32(bring + sb * WGB_B_BYTES), WGB_BOX, 1024u);
                if (tc::elect_one()) {
#pragma unroll
                    for (int k = 0; k < WGB_BK / 16; ++k) {
                        const uint64_t adv = (uint64_t)(k * (2048 >> 4));
                        if (!(dbg & 2)) umma_f16(tmem_base, a0 + adv, b0 + adv, idesc, (it | k) != 0 ? 1u : 0u);
                    }
                    if (nt == 2) {
#pragma unroll
                        for (int k = 0; k < WGB_BK / 16; ++k) {
                            const uint64_t adv = (uint64_t)(k * (2048 >> 4));
                            if (!(dbg & 2)) umma_f16(tmem_base + 256u, a1 + adv, b0 + adv, idesc, (it | k) != 0 ? 1u : 0u);
                        }
                    }
                    tc::umma_commit(&aempty[s]);
                    tc::umma_commit(&bempty[sb]);
                }
                __syncwarp();
            }
            if (tc::elect_one()) tc::umma_commit(tmem_full);
            __syncwarp();
        } else if (warp >= 6) {
            // ---- dz builders: lane = 8 consecutive channels, warp bw = rows [8 bw, 8 bw + 8) of the 64-row stage ----------------
            const int bw = warp - 6;
            const int cw = n_out >> 5;
            const bool ch_on = 8 * lane < n_out;
            const unsigned bsh = 8u * ((unsigned)lane & 3u);             // this lane's 8 channels = byte (lane % 4) of word lane / 4
            const uint32_t boxoff = (uint32_t)(lane >> 3) * WGB_BOX, jj = (uint32_t)(lane & 7);
            unsigned mrow[8];                                            // the 8 rows' mask words, loaded one stage ahead
            const uint32_t r_lo32 = (uint32_t)r_lo, rows32 = (uint32_t)rows;      // rows < 2^31 (checked by the launcher)
            const unsigned* mp0 = rowmask + ((int64_t)r_lo32 + 8 * bw) * cw + (lane >> 2);     // row 8 bw of stage 0
            auto prefetch = [&](int it) {
                const uint32_t R0 = r_lo32 + (uint32_t)it * WGB_BK + 8u * (uint32_t)bw;
                const unsigned* mp = mp0 + (int64_t)it * (WGB_BK * cw);
                if (ch_on && R0 + 8u <= rows32) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) mrow[i] = __ldg(mp + i * cw);
                } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i) mrow[i] = (ch_on && R0 + i < rows32) ? __ldg(mp + i * cw) : 0u;
                }
            };
            prefetch(0);
            for (int it = 0; it < num_kb; ++it) {
                const int s = it % WGB_NA, sbuf = it % WGB_NBUF;
                // the lane's 8 channel bits of every row as two nibbles (0 beyond `rows`)
                unsigned nlo[8], nhi[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) { nlo[i] = (mrow[i] >> bsh) & 15u; nhi[i] = (mrow[i] >> (bsh + 4u)) & 15u; }
                if (it + 1 < num_kb) prefetch(it + 1);
                const uint32_t Rs = r_lo32 + (uint32_t)it * WGB_BK;      // first row of the stage
                const uint32_t R0 = Rs + 8u * (uint32_t)bw;
                const uint32_t q0 = W == 8 ? R0 >> 3 : __umulhi(R0, 0x38E38E39u) >> 1;      // R0 / W
                const uint32_t rem0 = R0 - (uint32_t)W * q0;
                const uint32_t nrel0 = q0 - (W == 8 ? Rs >> 3 : __umulhi(Rs, 0x38E38E39u) >> 1);
                tc::mbar_wait<0, true>(&afull[s], (it / WGB_NA) & 1);             // the stage's g16 rows (and A) landed
                tc::mbar_wait<0, true>(&bempty[sbuf], ((it / WGB_NBUF) & 1) ^ 1); // the MMAs of the B slot's previous use completed
                const uint32_t sg = tc::smem_u32(smem + s * WGB_ASTAGE + 2 * WGB_A_BYTES) + (uint32_t)lane * 16u + nrel0 * (uint32_t)n_out * 2u;
                const uint32_t sb = tc::smem_u32(bring + sbuf * WGB_B_BYTES) + boxoff + (uint32_t)(8 * bw) * 128u;
                // 8 rows touch at most 2 nodes: their fp16 values are read ONCE each (shared-memory bandwidth, not issue slots, is
                // what a stage is short of: B written 32 KiB + read by the MMAs 48 KiB + the TMA's 20 KiB per ~900 cycles)
                uint32_t ga[4], gb2[4];
                asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(ga[0]), "=r"(ga[1]), "=r"(ga[2]), "=r"(ga[3]) : "r"(sg));
                asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(gb2[0]), "=r"(gb2[1]), "=r"(gb2[2]), "=r"(gb2[3])
                             : "r"(sg + (uint32_t)n_out * 2u));
                if (!(dbg & 1))
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const bool second = rem0 + (uint32_t)i >= (uint32_t)W;        // warp-uniform
                    // 4 channel bits -> the sign bits of 4 bytes (one multiply: bit m lands on bit 8 m + 7, no two partial
                    // products share a position), then one byte permute with sign replication per pair of channels gives the
                    // two half-word masks of an fp16x2 register
                    const uint32_t ylo = nlo[i] * 0x10204080u, yhi = nhi[i] * 0x10204080u;
                    uint32_t o[4];
                    o[0] = (second ? gb2[0] : ga[0]) & tc::prmt(ylo, 0u, 0x9988u);
                    o[1] = (second ? gb2[1] : ga[1]) & tc::prmt(ylo, 0u, 0xBBAAu);
                    o[2] = (second ? gb2[2] : ga[2]) & tc::prmt(yhi, 0u, 0x9988u);
                    o[3] = (second ? gb2[3] : ga[3]) & tc::prmt(yhi, 0u, 0xBBAAu);
                    // row rr = 8 bw + i: chunk position jj ^ (rr & 7) = jj ^ i
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sb + (uint32_t)i * 128u + ((jj ^ (uint32_t)i) << 4)),
                                 "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]) : "memory");
                }
                tc::fence_proxy_async();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&bfull[sbuf]);
                __syncwarp();
            }
        } else {
            float inv = gnb_pow2_scale(*dz_scale_bits).y;
            if (x_scale_bits != nullptr) inv *= gnb_pow2_scale(*x_scale_bits).y;
            tc::mbar_wait<200, true>(tmem_full, 0);
            tc::tcgen05_fence_after();
            const int q = warp & 3;
            for (int tt = 0; tt < nt; ++tt) {
                const int kin = kin0 + tt * WG_BM + q * 32 + lane;
                const bool ok = kin < k_in;
                for (int c = 0; c * 32 < n_out; ++c) {
                    uint32_t r[32];
                    tc::tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(tt * 256 + c * 32), r);
                    tc::tmem_ld_wait();
                    if (ok) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const int o = c * 32 + j;
                            if (o < n_out) atomicAdd(dw + (int64_t)o * lddw + kin, __uint_as_float(r[j]) * inv);
                        }
                    }
                }
            }
        }
    }
    tc::tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc<512>(tmem_base);
}

__global__ void __launch_bounds__(WGB_THREADS, 1)
gemm_f16_wgrad_build_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_x8,
                            const __half* __restrict__ g16, const unsigned* __restrict__ rowmask, float* __restrict__ dw,
                            int64_t lddw, int64_t n_nodes, int n_out, int k_in, int64_t rows_per_split,
                            const unsigned* __restrict__ dz_scale_bits, const unsigned* __restrict__ x_scale_bits, int dbg,
                            const int* __restrict__ full9) {
    gnb_pdl_begin();
    if (full9 == nullptr || *full9 != 0)
        wgrad_build_body<9>(&tm_x, g16, rowmask, dw, lddw, n_nodes * 9, n_out, k_in, rows_per_split, dz_scale_bits, x_scale_bits, dbg);
    else {
        const int64_t rows8 = n_nodes * 8;
        const int64_t rps8 = ((rows8 + gridDim.y - 1) / gridDim.y + WGB_BK - 1) / WGB_BK * WGB_BK;
        wgrad_build_body<8>(&tm_x8, g16, rowmask, dw, lddw, rows8, n_out, k_in, rps8, dz_scale_bits, x_scale_bits, dbg);
    }
}

// [rows, cols] fp32 row-major, box = 32 columns x 32 rows, 128-byte swizzle with 32-byte atoms, OOB -> 0
static int make_tmap_box32(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int64_t ld) {
    return gnb_make_tmap_f32(map, base, rows, cols, ld, 32u, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
}

}  // namespace

// dw[n_out, k_in] += dz[rows, n_out]^T x[rows, k_in]; operands pre-rounded to tf32, 16-byte aligned, pitches % 4 == 0.
// debug_swap != 0 exchanges the LBO/SBO descriptor fields (bring-up aid; production passes 0).
GNB_EXPORT int gnb_linear_bwd_weight_tf32(const float* dz, int64_t lddz, const float* x, int64_t ldx, float* dw,
                                          int64_t lddw, int64_t rows, int32_t n_out, int32_t k_in, int32_t debug_swap,
                                          void* stream) {
    if (rows < 0 || n_out < 1 || k_in < 1) return GNB_ERR_ARG;
    if (rows == 0) return GNB_OK;
    if (rows >= (int64_t)1 << 31) return GNB_ERR_ARG;
    CUtensorMap tx, tz;
    int rc = make_tmap_box32(&tx, x, rows, k_in, ldx);
    if (rc != 0) return rc == -2 ? GNB_ERR_UNSUPPORTED : GNB_ERR_ARG;
    rc = make_tmap_box32(&tz, dz, rows, n_out, lddz);
    if (rc != 0) return rc == -2 ? GNB_ERR_UNSUPPORTED : GNB_ERR_ARG;
    static unsigned long long attr_devs = 0ull;     // the attribute is per device: one bit per device ordinal
    int dev = 0;
    GNB_CHECK(cudaGetDevice(&dev));
    if (dev >= 64 || !((attr_devs >> dev) & 1ull)) {
        GNB_CHECK(cudaFuncSetAttribute(gemm_tc_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)WG_SMEM_BYTES));
        if (dev < 64) attr_devs |= 1ull << dev;
    }
    const int tiles_m = gnb_div_up(k_in, WG_BM), tiles_n = gnb_div_up(n_out, WG_BN);
    const int tiles = tiles_m * tiles_n;
    static const int wg_waves = getenv("GNB_WG_WAVES") ? atoi(getenv("GNB_WG_WAVES")) : 1;      // one CTA per SM: 567 -> 482 us per step against two waves
    int splits = (wg_waves * 148) / tiles;                                // one wave of 148 CTAs (every split-K slice ends in 32 k fp32 reductions into dw)
    const int64_t max_splits = (rows + 8 * WG_BK - 1) / (8 * WG_BK);     // at least 8 K-blocks per CTA
    if (splits > max_splits) splits = (int)max_splits;
    if (splits < 1) splits = 1;
    int64_t rps = (rows + splits - 1) / splits;
    rps = ((rps + WG_BK - 1) / WG_BK) * WG_BK;
    splits = (int)((rows + rps - 1) / rps);
    dim3 grid((unsigned)tiles, (unsigned)splits);
    gnb_launch(gemm_tc_wgrad_kernel, grid, WG_THREADS, WG_SMEM_BYTES, (cudaStream_t)stream)(tx, tz, dw, lddw, rows, n_out, k_in,
                                                                                    tiles_n, rps, debug_swap);
    GNB_RETURN_LAUNCH();
}

// dw[n_out, k_in] += dz^T x on 16-bit planes: dz planes [rows, n_out] (pitch lddz elements), x planes [rows, k_in] (pitch ldx);
// x1 / dz1 may be NULL (dz1 non-NULL requires x1). fmt bit 0: x is bf16 (else fp16), bit 1: dz is bf16 (else fp16).
// out_scale_bits (may be NULL): device word with the fp32 bits of the max|g| the dz producer scaled by -- the result is
// multiplied by the inverse power of two (gnb_absmax_bits). Pitches multiples of 8 elements. debug bit 0 swaps LBO / SBO.
static int wgrad16_impl(const void* dz0, const void* dz1, int64_t lddz, const void* x0, const void* x1, int64_t ldx, float* dw,
                        int64_t lddw, int64_t rows, int32_t n_out, int32_t k_in, int32_t debug, int32_t fmt,
                        const uint32_t* out_scale_bits, const uint32_t* out_scale_bits2, void* stream) {
    if (rows < 0 || n_out < 1 || k_in < 1 || dz0 == nullptr || x0 == nullptr || (dz1 != nullptr && x1 == nullptr)) return GNB_ERR_ARG;
    if ((lddz & 7) || (ldx & 7) || lddz < n_out || ldx < k_in) return GNB_ERR_ARG;
    if (rows == 0) return GNB_OK;
    if (rows >= (int64_t)1 << 31) return GNB_ERR_ARG;
    const int npa = x1 != nullptr ? 2 : 1, npb = dz1 != nullptr ? 2 : 1;
    const int bk = npa == 1 ? 64 : 32;
    const CUtensorMapDataType ta = (fmt & 1) ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    const CUtensorMapDataType tb = (fmt & 2) ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    CUtensorMap tx0, tx1, tz0, tz1;
    int rc = gnb_make_tmap_16(&tx0, x0, rows, k_in, ldx * 2, (uint32_t)bk, ta);
    if (rc == 0) rc = gnb_make_tmap_16(&tz0, dz0, rows, n_out, lddz * 2, (uint32_t)bk, tb);
    if (rc == 0 && npa == 2) rc = gnb_make_tmap_16(&tx1, x1, rows, k_in, ldx * 2, (uint32_t)bk, ta);
    if (rc == 0 && npb == 2) rc = gnb_make_tmap_16(&tz1, dz1, rows, n_out, lddz * 2, (uint32_t)bk, tb);
    if (rc != 0) return rc == -2 ? GNB_ERR_UNSUPPORTED : GNB_ERR_ARG;
    static unsigned long long attr_devs = 0ull;
    int dev = 0;
    GNB_CHECK(cudaGetDevice(&dev));
    if (dev >= 64 || !((attr_devs >> dev) & 1ull)) {
        GNB_CHECK(cudaFuncSetAttribute(gemm_bf_wgrad_kernel<1, 1, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WB_SMEM_BYTES));
        GNB_CHECK(cudaFuncSetAttribute(gemm_bf_wgrad_kernel<2, 1, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WB_SMEM_BYTES));
        GNB_CHECK(cudaFuncSetAttribute(gemm_bf_wgrad_kernel<2, 2, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WB_SMEM_BYTES));
        if (dev < 64) attr_devs |= 1ull << dev;
    }
    const int tiles_m = gnb_div_up(k_in, WG_BM), tiles_n = gnb_div_up(n_out, WG_BN);
    const int tiles = tiles_m * tiles_n;
    int splits = (2 * 148) / tiles;
    const int64_t max_splits = (rows + 8 * bk - 1) / (8 * bk);
    if (splits > max_splits) splits = (int)max_splits;
    if (splits < 1) splits = 1;
    int64_t rps = (rows + splits - 1) / splits;
    rps = ((rps + bk - 1) / bk) * bk;
    splits = (int)((rows + rps - 1) / rps);
    dim3 grid((unsigned)tiles, (unsigned)splits);
    const uint32_t fa = (fmt & 1) ? 1u : 0u, fb = (fmt & 2) ? 1u : 0u;
    cudaStream_t st = (cudaStream_t)stream;
    if (npa == 2 && npb == 2)
        gnb_launch(gemm_bf_wgrad_kernel<2, 2, 32>, grid, WG_THREADS, WB_SMEM_BYTES, st)(tx0, tx1, tz0, tz1, dw, lddw, rows, n_out, k_in, tiles_n,
                                                                              rps, debug, fa, fb, out_scale_bits, out_scale_bits2);
    else if (npa == 2)
        gnb_launch(gemm_bf_wgrad_kernel<2, 1, 32>, grid, WG_THREADS, WB_SMEM_BYTES, st)(tx0, tx1, tz0, tz0, dw, lddw, rows, n_out, k_in, tiles_n,
                                                                              rps, debug, fa, fb, out_scale_bits, out_scale_bits2);
    else
        gnb_launch(gemm_bf_wgrad_kernel<1, 1, 64>, grid, WG_THREADS, WB_SMEM_BYTES, st)(tx0, tx0, tz0, tz0, dw, lddw, rows, n_out, k_in, tiles_n,
                                                                              rps, debug, fa, fb, out_scale_bits, out_scale_bits2);
    GNB_RETURN_LAUNCH();
}
GNB_EXPORT int gnb_linear_bwd_weight_bf16(const void* dz0, const void* dz1, int64_t lddz, const void* x0, const void* x1,
                                          int64_t ldx, float* dw, int64_t lddw, int64_t rows, int32_t n_out, int32_t k_in,
                                          int32_t debug, void* stream) {
    if ((dz1 == nullptr) != (x1 == nullptr)) return GNB_ERR_ARG;
    return wgrad16_impl(dz0, dz1, lddz, x0, x1, ldx, dw, lddw, rows, n_out, k_in, debug, 3, nullptr, nullptr, stream);
}
// mixed16: dz as ONE fp16 plane scaled by gnb_pow2_scale(*dz_scale_bits).x, x as fp16 plane(s) of x * gnb_pow2_scale(
// *x_scale_bits).x (x1 may be NULL: fp16 = tf32's significand); the epilogue undoes both powers of two. A kind::f16 MMA takes
// ONE element format for both operands (bf16 against fp16 traps as an illegal instruction on sm_100a).
GNB_EXPORT int gnb_linear_bwd_weight_f16(const void* dz, int64_t lddz, const void* x0, const void* x1, int64_t ldx, float* dw,
                                         int64_t lddw, int64_t rows, int32_t n_out, int32_t k_in, const uint32_t* dz_scale_bits,
                                         const uint32_t* x_scale_bits, void* stream) {
    return wgrad16_impl(dz, nullptr, lddz, x0, x1, ldx, dw, lddw, rows, n_out, k_in, 0, 0, dz_scale_bits, x_scale_bits, stream);
}

static int g_wgrad_dbg = 0;     // profiling hook (results are garbage with any bit set): bit 0 builders store nothing, bit 1 no MMAs
GNB_EXPORT int gnb_wgrad_set_debug(int32_t flags) { g_wgrad_dbg = flags; return GNB_OK; }

// The same weight gradient with dz EXPANDED IN THE KERNEL from the outputs of gnb_edge_dz_prep: g16 [n, n_out] = fp16(g * 2^s)
// (contiguous rows) and the row-major ReLU bits rowmask[(i * 9 + s) * (n_out / 32) + c / 32] (bit c % 32):
//   dw[n_out, k_in] += sum_{i, s} (g16[i, :] * bit(i, s, :))^T x[(i, s), :] * 2^-s * 2^-sx        rows = 9 n (k = 8 tables)
// x = ONE fp16 plane of h * 2^sx [9 n, k_in]. n_out <= 256, n_out % 32 == 0.
// full9 (device, may be NULL = 9 slots): *full9 == 0 selects the 8-slot layout (rows i * 8 + s of x and rowmask, 8 n rows).
GNB_EXPORT int gnb_linear_bwd_weight_f16_masked_w(const void* g16, const uint32_t* rowmask, const void* x, int64_t ldx, float* dw,
                                                  int64_t lddw, int64_t n, int32_t n_out, int32_t k_in,
                                                  const uint32_t* dz_scale_bits, const uint32_t* x_scale_bits, const int32_t* full9,
                                                  void* stream) {
    if (n < 0 || n_out < 32 || n_out > 256 || (n_out & 31) || k_in < 1 || g16 == nullptr || rowmask == nullptr || x == nullptr ||
        dz_scale_bits == nullptr)
        return GNB_ERR_ARG;
    if ((ldx & 7) || ldx < k_in || (reinterpret_cast<uintptr_t>(g16) & 15u)) return GNB_ERR_ARG;
    if (n == 0) return GNB_OK;
    const int64_t rows = n * 9;
    if (rows >= (int64_t)1 << 31) return GNB_ERR_ARG;
    CUtensorMap tx;
    int rc = gnb_make_tmap_16(&tx, x, rows, k_in, ldx * 2, (uint32_t)WGB_BK, CU_TENSOR_MAP_DATA_TYPE_FLOAT16);
    if (rc != 0) return rc == -2 ? GNB_ERR_UNSUPPORTED : GNB_ERR_ARG;
    CUtensorMap tx8 = tx;
    if (full9 != nullptr) {
        rc = gnb_make_tmap_16(&tx8, x, n * 8, k_in, ldx * 2, (uint32_t)WGB_BK, CU_TENSOR_MAP_DATA_TYPE_FLOAT16);
        if (rc != 0) return rc == -2 ? GNB_ERR_UNSUPPORTED : GNB_ERR_ARG;
    }
    static unsigned long long attr_devs = 0ull;
    int dev = 0;
    GNB_CHECK(cudaGetDevice(&dev));
    if (dev >= 64 || !((attr_devs >> dev) & 1ull)) {
        GNB_CHECK(cudaFuncSetAttribute(gemm_f16_wgrad_build_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WGB_SMEM_BYTES));
        if (dev < 64) attr_devs |= 1ull << dev;
    }
    const int tiles = gnb_div_up(gnb_div_up(k_in, WG_BM), 2);           // CTAs along k_in: pairs of 128-wide tiles
    static const int waves = getenv("GNB_WGB_WAVES") ? atoi(getenv("GNB_WGB_WAVES")) : 1;      // one CTA per SM: 522 -> 458 us per step against two waves (epilogue atomics, prologues)
    int splits = (waves * 148) / tiles;
    const int64_t max_splits = (rows + 8 * WGB_BK - 1) / (8 * WGB_BK);
    if (splits > max_splits) splits = (int)max_splits;
    if (splits < 1) splits = 1;
    int64_t rps = (rows + splits - 1) / splits;
    rps = ((rps + WGB_BK - 1) / WGB_BK) * WGB_BK;
    splits = (int)((rows + rps - 1) / rps);
    gnb_launch(gemm_f16_wgrad_build_kernel, dim3((unsigned)tiles, (unsigned)splits), WGB_THREADS, WGB_SMEM_BYTES, (cudaStream_t)stream)(
        tx, tx8, (const __half*)g16, rowmask, dw, lddw, n, n_out, k_in, rps, dz_scale_bits, x_scale_bits, g_wgrad_dbg, full9);
    GNB_RETURN_LAUNCH();
}
GNB_EXPORT int gnb_linear_bwd_weight_f16_masked(const void* g16, const uint32_t* rowmask, const void* x, int64_t ldx, float* dw,
                                                int64_t lddw, int64_t n, int32_t n_out, int32_t k_in,
                                                const uint32_t* dz_scale_bits, const uint32_t* x_scale_bits, void* stream) {
    return gnb_linear_bwd_weight_f16_masked_w(g16, rowmask, x, ldx, dw, lddw, n, n_out, k_in, dz_scale_bits, x_scale_bits, nullptr, stream);
}
