// tcgen05 / TMEM / TMA Linear layer for the "tf32" precision mode of the DynEdge path.
//
//   y[rows, n_out] = act( sum_p x_p[rows, k_p] W[:, koff_p : koff_p + k_p]^T + bias )
//
// replaces torch.nn.Linear (+ReLU) at src/graphnet/models/gnn/dynedge.py:200-203, 226-229, 246-247 and, with
// p > 1, the skip-concatenation + first post-processing Linear of dynedge.py:328-331 (K-split over the
// per-layer outputs, [N,1043] never materialised). The same kernel computes dx = dz W in the backward pass
// (caller passes W^T as the weight operand).
//
// Orientation ("weights are A"): D[channel, row] = W_tile[128 ch x K] * X_tile[128 rows x K]^T, so that in
// TMEM lane = output channel and column = row. An epilogue warp then writes, per row, 32 consecutive
// channels = one coalesced 128-byte store straight from registers; no shared-memory staging of the output.
//
// Persistent CTAs (one per SM) loop over 128-row tiles; each tile computes up to TWO 128-channel tiles from
// the same activation tile (the big operand is read once), K in blocks of 32 tf32 (one 128-byte swizzle
// atom). 4-stage TMA -> mbarrier -> tcgen05.mma pipeline; the fp32 accumulators (2 x 128 columns) are
// double-buffered in TMEM (512 columns) so the epilogue of tile t overlaps the main loop of tile t+1.
// Warp roles: 0 = TMA producer, 1 = MMA issuer (+TMEM alloc), 2..9 = epilogue (lane quarter x column half).
// Operands are fp32 in HBM; weights and activations are expected pre-rounded to tf32 (cvt.rna) by their
// producers so that the tensor core's truncation is exact (unbiased rounding overall).
#include "common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int TC_BM = 128, TC_BN = 128, TC_BK = 32, TC_MT = 2, TC_STAGES = 4, TC_THREADS = 320;
constexpr int TC_MAX_PARTS = 6;
constexpr uint32_t TC_TILE_BYTES = TC_BM * TC_BK * 4;                           // 16 KiB: one 128 x 32 fp32 tile
constexpr uint32_t TC_STAGE_BYTES = (TC_MT + 1) * TC_TILE_BYTES;                // 2 weight tiles + 1 activation tile
constexpr int SC_MAX_MASK_LD = 16;                                              // mask words per row (hdim <= 512)
constexpr uint32_t SC_META_MASK_BYTES = 126 * SC_MAX_MASK_LD * 4;               // 8064: one tile of activation mask rows
constexpr uint32_t SC_META_BYTES = 8192 + 512;                                  // mask rows + 128 scatter offsets
constexpr uint32_t TC_SMEM_BYTES = TC_STAGES * TC_STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/ + 2 * SC_META_BYTES;
constexpr uint32_t TC_TMEM_COLS = 2 * TC_MT * TC_BN;                            // 512: two accumulator buffers

struct TmapArray { CUtensorMap m[TC_MAX_PARTS]; };
struct PartInfo { int nparts; int kblocks[TC_MAX_PARTS]; };
// Aggregating epilogue (EdgeConv second Linear): rows are padded edge slots (node i, slot s) = i * 9 + s; a tile is
// 14 nodes = 126 rows; the epilogue sums relu(acc + b) over the valid slots of every node, writes y[node, ch] and one
// bit per (slot, channel) = "pre-activation > 0" for the backward pass. The [E, C] message tensor is never stored.
struct AggInfo { const int* deg; int64_t n_nodes; unsigned* maskbits; int enabled; int dbg; unsigned long long* prof; };
constexpr int AGG_W = 9, AGG_NPT = 14, AGG_ROWS = AGG_W * AGG_NPT;   // k = 8 neighbour tables
// Scattering epilogue (backward of the hoisted EdgeConv hidden layer fused into the data-gradient GEMM): rows are padded
// edge slots as above, output channel c of row (i, s) is dh = (dz W2)[(i,s), c]; the epilogue applies the ReLU mask
// (bit mask of h > 0 written by the forward hidden-layer kernel), sums the slots of a node into dPQ[i, c] (P half) and
// adds each slot into dPQ[nbr[i,s], hdim + c] (Q half, fp32 `red.global.add`, 32 consecutive channels per warp
// instruction). dh [E, hdim] is never stored. Per tile the producer warp stages the 126 mask rows (one bulk copy) and
// the 126 scatter offsets nbr * ldpq in shared memory, so the epilogue touches no global metadata.
struct ScatInfo { const int* nbr; const unsigned* hmask; int mask_ld; float* dpq; int64_t ldpq; int hdim; int64_t n_nodes; int enabled; };

// Scatter epilogue for NN consecutive nodes whose 9 slot columns sit in r[J0 ...]; col0 = first column in the tile.
template <int NN, int J0>
__device__ __forceinline__ void scat_nodes(const uint32_t (&r)[32], int col0, const int* __restrict__ s_off,
                                           const unsigned* __restrict__ s_msk, int mask_ld, int lane, float* __restrict__ dq,
                                           float* __restrict__ dp, int64_t ldpq, int64_t nodes_left, bool ch_ok, int off_kill) {
    // metadata first (independent broadcast LDS, issued back to back), then a branch-free body: the reduction is a
    // predicated `red` (a C++ `if` around atomicAdd compiles to a divergence region per element that serialises the
    // shared-memory loads behind it: 90 cycles per element measured)
    int offr[NN * AGG_W];
    unsigned mwr[NN * AGG_W];
#pragma unroll
    for (int e = 0; e < NN * AGG_W; ++e) {
        offr[e] = s_off[col0 + e] | off_kill;            // negative = no reduction (padding slot / channel out of range)
        mwr[e] = s_msk[(col0 + e) * mask_ld];
    }
#pragma unroll
    for (int f = 0; f < NN; ++f) {
        float accp = 0.f;
#pragma unroll
        for (int sl = 0; sl < AGG_W; ++sl) {
            const int e = f * AGG_W + sl;
            const float v = ((mwr[e] >> lane) & 1u) ? __uint_as_float(r[J0 + e]) : 0.f;
            accp += v;
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ge.s32 p, %2, 0;\n\t@p red.global.add.f32 [%0], %1;\n\t}"
                         ::"l"(dq + offr[e]), "f"(v), "r"(offr[e]) : "memory");
        }
        if (ch_ok && f < nodes_left) dp[(int64_t)f * ldpq] = accp;
    }
}

// Epilogue store of one 32-row chunk: lane = output channel, r[j] = row j. One coalesced 128-byte store per row; the
// address is a running pointer and activation / rounding are resolved outside the unrolled loop (the naive per-element
// form compiled to ~30 instructions per store and made the epilogue warps the bottleneck of the kernel).
template <bool ROUND>
__device__ __forceinline__ void epi_store32(const uint32_t (&r)[32], float bv, float lo, float* __restrict__ yp, int64_t ldy) {
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        float v = fmaxf(__uint_as_float(r[j]) + bv, lo);
        if (ROUND) v = tc::round_tf32(v);
        *yp = v;
        yp += ldy;
    }
}
__device__ __noinline__ void epi_store_partial(const uint32_t (&r)[32], float bv, float lo, bool round, float* __restrict__ yp,
                                               int64_t ldy, int nvalid) {
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        if (j < nvalid) {
            float v = fmaxf(__uint_as_float(r[j]) + bv, lo);
            if (round) v = tc::round_tf32(v);
            *yp = v;
            yp += ldy;
        }
    }
}

__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_linear_kernel(const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ TmapArray tm_x,
                      const PartInfo parts, const float* __restrict__ bias, float* __restrict__ y, int64_t ldy,
                      int64_t rows, int n_out, int act, int round_out, int num_row_tiles, const AggInfo agg,
                      const ScatInfo sc) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + TC_STAGES * TC_STAGE_BYTES);
    uint64_t* empty = full + TC_STAGES;
    uint64_t* tmem_full = empty + TC_STAGES;      // [2]
    uint64_t* tmem_empty = tmem_full + 2;         // [2]
    uint64_t* meta_full = tmem_empty + 2;         // [2] scatter epilogue metadata of the tile in TMEM buffer b
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(meta_full + 2);
    uint8_t* meta = smem + TC_STAGES * TC_STAGE_BYTES + 256;      // [2] x {mask rows (8192 B) | offsets (512 B)}

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ch0 = blockIdx.y * (TC_MT * TC_BM);
    int mt = (n_out - ch0 + TC_BM - 1) / TC_BM;   // valid channel tiles of this CTA (1 or 2)
    if (mt > TC_MT) mt = TC_MT;

    if (warp == 0 && lane == 0) {
        tc::tma_prefetch_desc(&tm_w);
        for (int p = 0; p < parts.nparts; ++p) tc::tma_prefetch_desc(&tm_x.m[p]);
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < TC_STAGES; ++s) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], 1); }
            for (int b = 0; b < 2; ++b) {
                tc::mbar_init(&tmem_full[b], 1); tc::mbar_init(&tmem_empty[b], 8); tc::mbar_init(&meta_full[b], 1);
            }
            tc::fence_barrier_init();
            tc::fence_proxy_async();
        }
        __syncwarp();
        tc::tmem_alloc<TC_TMEM_COLS>(tmem_slot);
    }
    tc::tcgen05_fence_before();
    __syncthreads();
    tc::tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    int total_kb = 0;
    for (int p = 0; p < parts.nparts; ++p) total_kb += parts.kblocks[p];
    const int tile_rows = (agg.enabled || sc.enabled) ? AGG_ROWS : TC_BN;   // rows of the activation tile (TMA box rows)
    const bool prof_on = agg.prof != nullptr && blockIdx.x == 0 && blockIdx.y == 0;
    long long pw0 = 0, pw1 = 0;
    const long long pt0 = clock64();

    if (warp == 0) {
        {   // whole warp walks the pipeline (uniform control flow); one elected lane issues
            uint32_t it = 0, tile_i = 0;
            for (int t = blockIdx.x; t < num_row_tiles; t += gridDim.x, ++tile_i) {
                const int row0 = t * tile_rows;
                int offv[4] = {-1, -1, -1, -1};
                if (sc.enabled) {        // scatter offsets of the tile's 126 slots (consumed after the K loop: latency hidden)
#pragma unroll
                    for (int q4 = 0; q4 < 4; ++q4) {
                        const int col = lane + 32 * q4;
                        const int nb = (col < AGG_ROWS && (int64_t)row0 + col < rows) ? sc.nbr[(int64_t)row0 + col] : -1;
                        offv[q4] = nb >= 0 ? nb * (int)sc.ldpq : -1;
                    }
                }
                int kb_w = 0;
                for (int p = 0; p < parts.nparts; ++p) {
                    for (int kb = 0; kb < parts.kblocks[p]; ++kb, ++kb_w, ++it) {
                        const uint32_t s = it % TC_STAGES, ph = (it / TC_STAGES) & 1;
                        const long long c0 = prof_on ? clock64() : 0;
                        tc::mbar_wait(&empty[s], ph ^ 1);
                        if (prof_on) pw0 += clock64() - c0;
                        uint8_t* st = smem + s * TC_STAGE_BYTES;
                        const bool ld_w = !(agg.dbg & 4), ld_x = !(agg.dbg & 8);
                        if (tc::elect_one()) {
                            tc::mbar_arrive_expect_tx(&full[s], (ld_w ? (uint32_t)mt * TC_TILE_BYTES : 0u) +
                                                                    (ld_x ? (uint32_t)tile_rows * TC_BK * 4 : 0u));
                            if (ld_w)
                                for (int m = 0; m < mt; ++m)
                                    tc::tma_load_2d(st + m * TC_TILE_BYTES, &tm_w, &full[s], kb_w * TC_BK, ch0 + m * TC_BM);
                            if (ld_x) tc::tma_load_2d(st + TC_MT * TC_TILE_BYTES, &tm_x.m[p], &full[s], kb * TC_BK, row0);
                        }
                        __syncwarp();
                    }
                }
                if (sc.enabled) {
                    const uint32_t buf = tile_i & 1;
                    tc::mbar_wait(&tmem_empty[buf], ((tile_i >> 1) & 1) ^ 1);   // epilogue of tile_i - 2 is done with meta[buf]
                    uint8_t* mb = meta + buf * SC_META_BYTES;
                    int* so = reinterpret_cast<int*>(mb + 8192);
#pragma unroll
                    for (int q4 = 0; q4 < 4; ++q4) so[lane + 32 * q4] = offv[q4];
                    __syncwarp();
                    if (tc::elect_one()) {
                        const uint32_t bytes = (uint32_t)(AGG_ROWS * sc.mask_ld * 4);
                        tc::mbar_arrive_expect_tx(&meta_full[buf], bytes);
                        tc::bulk_load(mb, sc.hmask + (int64_t)t * AGG_ROWS * sc.mask_ld, bytes, &meta_full[buf]);
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp == 1) {
        {   // whole warp, uniform control flow; tcgen05.mma / commit issued by one elected lane
            constexpr uint32_t idesc = tc::umma_idesc_tf32(TC_BM, TC_BN);
            uint32_t it = 0, tile_i = 0;
            for (int t = blockIdx.x; t < num_row_tiles; t += gridDim.x, ++tile_i) {
                const uint32_t buf = tile_i & 1;
                const long long c1 = prof_on ? clock64() : 0;
                tc::mbar_wait_warp(&tmem_empty[buf], ((tile_i >> 1) & 1) ^ 1);      // epilogue drained this buffer
                if (prof_on) pw1 += clock64() - c1;
                tc::tcgen05_fence_after();
                const uint32_t acc = __shfl_sync(0xffffffffu, tmem_base, 0) + buf * (TC_MT * TC_BN);
                for (int kbi = 0; kbi < total_kb; ++kbi, ++it) {
                    const uint32_t s = it % TC_STAGES, ph = (it / TC_STAGES) & 1;
                    const long long c0 = prof_on ? clock64() : 0;
                    tc::mbar_wait_warp(&full[s], ph);
                    if (prof_on) pw0 += clock64() - c0;
                    tc::tcgen05_fence_after();
                    const uint32_t st = tc::smem_u32(smem + s * TC_STAGE_BYTES);
                    const uint64_t bdesc = tc::umma_desc_sw128_kmajor(st + TC_MT * TC_TILE_BYTES);
                    for (int m = 0; m < ((agg.dbg & 2) ? 0 : mt); ++m) {
                        const uint64_t adesc = tc::umma_desc_sw128_kmajor(st + m * TC_TILE_BYTES);
#pragma unroll
                        for (int k = 0; k < TC_BK / 8; ++k)   // UMMA_K = 8 tf32 = 32 B -> +2 in the 16-byte address field
                            if (tc::elect_one())
                                tc::umma_tf32(acc + m * TC_BN, adesc + 2 * k, bdesc + 2 * k, idesc, (kbi | k) != 0 ? 1u : 0u);
                    }
                    if (tc::elect_one()) tc::umma_commit(&empty[s]);
                    __syncwarp();
                }
                if (tc::elect_one()) tc::umma_commit(&tmem_full[buf]);
                __syncwarp();
            }
        }
    } else {
        const int ew = warp - 2;                    // 0..7
        const int q = warp & 3;                     // TMEM lane quarter this warp may access
        const int half = ew >> 2;                   // column half: rows [64*half, 64*half + 64) of the tile
        const float relu_lo = act == GNB_ACT_RELU ? 0.f : -INFINITY;
        uint32_t tile_i = 0;
        for (int t = blockIdx.x; t < num_row_tiles; t += gridDim.x, ++tile_i) {
            const uint32_t buf = tile_i & 1;
            const int64_t row0 = (int64_t)t * TC_BN;
            const long long c0 = prof_on ? clock64() : 0;
            tc::mbar_wait<100>(&tmem_full[buf], (tile_i >> 1) & 1);
            if (prof_on) pw0 += clock64() - c0;
            tc::tcgen05_fence_after();
            if (sc.enabled) {
                tc::mbar_wait<100>(&meta_full[buf], (tile_i >> 1) & 1);
                const int m = half;
                if (m < mt) {
                    const int ch = ch0 + m * TC_BM + q * 32 + lane;
                    const bool ch_ok = ch < n_out;
                    const int64_t node0 = (int64_t)t * AGG_NPT;
                    const uint8_t* mb = meta + buf * SC_META_BYTES;
                    const int* s_off = reinterpret_cast<const int*>(mb + 8192);
                    const unsigned* s_msk = reinterpret_cast<const unsigned*>(mb) + ((ch0 + m * TC_BM + q * 32) >> 5);
                    float* dq = sc.dpq + sc.hdim + ch;
                    float* dp = sc.dpq + node0 * sc.ldpq + ch;
                    const uint32_t tcol = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * (TC_MT * TC_BN) + m * TC_BN);
                    const int no_at = (!ch_ok || (agg.dbg & 16)) ? (int)0x80000000 : 0;
#pragma unroll 1
                    for (int c = 0; c < 4; ++c) {          // columns [27 c, 27 c + 27): three nodes
                        uint32_t r[32];
                        tc::tmem_ld_32x32b_x32(tcol + (uint32_t)(27 * c), r);
                        tc::tmem_ld_wait();
                        scat_nodes<3, 0>(r, 27 * c, s_off, s_msk, sc.mask_ld, lane, dq, dp + (int64_t)(3 * c) * sc.ldpq, sc.ldpq,
                                         sc.n_nodes - node0 - 3 * c, ch_ok, no_at);
                    }
                    {                                       // columns [108, 126): the last two nodes, registers 12..29
                        uint32_t r[32];
                        tc::tmem_ld_32x32b_x32(tcol + 96u, r);
                        tc::tmem_ld_wait();
                        scat_nodes<2, 12>(r, 108, s_off, s_msk, sc.mask_ld, lane, dq, dp + (int64_t)12 * sc.ldpq, sc.ldpq,
                                          sc.n_nodes - node0 - 12, ch_ok, no_at);
                    }
                }
            } else if (agg.enabled) {
                // warps 2-5 own channel tile 0, warps 6-9 channel tile 1; every thread walks all 126 slot columns
                const int m = half;
                if (m < mt) {
                    const int ch = ch0 + m * TC_BM + q * 32 + lane;
                    const bool ch_ok = ch < n_out;
                    const float bv = (bias != nullptr && ch_ok) ? bias[ch] : 0.f;
                    const int64_t node0 = (int64_t)t * AGG_NPT;
                    int dg[AGG_NPT];
#pragma unroll
                    for (int f = 0; f < AGG_NPT; ++f) dg[f] = (node0 + f < agg.n_nodes) ? agg.deg[node0 + f] : 0;
                    float acc = 0.f;
                    unsigned bits[4];
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        uint32_t r[32];
                        tc::tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) +
                                                   (uint32_t)(buf * (TC_MT * TC_BN) + m * TC_BN + c * 32), r);
                        tc::tmem_ld_wait();
                        unsigned w = 0u;
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const int col = c * 32 + j;
                            if (col < AGG_ROWS) {
                                const int f = col / AGG_W, sl = col % AGG_W;       // compile-time after unrolling
                                const float pre = __uint_as_float(r[j]) + bv;
                                const bool on = (sl < dg[f]) && (pre > 0.f);
                                acc += on ? pre : 0.f;
                                w |= on ? (1u << j) : 0u;
                                if (sl == AGG_W - 1) {
                                    float o = acc;
                                    if (round_out) o = tc::round_tf32(o);
                                    if (ch_ok && node0 + f < agg.n_nodes) y[(node0 + f) * ldy + ch] = o;
                                    acc = 0.f;
                                }
                            }
                        }
                        bits[c] = w;
                    }
                    if (ch_ok && agg.maskbits != nullptr)
                        reinterpret_cast<uint4*>(agg.maskbits)[(int64_t)t * n_out + ch] = make_uint4(bits[0], bits[1], bits[2], bits[3]);
                }
            } else
            for (int m = 0; m < mt; ++m) {
                const int ch = ch0 + m * TC_BM + q * 32 + lane;
                const bool ch_ok = ch < n_out;
                const float bv = (bias != nullptr && ch_ok) ? bias[ch] : 0.f;
#pragma unroll 1
                for (int c = 0; c < 2; ++c) {
                    const int col0 = half * 64 + c * 32;
                    uint32_t r[32];
                    const long long c1 = prof_on ? clock64() : 0;
                    tc::tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) +
                                               (uint32_t)(buf * (TC_MT * TC_BN) + m * TC_BN + col0), r);
                    tc::tmem_ld_wait();
                    if (prof_on) pw1 += clock64() - c1;
                    const int64_t left = rows - (row0 + col0);          // valid rows of this 32-row chunk
                    if (ch_ok && left > 0 && !(agg.dbg & 1)) {
                        float* yp = y + (row0 + col0) * ldy + ch;
                        if (left >= 32) {
                            if (round_out) epi_store32<true>(r, bv, relu_lo, yp, ldy);
                            else epi_store32<false>(r, bv, relu_lo, yp, ldy);
                        } else {
                            epi_store_partial(r, bv, relu_lo, round_out != 0, yp, ldy, (int)left);
                        }
                    }
                }
            }
            tc::tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&tmem_empty[buf]);
        }
    }
    __syncwarp();
    if (prof_on && lane == 0 && warp <= 2) {
        agg.prof[warp * 3] = (unsigned long long)pw0;
        agg.prof[warp * 3 + 1] = (unsigned long long)pw1;
        agg.prof[warp * 3 + 2] = (unsigned long long)(clock64() - pt0);
    }
    tc::tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc<TC_TMEM_COLS>(tmem_base);
}

// dst[r, c] = rna_tf32(src[r, c]) for c < cols, 0 for cols <= c < dst_cols
__global__ void round_pad_tf32_kernel(const float* __restrict__ src, int64_t lds, int64_t rows, int cols,
                                      float* __restrict__ dst, int64_t ldd, int dst_cols) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= rows * dst_cols) return;
    const int64_t r = t / dst_cols;
    const int c = (int)(t - r * dst_cols);
    dst[r * ldd + c] = c < cols ? tc::round_tf32(src[r * lds + c]) : 0.f;
}

int g_num_sms = 0;
unsigned long long* g_linear_prof = nullptr;
int g_linear_dbg = 0;   // tuning hook: see gnb_linear_set_debug

}  // namespace

// Tuning hook (profiling only): bit0 epilogue skips its global stores, bit1 no MMAs, bit2 no weight loads, bit3 no
// activation loads. Results are garbage with any bit set.
GNB_EXPORT int gnb_linear_set_debug(int32_t flags) { g_linear_dbg = flags; return GNB_OK; }

// Tuning aid: device buffer of 16 uint64 that CTA (0,0) fills with cycle counters
// {producer: wait-empty, total, -} {mma: wait-full, wait-tmem-empty, total} {epilogue warp 2: wait-tmem-full, tmem-ld, total}.
GNB_EXPORT int gnb_linear_set_profile_buffer(void* buf) { g_linear_prof = (unsigned long long*)buf; return GNB_OK; }

// xs / ldxs / ks: HOST arrays with one entry per part (device pointer, row pitch, width).
// w: [n_out, sum_p ceil(k_p/32)*32] fp32, part p's columns start at the 32-aligned running offset and are
// zero padded; weights and activations pre-rounded to tf32.
GNB_EXPORT int gnb_linear_fwd_tf32(const float* const* xs, const int64_t* ldxs, const int32_t* ks, int32_t nparts,
                                   const float* w, int64_t ldw, const float* bias, float* y, int64_t ldy, int64_t rows,
                                   int32_t n_out, int32_t act, int32_t round_out, void* stream) {
    if (nparts < 1 || nparts > TC_MAX_PARTS || rows < 0 || n_out < 1) return GNB_ERR_ARG;
    if (rows == 0) return GNB_OK;
    if (rows >= (int64_t)1 << 31) return GNB_ERR_ARG;
    PartInfo pi;
    TmapArray tx;
    pi.nparts = nparts;
    int64_t ktot = 0;
    for (int p = 0; p < nparts; ++p) {
        if (ks[p] < 1) return GNB_ERR_ARG;
        pi.kblocks[p] = (ks[p] + TC_BK - 1) / TC_BK;
        ktot += (int64_t)pi.kblocks[p] * TC_BK;
        int rc = gnb_make_tmap_f32(&tx.m[p], xs[p], rows, ks[p], ldxs[p], TC_BN);
        if (rc != 0) return rc == -2 ? GNB_ERR_UNSUPPORTED : GNB_ERR_ARG;
    }
    for (int p = nparts; p < TC_MAX_PARTS; ++p) { pi.kblocks[p] = 0; tx.m[p] = tx.m[0]; }
    if (ldw < ktot) return GNB_ERR_ARG;
    CUtensorMap tw;
    int rc = gnb_make_tmap_f32(&tw, w, n_out, ktot, ldw, TC_BM);
    if (rc != 0) return rc == -2 ? GNB_ERR_UNSUPPORTED : GNB_ERR_ARG;
    if (g_num_sms == 0) {
        int dev = 0;
        GNB_CHECK(cudaGetDevice(&dev));
        GNB_CHECK(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
        GNB_CHECK(cudaFuncSetAttribute(gemm_tc_linear_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)TC_SMEM_BYTES));
    }
    const int row_tiles = gnb_div_up(rows, TC_BN);
    const int groups = gnb_div_up(n_out, TC_MT * TC_BM);
    int ctas_x = g_num_sms / groups;                 // persistent: about one CTA per SM in total
    if (ctas_x < 1) ctas_x = 1;
    if (ctas_x > row_tiles) ctas_x = row_tiles;
    dim3 grid((unsigned)ctas_x, (unsigned)groups);
    AggInfo agg{nullptr, 0, nullptr, 0, g_linear_dbg, g_linear_prof};
    ScatInfo sc{nullptr, nullptr, 0, nullptr, 0, 0, 0, 0};
    gemm_tc_linear_kernel<<<grid, TC_THREADS, TC_SMEM_BYTES, (cudaStream_t)stream>>>(tw, tx, pi, bias, y, ldy, rows,
                                                                                     n_out, act, round_out, row_tiles, agg, sc);
    GNB_RETURN_LAUNCH();
}

// Second Linear of the EdgeConv MLP fused with ReLU and the k-neighbour SUM (k = 8 tables, width 9):
//   y[i, :] = sum_{s < deg[i]} relu(h[i*9 + s, :] w^T + bias),   maskbits[tile = i / 14][ch][4 x u32] = (pre-activation > 0)
// h: [n*9, k] tf32-rounded padded edge list, w: [n_out, ceil(k/32)*32] packed. n_out <= 512. maskbits may be NULL.
GNB_EXPORT int gnb_edge_linear_agg_fwd_tf32(const float* h, int64_t ldh, int32_t k, const float* w, int64_t ldw,
                                            const float* bias, const int32_t* deg, int64_t n, int32_t n_out,
                                            int32_t round_out, float* y, int64_t ldy, uint32_t* maskbits, void* stream) {
    if (n < 0 || n_out < 1 || k < 1) return GNB_ERR_ARG;
    if (n == 0) return GNB_OK;
    const int64_t rows = n * AGG_W;
    if (rows >= (int64_t)1 << 31) return GNB_ERR_ARG;
    PartInfo pi;
    TmapArray tx;
    pi.nparts = 1;
    pi.kblocks[0] = (k + TC_BK - 1) / TC_BK;
    const int64_t ktot = (int64_t)pi.kblocks[0] * TC_BK;
    int rc = gnb_make_tmap_f32(&tx.m[0], h, rows, k, ldh, AGG_ROWS);
    if (rc != 0) return rc == -2 ? GNB_ERR_UNSUPPORTED : GNB_ERR_ARG;
    for (int p = 1; p < TC_MAX_PARTS; ++p) { pi.kblocks[p] = 0; tx.m[p] = tx.m[0]; }
    if (ldw < ktot) return GNB_ERR_ARG;
    CUtensorMap tw;
    rc = gnb_make_tmap_f32(&tw, w, n_out, ktot, ldw, TC_BM);
    if (rc != 0) return rc == -2 ? GNB_ERR_UNSUPPORTED : GNB_ERR_ARG;
    if (g_num_sms == 0) {
        int dev = 0;
        GNB_CHECK(cudaGetDevice(&dev));
        GNB_CHECK(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
        GNB_CHECK(cudaFuncSetAttribute(gemm_tc_linear_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)TC_SMEM_BYTES));
    }
    const int row_tiles = gnb_div_up(n, AGG_NPT);
    const int groups = gnb_div_up(n_out, TC_MT * TC_BM);
    int ctas_x = g_num_sms / groups;
    if (ctas_x < 1) ctas_x = 1;
    if (ctas_x > row_tiles) ctas_x = row_tiles;
    dim3 grid((unsigned)ctas_x, (unsigned)groups);
    AggInfo agg{deg, n, maskbits, 1, g_linear_dbg, g_linear_prof};
    ScatInfo sc{nullptr, nullptr, 0, nullptr, 0, 0, 0, 0};
    gemm_tc_linear_kernel<<<grid, TC_THREADS, TC_SMEM_BYTES, (cudaStream_t)stream>>>(tw, tx, pi, bias, y, ldy, rows, n_out,
                                                                                     GNB_ACT_RELU, round_out, row_tiles, agg, sc);
    GNB_RETURN_LAUNCH();
}

// Data gradient of the EdgeConv second Linear fused with the backward of the hoisted hidden layer (k = 8 tables, width 9):
//   dh[(i,s), :] = dz[(i,s), :] wt^T          (wt = W2^T: [hdim, ceil(c_out/32)*32] tf32-rounded, zero padded)
//   da = dh * (h > 0);   dpq[i, 0:hdim] = sum_s da[(i,s)];   dpq[nbr[i,s], hdim:2 hdim] += da[(i,s)]
// dz: [n*9, c_out] tf32-rounded; hmask: [ceil(n/14)*126, mask_ld] activation bits from gnb_edge_hidden_fwd_mask (rows
// beyond n*9 are read but ignored; mask_ld % 4 == 0, mask_ld >= 4*ceil(hdim/128)); nbr: [n, 9] (-1 padded).
// dpq: [n, >= 2 hdim]; its Q half must be zero on entry (the P half is overwritten). hdim <= 512, n * ldpq < 2^31.
GNB_EXPORT int gnb_edge_hidden_dgrad_scatter_tf32(const float* dz, int64_t lddz, int32_t c_out, const float* wt, int64_t ldw,
                                                  const uint32_t* hmask, int32_t mask_ld, int32_t hdim, const int32_t* nbr,
                                                  int64_t n, float* dpq, int64_t ldpq, void* stream) {
    if (n < 0 || hdim < 1 || hdim > 512 || c_out < 1 || ldpq < 2 * (int64_t)hdim) return GNB_ERR_ARG;
    if ((mask_ld & 3) || mask_ld > SC_MAX_MASK_LD || mask_ld < 4 * ((hdim + 127) / 128)) return GNB_ERR_ARG;
    if ((reinterpret_cast<uintptr_t>(hmask) & 15u) || n * ldpq >= ((int64_t)1 << 31)) return GNB_ERR_ARG;
    if (n == 0) return GNB_OK;
    const int64_t rows = n * AGG_W;
    if (rows >= (int64_t)1 << 31) return GNB_ERR_ARG;
    PartInfo pi;
    TmapArray tx;
    pi.nparts = 1;
    pi.kblocks[0] = (c_out + TC_BK - 1) / TC_BK;
    const int64_t ktot = (int64_t)pi.kblocks[0] * TC_BK;
    int rc = gnb_make_tmap_f32(&tx.m[0], dz, rows, c_out, lddz, AGG_ROWS);
    if (rc != 0) return rc == -2 ? GNB_ERR_UNSUPPORTED : GNB_ERR_ARG;
    for (int p = 1; p < TC_MAX_PARTS; ++p) { pi.kblocks[p] = 0; tx.m[p] = tx.m[0]; }
    if (ldw < ktot) return GNB_ERR_ARG;
    CUtensorMap tw;
    rc = gnb_make_tmap_f32(&tw, wt, hdim, ktot, ldw, TC_BM);
    if (rc != 0) return rc == -2 ? GNB_ERR_UNSUPPORTED : GNB_ERR_ARG;
    if (g_num_sms == 0) {
        int dev = 0;
        GNB_CHECK(cudaGetDevice(&dev));
        GNB_CHECK(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
        GNB_CHECK(cudaFuncSetAttribute(gemm_tc_linear_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)TC_SMEM_BYTES));
    }
    const int row_tiles = gnb_div_up(n, AGG_NPT);
    const int groups = gnb_div_up(hdim, TC_MT * TC_BM);
    int ctas_x = g_num_sms / groups;
    if (ctas_x < 1) ctas_x = 1;
    if (ctas_x > row_tiles) ctas_x = row_tiles;
    dim3 grid((unsigned)ctas_x, (unsigned)groups);
    AggInfo agg{nullptr, 0, nullptr, 0, g_linear_dbg, g_linear_prof};
    ScatInfo sc{nbr, hmask, mask_ld, dpq, ldpq, hdim, n, 1};
    gemm_tc_linear_kernel<<<grid, TC_THREADS, TC_SMEM_BYTES, (cudaStream_t)stream>>>(tw, tx, pi, nullptr, nullptr, 0, rows, hdim,
                                                                                     GNB_ACT_NONE, 0, row_tiles, agg, sc);
    GNB_RETURN_LAUNCH();
}

// dst[rows, dst_cols] = [rna_tf32(src[rows, cols]) | 0]; used to pack weights / round activations.
GNB_EXPORT int gnb_round_pad_tf32(const float* src, int64_t lds, int64_t rows, int32_t cols, float* dst, int64_t ldd,
                                  int32_t dst_cols, void* stream) {
    if (dst_cols < cols || rows < 0) return GNB_ERR_ARG;
    if (rows == 0) return GNB_OK;
    round_pad_tf32_kernel<<<gnb_div_up(rows * dst_cols, 256), 256, 0, (cudaStream_t)stream>>>(src, lds, rows, cols, dst,
                                                                                              ldd, dst_cols);
    GNB_RETURN_LAUNCH();
}
