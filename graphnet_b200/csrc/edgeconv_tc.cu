// Fused dynamic EdgeConv forward on tcgen05 (tf32 mode, inference path):
//
//     y[i, :] = AGG_{s < deg[i]}  relu( W2 relu(P[i] + Q[nbr[i,s]]) + b2 )
//
// i.e. the per-edge half of PyG EdgeConv.propagate at src/graphnet/models/components/layers.py:60 for the DynEdge
// MLP (dynedge.py:192-210) after the first Linear has been hoisted to nodes (pq = [P | Q] = x [W1a-W1b ; W1b]^T).
// The gather, the hidden activation, the E x H x C contraction, bias + ReLU and the k-neighbour aggregation run
// in ONE kernel: neither the [E, H] hidden tensor nor the [E, C] message tensor ever exists in HBM.
//
// Mapping. Orientation is "weights are A": D[channel, edge] = W2_half[128 ch x H] * Hid[edges x H]^T, so a TMEM
// lane is an output channel and a TMEM column is an edge; an epilogue thread then sums its channel over the
// consecutive columns of a node entirely in registers (no shuffles, any degree) and writes y[node, ch]
// with 32 consecutive channels per warp store.
//   * a CTA owns one 128-channel half of W2 for its whole life: the half (<= 11 K-blocks of 128 x 32 tf32 =
//     176 KiB) is TMA-loaded ONCE into shared memory; no weight streaming in the main loop;
//   * persistent loop over tiles of `npt` consecutive nodes (npt * width <= 128 edges);
//   * warp 0 prepares per-tile edge lists (source / target node per row, node boundaries) up to two tiles ahead;
//   * 8 builder warps gather P[i] + Q[j] from L2/HBM (fully coalesced 128-byte row segments, loads issued two
//     K-blocks ahead), apply ReLU, round to tf32 and write the B operand tile (<= 128 edges x 32) straight into the 128-byte-
//     swizzled K-major layout (2-stage ring, mbarrier + proxy fence);
//   * 1 thread issues tcgen05.mma (kind::tf32, M = 128, N = edges rounded up to 16);
//   * accumulators double-buffered in TMEM (2 x 128 columns): 4 epilogue warps drain tile t while tile t+1 is
//     being built and multiplied.
#include "common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int EF_THREADS = 480;                 // 15 warps: 0 TMA/meta, 1 MMA, 2-5 epilogue, 6-13 builders, 14 signal
constexpr int EF_BUILDERS = 256;
constexpr int EF_MAX_KB = 11;                   // hidden width <= 352
constexpr int EF_STAGES = 2;
constexpr uint32_t EF_TILE_BYTES = 128 * 32 * 4;   // 16 KiB: 128 rows x 32 tf32

// Builders publish a finished B stage with a non-blocking named-barrier arrive; a dedicated signal warp (no global
// loads in flight) joins that named barrier and performs the mbarrier arrive. A release-arrive issued by a builder
// itself compiles to MEMBAR + ERRBAR and would wait for the builder's prefetched global loads every K step.
__device__ __forceinline__ void named_bar_arrive(int id, int count) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int count) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
constexpr int EF_SIGNAL_COUNT = EF_BUILDERS + 32;

struct TileMeta {
    int src[128];          // source node j of each edge row
    int node[128];         // target node i of each edge row
    unsigned last[4];      // bit e set: edge row e is the last edge of its node
    int flush_node[32];    // target node of the f-th flush
    float flush_scale[32]; // 1 (add) or 1/deg (mean)
    int n_edges, n_mma;
    int regular;           // every node of the tile has exactly 8 in-edges: node f owns edge rows [8f, 8f+8)
    int n_nodes;
};

struct EfParams {
    const float* pq; int64_t ldpq; int hdim; int kblocks;
    const int* nbr; const int* deg; int width; int64_t n;
    const float* b2; int c_out; int aggr; int round_out;
    float* y; int64_t ldy; int num_tiles; int npt;
    unsigned long long* prof;   // optional [16] cycle counters written by cluster 0 (tuning aid)
    int debug;                  // tuning experiments: bit0 Q rows := own node (no gather), bit1 no global loads, bit2 no epilogue, bit3 no MMA
};


// Epilogue of one tile for one thread (= one output channel): bias + ReLU on every edge column of the TMEM
// accumulator, k-neighbour sum in registers, one coalesced store per node. Fast path for tiles whose nodes all
// have 8 in-edges (the common case for k = 8): fixed groups of 8 columns, no per-element control flow.
__device__ __forceinline__ void ef_epilogue_tile(const TileMeta& m, uint32_t taddr, int64_t node0, int ch, bool ch_ok,
                                                 float bv, const EfParams& p) {
    const int n_edges = m.n_edges;
    if (m.regular) {
        const int nn = m.n_nodes;
        const float sc = p.aggr == GNB_AGGR_MEAN ? 0.125f : 1.f;
        for (int c = 0; c * 32 < n_edges; ++c) {
            uint32_t r[32];
            tc::tmem_ld_32x32b_x32(taddr + (uint32_t)(c * 32), r);
            tc::tmem_ld_wait();
#pragma unroll
            for (int f = 0; f < 4; ++f) {
                float a = 0.f;
#pragma unroll
                for (int i = 0; i < 8; ++i) a += fmaxf(__uint_as_float(r[8 * f + i]) + bv, 0.f);
                a *= sc;
                if (p.round_out) a = tc::round_tf32(a);
                const int node = 4 * c + f;
                if (ch_ok && node < nn) p.y[(node0 + node) * p.ldy + ch] = a;
            }
        }
        return;
    }
    float acc = 0.f;
    int fl = 0;
    for (int c = 0; c * 32 < n_edges; ++c) {
        uint32_t r[32];
        tc::tmem_ld_32x32b_x32(taddr + (uint32_t)(c * 32), r);
        tc::tmem_ld_wait();
        const unsigned lastbits = m.last[c];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            acc += fmaxf(__uint_as_float(r[j]) + bv, 0.f);
            if ((lastbits >> j) & 1u) {
                float o = acc * m.flush_scale[fl];
                if (p.round_out) o = tc::round_tf32(o);
                if (ch_ok) p.y[(int64_t)m.flush_node[fl] * p.ldy + ch] = o;
                ++fl;
                acc = 0.f;
            }
        }
    }
}

__global__ void __launch_bounds__(EF_THREADS, 1)
edgeconv_fused_fwd_kernel(const __grid_constant__ CUtensorMap tm_w2, const EfParams p) {
    gnb_pdl_begin();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* s_a = smem;                                            // [kblocks][16 KiB] resident weights
    uint8_t* s_b = smem + EF_MAX_KB * EF_TILE_BYTES;                // [2][16 KiB] hidden tiles
    TileMeta* meta = reinterpret_cast<TileMeta*>(s_b + EF_STAGES * EF_TILE_BYTES);   // [2]
    uint64_t* bars = reinterpret_cast<uint64_t*>(meta + 2);
    uint64_t* a_full = bars;
    uint64_t* b_full = bars + 1;      // [2]
    uint64_t* b_empty = bars + 3;     // [2]
    uint64_t* tmem_full = bars + 5;   // [2]
    uint64_t* tmem_empty = bars + 7;  // [2]
    uint64_t* meta_full = bars + 9;   // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int half = blockIdx.y;                    // 128-channel block of W2 owned by this CTA
    const int ch_base = half * 128;

    if (warp == 1) {
        if (lane == 0) {
            tc::mbar_init(a_full, 1);
            for (int s = 0; s < 2; ++s) {
                tc::mbar_init(&b_full[s], 1);
                tc::mbar_init(&b_empty[s], 1);
                tc::mbar_init(&tmem_full[s], 1);
                tc::mbar_init(&tmem_empty[s], 4);
                tc::mbar_init(&meta_full[s], 1);
            }
            tc::fence_barrier_init();
            tc::fence_proxy_async();
        }
        __syncwarp();
        tc::tmem_alloc<256>(tmem_slot);
    }
    tc::tcgen05_fence_before();
    __syncthreads();
    tc::tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {   // resident weights: one TMA box per K-block, a single transaction barrier
            tc::tma_prefetch_desc(&tm_w2);
            tc::mbar_arrive_expect_tx(a_full, (uint32_t)p.kblocks * EF_TILE_BYTES);
            for (int kb = 0; kb < p.kblocks; ++kb) tc::tma_load_2d(s_a + kb * EF_TILE_BYTES, &tm_w2, a_full, kb * 32, ch_base);
        }
        __syncwarp();
        // ---- tile metadata producer: runs up to two tiles ahead of the builders ---------------------------
        uint32_t ti = 0;
        for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++ti) {
            const uint32_t buf = ti & 1;
            tc::mbar_wait<200>(&tmem_empty[buf], ((ti >> 1) & 1) ^ 1); // epilogue of tile ti-2 has released meta[buf]
            TileMeta& m = meta[buf];
            const int64_t node0 = (int64_t)t * p.npt;
            const int nn = (int)((p.n - node0) < p.npt ? (p.n - node0) : p.npt);
            const int d = lane < nn ? p.deg[node0 + lane] : 0;
            int incl = d;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            const int off = incl - d;
            const int total = __shfl_sync(0xffffffffu, incl, 31);
            if (lane < 4) m.last[lane] = 0u;
            __syncwarp();
            const unsigned has = __ballot_sync(0xffffffffu, d > 0);
            if (d > 0) {
                const int pos = __popc(has & ((1u << lane) - 1u));
                m.flush_node[pos] = (int)(node0 + lane);
                m.flush_scale[pos] = p.aggr == GNB_AGGR_MEAN ? 1.f / (float)d : 1.f;
                for (int s = 0; s < d; ++s) {
                    m.src[off + s] = p.nbr[(node0 + lane) * p.width + s];
                    m.node[off + s] = (int)(node0 + lane);
                }
                const int e = off + d - 1;
                atomicOr(&m.last[e >> 5], 1u << (e & 31));
            }
            const unsigned all8 = __ballot_sync(0xffffffffu, lane >= nn || d == 8);
            if (lane == 0) {
                m.n_edges = total;
                const int nm = (total + 15) & ~15;
                m.n_mma = nm < 16 ? 16 : nm;
                m.regular = (all8 == 0xffffffffu) ? 1 : 0;
                m.n_nodes = nn;
            }
            // nodes without in-edges aggregate to 0 (PyG: empty neighbourhood -> 0)
            for (int l = 0; l < nn; ++l) {
                const int dl = __shfl_sync(0xffffffffu, d, l);
                if (dl == 0)
                    for (int c = lane; c < 128; c += 32)
                        if (ch_base + c < p.c_out) p.y[(node0 + l) * p.ldy + ch_base + c] = 0.f;
            }
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&meta_full[buf]);
        }
    } else if (warp == 1) {
        {   // whole warp in uniform control flow, one elected lane issues (see tc::elect_one)
            tc::mbar_wait(a_full, 0);
            uint32_t it = 0, ti = 0;
            for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++ti) {
                const uint32_t buf = ti & 1;
                tc::mbar_wait(&tmem_empty[buf], ((ti >> 1) & 1) ^ 1);
                tc::tcgen05_fence_after();
                const uint32_t acc = tmem_base + buf * 128;
                for (int kb = 0; kb < p.kblocks; ++kb, ++it) {
                    const uint32_t s = it & 1, ph = (it >> 1) & 1;
                    tc::mbar_wait<20>(&b_full[s], ph);
                    tc::tcgen05_fence_after();
                    const uint32_t idesc = tc::umma_idesc_tf32(128, (uint32_t)meta[buf].n_mma);
                    const uint64_t adesc = tc::umma_desc_sw128_kmajor(tc::smem_u32(s_a + kb * EF_TILE_BYTES));
                    const uint64_t bdesc = tc::umma_desc_sw128_kmajor(tc::smem_u32(s_b + s * EF_TILE_BYTES));
                    const bool lead = tc::elect_one();      // one election per K block: MMAs + commit issued back to back
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (lead) tc::umma_tf32(acc, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                    if (lead) tc::umma_commit(&b_empty[s]);
                    __syncwarp();
                }
                if (tc::elect_one()) tc::umma_commit(&tmem_full[buf]);
                __syncwarp();
            }
        }
    } else if (warp < 6) {
        // ---- epilogue: bias + ReLU + k-neighbour aggregation in registers, coalesced stores ----------------
        const int q = warp & 3;
        const int ch = ch_base + q * 32 + lane;
        const bool ch_ok = ch < p.c_out;
        const float bv = (ch_ok && p.b2 != nullptr) ? p.b2[ch] : 0.f;
        uint32_t ti = 0;
        for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++ti) {
            const uint32_t buf = ti & 1;
            tc::mbar_wait<100>(&tmem_full[buf], (ti >> 1) & 1);
            tc::tcgen05_fence_after();
            ef_epilogue_tile(meta[buf], tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * 128), (int64_t)t * p.npt, ch,
                             ch_ok, bv, p);
            tc::tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&tmem_empty[buf]);
        }
    } else if (warp < 14) {
        // ---- builders: gather P[i] + Q[j], ReLU, tf32 rounding, swizzled K-major store -------------------
        // 8 consecutive lanes cover the 128 bytes (one K-block) of ONE edge row, so every warp load touches 4 rows x
        // 128 B (full sectors); a thread owns 16-byte chunk `chunk` of rows 4*rg + {0,1,2,3}. Global loads run two
        // K-blocks ahead of the shared-memory stores (register ring of three, statically unrolled).
        const int bt = threadIdx.x - 192;            // 0..255
        const int chunk = bt & 7, rg = bt >> 3;
        const int kbs = p.kblocks;
        const float* __restrict__ pq = p.pq;
        const int tiles_cta = (p.num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
        const bool tail_chunk = (kbs - 1) * 32 + chunk * 4 >= p.hdim;      // this chunk is K padding in the last block
        uint32_t g = 0;                                                     // flat (tile, K-block) step counter

        struct Regs { float4 pv[4], qv[4]; };
        for (int tl = 0; tl < tiles_cta; ++tl) {
            const int b = tl & 1;
            tc::mbar_wait<50>(&meta_full[b], (tl >> 1) & 1);
            const TileMeta& m = meta[b];
            const int n_edges = m.n_edges;
            int offp[4], offq[4];
            bool ok[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int row = 4 * rg + j;
                ok[j] = row < n_edges;
                // rows past the edge list read a valid dummy location (row 0 of the tile); they are never stored
                const int rr = ok[j] ? row : 0;
                offp[j] = m.node[rr] * (int)p.ldpq + chunk * 4;
                offq[j] = m.src[rr] * (int)p.ldpq + p.hdim + chunk * 4;
            }
            if (n_edges == 0) { offp[0] = offp[1] = offp[2] = offp[3] = 0; offq[0] = offq[1] = offq[2] = offq[3] = 0; }

            auto load = [&](int kb, Regs& r) {
                if (kb >= kbs) return;
                if (kb == kbs - 1 && tail_chunk) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) { r.pv[j] = make_float4(0.f, 0.f, 0.f, 0.f); r.qv[j] = r.pv[j]; }
                    return;
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    r.pv[j] = __ldg(reinterpret_cast<const float4*>(pq + offp[j] + kb * 32));
                    r.qv[j] = __ldg(reinterpret_cast<const float4*>(pq + offq[j] + kb * 32));
                }
            };
            auto emit = [&](int kb, const Regs& r) {
                if (kb >= kbs) return;
                const uint32_t st = g & 1, ph = (g >> 1) & 1;
                ++g;
                tc::mbar_wait<20>(&b_empty[st], ph ^ 1);             // MMA has consumed this stage
                uint8_t* base = s_b + st * EF_TILE_BYTES;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (ok[j]) {
                        const int row = 4 * rg + j;
                        // relu, then round-to-nearest tf32: +half ulp (0x1000) on the bits; the MMA truncates the rest
                        uint4 v;
                        v.x = __float_as_uint(fmaxf(r.pv[j].x + r.qv[j].x, 0.f)) + 0x1000u;
                        v.y = __float_as_uint(fmaxf(r.pv[j].y + r.qv[j].y, 0.f)) + 0x1000u;
                        v.z = __float_as_uint(fmaxf(r.pv[j].z + r.qv[j].z, 0.f)) + 0x1000u;
                        v.w = __float_as_uint(fmaxf(r.pv[j].w + r.qv[j].w, 0.f)) + 0x1000u;
                        *reinterpret_cast<uint4*>(base + row * 128 + ((chunk ^ (row & 7)) << 4)) = v;
                    }
                }
                tc::fence_proxy_async();
                named_bar_arrive(1 + (int)st, EF_SIGNAL_COUNT);
            };

            Regs ra, rb, rc;
            load(0, ra);
            load(1, rb);
            for (int kb = 0; kb < kbs; kb += 3) {
                load(kb + 2, rc);
                emit(kb, ra);
                load(kb + 3, ra);
                emit(kb + 1, rb);
                load(kb + 4, rb);
                emit(kb + 2, rc);
            }
        }
    } else {
        // ---- signal warp: forwards "stage built" from the builders' named barrier to the MMA's mbarrier --------
        const int tiles_cta = (p.num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
        const uint32_t total = (uint32_t)tiles_cta * (uint32_t)p.kblocks;
        for (uint32_t g = 0; g < total; ++g) {
            const uint32_t st = g & 1;
            named_bar_sync(1 + (int)st, EF_SIGNAL_COUNT);
            if (lane == 0) tc::mbar_arrive(&b_full[st]);
        }
    }
    tc::tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc<256>(tmem_base);
}


// =====================================================================================================
// CTA-pair variant (cta_group::2) for c_out in (128, 256]: the two CTAs of a cluster own the two 128-channel
// halves of W2 (each resident in its own shared memory) and SHARE the gathered hidden tile: every CTA builds
// only 64 of the tile's 128 edge rows (half the gather / ALU work and half the B bytes per SM, 4 x 8 KiB
// stages instead of 2 x 16 KiB), the leader CTA issues one M = 256, N = 128 tcgen05.mma per K step, and each
// CTA's TMEM receives its own 128 channels x 128 edges.
// Cross-CTA signalling: builders of both CTAs arrive on the LEADER's b_full[s] (remote mbarrier arrive through
// mapa), the MMA commits are multicast to both CTAs' b_empty / tmem_full, epilogues of both CTAs arrive on the
// leader's tmem_empty.
constexpr int EP_STAGES = 4;
constexpr int EP_THREADS = EF_THREADS + 32;      // + a second signal warp (warps 14, 15 alternate K steps)
constexpr uint32_t EP_BTILE_BYTES = 64 * 32 * 4;    // 8 KiB: 64 rows x 32 tf32 per CTA

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(EP_THREADS, 1)
edgeconv_fused_pair_kernel(const __grid_constant__ CUtensorMap tm_w2, const EfParams p) {
    gnb_pdl_begin();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* s_a = smem;                                            // [kblocks][16 KiB] resident weights (this CTA's half)
    uint8_t* s_b = smem + EF_MAX_KB * EF_TILE_BYTES;                // [4][8 KiB] this CTA's 64 rows of the hidden tile
    TileMeta* meta = reinterpret_cast<TileMeta*>(s_b + EP_STAGES * EP_BTILE_BYTES);   // [2]
    uint64_t* bars = reinterpret_cast<uint64_t*>(meta + 2);
    uint64_t* a_full = bars;            // [1]
    uint64_t* b_full = bars + 1;        // [4]  (used on the leader: 16 warp arrivals)
    uint64_t* b_empty = bars + 5;       // [4]  (multicast commit)
    uint64_t* tmem_full = bars + 9;     // [2]  (multicast commit)
    uint64_t* tmem_empty = bars + 11;   // [2]  (used on the leader: 8 warp arrivals)
    uint64_t* epi_done = bars + 13;     // [2]  local: 4 epilogue warps (metadata slot reuse)
    uint64_t* meta_full = bars + 15;    // [2]  local
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = tc::cluster_ctarank();       // 0 = leader
    const int ch_base = (int)rank * 128;
    const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
    const bool prof_on = p.prof != nullptr && blockIdx.x == 0;
    long long pw0 = 0, pw1 = 0;                       // cycles this thread spent in its two kinds of waits
    const long long pt0 = clock64();
#define EF_TIMED(acc, stmt) do { const long long c0__ = clock64(); stmt; acc += clock64() - c0__; } while (0)

    if (warp == 1) {
        if (lane == 0) {
            tc::mbar_init(a_full, 1);
            for (int s = 0; s < EP_STAGES; ++s) { tc::mbar_init(&b_full[s], 2); tc::mbar_init(&b_empty[s], 1); }
            for (int s = 0; s < 2; ++s) {
                tc::mbar_init(&tmem_full[s], 1);
                tc::mbar_init(&tmem_empty[s], 8);
                tc::mbar_init(&epi_done[s], 4);
                tc::mbar_init(&meta_full[s], 1);
            }
            tc::fence_barrier_init();
            tc::fence_proxy_async();
        }
        __syncwarp();
        tc::tmem_alloc_2cta<256>(tmem_slot);
    }
    tc::tcgen05_fence_before();
    __syncthreads();                                          // barriers are initialised from here on
    if (threadIdx.x == 0) {   // this CTA's resident half of W2
        tc::tma_prefetch_desc(&tm_w2);
        tc::mbar_arrive_expect_tx(a_full, (uint32_t)p.kblocks * EF_TILE_BYTES);
        for (int kb = 0; kb < p.kblocks; ++kb) tc::tma_load_2d(s_a + kb * EF_TILE_BYTES, &tm_w2, a_full, kb * 32, ch_base);
        tc::mbar_wait<100>(a_full, 0);                        // weights landed in THIS CTA
    }
    __syncthreads();
    tc::cluster_sync_all();                                   // barriers initialised + weights resident in BOTH CTAs
    tc::tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ---- tile metadata producer (both CTAs compute the same edge list; each zero-fills its channel half) ----
        uint32_t ti = 0;
        for (int t = cluster_id; t < p.num_tiles; t += num_clusters, ++ti) {
            const uint32_t buf = ti & 1;
            EF_TIMED(pw0, tc::mbar_wait<200>(&epi_done[buf], ((ti >> 1) & 1) ^ 1));
            TileMeta& m = meta[buf];
            const int64_t node0 = (int64_t)t * p.npt;
            const int nn = (int)((p.n - node0) < p.npt ? (p.n - node0) : p.npt);
            const int d = lane < nn ? p.deg[node0 + lane] : 0;
            int incl = d;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            const int off = incl - d;
            const int total = __shfl_sync(0xffffffffu, incl, 31);
            if (lane < 4) m.last[lane] = 0u;
            __syncwarp();
            const unsigned has = __ballot_sync(0xffffffffu, d > 0);
            if (d > 0) {
                const int pos = __popc(has & ((1u << lane) - 1u));
                m.flush_node[pos] = (int)(node0 + lane);
                m.flush_scale[pos] = p.aggr == GNB_AGGR_MEAN ? 1.f / (float)d : 1.f;
                for (int s = 0; s < d; ++s) {
                    m.src[off + s] = p.nbr[(node0 + lane) * p.width + s];
                    m.node[off + s] = (int)(node0 + lane);
                }
                const int e = off + d - 1;
                atomicOr(&m.last[e >> 5], 1u << (e & 31));
            }
            const unsigned all8 = __ballot_sync(0xffffffffu, lane >= nn || d == 8);
            if (lane == 0) { m.n_edges = total; m.n_mma = 128; m.regular = (all8 == 0xffffffffu) ? 1 : 0; m.n_nodes = nn; }
            for (int l = 0; l < nn; ++l) {
                const int dl = __shfl_sync(0xffffffffu, d, l);
                if (dl == 0)
                    for (int c = lane; c < 128; c += 32)
                        if (ch_base + c < p.c_out) p.y[(node0 + l) * p.ldy + ch_base + c] = 0.f;
            }
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&meta_full[buf]);
        }
    } else if (warp == 1) {
        if (rank == 0) {   // whole warp in uniform control flow, one elected lane issues (see tc::elect_one)
            // ---- MMA issuer (leader CTA only): M = 256 over both CTAs, N = 128 edges ------------------------
            const uint32_t idesc = tc::umma_idesc_tf32(256, 128);
            uint32_t it = 0, ti = 0;
            for (int t = cluster_id; t < p.num_tiles; t += num_clusters, ++ti) {
                const uint32_t buf = ti & 1;
                EF_TIMED(pw1, tc::mbar_wait<20>(&tmem_empty[buf], ((ti >> 1) & 1) ^ 1));
                tc::tcgen05_fence_after();
                const uint32_t acc = tmem_base + buf * 128;
                for (int kb = 0; kb < p.kblocks; ++kb, ++it) {
                    const uint32_t s = it % EP_STAGES, ph = (it / EP_STAGES) & 1;
                    EF_TIMED(pw0, tc::mbar_wait<20>(&b_full[s], ph));
                    tc::tcgen05_fence_after();
                    const uint64_t adesc = tc::umma_desc_sw128_kmajor(tc::smem_u32(s_a + kb * EF_TILE_BYTES));
                    const uint64_t bdesc = tc::umma_desc_sw128_kmajor(tc::smem_u32(s_b + s * EP_BTILE_BYTES));
                    const bool lead = tc::elect_one();      // one election per K block: MMAs + commit issued back to back
                    if (!(p.debug & 8)) {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (lead)
                            tc::umma_tf32_2cta(acc, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    if (lead) tc::umma_commit_2cta(&b_empty[s], 3);
                    __syncwarp();
                }
                if (tc::elect_one()) tc::umma_commit_2cta(&tmem_full[buf], 3);
                __syncwarp();
            }
        }
    } else if (warp < 6) {
        // ---- epilogue (each CTA: its own 128 channels) ------------------------------------------------------
        const int q = warp & 3;
        const int ch = ch_base + q * 32 + lane;
        const bool ch_ok = ch < p.c_out;
        const float bv = (ch_ok && p.b2 != nullptr) ? p.b2[ch] : 0.f;
        uint32_t ti = 0;
        for (int t = cluster_id; t < p.num_tiles; t += num_clusters, ++ti) {
            const uint32_t buf = ti & 1;
            EF_TIMED(pw0, tc::mbar_wait<100>(&tmem_full[buf], (ti >> 1) & 1));
            tc::tcgen05_fence_after();
            if (!(p.debug & 4))
            ef_epilogue_tile(meta[buf], tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * 128), (int64_t)t * p.npt, ch,
                             ch_ok, bv, p);
            tc::tcgen05_fence_before();
            __syncwarp();
            EF_TIMED(pw1, if (lane == 0) {
                tc::mbar_arrive(&epi_done[buf]);                  // local: metadata slot may be rewritten
                tc::mbar_arrive_cluster_relaxed(&tmem_empty[buf], 0);     // leader: accumulator buffer may be overwritten (reads done: wait::ld)
            } __syncwarp());
        }
    } else if (warp < 14) {
        // ---- builders: this CTA's 64 edge rows (tile rows 64*rank + [0, 64)) --------------------------------
        const int bt = threadIdx.x - 192;            // 0..255
        const int chunk = bt & 7, rg = bt >> 3;      // local rows 2*rg + {0,1}
        const int kbs = p.kblocks;
        const float* __restrict__ pq = p.pq;
        const bool tail_chunk = (kbs - 1) * 32 + chunk * 4 >= p.hdim;
        uint32_t g = 0;
        struct Regs { float4 pv[2], qv[2]; };
        uint32_t tl = 0;
        for (int t = cluster_id; t < p.num_tiles; t += num_clusters, ++tl) {
            const int b = tl & 1;
            EF_TIMED(pw0, tc::mbar_wait<50>(&meta_full[b], (tl >> 1) & 1));
            const TileMeta& m = meta[b];
            const int n_edges = m.n_edges;
            int offp[2], offq[2];
            bool ok[2];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int row = 64 * (int)rank + 2 * rg + j;     // row of the 128-row tile
                ok[j] = row < n_edges;
                const int rr = (ok[j] && n_edges > 0) ? row : 0;
                offp[j] = n_edges > 0 ? m.node[rr] * (int)p.ldpq + chunk * 4 : 0;
                offq[j] = n_edges > 0 ? ((p.debug & 1) ? m.node[rr] : m.src[rr]) * (int)p.ldpq + p.hdim + chunk * 4 : 0;
            }
            auto load = [&](int kb, Regs& r) {
                if (kb >= kbs) return;
                if ((kb == kbs - 1 && tail_chunk) || (p.debug & 2)) {
#pragma unroll
                    for (int j = 0; j < 2; ++j) { r.pv[j] = make_float4(0.f, 0.f, 0.f, 0.f); r.qv[j] = r.pv[j]; }
                    return;
                }
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    r.pv[j] = __ldg(reinterpret_cast<const float4*>(pq + offp[j] + kb * 32));
                    r.qv[j] = __ldg(reinterpret_cast<const float4*>(pq + offq[j] + kb * 32));
                }
            };
            auto emit = [&](int kb, const Regs& r) {
                if (kb >= kbs) return;
                const uint32_t st = g % EP_STAGES, ph = (g / EP_STAGES) & 1;
                ++g;
                EF_TIMED(pw1, tc::mbar_wait<20>(&b_empty[st], ph ^ 1));
                uint8_t* base = s_b + st * EP_BTILE_BYTES;
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    if (ok[j]) {
                        const int lrow = 2 * rg + j;             // row inside this CTA's 64-row B tile
                        uint4 v;
                        v.x = __float_as_uint(fmaxf(r.pv[j].x + r.qv[j].x, 0.f)) + 0x1000u;
                        v.y = __float_as_uint(fmaxf(r.pv[j].y + r.qv[j].y, 0.f)) + 0x1000u;
                        v.z = __float_as_uint(fmaxf(r.pv[j].z + r.qv[j].z, 0.f)) + 0x1000u;
                        v.w = __float_as_uint(fmaxf(r.pv[j].w + r.qv[j].w, 0.f)) + 0x1000u;
                        *reinterpret_cast<uint4*>(base + lrow * 128 + ((chunk ^ (lrow & 7)) << 4)) = v;
                    }
                }
                // no per-thread proxy fence here: the named barrier orders these stores before the signal warp, which
                // issues ONE fence.proxy.async per stage (a fence in every builder warp costs ~400 cycles per K step)
                named_bar_arrive(1 + (int)st, EF_SIGNAL_COUNT);
            };
            Regs r0, r1, r2, r3;
            load(0, r0); load(1, r1); load(2, r2);
            for (int kb = 0; kb < kbs; kb += 4) {
                load(kb + 3, r3); emit(kb, r0);
                load(kb + 4, r0); emit(kb + 1, r1);
                load(kb + 5, r1); emit(kb + 2, r2);
                load(kb + 6, r2); emit(kb + 3, r3);
            }
        }
    } else {
        // ---- signal warp: this CTA's 64 rows of stage st are built -> arrive on the LEADER's b_full[st] ----------
        const int tiles_cl = (p.num_tiles - cluster_id + num_clusters - 1) / num_clusters;
        const uint32_t total = (uint32_t)tiles_cl * (uint32_t)p.kblocks;
        for (uint32_t g = (uint32_t)(warp - 14); g < total; g += 2) {      // warp 14: even steps, warp 15: odd steps
            const uint32_t st = g % EP_STAGES;
            EF_TIMED(pw0, named_bar_sync(1 + (int)st, EF_SIGNAL_COUNT));
            tc::fence_proxy_async();          // builders' generic-proxy stores -> visible to the tensor core (async proxy)
            EF_TIMED(pw1, if (lane == 0) tc::mbar_arrive_cluster_relaxed(&b_full[st], 0); __syncwarp());
        }
    }
    if (prof_on && lane == 0 && (warp == 0 || warp == 1 || warp == 2 || warp == 6 || warp == 14)) {
        const int slot = warp == 0 ? 0 : warp == 1 ? 3 : warp == 2 ? 6 : warp == 6 ? 9 : 12;
        p.prof[slot] = (unsigned long long)pw0;
        p.prof[slot + 1] = (unsigned long long)pw1;
        p.prof[slot + 2] = (unsigned long long)(clock64() - pt0);
    }
#undef EF_TIMED
    tc::tcgen05_fence_before();
    __syncthreads();
    tc::cluster_sync_all();            // no CTA may exit while its peer can still signal it
    if (warp == 1) tc::tmem_dealloc_2cta<256>(tmem_base);
}

constexpr uint32_t EP_SMEM_BYTES = EF_MAX_KB * EF_TILE_BYTES + EP_STAGES * EP_BTILE_BYTES + 2 * sizeof(TileMeta) + 256 + 1024;
int g_ef_variant = 0;      // 0 auto, 1 single-CTA kernel, 2 CTA-pair kernel
unsigned long long* g_ef_prof = nullptr;
int g_ef_debug = 0;

constexpr uint32_t EF_SMEM_BYTES = EF_MAX_KB * EF_TILE_BYTES + EF_STAGES * EF_TILE_BYTES + 2 * sizeof(TileMeta) + 128 + 1024;
int g_ef_sms = 0;

}  // namespace

// bring-up / A-B testing: 0 auto (CTA-pair kernel when c_out needs exactly two 128-channel halves), 1 single-CTA, 2 pair
GNB_EXPORT int gnb_edgeconv_set_variant(int32_t v) { g_ef_variant = v & 0xff; g_ef_debug = (v >> 8) & 0xff; return GNB_OK; }
// device buffer of 16 uint64 receiving per-role wait / total cycle counters of cluster 0 (nullptr = off)
GNB_EXPORT int gnb_edgeconv_set_profile_buffer(void* buf) { g_ef_prof = (unsigned long long*)buf; return GNB_OK; }

// w2p: [c_out, ceil(hdim/32)*32] fp32, tf32-rounded, zero padded columns. pq: [n, 2*hdim] (tf32-rounded P | Q).
// aggr: 0 add, 1 mean. hdim % 4 == 0, hdim <= 352, width <= 32.
GNB_EXPORT int gnb_edgeconv_fused_fwd_tf32(const float* pq, int64_t ldpq, int32_t hdim, const int32_t* nbr,
                                           const int32_t* deg, int32_t width, int64_t n, const float* w2p, int64_t ldw,
                                           const float* b2, int32_t c_out, int32_t aggr, int32_t round_out, float* y,
                                           int64_t ldy, void* stream) {
    const int kblocks = (hdim + 31) / 32;
    if ((hdim & 3) || hdim < 4 || kblocks > EF_MAX_KB || width < 1 || width > 32 || c_out < 1 || aggr < 0 || aggr > 1 ||
        (ldpq & 3) || (reinterpret_cast<uintptr_t>(pq) & 15u) || ldw < kblocks * 32)
        return hdim > EF_MAX_KB * 32 ? GNB_ERR_UNSUPPORTED : GNB_ERR_ARG;
    if (n == 0) return GNB_OK;
    CUtensorMap tw;
    int rc = gnb_make_tmap_f32(&tw, w2p, c_out, (int64_t)kblocks * 32, ldw, 128);
    if (rc != 0) return rc == -2 ? GNB_ERR_UNSUPPORTED : GNB_ERR_ARG;
    if (g_ef_sms == 0) {
        int dev = 0;
        GNB_CHECK(cudaGetDevice(&dev));
        GNB_CHECK(cudaDeviceGetAttribute(&g_ef_sms, cudaDevAttrMultiProcessorCount, dev));
        GNB_CHECK(cudaFuncSetAttribute(edgeconv_fused_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)EF_SMEM_BYTES));
    }
    EfParams p;
    p.pq = pq; p.ldpq = ldpq; p.hdim = hdim; p.kblocks = kblocks;
    p.nbr = nbr; p.deg = deg; p.width = width; p.n = n;
    p.b2 = b2; p.c_out = c_out; p.aggr = aggr; p.round_out = round_out;
    p.y = y; p.ldy = ldy; p.prof = g_ef_prof; p.debug = g_ef_debug;
    p.npt = 128 / width;
    if (p.npt > 32) p.npt = 32;
    p.num_tiles = gnb_div_up(n, p.npt);
    const int halves = gnb_div_up(c_out, 128);
    if (halves == 2 && g_ef_variant != 1) {   // CTA-pair kernel
        static bool pair_attr = false;
        if (!pair_attr) {
            GNB_CHECK(cudaFuncSetAttribute(edgeconv_fused_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)EP_SMEM_BYTES));
            pair_attr = true;
        }
        int clusters = g_ef_sms / 2;
        if (clusters > p.num_tiles) clusters = p.num_tiles;
        gnb_launch(edgeconv_fused_pair_kernel, dim3((unsigned)(2 * clusters)), EP_THREADS, EP_SMEM_BYTES, (cudaStream_t)stream)(tw, p);
        GNB_RETURN_LAUNCH();
    }
    int ctas = g_ef_sms / halves;
    if (ctas < 1) ctas = 1;
    if (ctas > p.num_tiles) ctas = p.num_tiles;
    dim3 grid((unsigned)ctas, (unsigned)halves);
    gnb_launch(edgeconv_fused_fwd_kernel, grid, EF_THREADS, EF_SMEM_BYTES, (cudaStream_t)stream)(tw, p);
    GNB_RETURN_LAUNCH();
}
