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
// CTA = 128 channels x 128 rows, K in blocks of 32 tf32 (= one 128-byte swizzle atom). 3-stage TMA ->
// mbarrier -> tcgen05.mma pipeline, accumulator in 128 TMEM columns, 2 CTAs resident per SM so one CTA's
// epilogue overlaps the other's main loop. Warp roles: 0 = TMA producer, 1 = MMA issuer (+TMEM alloc),
// 2..5 = epilogue (one TMEM lane quarter each).
// Operands are fp32 in HBM; weights and activations are expected pre-rounded to tf32 (cvt.rna) by their
// producers so that the tensor core's truncation is exact (unbiased rounding overall).
#include "common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int TC_BM = 128, TC_BN = 128, TC_BK = 32, TC_STAGES = 3, TC_THREADS = 192;
constexpr int TC_MAX_PARTS = 6;
constexpr uint32_t TC_STAGE_BYTES = (TC_BM + TC_BN) * TC_BK * 4;   // 32 KiB
constexpr uint32_t TC_SMEM_BYTES = TC_STAGES * TC_STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;

struct TmapArray { CUtensorMap m[TC_MAX_PARTS]; };
struct PartInfo { int nparts; int kblocks[TC_MAX_PARTS]; };

__global__ void __launch_bounds__(TC_THREADS, 2)
gemm_tc_linear_kernel(const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ TmapArray tm_x,
                      const PartInfo parts, const float* __restrict__ bias, float* __restrict__ y, int64_t ldy,
                      int64_t rows, int n_out, int act, int round_out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + TC_STAGES * TC_STAGE_BYTES);
    uint64_t* empty = full + TC_STAGES;
    uint64_t* tmem_full = empty + TC_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t row0 = (int64_t)blockIdx.x * TC_BN;
    const int ch0 = blockIdx.y * TC_BM;

    if (warp == 0 && lane == 0) {
        tc::tma_prefetch_desc(&tm_w);
        for (int p = 0; p < parts.nparts; ++p) tc::tma_prefetch_desc(&tm_x.m[p]);
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < TC_STAGES; ++s) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], 1); }
            tc::mbar_init(tmem_full, 1);
            tc::fence_barrier_init();
            tc::fence_proxy_async();
        }
        __syncwarp();
        tc::tmem_alloc<TC_BN>(tmem_slot);
    }
    tc::tcgen05_fence_before();
    __syncthreads();
    tc::tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    int total_kb = 0;
    for (int p = 0; p < parts.nparts; ++p) total_kb += parts.kblocks[p];

    if (warp == 0) {
        if (lane == 0) {
            int it = 0;
            for (int p = 0; p < parts.nparts; ++p) {
                for (int kb = 0; kb < parts.kblocks[p]; ++kb, ++it) {
                    const int s = it % TC_STAGES;
                    const uint32_t ph = (it / TC_STAGES) & 1;
                    tc::mbar_wait(&empty[s], ph ^ 1);
                    uint8_t* sa = smem + s * TC_STAGE_BYTES;
                    uint8_t* sb = sa + TC_BM * TC_BK * 4;
                    tc::mbar_arrive_expect_tx(&full[s], TC_STAGE_BYTES);
                    tc::tma_load_2d(sa, &tm_w, &full[s], it * TC_BK, ch0);
                    tc::tma_load_2d(sb, &tm_x.m[p], &full[s], kb * TC_BK, (int)row0);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = tc::umma_idesc_tf32(TC_BM, TC_BN);
            for (int it = 0; it < total_kb; ++it) {
                const int s = it % TC_STAGES;
                const uint32_t ph = (it / TC_STAGES) & 1;
                tc::mbar_wait(&full[s], ph);
                tc::tcgen05_fence_after();
                const uint32_t sa = tc::smem_u32(smem + s * TC_STAGE_BYTES);
                const uint64_t adesc = tc::umma_desc_sw128_kmajor(sa);
                const uint64_t bdesc = tc::umma_desc_sw128_kmajor(sa + TC_BM * TC_BK * 4);
#pragma unroll
                for (int k = 0; k < TC_BK / 8; ++k)   // UMMA_K = 8 tf32 = 32 bytes -> +2 in the 16-byte address field
                    tc::umma_tf32(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (it | k) != 0 ? 1u : 0u);
                tc::umma_commit(&empty[s]);
            }
            tc::umma_commit(tmem_full);
        }
    } else {
        tc::mbar_wait(tmem_full, 0);
        tc::tcgen05_fence_after();
        const int q = warp & 3;                       // TMEM lane quarter this warp may access
        const int ch = ch0 + q * 32 + lane;
        const bool ch_ok = ch < n_out;
        const float bv = (bias != nullptr && ch_ok) ? bias[ch] : 0.f;
#pragma unroll 1
        for (int c = 0; c < TC_BN / 32; ++c) {
            uint32_t r[32];
            tc::tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), r);
            tc::tmem_ld_wait();
            if (ch_ok) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int64_t row = row0 + c * 32 + j;
                    if (row < rows) {
                        float v = __uint_as_float(r[j]) + bv;
                        if (act == GNB_ACT_RELU) v = fmaxf(v, 0.f);
                        if (round_out) v = tc::round_tf32(v);
                        y[row * ldy + ch] = v;
                    }
                }
            }
        }
    }
    tc::tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc<TC_BN>(tmem_base);
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

}  // namespace

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
    static bool attr_set = false;
    if (!attr_set) {
        GNB_CHECK(cudaFuncSetAttribute(gemm_tc_linear_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)TC_SMEM_BYTES));
        attr_set = true;
    }
    dim3 grid((unsigned)gnb_div_up(rows, TC_BN), (unsigned)gnb_div_up(n_out, TC_BM));
    gemm_tc_linear_kernel<<<grid, TC_THREADS, TC_SMEM_BYTES, (cudaStream_t)stream>>>(tw, tx, pi, bias, y, ldy, rows,
                                                                                     n_out, act, round_out);
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
