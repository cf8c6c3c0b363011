// Native step executor for DynEdge: the whole forward (and backward) of
// src/graphnet/models/gnn/dynedge.py:295-349 enqueued from ONE C call.
//
// Why: the per-operator Python route launches ~500 kernels per training step through ctypes / autograd and
// is host-bound on B200 (13.8 ms of host time against ~11 ms of GPU time at 512 events). The executor walks
// the same operator sequence (same kernels, same arithmetic, same saved tensors) from C++ over one caller-
// provided workspace, so a step costs ~1-2 ms of host time.
//
// Supported model family ("fast path"): ReLU activation, no norm layers, every DynEdgeConv MLP is
// Linear-ReLU-Linear-ReLU with aggr = add, post-processing / read-out are Linear-ReLU chains, any pooling list
// (or none), global variables before or after pooling, skip_readout. Everything else stays on the
// per-operator route in graphnet_b200/models (generic `nn`, LayerNorm, GELU, max/mean aggregation).
#include "common.cuh"
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdlib.h>

// ---- launchers of the other translation units (C ABI, include/graphnet_b200.h) -----------------------
extern "C" {
int gnb_knn_table(const float*, int64_t, const int32_t*, int32_t, const int64_t*, int64_t, int64_t, int32_t, int32_t*,
                  int32_t*, void*);
int gnb_global_vars(const float*, int64_t, int32_t, const int32_t*, const int32_t*, int32_t, const int64_t*, int64_t,
                    const float*, float*, float*, int64_t, void*);
int gnb_edge_hidden_fwd(const float*, int64_t, int32_t, const int32_t*, const int32_t*, int32_t, int64_t, int32_t, float*,
                        int64_t, void*);
int gnb_edge_hidden_bwd(const float*, int64_t, const float*, int64_t, int32_t, const int32_t*, const int32_t*, int32_t,
                        int64_t, int32_t, float*, int64_t, void*);
int gnb_edge_aggregate_fwd(const float*, int64_t, int32_t, const int32_t*, int32_t, int64_t, int32_t, float*, int64_t,
                           int8_t*, void*);
int gnb_segment_pool_fwd(const float*, int64_t, int32_t, const int64_t*, int64_t, const int32_t*, int32_t, float*,
                         int32_t*, void*);
int gnb_edgeconv_fused_fwd_tf32(const float*, int64_t, int32_t, const int32_t*, const int32_t*, int32_t, int64_t,
                                const float*, int64_t, const float*, int32_t, int32_t, int32_t, float*, int64_t, void*);
int gnb_edge_linear_agg_fwd_tf32(const float*, int64_t, int32_t, const float*, int64_t, const float*, const int32_t*, int64_t,
                                 int32_t, int32_t, float*, int64_t, uint32_t*, void*);
int gnb_edge_mask_bwd_colsum(const float*, int64_t, const uint32_t*, int64_t, int32_t, const int32_t*, float*, int64_t, float*,
                             int32_t, void*);
int gnb_edge_hidden_dgrad_scatter_tf32(const float*, int64_t, int32_t, const float*, int64_t, const uint32_t*, int32_t, int32_t,
                                       const int32_t*, int64_t, float*, int64_t, void*);
int gnb_edge_hidden_dgrad_scatter_split_tf32(const float*, int64_t, int32_t, const float*, int64_t, const uint32_t*, int32_t, int32_t,
                                             const int32_t*, int64_t, float*, int64_t, float*, int64_t, float*, int32_t, void*);
int gnb_edge_hidden_fwd_mask(const float*, int64_t, int32_t, const int32_t*, const int32_t*, int32_t, int64_t, int32_t, float*,
                             int64_t, uint32_t*, int32_t, void*);
int gnb_segment_pool_bwd(const float*, int64_t, const int32_t*, int32_t, const int64_t*, int64_t, int64_t, const int32_t*,
                         int32_t, float*, int64_t, void*);
int gnb_act_bwd_colsum(const float*, int64_t, const float*, int64_t, int64_t, int32_t, float*, int64_t, float*, int32_t,
                       const int32_t*, int32_t, int32_t, void*);
int gnb_linear_fwd_tf32(const float* const*, const int64_t*, const int32_t*, int32_t, const float*, int64_t, const float*,
                        float*, int64_t, int64_t, int32_t, int32_t, int32_t, void*);
int gnb_linear_bwd_weight_tf32(const float*, int64_t, const float*, int64_t, float*, int64_t, int64_t, int32_t, int32_t,
                               int32_t, void*);
int gnb_linear_fwd_tf32x3(const float* const*, const int64_t*, const int32_t*, int32_t, const float*, const float*, int64_t,
                          const float*, float*, int64_t, int64_t, int32_t, int32_t, void*);
int gnb_edge_linear_agg_fwd_tf32x3(const float*, int64_t, int32_t, const float*, const float*, int64_t, const float*,
                                   const int32_t*, int64_t, int32_t, float*, int64_t, uint32_t*, void*);
int gnb_edge_hidden_fwd_bf16(const float*, int64_t, int32_t, const int32_t*, const int32_t*, int32_t, int64_t, void*, void*, int64_t,
                             uint32_t*, int32_t, void*);
int gnb_edge_linear_agg_fwd_bf16(const void*, const void*, int64_t, int32_t, const void*, const void*, int64_t, const float*,
                                 const int32_t*, int64_t, int32_t, int32_t, float*, int64_t, uint32_t*, void*);
int gnb_edge_mask_bwd_colsum_bf16(const float*, int64_t, const uint32_t*, int64_t, int32_t, void*, void*, int64_t, float*, void*);
int gnb_linear_bwd_weight_bf16(const void*, const void*, int64_t, const void*, const void*, int64_t, float*, int64_t, int64_t, int32_t,
                               int32_t, int32_t, void*);
int gnb_edge_hidden_dgrad_scatter_bf16(const void*, const void*, int64_t, int32_t, const void*, const void*, int64_t, const uint32_t*,
                                       int32_t, int32_t, const int32_t*, int64_t, float*, int64_t, float*, int64_t, float*, int32_t,
                                       void*);
int gnb_to_bf16_planes(const float*, int64_t, int64_t, int32_t, void*, void*, int64_t, int32_t, int32_t, void*);
int gnb_zero_block(float*, int64_t, int64_t, int32_t, void*);
int gnb_linear_next_absmax(uint32_t*, int32_t);
int gnb_linear_bwd_weight_f16_masked(const void*, const uint32_t*, const void*, int64_t, float*, int64_t, int64_t, int32_t, int32_t,
                                     const uint32_t*, const uint32_t*, void*);
int gnb_edge_hidden_dgrad_scatter_f16_masked(const void*, const uint32_t*, int32_t, const void*, int64_t, const uint32_t*, int32_t,
                                             int32_t, const int32_t*, int64_t, float*, int64_t, float*, int64_t, float*, int32_t,
                                             const uint32_t*, void*);
int gnb_edgeconv_fused_fwd_f16(const float*, int64_t, int32_t, const int32_t*, const int32_t*, int64_t, const void*, const void*,
                               int64_t, const float*, int32_t, int32_t, float*, int64_t, uint32_t*, void*, int64_t, uint8_t*, int64_t,
                               const uint32_t*, int32_t, void*);
int gnb_edge_dz_prep(const float*, int64_t, const uint32_t*, int64_t, int32_t, const uint32_t*, void*, uint32_t*, float*, void*);
// the same four with the device-side layout switch (full9: *full9 == 0 selects the 8-slot layout; gnb_edge_slot_flag writes it)
int gnb_edge_slot_flag_or(const int32_t*, int64_t, int32_t, int32_t*, void*);
int gnb_linear_bwd_weight_f16_masked_w(const void*, const uint32_t*, const void*, int64_t, float*, int64_t, int64_t, int32_t, int32_t,
                                       const uint32_t*, const uint32_t*, const int32_t*, void*);
int gnb_edge_hidden_dgrad_scatter_f16_masked_w(const void*, const uint32_t*, int32_t, const void*, int64_t, const uint32_t*, int32_t,
                                               int32_t, const int32_t*, int64_t, float*, int64_t, float*, int64_t, float*, int32_t,
                                               const uint32_t*, const int32_t*, void*);
int gnb_edgeconv_fused_fwd_f16_w(const float*, int64_t, int32_t, const int32_t*, const int32_t*, int64_t, const void*, const void*,
                                 int64_t, const float*, int32_t, int32_t, float*, int64_t, uint32_t*, void*, int64_t, uint8_t*, int64_t,
                                 const uint32_t*, int32_t, const int32_t*, void*);
int gnb_edge_dz_prep_wz(const float*, int64_t, const uint32_t*, int64_t, int32_t, const uint32_t*, void*, uint32_t*, float*,
                        const int32_t*, float*, int64_t, int64_t, int32_t, float*, int32_t, void*);
int gnb_to_f16_planes(const float*, int64_t, int64_t, int32_t, void*, void*, int64_t, int32_t, int32_t, void*);
int gnb_absmax_bits(const float*, int64_t, int64_t, int32_t, int32_t, uint32_t*, void*);
int gnb_edge_hidden_fwd_f16(const float*, int64_t, int32_t, const int32_t*, const int32_t*, int32_t, int64_t, void*, void*, int64_t,
                            uint32_t*, int32_t, const uint32_t*, void*);
int gnb_edge_linear_agg_fwd_f16(const void*, const void*, int64_t, int32_t, const void*, const void*, int64_t, const float*,
                                const int32_t*, int64_t, int32_t, int32_t, float*, int64_t, uint32_t*, const uint32_t*, void*);
int gnb_edge_mask_bwd_colsum_f16(const float*, int64_t, const uint32_t*, int64_t, int32_t, void*, int64_t, float*, const uint32_t*, void*);
int gnb_linear_bwd_weight_f16(const void*, int64_t, const void*, const void*, int64_t, float*, int64_t, int64_t, int32_t, int32_t,
                              const uint32_t*, const uint32_t*, void*);
int gnb_edge_hidden_dgrad_scatter_f16(const void*, int64_t, int32_t, const void*, int64_t, const uint32_t*, int32_t, int32_t,
                                          const int32_t*, int64_t, float*, int64_t, float*, int64_t, float*, int32_t, const uint32_t*,
                                          void*);
int gnb_linear_fwd_f32(const float*, int64_t, const float*, int64_t, const float*, float*, int64_t, int64_t, int64_t,
                       int64_t, int32_t, int32_t, void*);
int gnb_linear_bwd_data_f32(const float*, int64_t, const float*, int64_t, float*, int64_t, int64_t, int64_t, int64_t,
                            int32_t, void*);
int gnb_linear_bwd_weight_f32(const float*, int64_t, const float*, int64_t, float*, int64_t, int64_t, int64_t, int64_t,
                              void*);
}

#define GNB_MAX_LAYERS 8
#define GNB_MAX_KNN_COLS 16

struct gnb_dynedge_config {
    int32_t nb_inputs, k, precision;               // precision: 0 = fp32 SIMT, 1 = tf32 tcgen05, 2 = tf32x3 (split-operand
                                                   // forward GEMMs = fp32 grade, single-pass tf32 backward GEMMs),
                                                   // 3 = bf16 (per-edge tensors h / dz stored as ONE bf16 plane, per-edge GEMMs
                                                   // kind::f16; node-level GEMMs as in 1), 4 = bf16x3 (per-edge tensors as TWO
                                                   // bf16 planes, three products per per-edge GEMM; node-level GEMMs as in 2),
                                                   // 5 = mixed16: per-edge tensors as fp16 planes scaled per layer by a power of
                                                   // two (fp16 = tf32's significand in half the bytes): forward on two planes of
                                                   // h and W2 (three products, fp32 grade), backward on ONE plane of dz, h, W2^T
                                                   // (tf32 grade); node-level GEMMs as in 2;
                                                   // 6 = f16: ONE scaled fp16 plane forward and backward (tf32 grade at half the
                                                   // per-edge bytes); node-level GEMMs as in 1
    int32_t n_conv, conv_hidden[GNB_MAX_LAYERS], conv_out[GNB_MAX_LAYERS];
    int32_t n_post, post_out[GNB_MAX_LAYERS];
    int32_t n_readout, readout_out[GNB_MAX_LAYERS];
    int32_t n_pool, pool[4];
    int32_t globals_after_pooling, skip_readout;
    int32_t n_knn_cols, knn_cols[GNB_MAX_KNN_COLS];
    int32_t flags;                                 // bit 0: do not use the fused tcgen05 EdgeConv kernels
};

extern "C" __attribute__((visibility("default"))) long long gnb_launch_counter = 0;
// Optional marker inside the backward pass (multi-GPU overlap): once the backward of conv layer `g_bwd_event_layer` has been
// enqueued, the gradients of that layer, of the post-processing and of the read-out are final -- the event lets a side
// stream start their all-reduce while the earlier layers' backward still runs.
static cudaEvent_t g_bwd_event = nullptr;
static int g_bwd_event_layer = -1;
GNB_EXPORT int gnb_dynedge_set_backward_event(void* event, int32_t after_conv_layer) {
    g_bwd_event = (cudaEvent_t)event;
    g_bwd_event_layer = event != nullptr ? after_conv_layer : -1;
    return GNB_OK;
}
GNB_EXPORT int64_t gnb_launch_count(void) { return (int64_t)gnb_launch_counter; }

namespace {

inline int64_t up(int64_t v, int64_t m) { return (v + m - 1) / m * m; }

// ---- small kernels private to the executor ---------------------------------------------------------
// dst[r, c] = (c < cols ? maybe_round(src[r, c]) : 0)        (strided copy / pad / tf32 rounding)
// bf16 correction operand of the split GEMM (csrc/gemm_tc.cu::split_pad_tf32_kernel): row `crow` viewed as bf16, per 32-wide K
// block [bf16(v - hi) x 32 | bf16(hi) x 32]
__device__ __forceinline__ void store_corr(float* crow, int c, float v, float hi) {
    __nv_bfloat16* b = reinterpret_cast<__nv_bfloat16*>(crow);
    b[(c >> 5) * 64 + (c & 31)] = __float2bfloat16_rn(v - hi);
    b[(c >> 5) * 64 + 32 + (c & 31)] = __float2bfloat16_rn(hi);
}
// ---- weight re-packing: every packed / transposed / split weight operand of a step in ONE launch ---------------------------
// The GEMMs read weights re-packed per step (K padded to 32, tf32 rounding, the split operands' correction planes, fp16 / bf16
// planes, W^T for the data-gradient GEMMs, Wcat = [W1a - W1b ; W1b]). These used to be ~35 launches of a few microseconds
// each, interleaved with the layers and each followed by ~2 us of idle stream (scripts/r02/gaps.py): ~150 us of an 6.5 ms
// step. The forward now collects them as jobs and runs them in one kernel before the first layer; the backward finds its
// transposed operands in the workspace.
enum : int { JOB_COPY_PAD = 0, JOB_TRANSPOSE_PAD = 1, JOB_PACK_CONV = 2, JOB_PLANES = 3 };
struct PackJob {
    const float* src; const float* src2;      // PACK_CONV: W1, b1
    void* dst; void* dst2; void* dst3;        // COPY_PAD: dst, lo | TRANSPOSE_PAD: dst | PACK_CONV: Wcat, bcat, Wcat_lo | PLANES: p0, p1
    int64_t lds, ldd;
    int rows, cols, dst_cols, flags;          // flags: bit 0 round to tf32, bit 1 transpose (PLANES), bit 2 bf16 instead of fp16 (PLANES)
    int kind; unsigned first_block;
};
constexpr int PACK_MAX_JOBS = 40;
struct PackJobs { PackJob j[PACK_MAX_JOBS]; int n; unsigned total_blocks; };

// lo != nullptr (tf32x3 weights; dst_cols % 32 == 0): dst = rna_tf32(v), lo = the bf16 correction operand
__device__ __forceinline__ void copy_pad_elem(int64_t t, const float* __restrict__ src, int64_t lds, int64_t rows, int cols,
                                              float* __restrict__ dst, int64_t ldd, int dst_cols, int rnd, float* __restrict__ lo) {
    if (t >= rows * dst_cols) return;
    const int64_t r = t / dst_cols;
    const int c = (int)(t - r * dst_cols);
    float v = c < cols ? src[r * lds + c] : 0.f;
    const float hi = rnd ? gnb_round_tf32(v) : v;
    dst[r * ldd + c] = hi;
    if (lo != nullptr) store_corr(lo + r * ldd, c, v, hi);
}
// dst[c, r] = maybe_round(src[r, c]) (zero padded to dst_cols): W^T for the backward-data GEMM
__device__ __forceinline__ void transpose_pad_elem(int64_t t, const float* __restrict__ src, int64_t lds, int rows, int cols,
                                                   float* __restrict__ dst, int64_t ldd, int dst_cols, int rnd, int perm = 0) {
    if (t >= (int64_t)cols * dst_cols) return;
    const int c = (int)(t / dst_cols);          // dst row = src column
    const int r = (int)(t - (int64_t)c * dst_cols);
    // perm: the source rows (2 hid rows of a hoisted Linear's packed weight) are stored lane-interleaved; the transpose is natural
    float v = r < rows ? src[(int64_t)(perm ? gnb_pq_map_row(r, rows / 2, false) : r) * lds + c] : 0.f;
    dst[(int64_t)c * ldd + r] = rnd ? gnb_round_tf32(v) : v;
}
// First Linear of an EdgeConv MLP hoisted to nodes: W1 = [Wa | Wb] ([H, 2C]) -> Wcat = [Wa - Wb ; Wb] ([2H, ld]),
// bcat = [b1 ; 0]
__device__ __forceinline__ void pack_conv_elem(int64_t t, const float* __restrict__ w1, const float* __restrict__ b1, int h, int c,
                                               float* __restrict__ wcat, int64_t ld, float* __restrict__ bcat, int rnd,
                                               float* __restrict__ wlo, int perm = 0) {
    if (t >= 2 * (int64_t)h * ld) return;
    const int rd = (int)(t / ld);               // row of Wcat as stored
    const int col = (int)(t - (int64_t)rd * ld);
    // perm: the output columns of PQ = x Wcat^T (= rows of Wcat) in the lane-interleaved order the fused EdgeConv forward
    // gathers best (gnb_edgeconv_fused_fwd_f16, pq_layout 1); the hidden units of an MLP have no intrinsic order
    const int r = perm ? gnb_pq_map_row(rd, h, true) : rd;        // the hidden unit (natural row) stored there
    float v = 0.f;
    if (col < c) v = r < h ? w1[(int64_t)r * 2 * c + col] - w1[(int64_t)r * 2 * c + c + col] : w1[(int64_t)(r - h) * 2 * c + c + col];
    const float hi = rnd ? gnb_round_tf32(v) : v;
    wcat[(int64_t)rd * ld + col] = hi;
    if (wlo != nullptr) store_corr(wlo + (int64_t)rd * ld, col, v, hi);
    if (col == 0) bcat[rd] = r < h ? b1[r] : 0.f;
}
// one or two 16-bit planes of a weight matrix (v ~ p0 + p1), optionally transposed, zero padded to dst_cols
// (same arithmetic as gnb_to_f16_planes / gnb_to_bf16_planes in gemm_tc.cu)
template <class T16>
__device__ __forceinline__ void planes_elem(int64_t t, const float* __restrict__ src, int64_t lds, int64_t rows, int cols,
                                            T16* __restrict__ p0, T16* __restrict__ p1, int64_t ldd, int dst_cols, int transpose) {
    const int64_t drows = transpose ? cols : rows;
    if (t >= drows * dst_cols) return;
    const int64_t r = t / dst_cols;
    const int c = (int)(t - r * dst_cols);
    float v = 0.f;
    if (transpose) { if (c < rows) v = src[(int64_t)c * lds + r]; }
    else if (c < cols) v = src[r * lds + c];
    const T16 b0 = T16(v);
    p0[r * ldd + c] = b0;
    if (p1 != nullptr) p1[r * ldd + c] = T16(v - float(b0));
}
__global__ void copy_pad_kernel(const float* __restrict__ src, int64_t lds, int64_t rows, int cols,
                                float* __restrict__ dst, int64_t ldd, int dst_cols, int rnd, float* __restrict__ lo) {
    gnb_pdl_begin();
    copy_pad_elem((int64_t)blockIdx.x * blockDim.x + threadIdx.x, src, lds, rows, cols, dst, ldd, dst_cols, rnd, lo);
}
__global__ void __launch_bounds__(256) pack_jobs_kernel(const __grid_constant__ PackJobs jobs) {
    gnb_pdl_begin();
    int k = 0;
    while (k + 1 < jobs.n && blockIdx.x >= jobs.j[k + 1].first_block) ++k;
    const PackJob& jb = jobs.j[k];
    const int64_t t = (int64_t)(blockIdx.x - jb.first_block) * 256 + threadIdx.x;
    switch (jb.kind) {
        case JOB_COPY_PAD:
            copy_pad_elem(t, jb.src, jb.lds, jb.rows, jb.cols, (float*)jb.dst, jb.ldd, jb.dst_cols, jb.flags & 1, (float*)jb.dst2);
            break;
        case JOB_TRANSPOSE_PAD:
            transpose_pad_elem(t, jb.src, jb.lds, jb.rows, jb.cols, (float*)jb.dst, jb.ldd, jb.dst_cols, jb.flags & 1, (jb.flags >> 3) & 1);
            break;
        case JOB_PACK_CONV:
            pack_conv_elem(t, jb.src, jb.src2, jb.rows, jb.cols, (float*)jb.dst, jb.ldd, (float*)jb.dst2, jb.flags & 1, (float*)jb.dst3, (jb.flags >> 3) & 1);
            break;
        default:
            if (jb.flags & 4)
                planes_elem<__nv_bfloat16>(t, jb.src, jb.lds, jb.rows, jb.cols, (__nv_bfloat16*)jb.dst, (__nv_bfloat16*)jb.dst2, jb.ldd,
                                           jb.dst_cols, (jb.flags >> 1) & 1);
            else
                planes_elem<__half>(t, jb.src, jb.lds, jb.rows, jb.cols, (__half*)jb.dst, (__half*)jb.dst2, jb.ldd, jb.dst_cols,
                                    (jb.flags >> 1) & 1);
            break;
    }
}
// dst[r, c] += src[r, c]
__global__ void add2d_kernel(const float* __restrict__ src, int64_t lds, int64_t rows, int cols,
                             float* __restrict__ dst, int64_t ldd) {
    gnb_pdl_begin();
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= rows * cols) return;
    const int64_t r = t / cols;
    const int c = (int)(t - r * cols);
    dst[r * ldd + c] += src[r * lds + c];
}
// dW1[r, col] += dWcat[r, col];  dW1[r, C + col] += dWcat[H + r, col] - dWcat[r, col];  db1 += dbcat[0:H]
__global__ void unpack_conv_grad_kernel(const float* __restrict__ dwcat, int64_t ld, const float* __restrict__ dbcat,
                                        int h, int c, float* __restrict__ dw1, float* __restrict__ db1) {
    gnb_pdl_begin();
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)h * c) return;
    const int r = (int)(t / c);
    const int col = (int)(t - (int64_t)r * c);
    const float gp = dwcat[(int64_t)r * ld + col], gq = dwcat[(int64_t)(h + r) * ld + col];
    dw1[(int64_t)r * 2 * c + col] += gp;
    dw1[(int64_t)r * 2 * c + c + col] += gq - gp;
    if (col == 0) db1[r] += dbcat[r];
}

#define EX(call)                        \
    do {                                \
        int rc__ = (call);              \
        if (rc__ != 0) return rc__;     \
    } while (0)
#define EXL()                                         \
    do {                                              \
        ++gnb_launch_counter;                         \
        cudaError_t e__ = cudaGetLastError();         \
        if (e__ != cudaSuccess) return (int)e__;      \
    } while (0)

struct Arena {
    char* base;
    int64_t off = 0, cap;
    Arena(void* b, int64_t c) : base((char*)b), cap(c) {}
    template <class T> T* get(int64_t count) {
        T* p = base ? (T*)(base + off) : nullptr;
        off += up(count * (int64_t)sizeof(T), 256);
        return p;
    }
};

struct ConvBuf { float *wcat, *wcat_lo, *w2p_lo, *bcat, *w2p, *pq, *h, *m, *y; float *wcat_t, *w2t;   // (training, tensor-core modes: Wcat^T, W2^T for the data-gradient GEMMs)
                 int32_t *nbr, *deg; uint32_t *mask, *hmask; bool nodz, fused, pq_perm, slots8; int cin, cin_ld, kld, hid, hld, cout, mld;
                 // bf16 modes: h planes [n * 9, hid], W2 planes [cout, hld64], W2^T planes [hid, cld64]
                 __nv_bfloat16 *hb[2], *w2b[2], *w2tb[2]; int hld64, cld64; };
struct DenseBuf { float *wp, *wp_lo, *z; int k_total, kld, n_out; float* wt; float* wt_part[GNB_MAX_LAYERS + 1]; };   // wt: W^T (per K-split part for the first post-processing layer)

struct Plan {
    int64_t n, nseg;
    int w0, width;                       // width = k + 1
    bool agg;                            // training: second EdgeConv Linear fused with ReLU + k-sum (mask bits instead of m)
    int node_width, x0_ld;
    float *g, *x0;
    ConvBuf conv[GNB_MAX_LAYERS];
    int post_parts, part_k[GNB_MAX_LAYERS + 1], part_off[GNB_MAX_LAYERS + 1];
    DenseBuf post[GNB_MAX_LAYERS], ro[GNB_MAX_LAYERS];
    float *pooled, *rin; int32_t* parg; int pool_c, rin_cols, rin_ld; int64_t out_rows;
    // backward scratch
    float *gnode[GNB_MAX_LAYERS + 1], *dz_big, *dh_big, *dpq, *dzq, *dwp, *wt, *dbtmp, *gro_a, *gro_b;
    __nv_bfloat16* dzb[2];               // bf16 modes: dz planes [n * 9, max_c]
    int bf;                              // bf16 planes per per-edge tensor (0: fp32 / tf32 tensors)
    bool mixed;                          // mode 5: dz / W2^T as one fp16 plane in the backward pass
    __nv_bfloat16* g16; uint32_t* rowmask;   // dz-free backward: fp16(g_y 2^s) [n, max_c] and row-major ReLU bits [tiles * 126, max_c / 32]
    uint32_t* scale_bits;                // mixed16: [2][GNB_MAX_LAYERS] fp32 bits of 2 max|PQ| (>= max h) and of max|g_y| per layer
    int32_t* full9;                      // mixed16: [GNB_MAX_LAYERS] 1 = the layer's graph holds a node with k + 1 neighbours (9-slot layout), 0 = 8-slot layout
    int64_t bytes;
};

// Deterministic layout of the workspace: identical in forward and backward.
int make_plan(const gnb_dynedge_config& c, int64_t n, int64_t nseg, int w0, bool training, void* ws, int64_t cap,
              Plan& p) {
    if (c.n_conv < 1 || c.n_conv > GNB_MAX_LAYERS || c.n_post < 1 || c.n_post > GNB_MAX_LAYERS || c.n_readout < 1 ||
        c.n_readout > GNB_MAX_LAYERS || c.n_pool < 0 || c.n_pool > 4 || c.nb_inputs < 4 || c.nb_inputs > 32 ||
        c.n_knn_cols < 1 || c.n_knn_cols > GNB_MAX_KNN_COLS || c.k < 1 || c.k > 100)
        return GNB_ERR_UNSUPPORTED;
    if (c.globals_after_pooling && c.n_pool == 0) return GNB_ERR_ARG;
    Arena a(ws, cap);
    p.n = n; p.nseg = nseg; p.w0 = w0; p.width = c.k + 1;
    const bool split = c.precision == 2 || c.precision == 4 || c.precision == 5;   // pre-split weight operands: a lo buffer behind every packed forward weight
    if (c.precision < 0 || c.precision > 6) return GNB_ERR_ARG;
    p.bf = c.precision >= 3 ? ((c.precision == 3 || c.precision == 6) ? 1 : 2) : 0;
    p.mixed = c.precision >= 5;           // fp16 planes with power-of-two scale words
    p.agg = c.precision >= 1 && c.k == 8 && w0 == 9 && !(c.flags & 1) && (training || (c.flags & 2) || split || p.bf);
    if (p.bf && !p.agg) return GNB_ERR_UNSUPPORTED;               // the bf16 modes exist on the k = 8 tensor-core route only
    const int f = c.nb_inputs, ng = f + 5;
    const bool distribute = !c.globals_after_pooling;
    p.node_width = f + (distribute ? ng : 0);
    p.x0_ld = (int)up(p.node_width, 32);
    p.g = a.get<float>(nseg * ng);
    p.x0 = a.get<float>(n * p.x0_ld);
    p.scale_bits = p.mixed ? a.get<uint32_t>(2 * GNB_MAX_LAYERS) : nullptr;
    p.full9 = p.mixed ? a.get<int32_t>(GNB_MAX_LAYERS) : nullptr;
    // edge-slot rows of a per-edge bit-mask buffer, whole tiles of either layout (14 x 9 or 16 x 8)
    const int64_t tile_rows = ((n + 13) / 14 * 126 > (n + 15) / 16 * 128) ? (n + 13) / 14 * 126 : (n + 15) / 16 * 128;
    const int64_t max_w = w0 > p.width ? w0 : p.width;
    int max_h = 0, max_c = 0;
    for (int l = 0; l < c.n_conv; ++l) {
        if ((c.conv_hidden[l] & 3) || (c.conv_out[l] & 3) || c.conv_hidden[l] < 4 || c.conv_out[l] < 4) return GNB_ERR_UNSUPPORTED;
        if (c.conv_hidden[l] > max_h) max_h = c.conv_hidden[l];
        if (c.conv_out[l] > max_c) max_c = c.conv_out[l];
    }
    float *h_shared = nullptr, *m_shared = nullptr, *pq_shared = nullptr;
    __nv_bfloat16* hb_shared[2] = {nullptr, nullptr};
    if (!training) {   // inference: per-edge tensors are transient, share them across layers
        if (!p.bf) {
            h_shared = a.get<float>(n * max_w * max_h);
            m_shared = a.get<float>(n * max_w * max_c);
        }
        static const bool f16_inf_fused0 = !(getenv("GNB_F16_INFER_FUSED") != nullptr && atoi(getenv("GNB_F16_INFER_FUSED")) == 0);
        bool all_fused = (c.precision == 5 || (c.precision == 6 && f16_inf_fused0)) && !(c.flags & 8) && w0 == 9 && max_h <= 512 && max_c <= 256;      // (see b.fused below)
        for (int pl = 0; pl < p.bf && !all_fused; ++pl) hb_shared[pl] = a.get<__nv_bfloat16>(n * max_w * max_h);
        pq_shared = a.get<float>(n * 2 * max_h);
    }
    for (int l = 0; l < c.n_conv; ++l) {
        ConvBuf& b = p.conv[l];
        const int64_t wl = l == 0 ? w0 : p.width;
        b.cin = l == 0 ? p.node_width : c.conv_out[l - 1];
        b.cin_ld = l == 0 ? p.x0_ld : c.conv_out[l - 1];
        b.kld = (int)up(b.cin_ld, 32);
        b.hid = c.conv_hidden[l]; b.hld = (int)up(b.hid, 32); b.cout = c.conv_out[l];
        b.wcat = a.get<float>((int64_t)2 * b.hid * b.kld);
        b.wcat_lo = split ? a.get<float>((int64_t)2 * b.hid * b.kld) : nullptr;
        b.bcat = a.get<float>(2 * b.hid);
        b.w2p = a.get<float>((int64_t)b.cout * b.hld);
        b.w2p_lo = split ? a.get<float>((int64_t)b.cout * b.hld) : nullptr;
        const bool tcw = training && c.precision >= 1;          // transposed operands of the tensor-core data-gradient GEMMs
        b.wcat_t = (tcw && l > 0) ? a.get<float>((int64_t)b.kld * up(2 * b.hid, 32)) : nullptr;
        b.w2t = (tcw && c.precision <= 2) ? a.get<float>((int64_t)b.hld * up(b.cout, 32)) : nullptr;
        b.pq = training ? a.get<float>(n * 2 * b.hid) : pq_shared;
        b.hld64 = (int)up(b.hid, 64); b.cld64 = (int)up(b.cout, 64);
        // fp16-plane modes: the backward GEMMs expand dz themselves (no stored dz) where the shapes allow it, and then the
        // forward is ONE kernel (gather + hidden layer + second Linear + aggregation: h never crosses HBM except plane 0
        // as the weight gradient's operand); inference fuses whenever the shapes allow it. flags bit 3 keeps the two-kernel forward.
        b.nodz = p.mixed && training && !(c.flags & 4) && b.cout <= 256 && (b.cout & 63) == 0 && wl == 9;
        // (inference: the fused kernel wins with two planes, 5.55 against 6.27 ms per 1024 events)
        // (f16 inference, one plane: with W2's plane resident in shared memory the fused kernel moves no weights at all and wins,
        // 3.88 against 4.00 ms per 1024 events; GNB_F16_INFER_FUSED=0 keeps the two-kernel forward there)
        static const bool f16_inf_fused = !(getenv("GNB_F16_INFER_FUSED") != nullptr && atoi(getenv("GNB_F16_INFER_FUSED")) == 0);
        b.fused = p.mixed && !(c.flags & 8) && b.cout <= 256 && wl == 9 && b.hid <= 512 && (training ? b.nodz : (c.precision == 5 || f16_inf_fused));
        // the fused forward is the only reader of this layer's PQ: its columns are stored in the order the builders gather best
        // (GNB_PQ_NATURAL=1 keeps the natural order: timing comparisons)
        static const bool pq_natural = getenv("GNB_PQ_NATURAL") != nullptr && atoi(getenv("GNB_PQ_NATURAL")) != 0;
        b.pq_perm = b.fused && !pq_natural;
        // 8-slot layout of the layer's per-edge tensors when its graph holds no node with k + 1 neighbours (decided on the device per
        // step: p.full9[l]); needs every per-edge kernel of the layer to know both layouts. flags bit 4 / GNB_SLOTS9=1: always 9 slots
        static const bool slots9 = getenv("GNB_SLOTS9") != nullptr && atoi(getenv("GNB_SLOTS9")) != 0;
        b.slots8 = b.fused && (!training || b.nodz) && !(c.flags & 16) && !slots9;
        for (int pl = 0; pl < 2; ++pl) {
            const bool on = pl < p.bf;
            const bool h_on = on && !(b.fused && (pl == 1 || !training));      // fused forward: only plane 0, only for the backward pass
            b.hb[pl] = h_on ? (training ? a.get<__nv_bfloat16>(n * wl * b.hid) : hb_shared[pl]) : nullptr;
            b.w2b[pl] = on ? a.get<__nv_bfloat16>((int64_t)b.cout * b.hld64) : nullptr;
            b.w2tb[pl] = (on && training && !(p.mixed && pl == 1)) ? a.get<__nv_bfloat16>((int64_t)b.hid * b.cld64) : nullptr;
        }
        if (p.bf && ((b.hid & 7) || (b.cout & 7))) return GNB_ERR_UNSUPPORTED;
        b.h = p.bf ? nullptr : (training ? a.get<float>(n * wl * b.hid) : h_shared);
        b.m = training ? (p.agg ? nullptr : a.get<float>(n * wl * b.cout)) : m_shared;
        b.mask = (p.agg && training) ? a.get<uint32_t>((n + 13) / 14 * (int64_t)b.cout * 4) : nullptr;

        // activation bits of h for the scattering data-gradient epilogue (whole 14-node tiles, mld words per slot row)
        b.mld = 4 * ((b.hid + 127) / 128);
        const bool scat = p.agg && training && b.hid <= 512 && n * 2 * b.hid < ((int64_t)1 << 31);
        b.hmask = scat ? a.get<uint32_t>(tile_rows * (int64_t)b.mld) : nullptr;
        if (p.bf && training && !scat) return GNB_ERR_UNSUPPORTED;
        b.y = a.get<float>(n * b.cout);
        b.nbr = (l + 1 < c.n_conv) ? a.get<int32_t>(n * p.width) : nullptr;
        b.deg = (l + 1 < c.n_conv) ? a.get<int32_t>(n) : nullptr;
    }
    // post-processing: layer 0 is a K-split over [x0 | y_1 .. y_L]
    p.post_parts = c.n_conv + 1;
    int off = 0;
    for (int q = 0; q < p.post_parts; ++q) {
        p.part_k[q] = q == 0 ? p.x0_ld : c.conv_out[q - 1];
        p.part_off[q] = off;
        off += (int)up(p.part_k[q], 32);
    }
    int prev = 0;
    for (int j = 0; j < c.n_post; ++j) {
        if (c.post_out[j] & 3) return GNB_ERR_UNSUPPORTED;
        DenseBuf& d = p.post[j];
        d.n_out = c.post_out[j];
        d.k_total = j == 0 ? off : prev;
        d.kld = (int)up(d.k_total, 32);
        d.wp = a.get<float>((int64_t)d.n_out * d.kld);
        d.wp_lo = split ? a.get<float>((int64_t)d.n_out * d.kld) : nullptr;
        d.wt = nullptr;
        for (int q = 0; q <= GNB_MAX_LAYERS; ++q) d.wt_part[q] = nullptr;
        if (training && c.precision >= 1) {
            if (j > 0) d.wt = a.get<float>((int64_t)d.kld * up(d.n_out, 32));
            else for (int q = 1; q < p.post_parts; ++q) d.wt_part[q] = a.get<float>((int64_t)up(p.part_k[q], 32) * up(d.n_out, 32));
        }
        d.z = a.get<float>(n * d.n_out);
        prev = d.n_out;
    }
    p.pool_c = prev;
    p.out_rows = n;
    p.pooled = nullptr; p.parg = nullptr; p.rin = nullptr; p.rin_cols = prev; p.rin_ld = prev;
    if (!c.skip_readout) {
        if (c.n_pool > 0) {
            p.out_rows = nseg;
            p.pooled = a.get<float>(nseg * (int64_t)c.n_pool * prev);
            p.parg = a.get<int32_t>(nseg * (int64_t)c.n_pool * prev);
            p.rin_cols = c.n_pool * prev + (c.globals_after_pooling ? ng : 0);
            p.rin_ld = (int)up(p.rin_cols, 4);
            p.rin = a.get<float>(nseg * (int64_t)p.rin_ld);
        }
        int rprev = p.rin_cols;
        for (int j = 0; j < c.n_readout; ++j) {
            if (c.readout_out[j] & 3) return GNB_ERR_UNSUPPORTED;
            DenseBuf& d = p.ro[j];
            d.n_out = c.readout_out[j];
            d.k_total = rprev;
            d.kld = (int)up(rprev, 32);
            d.wp = a.get<float>((int64_t)d.n_out * d.kld);
            d.wp_lo = split ? a.get<float>((int64_t)d.n_out * d.kld) : nullptr;
            d.wt = (training && c.precision >= 1) ? a.get<float>((int64_t)d.kld * up(d.n_out, 32)) : nullptr;
            d.z = a.get<float>(p.out_rows * d.n_out);
            rprev = d.n_out;
        }
    }
    if (training) {   // backward scratch
        for (int l = 0; l <= c.n_conv; ++l) p.gnode[l] = l == 0 ? nullptr : a.get<float>(n * c.conv_out[l - 1]);
        p.dz_big = p.bf ? nullptr : a.get<float>(n * max_w * max_c);
        p.dh_big = p.bf ? nullptr : a.get<float>(n * max_w * max_h);
        bool any_dz = false, any_nodz = false;
        for (int l = 0; l < c.n_conv; ++l) { any_dz = any_dz || !p.conv[l].nodz; any_nodz = any_nodz || p.conv[l].nodz; }
        for (int pl = 0; pl < 2; ++pl)
            p.dzb[pl] = (pl < p.bf && !(p.mixed && pl == 1) && any_dz) ? a.get<__nv_bfloat16>(n * max_w * max_c) : nullptr;
        p.g16 = any_nodz ? a.get<__nv_bfloat16>(n * (int64_t)max_c) : nullptr;
        p.rowmask = any_nodz ? a.get<uint32_t>(tile_rows * (int64_t)(max_c / 32)) : nullptr;
        p.dpq = a.get<float>(n * 2 * max_h);
        p.dzq = a.get<float>(n * 2 * max_h);
        int64_t max_wp = 0, max_dense = 0;
        for (int l = 0; l < c.n_conv; ++l) {
            int64_t s1 = (int64_t)2 * p.conv[l].hid * p.conv[l].kld, s2 = (int64_t)p.conv[l].cout * p.conv[l].hld;
            int64_t t1 = (int64_t)p.conv[l].kld * up(2 * p.conv[l].hid, 32), t2 = (int64_t)p.conv[l].hld * up(p.conv[l].cout, 32);
            max_wp = s1 > max_wp ? s1 : max_wp; max_wp = s2 > max_wp ? s2 : max_wp;
            max_wp = t1 > max_wp ? t1 : max_wp; max_wp = t2 > max_wp ? t2 : max_wp;
        }
        for (int j = 0; j < c.n_post; ++j) {
            int64_t s = (int64_t)p.post[j].n_out * p.post[j].kld, t = (int64_t)p.post[j].kld * up(p.post[j].n_out, 32);
            max_wp = s > max_wp ? s : max_wp; max_wp = t > max_wp ? t : max_wp;
            if (n * p.post[j].n_out > max_dense) max_dense = n * p.post[j].n_out;
            if (n * (int64_t)p.post[j].k_total > max_dense && j > 0) max_dense = n * (int64_t)p.post[j].k_total;
        }
        if (!c.skip_readout)
            for (int j = 0; j < c.n_readout; ++j) {
                int64_t s = (int64_t)p.ro[j].n_out * p.ro[j].kld, t = (int64_t)p.ro[j].kld * up(p.ro[j].n_out, 32);
                max_wp = s > max_wp ? s : max_wp; max_wp = t > max_wp ? t : max_wp;
                if (p.out_rows * p.ro[j].n_out > max_dense) max_dense = p.out_rows * p.ro[j].n_out;
                if (p.out_rows * (int64_t)up(p.ro[j].k_total, 4) > max_dense) max_dense = p.out_rows * (int64_t)up(p.ro[j].k_total, 4);
            }
        p.dwp = a.get<float>(max_wp);
        p.wt = a.get<float>(max_wp);
        p.dbtmp = a.get<float>(2 * max_h > 4096 ? 2 * max_h : 4096);
        p.gro_a = a.get<float>(max_dense);
        p.gro_b = a.get<float>(max_dense);
    }
    p.bytes = a.off;
    if (ws != nullptr && a.off > cap) return GNB_ERR_ARG;
    return GNB_OK;
}

struct Exec {
    const gnb_dynedge_config& c;
    cudaStream_t st;
    bool tf32;        // tensor-core GEMMs (precision 1 and 2)
    bool split;       // precision 2: forward GEMMs on split operands (activations stay plain fp32, weights come as hi + lo)
    bool fround;      // forward activations are stored rounded to tf32 (precision 1 only)
    int rnd;          // backward: gradients that feed a tensor-core GEMM are stored rounded
    int frnd;         // forward flag for the producers of GEMM operands
    Exec(const gnb_dynedge_config& cfg, void* s)
        : c(cfg), st((cudaStream_t)s), tf32(cfg.precision >= 1), split(cfg.precision == 2 || cfg.precision == 4 || cfg.precision == 5),
          fround(cfg.precision == 1 || cfg.precision == 3 || cfg.precision == 6), rnd(cfg.precision >= 1 ? GNB_FLAG_ROUND_TF32 : 0),
          frnd((cfg.precision == 1 || cfg.precision == 3 || cfg.precision == 6) ? GNB_FLAG_ROUND_TF32 : 0) {
        jobs.n = 0; jobs.total_blocks = 0;
    }

    // ---- weight re-packing jobs (one launch per PACK_MAX_JOBS jobs) --------------------------------------------------------
    PackJobs jobs;
    void job_add(int kind, const float* src, const float* src2, void* dst, void* dst2, void* dst3, int64_t lds, int64_t ldd,
                 int rows, int cols, int dst_cols, int flags, int64_t total, int* rc) {
        if (total <= 0 || *rc != 0) return;
        if (jobs.n == PACK_MAX_JOBS) *rc = job_flush();
        PackJob& j = jobs.j[jobs.n++];
        j.src = src; j.src2 = src2; j.dst = dst; j.dst2 = dst2; j.dst3 = dst3; j.lds = lds; j.ldd = ldd;
        j.rows = rows; j.cols = cols; j.dst_cols = dst_cols; j.flags = flags; j.kind = kind; j.first_block = jobs.total_blocks;
        jobs.total_blocks += (unsigned)gnb_div_up(total, 256);
    }
    int job_flush() {
        if (jobs.n == 0) return 0;
        gnb_launch(pack_jobs_kernel, jobs.total_blocks, 256, 0, st)(jobs);
        jobs.n = 0; jobs.total_blocks = 0;
        EXL(); return 0;
    }
    void job_copy_pad(const float* src, int64_t lds, int64_t rows, int cols, float* dst, int64_t ldd, int dst_cols, bool round,
                      float* lo, int* rc) {
        job_add(JOB_COPY_PAD, src, nullptr, dst, lo, nullptr, lds, ldd, (int)rows, cols, dst_cols, round ? 1 : 0, rows * dst_cols, rc);
    }
    // dst[k, nld] = (wp + off)[n_out, k]^T for lin_bwd_data
    void job_transpose(const float* wp, int64_t ldw, int off, int k, int n_out, float* dst, int* rc, bool perm = false) {
        const int nld = (int)up(n_out, 32);
        job_add(JOB_TRANSPOSE_PAD, wp + off, nullptr, dst, nullptr, nullptr, ldw, nld, n_out, k, nld, (tf32 ? 1 : 0) | (perm ? 8 : 0),
                (int64_t)k * nld, rc);
    }
    void job_planes(const float* src, int64_t lds, int rows, int cols, void* p0, void* p1, int64_t ldd, int dst_cols, bool transpose,
                    bool bf16, int* rc) {
        job_add(JOB_PLANES, src, nullptr, p0, p1, nullptr, lds, ldd, rows, cols, dst_cols, (transpose ? 2 : 0) | (bf16 ? 4 : 0),
                (int64_t)(transpose ? cols : rows) * dst_cols, rc);
    }

    int copy_pad(const float* src, int64_t lds, int64_t rows, int cols, float* dst, int64_t ldd, int dst_cols, bool round,
                 float* lo = nullptr) {
        if (rows * dst_cols == 0) return 0;
        gnb_launch(copy_pad_kernel, gnb_div_up(rows * dst_cols, 256), 256, 0, st)(src, lds, rows, cols, dst, ldd, dst_cols, round ? 1 : 0, lo);
        EXL(); return 0;
    }
    int add2d(const float* src, int64_t lds, int64_t rows, int cols, float* dst, int64_t ldd) {
        if (rows * cols == 0) return 0;
        gnb_launch(add2d_kernel, gnb_div_up(rows * cols, 256), 256, 0, st)(src, lds, rows, cols, dst, ldd);
        EXL(); return 0;
    }
    // y = act(sum_p x_p wp[:, off_p : off_p + k_p]^T + b)
    // round_out: round y to tf32 -- only needed when y itself feeds a tensor-core GEMM
    int lin_fwd(int nparts, const float* const* xs, const int64_t* lds, const int32_t* ks, const int* offs, const float* wp,
                int64_t ldw, const float* bias, float* y, int64_t rows, int n_out, int act, int round_out = 1,
                const float* wp_lo = nullptr) {
        if (rows == 0) return 0;
        if (split) return gnb_linear_fwd_tf32x3(xs, lds, ks, nparts, wp, wp_lo, ldw, bias, y, n_out, rows, n_out, act, st);
        if (tf32) return gnb_linear_fwd_tf32(xs, lds, ks, nparts, wp, ldw, bias, y, n_out, rows, n_out, act, round_out, st);
        for (int q = 0; q < nparts; ++q)
            EX(gnb_linear_fwd_f32(xs[q], lds[q], wp + offs[q], ldw, q == nparts - 1 ? bias : nullptr, y, n_out, rows, n_out,
                                  ks[q], q == nparts - 1 ? act : GNB_ACT_NONE, q > 0 ? 1 : 0, st));
        return 0;
    }
    // dx[rows, k] (+)= dz[rows, n_out] wp[:, off : off + k]     (wt = W^T [k, up(n_out, 32)], packed by the forward pass, tensor-core modes)
    int lin_bwd_data(const float* dz, int64_t lddz, const float* wp, int64_t ldw, int off, int k, int n_out, float* dx,
                     int64_t lddx, int64_t rows, bool accumulate, const float* wt) {
        if (rows == 0) return 0;
        if (!tf32) return gnb_linear_bwd_data_f32(dz, lddz, wp + off, ldw, dx, lddx, rows, n_out, k, accumulate ? 1 : 0, st);
        if (wt == nullptr) return GNB_ERR_ARG;
        const int nld = (int)up(n_out, 32);
        const float* xs[1] = {dz}; const int64_t l1[1] = {lddz}; const int32_t k1[1] = {n_out};
        // dx += dz W is accumulated in the GEMM epilogue (flag 0x200), no temporary
        return gnb_linear_fwd_tf32(xs, l1, k1, 1, wt, nld, nullptr, dx, lddx, rows, k,
                                   GNB_ACT_NONE | (accumulate ? GNB_FLAG_ACCUMULATE : 0), 0, st);
    }
    // dwp[:, off : off + k] += dz^T x     (dwp zeroed by the caller)
    int lin_bwd_weight(const float* dz, int64_t lddz, const float* x, int64_t ldx, float* dwp, int64_t ldw, int off, int k,
                       int n_out, int64_t rows) {
        if (rows == 0) return 0;
        if (tf32) return gnb_linear_bwd_weight_tf32(dz, lddz, x, ldx, dwp + off, ldw, rows, n_out, k, 0, st);
        return gnb_linear_bwd_weight_f32(dz, lddz, x, ldx, dwp + off, ldw, rows, n_out, k, st);
    }
};

}  // namespace

GNB_EXPORT int64_t gnb_dynedge_workspace_bytes(const gnb_dynedge_config* cfg, int64_t n, int64_t nseg, int32_t w0,
                                               int32_t training) {
    Plan p;
    int rc = make_plan(*cfg, n, nseg, w0, training != 0, nullptr, 0, p);
    return rc == 0 ? p.bytes : (int64_t)rc;
}

GNB_EXPORT int gnb_dynedge_layout(const gnb_dynedge_config* cfg, int64_t n, int64_t nseg, int32_t w0, int32_t training,
                                  int64_t* offsets) {
    Plan p;
    char* base = reinterpret_cast<char*>(0x1000);     // fake base: only the offsets are of interest
    int rc = make_plan(*cfg, n, nseg, w0, training != 0, base, (int64_t)1 << 60, p);
    if (rc != 0) return rc;
    for (int l = 0; l < cfg->n_conv; ++l) {
        offsets[3 * l + 0] = (char*)p.conv[l].y - base;
        offsets[3 * l + 1] = p.conv[l].nbr ? (char*)p.conv[l].nbr - base : -1;
        offsets[3 * l + 2] = p.conv[l].deg ? (char*)p.conv[l].deg - base : -1;
    }
    return GNB_OK;
}

// params: device pointers in state_dict order: per conv {W1, b1, W2, b2}, per post layer {W, b}, per read-out
// layer {W, b}. out: [nseg or n, last width]. n_pulses: fp32[nseg].
GNB_EXPORT int gnb_dynedge_forward(const gnb_dynedge_config* cfg, const float* const* params, const float* x, int64_t ldx,
                                   const int64_t* ptr, const float* n_pulses, const int32_t* nbr0, const int32_t* deg0,
                                   int32_t w0, const int32_t* knn_cols_dev, int64_t n, int64_t nseg, void* workspace,
                                   int64_t workspace_bytes, float* out, int32_t training, void* stream) {
    const gnb_dynedge_config& c = *cfg;
    Plan p;
    EX(make_plan(c, n, nseg, w0, (training & 1) != 0, workspace, workspace_bytes, p));
    Exec e(c, stream);
    const int f = c.nb_inputs;
    const bool distribute = !c.globals_after_pooling;
    const bool fused_edge = (training & 2) == 0 && !(c.flags & 1);   // fused inference EdgeConv kernel
    training &= 1;
    {   // every packed weight operand of the step (and, in training, of its backward pass): one launch
        int rc = 0, pj = 0;
        for (int l = 0; l < c.n_conv; ++l, pj += 4) {
            ConvBuf& b = p.conv[l];
            const float *w1 = params[pj], *b1 = params[pj + 1], *w2 = params[pj + 2];
            e.job_add(JOB_PACK_CONV, w1, b1, b.wcat, b.bcat, b.wcat_lo, 0, b.kld, b.hid, b.cin, 0, (e.tf32 ? 1 : 0) | (b.pq_perm ? 8 : 0), 2 * (int64_t)b.hid * b.kld, &rc);
            if (!p.bf) e.job_copy_pad(w2, b.hid, b.cout, b.hid, b.w2p, b.hld, b.hld, e.tf32, b.w2p_lo, &rc);
            if (p.bf) {
                const bool bf16 = !p.mixed;
                e.job_planes(w2, b.hid, b.cout, b.hid, b.w2b[0], b.w2b[1], b.hld64, b.hld64, false, bf16, &rc);
                if (training & 1) e.job_planes(w2, b.hid, b.cout, b.hid, b.w2tb[0], b.w2tb[1], b.cld64, b.cld64, true, bf16, &rc);
            }
        }
        {
            DenseBuf& d = p.post[0];
            const float* w = params[pj];
            int src_col = 0;
            int64_t src_ld = p.node_width;
            for (int l = 0; l < c.n_conv; ++l) src_ld += c.conv_out[l];
            for (int q = 0; q < p.post_parts; ++q) {
                const int valid = q == 0 ? p.node_width : c.conv_out[q - 1];
                e.job_copy_pad(w + src_col, src_ld, d.n_out, valid, d.wp + p.part_off[q], d.kld, (int)up(p.part_k[q], 32), e.tf32,
                               d.wp_lo ? d.wp_lo + p.part_off[q] : nullptr, &rc);
                src_col += valid;
            }
            pj += 2;
        }
        for (int j = 1; j < c.n_post; ++j, pj += 2) {
            DenseBuf& d = p.post[j];
            e.job_copy_pad(params[pj], d.k_total, d.n_out, d.k_total, d.wp, d.kld, d.kld, e.tf32, d.wp_lo, &rc);
        }
        if (!c.skip_readout)
            for (int j = 0; j < c.n_readout; ++j, pj += 2) {
                DenseBuf& d = p.ro[j];
                e.job_copy_pad(params[pj], d.k_total, d.n_out, d.k_total, d.wp, d.kld, d.kld, e.tf32, d.wp_lo, &rc);
            }
        EX(rc);
        EX(e.job_flush());      // (the transposes below read the packed buffers written above)
        if ((training & 1) && e.tf32) {
            for (int l = 0; l < c.n_conv; ++l) {
                ConvBuf& b = p.conv[l];
                if (b.wcat_t != nullptr) e.job_transpose(b.wcat, b.kld, 0, b.cin_ld, 2 * b.hid, b.wcat_t, &rc, b.pq_perm);
                if (b.w2t != nullptr) e.job_transpose(b.w2p, b.hld, 0, b.hid, b.cout, b.w2t, &rc);
            }
            for (int q = 1; q < p.post_parts; ++q)
                e.job_transpose(p.post[0].wp, p.post[0].kld, p.part_off[q], p.part_k[q], p.post[0].n_out, p.post[0].wt_part[q], &rc);
            for (int j = 1; j < c.n_post; ++j)
                e.job_transpose(p.post[j].wp, p.post[j].kld, 0, p.post[j].k_total, p.post[j].n_out, p.post[j].wt, &rc);
            if (!c.skip_readout)
                for (int j = 0; j < c.n_readout; ++j)
                    e.job_transpose(p.ro[j].wp, p.ro[j].kld, 0, p.ro[j].k_total, p.ro[j].n_out, p.ro[j].wt, &rc);
            EX(rc);
            EX(e.job_flush());
        }
    }
    // global variables (+ x0 = [x | g[batch] | 0])
    EX(gnb_global_vars(x, ldx, f, nbr0, deg0, w0, ptr, nseg, n_pulses, p.g, distribute ? p.x0 : nullptr, p.x0_ld, stream));
    if (!distribute) EX(e.copy_pad(x, ldx, n, f, p.x0, p.x0_ld, p.x0_ld, e.fround));
    else if (e.fround) EX(e.copy_pad(p.x0, p.x0_ld, n, p.x0_ld, p.x0, p.x0_ld, p.x0_ld, true));
    // DynEdgeConv layers
    const float* xin = p.x0;
    const int32_t *nbr = nbr0, *deg = deg0;
    int wl = w0;
    int pi = 0;
    for (int l = 0; l < c.n_conv; ++l) {
        ConvBuf& b = p.conv[l];
        const float* b2 = params[pi + 3];      // (W1, b1, W2 were packed by the job launch above)
        pi += 4;
        {   // PQ = xin Wcat^T + bcat
            if (p.mixed) {      // fp16-plane modes: the GEMM epilogue also yields the scale word of h = relu(P_i + Q_j) <= 2 max|PQ|
                // (one memset: the scale words and, right behind them in the arena, the layers' edge-slot flag words)
                if (l == 0) GNB_CHECK(cudaMemsetAsync(p.scale_bits, 0, (size_t)((char*)(p.full9 + GNB_MAX_LAYERS) - (char*)p.scale_bits), e.st));
                EX(gnb_linear_next_absmax(p.scale_bits + l, 1));
            }
            const float* xs[1] = {xin}; const int64_t lds[1] = {b.cin_ld}; const int32_t ks[1] = {b.cin_ld}; const int offs[1] = {0};
            EX(e.lin_fwd(1, xs, lds, ks, offs, b.wcat, b.kld, b.bcat, b.pq, n, 2 * b.hid, GNB_ACT_NONE, 0, b.wcat_lo));   // P+Q is added in fp32
        }
        if (!training && e.tf32 && !e.split && fused_edge && !p.agg && b.hid <= 352 && wl <= 32) {
            // inference: gather + hidden ReLU + E x H x C contraction + bias/ReLU + aggregation in one tcgen05 kernel
            EX(gnb_edgeconv_fused_fwd_tf32(b.pq, 2 * b.hid, b.hid, nbr, deg, wl, n, b.w2p, b.hld, b2, b.cout, GNB_AGGR_ADD, 1,
                                           b.y, b.cout, stream));
        } else if (p.bf) {
            // bf16 / bf16x3: h as bf16 plane(s) straight from the hidden-layer kernel, second Linear + ReLU + k-sum on kind::f16
            if (p.mixed) {
                uint32_t* hs = p.scale_bits + l;          // written by the PQ GEMM's epilogue (gnb_linear_next_absmax above)
                if (b.fused) {
                    static const int dbgf = getenv("GNB_FUSED_DBG") ? atoi(getenv("GNB_FUSED_DBG")) : 0;
                    if (b.slots8) EX(gnb_edge_slot_flag_or(deg, n, c.k, p.full9 + l, stream));      // (the word was zeroed with the scale words)
                    EX(gnb_edgeconv_fused_fwd_f16_w(b.pq, 2 * b.hid, b.hid, nbr, deg, n, b.w2b[0], b.w2b[1], b.hld64, b2, b.cout,
                                                    e.fround ? 1 : 0, b.y, b.cout, (dbgf & 4) ? nullptr : b.mask,
                                                    (training && !(dbgf & 1)) ? (void*)b.hb[0] : nullptr, b.hid,
                                                    (training && !(dbgf & 2)) ? (uint8_t*)b.hmask : nullptr, (int64_t)b.mld * 4, hs,
                                                    b.pq_perm ? 1 : 0, b.slots8 ? p.full9 + l : nullptr, stream));
                } else {
                    EX(gnb_edge_hidden_fwd_f16(b.pq, 2 * b.hid, b.hid, nbr, deg, wl, n, b.hb[0], b.hb[1], b.hid, b.hmask, b.mld, hs, stream));
                    EX(gnb_edge_linear_agg_fwd_f16(b.hb[0], b.hb[1], b.hid, b.hid, b.w2b[0], b.w2b[1], b.hld64, b2, deg, n, b.cout,
                                                   e.fround ? 1 : 0, b.y, b.cout, b.mask, hs, stream));
                }
            } else {
                EX(gnb_edge_hidden_fwd_bf16(b.pq, 2 * b.hid, b.hid, nbr, deg, wl, n, b.hb[0], b.hb[1], b.hid, b.hmask, b.mld, stream));
                EX(gnb_edge_linear_agg_fwd_bf16(b.hb[0], b.hb[1], b.hid, b.hid, b.w2b[0], b.w2b[1], b.hld64, b2, deg, n, b.cout,
                                                e.fround ? 1 : 0, b.y, b.cout, b.mask, stream));
            }
        } else if (p.agg) {
            // training (and inference with flags bit 1): the second Linear, ReLU and the k-sum run in one tcgen05 kernel; h is kept for the backward pass;
            // whose epilogue writes y and one ReLU bit per (slot, channel) -- the [E, C] message tensor is never stored
            // (tf32x3: h, y stay plain fp32 -- the split happens inside the GEMMs; the weight gradient later truncates h)
            if (b.hmask != nullptr)
                EX(gnb_edge_hidden_fwd_mask(b.pq, 2 * b.hid, b.hid, nbr, deg, wl, n, GNB_ACT_RELU | e.frnd, b.h, b.hid, b.hmask,
                                            b.mld, stream));
            else
                EX(gnb_edge_hidden_fwd(b.pq, 2 * b.hid, b.hid, nbr, deg, wl, n, GNB_ACT_RELU | e.frnd, b.h, b.hid, stream));
            if (e.split)
                EX(gnb_edge_linear_agg_fwd_tf32x3(b.h, b.hid, b.hid, b.w2p, b.w2p_lo, b.hld, b2, deg, n, b.cout, b.y, b.cout, b.mask,
                                                  stream));
            else
                EX(gnb_edge_linear_agg_fwd_tf32(b.h, b.hid, b.hid, b.w2p, b.hld, b2, deg, n, b.cout, 1, b.y, b.cout, b.mask, stream));
        } else {
            EX(gnb_edge_hidden_fwd(b.pq, 2 * b.hid, b.hid, nbr, deg, wl, n, GNB_ACT_RELU | e.frnd, b.h, b.hid, stream));
            {   // m = relu(h W2^T + b2)
                const float* xs[1] = {b.h}; const int64_t lds[1] = {b.hid}; const int32_t ks[1] = {b.hid}; const int offs[1] = {0};
                EX(e.lin_fwd(1, xs, lds, ks, offs, b.w2p, b.hld, b2, b.m, n * wl, b.cout, GNB_ACT_RELU, 0, b.w2p_lo));   // summed in fp32
            }
            EX(gnb_edge_aggregate_fwd(b.m, b.cout, b.cout, deg, wl, n, GNB_AGGR_ADD | e.frnd, b.y, b.cout, nullptr, stream));
        }
        if (l + 1 < c.n_conv) {
            EX(gnb_knn_table(b.y, b.cout, knn_cols_dev, c.n_knn_cols, ptr, nseg, n, c.k, b.nbr, b.deg, stream));
            nbr = b.nbr; deg = b.deg; wl = p.width;
        }
        xin = b.y;
    }
    // post-processing
    const float* zin = nullptr;
    for (int j = 0; j < c.n_post; ++j) {
        DenseBuf& d = p.post[j];
        const float* bias = params[pi + 1];
        pi += 2;
        if (j == 0) {
            const float* xs[GNB_MAX_LAYERS + 1]; int64_t lds[GNB_MAX_LAYERS + 1]; int32_t ks[GNB_MAX_LAYERS + 1];
            for (int q = 0; q < p.post_parts; ++q) { xs[q] = q == 0 ? p.x0 : p.conv[q - 1].y; lds[q] = p.part_k[q]; ks[q] = p.part_k[q]; }
            EX(e.lin_fwd(p.post_parts, xs, lds, ks, p.part_off, d.wp, d.kld, bias, d.z, n, d.n_out, GNB_ACT_RELU, 1, d.wp_lo));
        } else {
            const float* xs[1] = {zin}; const int64_t lds[1] = {d.k_total}; const int32_t ks[1] = {d.k_total}; const int offs[1] = {0};
            EX(e.lin_fwd(1, xs, lds, ks, offs, d.wp, d.kld, bias, d.z, n, d.n_out, GNB_ACT_RELU, 1, d.wp_lo));
        }
        zin = d.z;
    }
    const int last_post = p.post[c.n_post - 1].n_out;
    if (c.skip_readout) return e.copy_pad(zin, last_post, n, last_post, out, last_post, last_post, false);
    // pooling + read-out
    const float* rin = zin;
    int64_t rin_ld = last_post;
    if (c.n_pool > 0) {
        EX(gnb_segment_pool_fwd(zin, last_post, last_post, ptr, nseg, c.pool, c.n_pool, p.pooled, p.parg, stream));
        const int pc = c.n_pool * last_post;
        EX(e.copy_pad(p.pooled, pc, nseg, pc, p.rin, p.rin_ld, c.globals_after_pooling ? pc : p.rin_ld, e.fround));
        if (c.globals_after_pooling)
            EX(e.copy_pad(p.g, f + 5, nseg, f + 5, p.rin + pc, p.rin_ld, p.rin_ld - pc, e.fround));
        rin = p.rin; rin_ld = p.rin_ld;
    }
    for (int j = 0; j < c.n_readout; ++j) {
        DenseBuf& d = p.ro[j];
        const float* bias = params[pi + 1];
        pi += 2;
        const float* xs[1] = {rin}; const int64_t lds[1] = {rin_ld}; const int32_t ks[1] = {d.k_total}; const int offs[1] = {0};
        EX(e.lin_fwd(1, xs, lds, ks, offs, d.wp, d.kld, bias, d.z, p.out_rows, d.n_out, GNB_ACT_RELU, 1, d.wp_lo));
        rin = d.z; rin_ld = d.n_out;
    }
    const int last = p.ro[c.n_readout - 1].n_out;
    return e.copy_pad(rin, last, p.out_rows, last, out, last, last, false);
}

// grads: device pointers parallel to params; gradients are ACCUMULATED onto them. gout: [out_rows, last width].
// Must follow a forward with training = 1 on the same workspace.
GNB_EXPORT int gnb_dynedge_backward(const gnb_dynedge_config* cfg, float* const* grads, const int64_t* ptr,
                                    const int32_t* nbr0, const int32_t* deg0, int32_t w0, int64_t n, int64_t nseg,
                                    void* workspace, int64_t workspace_bytes, const float* gout, void* stream) {
    const gnb_dynedge_config& c = *cfg;
    Plan p;
    EX(make_plan(c, n, nseg, w0, true, workspace, workspace_bytes, p));
    Exec e(c, stream);
    const int n_params = 4 * c.n_conv + 2 * c.n_post + (c.skip_readout ? 0 : 2 * c.n_readout);
    int pi = n_params;
    const float* gcur = gout;
    int64_t gcur_ld = 0;
    float* gz = p.gro_a;      // ping-pong buffers for node/event level gradients
    float* gz2 = p.gro_b;
    const int last_post = p.post[c.n_post - 1].n_out;
    // ---- read-out chain --------------------------------------------------------------------------
    if (!c.skip_readout) {
        gcur_ld = p.ro[c.n_readout - 1].n_out;
        for (int j = c.n_readout - 1; j >= 0; --j) {
            DenseBuf& d = p.ro[j];
            pi -= 2;
            float *gw = grads[pi], *gb = grads[pi + 1];
            const float* xin = j == 0 ? (c.n_pool > 0 ? p.rin : p.post[c.n_post - 1].z) : p.ro[j - 1].z;
            const int64_t xin_ld = j == 0 ? (c.n_pool > 0 ? p.rin_ld : last_post) : p.ro[j - 1].n_out;
            float* dz = (gcur == gz) ? gz2 : gz;      // [out_rows, n_out]; never the buffer holding gcur
            float* dx = (dz == gz) ? gz2 : gz;
            EX(gnb_act_bwd_colsum(gcur, gcur_ld, d.z, d.n_out, p.out_rows, d.n_out, dz, d.n_out, gb, GNB_ACT_RELU | e.rnd,
                                  nullptr, 1, 0, stream));
            // weight gradients accumulate straight into the caller's gradient buffer (both GEMM back ends add)
            EX(e.lin_bwd_weight(dz, d.n_out, xin, xin_ld, gw, d.k_total, 0, d.k_total, d.n_out, p.out_rows));
            const int64_t dx_ld = up(d.k_total, 4);
            EX(e.lin_bwd_data(dz, d.n_out, d.wp, d.kld, 0, d.k_total, d.n_out, dx, dx_ld, p.out_rows, false, d.wt));
            gcur = dx; gcur_ld = dx_ld;
        }
        if (c.n_pool > 0) {   // gradient of the pooled block -> nodes (the appended global variables carry no gradient)
            float* o = (gcur == gz) ? gz2 : gz;
            EX(gnb_segment_pool_bwd(gcur, gcur_ld, p.parg, last_post, ptr, nseg, n, c.pool, c.n_pool, o, last_post, stream));
            gcur = o; gcur_ld = last_post;
        }
    } else {
        gcur_ld = last_post;
    }
    // ---- post-processing chain -------------------------------------------------------------------
    for (int j = c.n_post - 1; j >= 0; --j) {
        DenseBuf& d = p.post[j];
        pi -= 2;
        float *gw = grads[pi], *gb = grads[pi + 1];
        float* dz = (gcur == gz) ? gz2 : gz;      // a buffer that is not the current gradient
        EX(gnb_act_bwd_colsum(gcur, gcur_ld, d.z, d.n_out, n, d.n_out, dz, d.n_out, gb, GNB_ACT_RELU | e.rnd, nullptr, 1, 0, stream));
        if (j > 0) {
            EX(e.lin_bwd_weight(dz, d.n_out, p.post[j - 1].z, p.post[j - 1].n_out, gw, d.k_total, 0, d.k_total, d.n_out, n));
            float* dx = (dz == gz) ? gz2 : gz;
            EX(e.lin_bwd_data(dz, d.n_out, d.wp, d.kld, 0, d.k_total, d.n_out, dx, d.k_total, n, false, d.wt));
            gcur = dx; gcur_ld = d.k_total;
        } else {
            int dst_col = 0;
            int64_t dst_ld = p.node_width;
            for (int l = 0; l < c.n_conv; ++l) dst_ld += c.conv_out[l];
            for (int q = 0; q < p.post_parts; ++q) {
                const int valid = q == 0 ? p.node_width : c.conv_out[q - 1];
                const float* xq = q == 0 ? p.x0 : p.conv[q - 1].y;
                EX(e.lin_bwd_weight(dz, d.n_out, xq, p.part_k[q], gw, dst_ld, dst_col, valid, d.n_out, n));
                dst_col += valid;
                // fp16-plane modes: the GEMM that writes the FINAL value of a conv layer's output gradient also yields its
                // max|g| (the scale word of dz): here for the last conv layer, in the conv loop below for the others
                if (q == c.n_conv && p.mixed && e.tf32) EX(gnb_linear_next_absmax(p.scale_bits + GNB_MAX_LAYERS + (q - 1), 0));
                if (q > 0)   // gradient w.r.t. the output of conv q-1 (x0 carries none)
                    EX(e.lin_bwd_data(dz, d.n_out, d.wp, d.kld, p.part_off[q], p.part_k[q], d.n_out, p.gnode[q], p.part_k[q], n,
                                      false, d.wt_part[q]));
            }
        }
    }
    // ---- DynEdgeConv layers, last to first -------------------------------------------------------
    for (int l = c.n_conv - 1; l >= 0; --l) {
        ConvBuf& b = p.conv[l];
        pi -= 4;
        float *gw1 = grads[pi], *gb1 = grads[pi + 1], *gw2 = grads[pi + 2], *gb2 = grads[pi + 3];
        const int32_t* nbr = l == 0 ? nbr0 : p.conv[l - 1].nbr;
        const int32_t* deg = l == 0 ? deg0 : p.conv[l - 1].deg;
        const int wl = l == 0 ? w0 : p.width;
        const int64_t rows = n * wl;
        const float* gy = p.gnode[l + 1];
        // fp16-plane modes: dz is expanded inside the two GEMMs instead of being stored (flags bit 2 keeps the stored-dz route)
        const bool nodz = b.nodz;
        // (aggregate-bwd + ReLU-bwd + bias grad) in one pass
        if (p.mixed) {
            uint32_t* gs = p.scale_bits + GNB_MAX_LAYERS + l;      // zeroed by the forward pass, filled by the epilogue of the GEMM
                                                                   // that wrote the last contribution to gy (gnb_linear_next_absmax)
            if (!e.tf32) EX(gnb_absmax_bits(gy, b.cout, n, b.cout, 0, gs, stream));
            if (nodz) {
                // dz = g * mask bit is never stored: both GEMMs expand it in shared memory from g_y and the mask words
                // (one small pass writes fp16(g_y 2^s), the row-major bits and the bias gradient)
                const int32_t* f9 = b.slots8 ? p.full9 + l : nullptr;          // written by the forward pass
                // (the launch also zeroes the Q half of dPQ and the bias-gradient scratch for the scattering kernel below)
                EX(gnb_edge_dz_prep_wz(gy, b.cout, b.mask, n, b.cout, gs, p.g16, p.rowmask, gb2, f9, p.dzq + b.hid, 2 * b.hid, n, b.hid,
                                       p.dbtmp, 2 * b.hid, stream));
                EX(gnb_linear_bwd_weight_f16_masked_w(p.g16, p.rowmask, b.hb[0], b.hid, gw2, b.hid, n, b.cout, b.hid, gs, p.scale_bits + l, f9, stream));
            } else {
                EX(gnb_edge_mask_bwd_colsum_f16(gy, b.cout, b.mask, n, b.cout, p.dzb[0], b.cout, gb2, gs, stream));
                EX(gnb_linear_bwd_weight_f16(p.dzb[0], b.cout, b.hb[0], nullptr, b.hid, gw2, b.hid, rows, b.cout, b.hid, gs, p.scale_bits + l, stream));
            }
        } else if (p.bf) {
            EX(gnb_edge_mask_bwd_colsum_bf16(gy, b.cout, b.mask, n, b.cout, p.dzb[0], p.dzb[1], b.cout, gb2, stream));
            EX(gnb_linear_bwd_weight_bf16(p.dzb[0], p.dzb[1], b.cout, b.hb[0], b.hb[1], b.hid, gw2, b.hid, rows, b.cout, b.hid, 0, stream));
        } else if (p.agg)
            EX(gnb_edge_mask_bwd_colsum(gy, b.cout, b.mask, n, b.cout, deg, p.dz_big, b.cout, gb2, e.rnd, stream));
        else
            EX(gnb_act_bwd_colsum(gy, b.cout, b.m, b.cout, rows, b.cout, p.dz_big, b.cout, gb2, GNB_ACT_RELU | e.rnd, deg, wl,
                                  GNB_AGGR_ADD, stream));
        if (!p.bf) EX(e.lin_bwd_weight(p.dz_big, b.cout, b.h, b.hid, gw2, b.hid, 0, b.hid, b.cout, rows));
        const bool zeroed = p.mixed && nodz;          // by the dz prep launch above
        if (!zeroed) GNB_CHECK(cudaMemsetAsync(p.dbtmp, 0, (size_t)2 * b.hid * 4, e.st));
        const float* dzq = p.dzq;
        if (p.bf) {
            if (!zeroed) EX(gnb_zero_block(p.dzq + b.hid, 2 * b.hid, n, b.hid, stream));
            if (p.mixed && nodz)
                EX(gnb_edge_hidden_dgrad_scatter_f16_masked_w(p.g16, p.rowmask, b.cout, b.w2tb[0], b.cld64, b.hmask, b.mld, b.hid, nbr, n,
                                                              p.dzq + b.hid, 2 * b.hid, p.dzq, 2 * b.hid, p.dbtmp,
                                                              e.rnd | (b.fused ? GNB_FLAG_HMASK_ROWMAJOR : 0),
                                                              p.scale_bits + GNB_MAX_LAYERS + l, b.slots8 ? p.full9 + l : nullptr, stream));
            else if (p.mixed)
                EX(gnb_edge_hidden_dgrad_scatter_f16(p.dzb[0], b.cout, b.cout, b.w2tb[0], b.cld64, b.hmask, b.mld, b.hid, nbr, n,
                                                     p.dzq + b.hid, 2 * b.hid, p.dzq, 2 * b.hid, p.dbtmp, e.rnd,
                                                     p.scale_bits + GNB_MAX_LAYERS + l, stream));
            else
                EX(gnb_edge_hidden_dgrad_scatter_bf16(p.dzb[0], p.dzb[1], b.cout, b.cout, b.w2tb[0], b.w2tb[1], b.cld64, b.hmask, b.mld,
                                                      b.hid, nbr, n, p.dzq + b.hid, 2 * b.hid, p.dzq, 2 * b.hid, p.dbtmp, e.rnd, stream));
        } else if (b.hmask != nullptr) {
            // data gradient + ReLU mask + scatter in one kernel (dh [E, hid] is never materialised). Both halves land straight
            // in dzq, the operand buffer of the two GEMMs that follow: the P half (slot sums, plain stores) rounded to tf32
            // with its column sums (bias gradient) accumulated by the epilogue, the Q half by fp32 reductions from all over
            // the event onto a zeroed half -- unrounded: the tensor core reads its truncation, which leaves the gradient
            // error where it is (tests/studies/bf16_storage_study.py: 2.1e-3 / 7.0e-4 against 2.2e-3 / 7.6e-4 with the
            // rounding pass this replaces: 12 B per value read + written + zeroed, ~55 us per layer, for a 4 B memset).
            EX(gnb_zero_block(p.dzq + b.hid, 2 * b.hid, n, b.hid, stream));
            const int nld = (int)up(b.cout, 32);
            EX(gnb_edge_hidden_dgrad_scatter_split_tf32(p.dz_big, b.cout, b.cout, b.w2t, nld, b.hmask, b.mld, b.hid, nbr, n,
                                                        p.dzq + b.hid, 2 * b.hid, p.dzq, 2 * b.hid, p.dbtmp, e.rnd, stream));
        } else {
            GNB_CHECK(cudaMemsetAsync(p.dpq, 0, (size_t)n * 2 * b.hid * 4, e.st));
            EX(e.lin_bwd_data(p.dz_big, b.cout, b.w2p, b.hld, 0, b.hid, b.cout, p.dh_big, b.hid, rows, false, b.w2t));
            EX(gnb_edge_hidden_bwd(p.dh_big, b.hid, b.h, b.hid, b.hid, nbr, deg, wl, n, GNB_ACT_RELU, p.dpq, 2 * b.hid, stream));
            // PQ = xin Wcat^T + bcat: rounded copy for the tensor cores (tf32 mode) + bias gradient in the same pass
            EX(gnb_act_bwd_colsum(p.dpq, 2 * b.hid, nullptr, 0, n, 2 * b.hid, p.dzq, 2 * b.hid, p.dbtmp, GNB_ACT_NONE | e.rnd,
                                  nullptr, 1, 0, stream));
        }
        const float* xin = l == 0 ? p.x0 : p.conv[l - 1].y;
        GNB_CHECK(cudaMemsetAsync(p.dwp, 0, (size_t)2 * b.hid * b.kld * 4, e.st));
        EX(e.lin_bwd_weight(dzq, 2 * b.hid, xin, b.cin_ld, p.dwp, b.kld, 0, b.cin_ld, 2 * b.hid, n));
        gnb_launch(unpack_conv_grad_kernel, gnb_div_up((int64_t)b.hid * b.cin, 256), 256, 0, e.st)(p.dwp, b.kld, p.dbtmp, b.hid, b.cin, gw1, gb1);
        EXL();
        if (l > 0 && p.mixed && e.tf32) EX(gnb_linear_next_absmax(p.scale_bits + GNB_MAX_LAYERS + (l - 1), 0));
        if (l > 0)   // gradient into the previous layer's output: accumulate onto the post-processing part
            EX(e.lin_bwd_data(dzq, 2 * b.hid, b.wcat, b.kld, 0, b.cin_ld, 2 * b.hid, p.gnode[l], b.cin_ld, n, true, b.wcat_t));
        if (g_bwd_event != nullptr && l == g_bwd_event_layer) GNB_CHECK(cudaEventRecord(g_bwd_event, e.st));
    }
    return GNB_OK;
}
