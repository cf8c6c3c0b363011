/* graphnet_b200 -- C ABI of the B200 (sm_100a) DynEdge hot path.
 *
 * The reference (ArturoLlorente/graphnet) has no FFI of its own: its hot path calls third-party
 * torch operators from Python. The entry points below are the operator boundary a binding would
 * target; each comment names the reference call site (path:line relative to the reference tree) and
 * the third-party operator it replaces. INTEGRATION.md shows the ctypes / torch-extension stub.
 *
 * Conventions
 *   - plain pointers + sizes; all pointers are DEVICE pointers unless marked (host);
 *   - fp32 data, row-major, `ld*` = row pitch in elements; int32 neighbour tables; int64 ptr/batch;
 *   - `stream` is a cudaStream_t passed as void*; work is enqueued, never synchronised;
 *   - return value: 0 = OK, >0 = cudaError_t of the failed launch, -1 = invalid argument,
 *     -2 = unsupported configuration. Nothing throws; inputs are never modified.
 *
 * Graph format: `nbr[N, W]` (int32, -1 padded) holds for node i the sources j of the edges j->i in
 * ascending (distance, index) order, `deg[N]` their count. For kNN graphs W = k+1 (a node with more
 * than k exact duplicates at lower index keeps k+1 edges, as torch_cluster does). Row r = i*W + s of
 * a "padded edge list" tensor belongs to edge slot s of node i.
 */
#ifndef GRAPHNET_B200_H
#define GRAPHNET_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- graph construction ------------------------------------------------------------------ */

/* ptr[b] = first i with batch[i] >= b, b in [0, nseg]; batch sorted ascending.
 * Replaces torch_cluster.knn's `ptr = bucketize(arange(B+1), batch)`. */
int gnb_batch_to_ptr(const int64_t* batch, int64_t n, int64_t nseg, int64_t* ptr, void* stream);

/* Batched per-event k-nearest-neighbour table on columns cols[0..d) of x[n, ldx].
 * Replaces torch_geometric.nn.pool.knn_graph(x[:, cols], k, batch) at
 *   src/graphnet/models/components/layers.py:63-67 and src/graphnet/models/graphs/edges/edges.py:74-78.
 * Bit-exact total order (L2^2 in fp32 without FMA, index). k <= 100, d <= 512. */
int gnb_knn_table(const float* x, int64_t ldx, const int32_t* cols, int32_t d, const int64_t* ptr, int64_t nseg,
                  int64_t n, int32_t k, int32_t* nbr, int32_t* deg, void* stream);
/* Kernel selection for gnb_knn_table: 0 auto (8 lanes per query for k = 8 on 3 columns), 1 one thread per query,
 * 2 split. Both produce the same table bit for bit. */
int gnb_knn_set_variant(int32_t v);

/* Expand a table into the PyG edge_index[2, n_edges] (row 0 = source/neighbour, row 1 = target);
 * rowptr = exclusive prefix sum of deg (n+1 entries). */
int gnb_table_to_edge_index(const int32_t* nbr, const int32_t* deg, const int64_t* rowptr, int64_t n, int32_t width,
                            int64_t n_edges, int64_t* edge_index, void* stream);

/* ---- DynEdge global variables ------------------------------------------------------------ */

/* g[nseg, nf+5] = [scatter_mean(x) | homophily of columns 0..3 | log10(n_pulses)] and, when x0 != NULL,
 * x0[n, ld0] = [x | g[event of node] | 0...]. 4 <= nf <= 32; n_pulses is fp32[nseg].
 * Replaces DynEdge._calculate_global_variables + the dense distribute/concat at
 *   src/graphnet/models/gnn/dynedge.py:266-293, 300-319 and src/graphnet/models/utils.py:13-29. */
int gnb_global_vars(const float* x, int64_t ldx, int32_t nf, const int32_t* nbr, const int32_t* deg, int32_t width,
                    const int64_t* ptr, int64_t nseg, const float* n_pulses, float* g, float* x0, int64_t ld0,
                    void* stream);

/* ---- EdgeConv pieces (PyG EdgeConv.propagate at models/components/layers.py:60) ----------- */

/* h[(i,s), 0:hdim] = act(P[i] + Q[nbr[i,s]]) with pq[n, 2*hdim] = [P | Q]; zero row for s >= deg[i].
 * First Linear of the edge MLP hoisted to nodes. hdim % 4 == 0, 16-byte aligned rows.
 * act: 0 none, 1 relu, | 0x100 round h to tf32. */
int gnb_edge_hidden_fwd(const float* pq, int64_t ldpq, int32_t hdim, const int32_t* nbr, const int32_t* deg,
                        int32_t width, int64_t n, int32_t act, float* h, int64_t ldh, void* stream);
/* dpq[n, 2*hdim] (Q half zero on entry) from gh = dL/dh. */
int gnb_edge_hidden_bwd(const float* gh, int64_t ldg, const float* h, int64_t ldh, int32_t hdim, const int32_t* nbr,
                        const int32_t* deg, int32_t width, int64_t n, int32_t act, float* dpq, int64_t ldpq,
                        void* stream);

/* Fused EdgeConv forward on tcgen05 (tf32, no autograd state): y[i] = AGG_s relu(W2 relu(P[i] + Q[nbr[i,s]]) + b2)
 * without materialising the [E, hdim] / [E, c_out] per-edge tensors. w2p: [c_out, ceil(hdim/32)*32] tf32-rounded,
 * zero padded; pq tf32-rounded; aggr 0 add / 1 mean; hdim % 4 == 0, hdim <= 352, width <= 32. */
int gnb_edgeconv_fused_fwd_tf32(const float* pq, int64_t ldpq, int32_t hdim, const int32_t* nbr, const int32_t* deg,
                                int32_t width, int64_t n, const float* w2p, int64_t ldw, const float* b2, int32_t c_out,
                                int32_t aggr, int32_t round_out, float* y, int64_t ldy, void* stream);

/* Tuning aid for gnb_linear_fwd_tf32 / gnb_edge_linear_agg_fwd_tf32 (results are garbage with any bit set): bit0 the
 * epilogue skips its global stores, bit1 no MMAs, bit2 no weight loads, bit3 no activation loads. 0 = normal. */
int gnb_linear_set_debug(int32_t flags);
/* Kernel selection for gnb_linear_fwd_tf32 / gnb_edge_linear_agg_fwd_tf32 / gnb_edge_hidden_dgrad_scatter(_split)_tf32:
 * 0 auto (CTA-pair cta_group::2 kernel for >= 296 row tiles, dual-group scattering kernel where the hidden width needs two
 * 256-channel groups), 1 single-CTA kernel, 2 CTA-pair kernel with one cluster set per channel group, 3 CTA-pair kernel
 * with the dual-group scattering kernel forced wherever it applies. */
int gnb_linear_set_variant(int32_t v);
/* CTA-pair kernel shared-memory plan: 0 = stream the weight tiles with the activations; n >= 2 = keep the CTA's 128
 * weight rows resident whenever a single-part K leaves at least n activation stages (halves the L2->SM traffic). */
int gnb_linear_set_pair_resident(int32_t min_stages);
/* Tuning aid: device buffer of 16 uint64 that CTA (0,0) of the Linear kernel fills with cycle counters {producer:
 * wait-empty, -, total} {mma: wait-full, wait-tmem-empty, total} {epilogue warp 2: wait-tmem-full, tmem-ld, total}. */
int gnb_linear_set_profile_buffer(void* buf);

/* Kernel selection for gnb_edgeconv_fused_fwd_tf32: 0 auto (CTA-pair cta_group::2 kernel when 128 < c_out <= 256),
 * 1 single-CTA kernel, 2 CTA-pair kernel. */
int gnb_edgeconv_set_variant(int32_t v);
/* Tuning aid: device buffer of 16 uint64 that cluster 0 of the CTA-pair kernel fills with per-role wait / total cycle
 * counters {meta, mma, epilogue, builder, signal} x {wait0, wait1, total}; NULL switches it off. */
int gnb_edgeconv_set_profile_buffer(void* buf);

/* Generic message input u[(i,s)] = [x_i | x_j - x_i] and its backward (dx zero on entry). */
int gnb_edge_cat_fwd(const float* x, int64_t ldx, int32_t c_in, const int32_t* nbr, const int32_t* deg, int32_t width,
                     int64_t n, float* u, int64_t ldu, void* stream);
int gnb_edge_cat_bwd(const float* du, int64_t ldu, int32_t c_in, const int32_t* nbr, const int32_t* deg, int32_t width,
                     int64_t n, float* dx, int64_t ldx, void* stream);

/* y[i] = AGG_{s<deg[i]} m[(i,s)], aggr: 0 add, 1 mean, 2 max (arg = winning slot, int8[n, c_out]), | 0x100 round y
 * to tf32.
 * Replaces the scatter in MessagePassing.aggregate; max routes the gradient to one arg (torch_scatter). */
int gnb_edge_aggregate_fwd(const float* m, int64_t ldm, int32_t c_out, const int32_t* deg, int32_t width, int64_t n,
                           int32_t aggr, float* y, int64_t ldy, int8_t* arg, void* stream);
int gnb_edge_aggregate_bwd(const float* gy, int64_t ldy, int32_t c_out, const int32_t* deg, int32_t width, int64_t n,
                           int32_t aggr, const int8_t* arg, float* gm, int64_t ldm, void* stream);

/* ---- global pooling ----------------------------------------------------------------------- */

/* out[nseg, np*c] = cat_p scatter_<schemes[p]>(x, batch); schemes (host) in {0 min, 1 max, 2 sum, 3 mean};
 * arg[nseg, np*c] = node index of the min/max (lowest index on ties), -1 otherwise.
 * Replaces DynEdge._global_pooling, src/graphnet/models/gnn/dynedge.py:251-264 (torch_scatter). */
int gnb_segment_pool_fwd(const float* x, int64_t ldx, int32_t c, const int64_t* ptr, int64_t nseg,
                         const int32_t* schemes, int32_t np, float* out, int32_t* arg, void* stream);
int gnb_segment_pool_bwd(const float* gout, int64_t ldg, const int32_t* arg, int32_t c, const int64_t* ptr, int64_t nseg, int64_t n,
                         const int32_t* schemes, int32_t np, float* gx, int64_t ldx, void* stream);

/* ---- dense layers (torch.nn.Linear + ReLU at dynedge.py:200-203, 226-229, 246-247) -------- */

/* dz = g * (y > 0), all [rows, cols], cols % 4 == 0; flags & 0x100: round dz to tf32. */
int gnb_relu_bwd(const float* g, int64_t ldg, const float* y, int64_t ldy, int64_t rows, int32_t cols, float* dz,
                 int64_t ldz, int32_t flags, void* stream);
/* Fused backward of bias+activation (and optionally of the k-neighbour aggregation in front of it):
 * dz[r] = grow(r) * act'(y[r]), db += colsum(dz); grow(r) = g[r] or, with deg != NULL, g[r / width] (add) or
 * g[r / width] / deg (mean) for valid slots and 0 for padding slots. flags: act | 0x100 (round dz to tf32) | 0x400 (only with
 * act = none, deg = NULL, db = NULL: g is zeroed behind the read -- an accumulation buffer its next user expects empty). */
int gnb_act_bwd_colsum(const float* g, int64_t ldg, const float* y, int64_t ldy, int64_t rows, int32_t cols, float* dz,
                       int64_t ldz, float* db, int32_t flags, const int32_t* deg, int32_t width, int32_t aggr,
                       void* stream);
/* out[c] += sum_r a[r, c]. */
int gnb_colsum(const float* a, int64_t lda, int64_t rows, int32_t cols, float* out, void* stream);

/* tcgen05 (kind::tf32, fp32 accumulate in TMEM, TMA-fed) backend of the same Linear:
 * y[rows, n_out] = act(sum_p xs[p][rows, ks[p]] w[:, koff_p : koff_p + ks[p]]^T + bias), koff_p = running sum of
 * ceil(ks[p]/32)*32. xs / ldxs / ks are HOST arrays (nparts <= 6); pointers 16-byte aligned, pitches % 4 == 0.
 * Operands are expected pre-rounded to tf32 (gnb_round_pad_tf32 or a producer's 0x100 flag); round_out rounds y.
 * With nparts > 1 this is the skip-concatenation + first post-processing Linear of dynedge.py:328-331.
 * act = GNB_ACT_* | 0x200: with bit 0x200 the result is ADDED to y (gradient accumulation; round_out ignored). */
int gnb_linear_fwd_tf32(const float* const* xs, const int64_t* ldxs, const int32_t* ks, int32_t nparts, const float* w,
                        int64_t ldw, const float* bias, float* y, int64_t ldy, int64_t rows, int32_t n_out, int32_t act,
                        int32_t round_out, void* stream);
/* The same Linear at fp32 grade on the tensor cores (precision mode "tf32x3", the forward pass of training): split
 * operands, W = w_hi + w_lo and x = x_hi + x_lo (hi parts tf32-exact; plain fp32 activations, split inside the kernel):
 * w_hi x_hi as kind::tf32 plus the two correction products w_lo x_hi + w_hi x_lo as ONE bf16 contraction (kind::f16) over
 * the concatenated K axis -- w_lo here is that bf16 operand as written by gnb_split_pad_tf32 (same bytes and pitch as a
 * [n_out, ldw] fp32 matrix) --, fp32 accumulation in TMEM; y is written unrounded. Replaces torch.nn.Linear's fp32 arithmetic (dynedge.py:200-203, 226-229,
 * 246-247; the reference computes in fp32, graphs/graphs.py:21). Always the CTA-pair kernel; n_out <= 1024. */
int gnb_linear_fwd_tf32x3(const float* const* xs, const int64_t* ldxs, const int32_t* ks, int32_t nparts,
                          const float* w_hi, const float* w_lo, int64_t ldw, const float* bias, float* y, int64_t ldy,
                          int64_t rows, int32_t n_out, int32_t act, void* stream);
/* hi[rows, dst_cols] = [rna_tf32(src) | 0] and the bf16 correction operand of the tf32x3 GEMMs in `lo`: row r viewed as
 * bf16[2 ldd] holds per 32-wide K block kb, at bf16 index 64 kb, [bf16(src - hi) x 32 | bf16(hi) x 32]. dst_cols % 32 == 0. */
int gnb_split_pad_tf32(const float* src, int64_t lds, int64_t rows, int32_t cols, float* hi, float* lo, int64_t ldd,
                       int32_t dst_cols, void* stream);
/* dw[n_out, k_in] += dz[rows, n_out]^T x[rows, k_in] on tcgen05 (split over rows, fp32 red.add into dw).
 * TMA-fed MN-major operands; same alignment rules as gnb_linear_fwd_tf32. debug_swap: 0 in production. */
int gnb_linear_bwd_weight_tf32(const float* dz, int64_t lddz, const float* x, int64_t ldx, float* dw, int64_t lddw,
                               int64_t rows, int32_t n_out, int32_t k_in, int32_t debug_swap, void* stream);
/* Second EdgeConv Linear + ReLU + k-neighbour SUM in one tcgen05 kernel (k = 8 tables: width 9, tiles of 14 nodes):
 * y[i] = sum_{s<deg[i]} relu(h[i*9+s] w^T + bias); maskbits[(i/14) * n_out + ch][4 x u32]: bit (i%14)*9+s = pre-act > 0.
 * The [E, n_out] message tensor is never stored. h / w as for gnb_linear_fwd_tf32 (single part). */
int gnb_edge_linear_agg_fwd_tf32(const float* h, int64_t ldh, int32_t k, const float* w, int64_t ldw, const float* bias,
                                 const int32_t* deg, int64_t n, int32_t n_out, int32_t round_out, float* y, int64_t ldy,
                                 uint32_t* maskbits, void* stream);
/* gnb_edge_linear_agg_fwd_tf32 with split operands (see gnb_linear_fwd_tf32x3): h plain fp32, w_hi / w_lo packed like w;
 * y unrounded. PyG EdgeConv's second Linear + aggr="add" (layers.py:55-62) at fp32 grade. */
int gnb_edge_linear_agg_fwd_tf32x3(const float* h, int64_t ldh, int32_t k, const float* w_hi, const float* w_lo, int64_t ldw,
                                   const float* bias, const int32_t* deg, int64_t n, int32_t n_out, float* y, int64_t ldy,
                                   uint32_t* maskbits, void* stream);
/* Second Linear of an EdgeConv MLP + activation + k-neighbour MAX in one tcgen05 kernel: replaces PyG MessagePassing
 * (aggr="max") over EdgeConvTito.message (reference models/components/layers.py:72-114, instantiated with aggr="max" at
 * models/gnn/dynedge_kaggle_tito.py:157-162): y[i] = max_{s<deg[i]} act(h[i*9+s] w^T + bias), 0 when deg[i] = 0;
 * arg[i * ldarg + ch] = winning slot | 0x40 if its pre-activation was > 0, -1 when deg[i] = 0; the first slot wins ties
 * (torch_scatter scatter_max). act = GNB_ACT_NONE (0) / RELU (1) / LEAKY (2: torch.nn.LeakyReLU() default slope 0.01).
 * The [E, n_out] message tensor is never stored. h / w as for gnb_edge_linear_agg_fwd_tf32. */
int gnb_edge_linear_aggmax_fwd_tf32(const float* h, int64_t ldh, int32_t k, const float* w, int64_t ldw, const float* bias,
                                    const int32_t* deg, int64_t n, int32_t n_out, int32_t act, int32_t round_out, float* y,
                                    int64_t ldy, int8_t* arg, int64_t ldarg, void* stream);
/* The same on split operands (fp32-grade forward, see gnb_linear_fwd_tf32x3): h plain fp32, y unrounded. */
int gnb_edge_linear_aggmax_fwd_tf32x3(const float* h, int64_t ldh, int32_t k, const float* w_hi, const float* w_lo, int64_t ldw,
                                      const float* bias, const int32_t* deg, int64_t n, int32_t n_out, int32_t act, float* y,
                                      int64_t ldy, int8_t* arg, int64_t ldarg, void* stream);
/* Arg-routed backward of the max aggregation and its activation (autograd of scatter_max + LeakyReLU / ReLU in the reference):
 * dz[i*width+s, c] = (s == slot(arg[i,c])) ? gy[i,c] * (arg[i,c] & 0x40 ? 1 : slope(act)) : 0; db[c] += column sums (db may
 * be NULL); flags & 0x100 rounds dz to tf32. c_out <= 512. */
int gnb_edge_argmax_bwd(const float* gy, int64_t ldy, const int8_t* arg, int64_t ldarg, int32_t c_out, int32_t width, int64_t n,
                        int32_t act, int32_t flags, float* dz, int64_t ldz, float* db, void* stream);
/* Backward of the above up to the pre-activation: dz[i*9+s] = g[i] * maskbit, db += colsum(dz); flags & 0x100 rounds dz. */
int gnb_edge_mask_bwd_colsum(const float* g, int64_t ldg, const uint32_t* maskbits, int64_t n, int32_t cols,
                             const int32_t* deg, float* dz, int64_t ldz, float* db, int32_t flags, void* stream);
/* Data gradient of the EdgeConv second Linear fused with the backward of the hoisted hidden layer (autograd of
 * torch.nn.Linear + ReLU + the x_i / x_j gather at dynedge.py:200-203 / PyG EdgeConv.message), k = 8 tables (width 9):
 *   dh = dz wt^T, da = dh * (h > 0), dpq[i, 0:hdim] = sum_s da[(i,s)], dpq[nbr[i,s], hdim:2*hdim] += da[(i,s)].
 * wt = W2^T [hdim, ceil(c_out/32)*32] tf32-rounded, zero padded; dz [n*9, c_out] tf32-rounded; hmask = activation bits
 * written by gnb_edge_hidden_fwd_mask, [ceil(n/14)*126, mask_ld] words (mask_ld % 4 == 0, >= 4*ceil(hdim/128), 16-byte
 * aligned); nbr [n, 9] (-1 padded); dpq [n, ldpq >= 2*hdim], Q half zero on entry; n*ldpq < 2^31. dh is never stored. */
int gnb_edge_hidden_dgrad_scatter_tf32(const float* dz, int64_t lddz, int32_t c_out, const float* wt, int64_t ldw,
                                       const uint32_t* hmask, int32_t mask_ld, int32_t hdim, const int32_t* nbr, int64_t n,
                                       float* dpq, int64_t ldpq, void* stream);
/* Same kernel with the two halves in separate tensors: dq [n, lddq >= hdim] receives the reductions (zero on entry,
 * n*lddq < 2^31), dp [n, lddp >= hdim] is overwritten with the slot sums (rounded to tf32 with flags & 0x100, i.e. ready
 * to feed the next tensor-core GEMM) and, if dbias != NULL, dbias[0:hdim] += column sums of the unrounded dp (= gradient
 * of the hoisted Linear's bias; torch.nn.Linear autograd at dynedge.py:200-203). */
int gnb_edge_hidden_dgrad_scatter_split_tf32(const float* dz, int64_t lddz, int32_t c_out, const float* wt, int64_t ldw,
                                             const uint32_t* hmask, int32_t mask_ld, int32_t hdim, const int32_t* nbr,
                                             int64_t n, float* dq, int64_t lddq, float* dp, int64_t lddp, float* dbias,
                                             int32_t flags, void* stream);
/* gnb_edge_hidden_fwd that also writes the activation bits hmask[r, mask_ld] (mask_ld % 4 == 0, mask_ld * 32 >= hdim):
 * (h[r, c] > 0) is bit (c % 128) / 4 of word 4 * (c / 128) + c % 4; bits beyond hdim are zero. */
int gnb_edge_hidden_fwd_mask(const float* pq, int64_t ldpq, int32_t hdim, const int32_t* nbr, const int32_t* deg,
                             int32_t width, int64_t n, int32_t act, float* h, int64_t ldh, uint32_t* hmask,
                             int32_t mask_ld, void* stream);
/* dst[rows, dst_cols] = [rna_tf32(src[rows, cols]) | 0]. */
int gnb_round_pad_tf32(const float* src, int64_t lds, int64_t rows, int32_t cols, float* dst, int64_t ldd,
                       int32_t dst_cols, void* stream);

/* ---- bf16-plane per-edge GEMMs: precision modes "bf16" (one plane, north_star's looser mode) and "bf16x3" (two planes) ----
 * A per-edge tensor v is stored as bf16 planes v0 = bf16(v), v1 = bf16(v - v0) written by its producer; plane-1 pointers
 * NULL (for every operand of a call) = one plane. Pitches are in bf16 ELEMENTS, multiples of 8; pointers 16-byte aligned.
 * One plane: a single kind::f16 product (half the bytes, twice the MMA rate of tf32). Two planes: v0 w0 + v1 w0 + v0 w1
 * (dropped terms ~2^-17) -- the bytes of the fp32 tensors, no in-kernel operand split. Same reference arithmetic as the tf32
 * entry points they mirror (torch.nn.Linear / PyG EdgeConv in fp32: layers.py:55-62, dynedge.py:200-203). */
/* hidden layer of the hoisted EdgeConv MLP as planes: h = relu(P_i + Q_j) (gnb_edge_hidden_fwd_mask), hmask may be NULL. */
int gnb_edge_hidden_fwd_bf16(const float* pq, int64_t ldpq, int32_t hdim, const int32_t* nbr, const int32_t* deg,
                             int32_t width, int64_t n, void* h0, void* h1, int64_t ldh, uint32_t* hmask, int32_t mask_ld,
                             void* stream);
/* gnb_edge_linear_agg_fwd_tf32 on planes: h planes [n*9, k], w planes [n_out, ldw >= k] (zero beyond k). */
int gnb_edge_linear_agg_fwd_bf16(const void* h0, const void* h1, int64_t ldh, int32_t k, const void* w0, const void* w1,
                                 int64_t ldw, const float* bias, const int32_t* deg, int64_t n, int32_t n_out,
                                 int32_t round_out, float* y, int64_t ldy, uint32_t* maskbits, void* stream);
/* gnb_edge_mask_bwd_colsum writing dz as planes [n*9, ldz >= cols]; db sums the fp32 values. */
int gnb_edge_mask_bwd_colsum_bf16(const float* g, int64_t ldg, const uint32_t* maskbits, int64_t n, int32_t cols, void* dz0,
                                  void* dz1, int64_t ldz, float* db, void* stream);
/* gnb_linear_bwd_weight_tf32 on planes: dw[n_out, k_in] += dz^T x (fp32 red.add); debug: 0 in production. */
int gnb_linear_bwd_weight_bf16(const void* dz0, const void* dz1, int64_t lddz, const void* x0, const void* x1, int64_t ldx,
                               float* dw, int64_t lddw, int64_t rows, int32_t n_out, int32_t k_in, int32_t debug,
                               void* stream);
/* gnb_edge_hidden_dgrad_scatter_split_tf32 on planes: dz planes [n*9, c_out], wt planes = W2^T [hdim, ldw >= c_out]. */
int gnb_edge_hidden_dgrad_scatter_bf16(const void* dz0, const void* dz1, int64_t lddz, int32_t c_out, const void* wt0,
                                       const void* wt1, int64_t ldw, const uint32_t* hmask, int32_t mask_ld, int32_t hdim,
                                       const int32_t* nbr, int64_t n, float* dq, int64_t lddq, float* dp, int64_t lddp,
                                       float* dbias, int32_t flags, void* stream);
/* Plain Linear on planes (CTA-pair kernel): y = act(x w^T + bias), fp32 output. */
int gnb_linear_fwd_bf16(const void* x0, const void* x1, int64_t ldx, int32_t k, const void* w0, const void* w1, int64_t ldw,
                        const float* bias, float* y, int64_t ldy, int64_t rows, int32_t n_out, int32_t act,
                        int32_t round_out, void* stream);
/* ---- precision mode "mixed16": the per-edge tensors as fp16 planes -- fp16 carries tf32's 11-bit significand in half the
 * bytes -- scaled per layer by a power of two so that they sit inside fp16's range. A scale word (device uint32) holds the
 * fp32 bits of (a bound on) max|v|: gnb_absmax_bits over the producer's input; scale 2^s = 2^(14 - floor(log2 max)), applied by
 * the producer and undone exactly (power of two) in the consuming epilogues. Forward: two planes of h and W2, three products
 * (fp32 grade); backward: ONE plane of dz, h, W2^T (tf32 grade). (bf16 against fp16 operands in one MMA is an illegal
 * instruction on sm_100a, so every operand of these GEMMs is fp16.) */
/* *out_bits = max(*out_bits, fp32 bits of 2^shift * max|a|); *out_bits zero before the first call. */
int gnb_absmax_bits(const float* a, int64_t lda, int64_t rows, int32_t cols, int32_t shift, uint32_t* out_bits, void* stream);
/* One-shot: the next gnb_linear_fwd_tf32 / _tf32x3 launch also folds bits of 2^shift max|y| into *bits (as gnb_absmax_bits). */
int gnb_linear_next_absmax(uint32_t* bits, int32_t shift);
/* gnb_edge_hidden_fwd_bf16 with fp16 planes of h * 2^s; *scale_bits >= bits of max h (absmax of PQ with shift 1). */
int gnb_edge_hidden_fwd_f16(const float* pq, int64_t ldpq, int32_t hdim, const int32_t* nbr, const int32_t* deg,
                            int32_t width, int64_t n, void* h0, void* h1, int64_t ldh, uint32_t* hmask, int32_t mask_ld,
                            const uint32_t* scale_bits, void* stream);
/* gnb_edge_linear_agg_fwd_bf16 on fp16 planes; the epilogue folds 2^-s of h into its bias FMA. */
int gnb_edge_linear_agg_fwd_f16(const void* h0, const void* h1, int64_t ldh, int32_t k, const void* w0, const void* w1,
                                int64_t ldw, const float* bias, const int32_t* deg, int64_t n, int32_t n_out,
                                int32_t round_out, float* y, int64_t ldy, uint32_t* maskbits, const uint32_t* scale_bits,
                                void* stream);
/* gnb_edge_mask_bwd_colsum with dz as ONE fp16 plane of dz * 2^s (*scale_bits = bits of max|g|). */
int gnb_edge_mask_bwd_colsum_f16(const float* g, int64_t ldg, const uint32_t* maskbits, int64_t n, int32_t cols, void* dz,
                                 int64_t ldz, float* db, const uint32_t* scale_bits, void* stream);
/* dw += dz^T x: dz the scaled fp16 plane, x fp16 plane(s) of the scaled x (x1 may be NULL); both scales undone. */
int gnb_linear_bwd_weight_f16(const void* dz, int64_t lddz, const void* x0, const void* x1, int64_t ldx, float* dw,
                              int64_t lddw, int64_t rows, int32_t n_out, int32_t k_in, const uint32_t* dz_scale_bits,
                              const uint32_t* x_scale_bits, void* stream);
/* gnb_edge_hidden_dgrad_scatter_split_tf32 with dz (scaled) and wt = W2^T as single fp16 planes. */
int gnb_edge_hidden_dgrad_scatter_f16(const void* dz, int64_t lddz, int32_t c_out, const void* wt, int64_t ldw,
                                      const uint32_t* hmask, int32_t mask_ld, int32_t hdim, const int32_t* nbr, int64_t n,
                                      float* dq, int64_t lddq, float* dp, int64_t lddp, float* dbias, int32_t flags,
                                      const uint32_t* scale_bits, void* stream);
/* Fused EdgeConv forward of the fp16-plane modes: gather + hidden layer + second Linear + ReLU + k-sum in ONE tcgen05 kernel
 * (PyG EdgeConv.message / aggregate, layers.py:55-62, on the hoisted first Linear): builder warps produce the B operand tile
 * h = relu(P_i + Q_j) * 2^s (fp16 plane 0, and plane 1 = remainder when w1 != NULL) in shared memory from PQ [n, 2 hid], so h
 * never crosses HBM on the forward path. y / maskbits as gnb_edge_linear_agg_fwd_f16. Training side outputs (may be NULL):
 * h0_out [9 n, ldh] = plane 0 of h (x operand of the weight gradient), hbytes [ceil(n / 14) * 126, ldhb] = bits of h > 0,
 * byte c / 8 bit c % 8 (= row-major words; flags 0x800 of gnb_edge_hidden_dgrad_scatter_f16_masked). n_out <= 256, hid % 8 == 0.
 * pq_layout: 0 = natural column order of the P and Q halves; 1 = every full 64-column block of each half lane-interleaved
 * (stored 16-byte piece j < 8 holds hidden units 8 j .. 8 j + 3, piece 8 + j units 8 j + 4 .. 8 j + 7: the eight lanes of a row
 * then gather 128 contiguous bytes per load; the caller packs the hoisted Linear's weight rows in that order). Outputs natural. */
int gnb_edgeconv_fused_fwd_f16(const float* pq, int64_t ldpq, int32_t hid, const int32_t* nbr, const int32_t* deg, int64_t n,
                               const void* w0, const void* w1, int64_t ldw, const float* bias, int32_t n_out, int32_t round_out,
                               float* y, int64_t ldy, uint32_t* maskbits, void* h0_out, int64_t ldh, uint8_t* hbytes,
                               int64_t ldhb, const uint32_t* scale_bits, int32_t pq_layout, void* stream);
/* The backward of the aggregating Linear WITHOUT a stored dz (fp16-plane modes): dz[(i, s), :] = g[i, :] * bit(i, s, :) is a
 * 9-fold redundant function of the node-level gradient and the ReLU bits, so the two GEMMs expand it in shared memory (builder
 * warps write the tensor-core operand tile) instead of reading a stored [9 n, c_out] tensor (autograd of PyG EdgeConv's
 * aggr="add" + ReLU, layers.py:55-62). gnb_edge_dz_prep makes their inputs in one pass over g [n, cols] and the tile-major
 * maskbits of gnb_edge_linear_agg_fwd_*: g16 [n, cols] = fp16(g * 2^s) (2^s from *scale_bits = bits of max|g|), rowmask =
 * the same bits ROW-major ([ceil(n / 14) * 126, cols / 32] words: bit c % 32 of word c / 32 of row i * 9 + s), and db[c] +=
 * column sums of dz (the bias gradient). cols % 32 == 0 (scatter: % 64), cols <= 256, k = 8 tables. */
int gnb_edge_dz_prep(const float* g, int64_t ldg, const uint32_t* maskbits, int64_t n, int32_t cols, const uint32_t* scale_bits,
                     void* g16, uint32_t* rowmask, float* db, void* stream);
/* profiling hook of gnb_linear_bwd_weight_f16_masked (0 in production; results are garbage otherwise) */
int gnb_wgrad_set_debug(int32_t flags);
/* dw[n_out, k_in] += dz^T x with x = one fp16 plane of h * 2^sx [9 n, k_in]; both scales undone in the epilogue. */
int gnb_linear_bwd_weight_f16_masked(const void* g16, const uint32_t* rowmask, const void* x, int64_t ldx, float* dw, int64_t lddw,
                                     int64_t n, int32_t n_out, int32_t k_in, const uint32_t* dz_scale_bits,
                                     const uint32_t* x_scale_bits, void* stream);
/* gnb_edge_hidden_dgrad_scatter_f16 on the expanded dz; wt = W2^T as one fp16 plane [hdim, ldw >= c_out]. */
int gnb_edge_hidden_dgrad_scatter_f16_masked(const void* g16, const uint32_t* rowmask, int32_t c_out, const void* wt, int64_t ldw,
                                             const uint32_t* hmask, int32_t mask_ld, int32_t hdim, const int32_t* nbr, int64_t n,
                                             float* dq, int64_t lddq, float* dp, int64_t lddp, float* dbias, int32_t flags,
                                             const uint32_t* scale_bits, void* stream);
/* The 8-slot edge layout (fp16-plane modes, k = 8). The neighbour table is k + 1 = 9 wide because a node with more than k exact
 * duplicates at lower index keeps k + 1 edges (torch_cluster's knn_graph behind edges.py:72-80, layers.py:63-67); on graphs where
 * no node does, slot 8 of every node is padding and costs 1/9 of every per-edge kernel. gnb_edge_slot_flag writes *flag = 1 when
 * some deg[i] > k, else 0, on the device (no host synchronisation); the four _w entry points below take that word as `full9`
 * (NULL or *full9 != 0: the 9-slot layout documented above) and, when *full9 == 0, run on 16-node x 8-slot tiles: per-edge rows
 * i * 8 + s (h0_out / x [8 n, .], hbytes / hmask / rowmask [ceil(n / 16) * 128, .]), maskbits[(i / 16) * cols + c] bit
 * 8 (i % 16) + s. A layer's four kernels must be given the same word. Results are identical in both layouts (the padding slot
 * contributes exact zeros). Buffers sized for max(ceil(n / 14) * 126, ceil(n / 16) * 128) rows fit either layout. */
int gnb_edge_slot_flag(const int32_t* deg, int64_t n, int32_t k, int32_t* flag, void* stream);
/* the same for a flag word the caller has zeroed (*flag = 1 is stored when some deg[i] > k, nothing otherwise): many CTAs */
int gnb_edge_slot_flag_or(const int32_t* deg, int64_t n, int32_t k, int32_t* flag, void* stream);
int gnb_edgeconv_fused_fwd_f16_w(const float* pq, int64_t ldpq, int32_t hid, const int32_t* nbr, const int32_t* deg, int64_t n,
                                 const void* w0, const void* w1, int64_t ldw, const float* bias, int32_t n_out, int32_t round_out,
                                 float* y, int64_t ldy, uint32_t* maskbits, void* h0_out, int64_t ldh, uint8_t* hbytes,
                                 int64_t ldhb, const uint32_t* scale_bits, int32_t pq_layout, const int32_t* full9, void* stream);
int gnb_edge_dz_prep_w(const float* g, int64_t ldg, const uint32_t* maskbits, int64_t n, int32_t cols, const uint32_t* scale_bits,
                       void* g16, uint32_t* rowmask, float* db, const int32_t* full9, void* stream);
/* gnb_edge_dz_prep_w with a zero job riding on the launch: zero_a[r, 0 : zcols] = 0 for r < zrows (pitch lda; the scatter target of
 * the data-gradient kernel that follows) and zero_b[0 : zb_count] = 0; either pointer may be NULL. */
int gnb_edge_dz_prep_wz(const float* g, int64_t ldg, const uint32_t* maskbits, int64_t n, int32_t cols, const uint32_t* scale_bits,
                        void* g16, uint32_t* rowmask, float* db, const int32_t* full9, float* zero_a, int64_t lda, int64_t zrows,
                        int32_t zcols, float* zero_b, int32_t zb_count, void* stream);
int gnb_linear_bwd_weight_f16_masked_w(const void* g16, const uint32_t* rowmask, const void* x, int64_t ldx, float* dw,
                                       int64_t lddw, int64_t n, int32_t n_out, int32_t k_in, const uint32_t* dz_scale_bits,
                                       const uint32_t* x_scale_bits, const int32_t* full9, void* stream);
int gnb_edge_hidden_dgrad_scatter_f16_masked_w(const void* g16, const uint32_t* rowmask, int32_t c_out, const void* wt, int64_t ldw,
                                               const uint32_t* hmask, int32_t mask_ld, int32_t hdim, const int32_t* nbr, int64_t n,
                                               float* dq, int64_t lddq, float* dp, int64_t lddp, float* dbias, int32_t flags,
                                               const uint32_t* scale_bits, const int32_t* full9, void* stream);
/* fp32 [rows, cols] -> fp16 planes (round to nearest; zero padded to dst_cols; p1 may be NULL); transpose != 0: of src^T. */
int gnb_to_f16_planes(const float* src, int64_t lds, int64_t rows, int32_t cols, void* p0, void* p1, int64_t ldd,
                      int32_t dst_cols, int32_t transpose, void* stream);
/* fp32 [rows, cols] -> planes [rows, dst_cols] (zero beyond cols); transpose != 0: planes of src^T ([cols, dst_cols >= rows]). */
int gnb_to_bf16_planes(const float* src, int64_t lds, int64_t rows, int32_t cols, void* p0, void* p1, int64_t ldd,
                       int32_t dst_cols, int32_t transpose, void* stream);

/* a[r, 0:cols] = 0 over a [rows, cols] block of pitch lda (the zero-on-entry scatter target dq of the fused data-gradient kernels). */
int gnb_zero_block(float* a, int64_t lda, int64_t rows, int32_t cols, void* stream);

/* fp32 SIMT backend: y[m,n] = act(x[m,k] w[n,k]^T + bias (+ y if accumulate)); act: 0 none, 1 relu. */
int gnb_linear_fwd_f32(const float* x, int64_t ldx, const float* w, int64_t ldw, const float* bias, float* y,
                       int64_t ldy, int64_t m, int64_t n, int64_t k, int32_t act, int32_t accumulate, void* stream);
/* dx[m,k] = dz[m,n] w[n,k] (+ dx if accumulate). */
int gnb_linear_bwd_data_f32(const float* dz, int64_t lddz, const float* w, int64_t ldw, float* dx, int64_t lddx,
                            int64_t m, int64_t n, int64_t k, int32_t accumulate, void* stream);
/* dw[n,k] += dz[m,n]^T x[m,k]. */
int gnb_linear_bwd_weight_f32(const float* dz, int64_t lddz, const float* x, int64_t ldx, float* dw, int64_t lddw,
                              int64_t m, int64_t n, int64_t k, void* stream);

/* Number of kernels launched by this library so far (monotonic; bench.py reports per-step deltas). */
int64_t gnb_launch_count(void);

/* ---- native step executor (whole DynEdge.forward / backward from one call) ------------------- */

#define GNB_MAX_LAYERS 8
#define GNB_MAX_KNN_COLS 16
/* Mirrors the constructor arguments of DynEdge (src/graphnet/models/gnn/dynedge.py:24-38) for the fast-path
 * family: ReLU, no norm layers, 2-Linear conv MLPs with aggr=add, Linear-ReLU post-processing / read-out chains. */
typedef struct {
    int32_t nb_inputs, k, precision;                 /* precision: 0 fp32 SIMT GEMMs, 1 tf32 tcgen05 GEMMs, 2 tf32x3: forward GEMMs split-operand (fp32 grade), backward GEMMs tf32,
                                                        3 bf16: per-edge tensors as one bf16 plane (kind::f16), node-level GEMMs as 1,
                                                        4 bf16x3: per-edge tensors as two bf16 planes, node-level GEMMs as 2,
                                                        5 mixed16: per-edge tensors as scaled fp16 planes (two forward, one backward), node-level GEMMs as 2,
                                                        6 f16: one scaled fp16 plane forward and backward (tf32 grade), node-level GEMMs as 1 */
    int32_t n_conv, conv_hidden[GNB_MAX_LAYERS], conv_out[GNB_MAX_LAYERS];
    int32_t n_post, post_out[GNB_MAX_LAYERS];
    int32_t n_readout, readout_out[GNB_MAX_LAYERS];
    int32_t n_pool, pool[4];                         /* 0 min, 1 max, 2 sum, 3 mean, caller order */
    int32_t globals_after_pooling, skip_readout;
    int32_t n_knn_cols, knn_cols[GNB_MAX_KNN_COLS];  /* features_subset */
    int32_t flags;                                   /* bit 0: do not use the fused tcgen05 EdgeConv kernels / epilogues;
                                                        bit 1: inference runs hidden layer + aggregating GEMM (two kernels per
                                                        layer, faster at k = 8) instead of the single fused EdgeConv kernel;
                                                        bit 2: fp16-plane modes store dz (mask-backward kernel) instead of
                                                        expanding it inside the two backward GEMMs;
                                                        bit 3: fp16-plane modes keep the two-kernel forward (hidden-layer kernel
                                                        + aggregating GEMM) instead of the fused EdgeConv forward;
                                                        bit 4: fp16-plane modes always use the 9-slot edge layout (default: the
                                                        8-slot layout for every graph without a k + 1-neighbour node) */
} gnb_dynedge_config;

/* Bytes of workspace for a batch of n nodes / nseg events whose initial graph has table width w0; < 0: error. */
int64_t gnb_dynedge_workspace_bytes(const gnb_dynedge_config* cfg, int64_t n, int64_t nseg, int32_t w0, int32_t training);
/* Byte offsets into the workspace of, per conv layer l: its output y_l [n, conv_out[l]], and the graph recomputed
 * from it (nbr [n, k+1], deg [n]; -1 for the last layer): offsets[3*l + {0,1,2}]. Test / debug aid. */
int gnb_dynedge_layout(const gnb_dynedge_config* cfg, int64_t n, int64_t nseg, int32_t w0, int32_t training,
                       int64_t* offsets);
/* DynEdge.forward(data) (dynedge.py:295-349). params: HOST array of device pointers in state_dict order
 * ({W1,b1,W2,b2} per conv, {W,b} per post layer, {W,b} per read-out layer). out: [nseg or n, last width]. */
int gnb_dynedge_forward(const gnb_dynedge_config* cfg, const float* const* params, const float* x, int64_t ldx,
                        const int64_t* ptr, const float* n_pulses, const int32_t* nbr0, const int32_t* deg0, int32_t w0,
                        const int32_t* knn_cols_dev, int64_t n, int64_t nseg, void* workspace, int64_t workspace_bytes,
                        float* out, int32_t training, void* stream);
/* Backward of the above (after a forward with training = 1 on the same workspace); gradients are accumulated
 * onto grads[] (HOST array of device pointers parallel to params). */
int gnb_dynedge_backward(const gnb_dynedge_config* cfg, float* const* grads, const int64_t* ptr, const int32_t* nbr0,
                         const int32_t* deg0, int32_t w0, int64_t n, int64_t nseg, void* workspace,
                         int64_t workspace_bytes, const float* gout, void* stream);
/* Multi-GPU overlap hook: `event` (a cudaEvent_t, NULL to clear) is recorded on the backward's stream right after the
 * backward of conv layer `after_conv_layer` has been enqueued; from then on the gradients of that layer, of every later conv
 * layer, of the post-processing and of the read-out are final, so their slice of the flat gradient buffer can be all-reduced
 * on a side stream while the earlier layers' backward runs (DDP's bucketed overlap, easy_model.py:90-110). Process-wide. */
int gnb_dynedge_set_backward_event(void* event, int32_t after_conv_layer);

/* Device-side graph definition (raw pulses -> DynEdge input). gnb_standardize replaces Detector._standardize
 * (models/detector/detector.py:63-77; IceCube86 table at detector/icecube.py:21-48): out[r, c] = x (kind 0),
 * (x - sub[c]) / div[c] (kind 1) or log10(x) (kind 2), fp32, same operation order. kind / sub / div: HOST arrays of f <= 32
 * entries. gnb_ptr_to_batch builds the `batch` vector of a collated Batch (data/dataloader.py:12-18) from ptr[nseg + 1]. */
int gnb_standardize(const float* x, int64_t ldx, int64_t n, int32_t f, const int32_t* kind, const float* sub, const float* div,
                    float* out, int64_t ldo, void* stream);
int gnb_ptr_to_batch(const int64_t* ptr, int64_t nseg, int64_t n, int64_t* batch, void* stream);
/* Adam step over one flat fp32 parameter buffer (torch.optim.Adam semantics as configured by the reference:
 * easy_model.py:215-219, examples/04_training/01_train_dynedge.py:128-129): g' = g + weight_decay p, m += (1-beta1)(g'-m),
 * v = beta2 v + (1-beta2) g'^2, p -= step_size m / (sqrt(v) inv_sqrt_bc2 + eps) with step_size = lr / (1 - beta1^t) and
 * inv_sqrt_bc2 = 1 / sqrt(1 - beta2^t) computed by the caller. p, g, m, v: [n], 16-byte aligned; zero_grad != 0 leaves g = 0.
 * g is multiplied by grad_scale on the way in (1 / world_size = the mean of DDP, easy_model.py:90-110, after a SUM all-reduce). */
int gnb_adam_flat(float* p, float* g, float* m, float* v, int64_t n, float step_size, float beta1, float beta2, float eps,
                  float inv_sqrt_bc2, float weight_decay, float grad_scale, int32_t zero_grad, void* stream);
/* ---- task heads + losses (the O(B) step right after the path) -------------------------------------
 * EnergyReconstruction (task/reconstruction.py:101-112) + LogCoshLoss on log10 (training/loss_functions.py:93-112) and
 * DirectionReconstructionWithKappa (reconstruction.py:49-70) + VonMisesFisher3DLoss (loss_functions.py:281-353, 424-447;
 * log C_3 in closed form, no host round trip). feat [nev, hdim]; we [hdim], be [1]; wd [3, hdim], bd [3]; energy [nev];
 * direction [nev, 3]. pred_e [nev], pred_d [nev, 4] = (unit vector, kappa); dz [nev, 4] = d(loss_e + loss_d) / d(affine
 * outputs), already divided by nev; loss[2] += (mean log-cosh, mean vMF NLL), zero on entry. */
int gnb_task_heads_fwd(const float* feat, int64_t ldf, int32_t hdim, const float* we, const float* be, const float* wd,
                       const float* bd, const float* energy, const float* direction, int64_t nev, float* pred_e,
                       float* pred_d, float* dz, float* loss, void* stream);
/* Backward for an upstream scalar gradient gout[0] (NULL = 1): dfeat [nev, hdim] written (may be NULL); dwe [hdim],
 * dbe [1], dwd [3, hdim], dbd [3] ACCUMULATED (+=). hdim <= 2048. */
int gnb_task_heads_bwd(const float* feat, int64_t ldf, int32_t hdim, const float* we, const float* wd, const float* dz,
                       const float* gout, int64_t nev, float* dfeat, int64_t lddf, float* dwe, float* dbe, float* dwd,
                       float* dbd, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GRAPHNET_B200_H */
