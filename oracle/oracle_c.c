/* CPU oracle (plain C) for the integer/index part of the DynEdge hot path.
 * TEST INFRASTRUCTURE ONLY: built by oracle/Makefile into oracle/_build/liboracle_c.so and
 * loaded by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg. Never linked
 * into or called from graphnet_b200/.
 *
 * Restates (paths relative to /root/reference):
 *   - knn_graph as called at src/graphnet/models/components/layers.py:63-67 and
 *     src/graphnet/models/graphs/edges/edges.py:74-78. The algorithm lives in the
 *     un-vendored dependency torch-cluster (>=1.6, setup.py:52-59): torch_cluster.knn(x, x,
 *     k+1, ptr, ptr) followed by removal of self pairs. This is the published CUDA kernel's
 *     formulation: one scan per query over its own event, running sorted list of the k+1
 *     best, insertion only on strict improvement (ties keep the lower index), initial
 *     distance 1e10, unfilled slots dropped.
 *   - homophily as called at src/graphnet/models/utils.py:25-28 (PyG, edge method).
 *   - scatter_{min,max,sum,mean} as called at src/graphnet/models/gnn/dynedge.py:251-264.
 * Parity at this third-party boundary is UNPINNED (no reference vectors exist); this file
 * is cross-checked against the independent sort-based torch formulation in
 * oracle/dynedge_oracle.py.
 *
 * Build with -ffp-contract=off: distances must be ((d0*d0 + d1*d1) + d2*d2) with every
 * operation rounded to fp32 (no FMA).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#define ORACLE_MAX_K1 101

/* x: [n_total, ld] fp32, distance columns cols[0..d). ptr: [nseg+1].
 * out_nbr: [n_total, k+1] (-1 padded), out_deg: [n_total]; only segments [seg_lo, seg_hi) are
 * processed (callers may thread over ranges). Returns the edge count of that range. */
int64_t oracle_knn_table(const float* x, int64_t ld, const int32_t* cols, int32_t d,
                         const int64_t* ptr, int64_t seg_lo, int64_t seg_hi, int32_t k,
                         int32_t* out_nbr, int32_t* out_deg) {
    const int32_t k1 = k + 1;
    if (k1 > ORACLE_MAX_K1) return -1;
    int64_t total = 0;
    for (int64_t b = seg_lo; b < seg_hi; ++b) {   /* callers thread over segment ranges */
        const int64_t lo = ptr[b], hi = ptr[b + 1];
        for (int64_t q = lo; q < hi; ++q) {
            float best_d[ORACLE_MAX_K1];
            int64_t best_i[ORACLE_MAX_K1];
            for (int32_t s = 0; s < k1; ++s) { best_d[s] = 1e10f; best_i[s] = -1; }
            for (int64_t c = lo; c < hi; ++c) {
                float acc = 0.0f;
                for (int32_t j = 0; j < d; ++j) {
                    const float diff = x[c * ld + cols[j]] - x[q * ld + cols[j]];
                    const float sq = diff * diff;
                    acc = (j == 0) ? sq : acc + sq;
                }
                for (int32_t e1 = 0; e1 < k1; ++e1) {
                    if (best_d[e1] > acc) {
                        for (int32_t e2 = k1 - 1; e2 > e1; --e2) {
                            best_d[e2] = best_d[e2 - 1];
                            best_i[e2] = best_i[e2 - 1];
                        }
                        best_d[e1] = acc;
                        best_i[e1] = c;
                        break;
                    }
                }
            }
            int32_t deg = 0;
            for (int32_t s = 0; s < k1; ++s) {
                if (best_i[s] >= 0 && best_i[s] != q) out_nbr[q * k1 + deg++] = (int32_t)best_i[s];
            }
            for (int32_t s = deg; s < k1; ++s) out_nbr[q * k1 + s] = -1;
            out_deg[q] = deg;
            total += deg;
        }
    }
    return total;
}

/* Expand the table into the int64 [2,E] edge_index (row 0 = neighbour, row 1 = query). */
int64_t oracle_knn_edge_index(const int32_t* nbr, const int32_t* deg, int64_t n_total, int32_t k,
                              int64_t* out_src, int64_t* out_dst) {
    const int32_t k1 = k + 1;
    int64_t e = 0;
    for (int64_t q = 0; q < n_total; ++q)
        for (int32_t s = 0; s < deg[q]; ++s) { out_src[e] = nbr[q * k1 + s]; out_dst[e] = q; ++e; }
    return e;
}

/* Per-segment homophily of column `col`: fraction of edges with equal endpoints, grouped by
 * the segment of the edge's target; 0 when a segment has no edges. */
void oracle_homophily(const float* x, int64_t ld, int32_t col, const int64_t* src, const int64_t* dst,
                      int64_t n_edges, const int64_t* batch, int64_t nseg, float* out) {
    float* tot = (float*)calloc((size_t)nseg, sizeof(float));
    float* cnt = (float*)calloc((size_t)nseg, sizeof(float));
    for (int64_t e = 0; e < n_edges; ++e) {
        const int64_t b = batch[dst[e]];
        tot[b] += (x[src[e] * ld + col] == x[dst[e] * ld + col]) ? 1.0f : 0.0f;
        cnt[b] += 1.0f;
    }
    for (int64_t b = 0; b < nseg; ++b) out[b] = tot[b] / (cnt[b] < 1.0f ? 1.0f : cnt[b]);
    free(tot); free(cnt);
}

/* scheme: 0=min 1=max 2=sum 3=mean. out [nseg, c]; arg [nseg, c] (min/max; first occurrence; -1
 * for an empty segment, whose value is 0 as in torch_scatter). */
void oracle_segment_pool(const float* x, int64_t ld, int64_t c, const int64_t* ptr, int64_t nseg,
                         int32_t scheme, float* out, int64_t* arg) {
    for (int64_t b = 0; b < nseg; ++b) {
        const int64_t lo = ptr[b], hi = ptr[b + 1];
        for (int64_t j = 0; j < c; ++j) {
            if (hi <= lo) { out[b * c + j] = 0.0f; if (arg) arg[b * c + j] = -1; continue; }
            if (scheme <= 1) {
                float best = x[lo * ld + j]; int64_t bi = lo;
                for (int64_t i = lo + 1; i < hi; ++i) {
                    const float v = x[i * ld + j];
                    if (scheme == 0 ? (v < best) : (v > best)) { best = v; bi = i; }
                }
                out[b * c + j] = best; if (arg) arg[b * c + j] = bi;
            } else {
                float s = 0.0f;
                for (int64_t i = lo; i < hi; ++i) s += x[i * ld + j];
                if (scheme == 3) s = s / (float)(hi - lo);
                out[b * c + j] = s;
            }
        }
    }
}
