// Batched per-event k-nearest-neighbour graph build (bit-exact against oracle/).
//
// Replaces torch_geometric.nn.pool.knn_graph -> torch_cluster.knn as called by the reference at
//   src/graphnet/models/components/layers.py:63-67      (latent-space recompute after every conv)
//   src/graphnet/models/graphs/edges/edges.py:74-78      (initial graph, KNNEdges)
//
// Semantics: for every query q of event b the candidates are all nodes of
// b, q included; the k+1 best under the total order (distance, index) are kept, distance =
// ((d0*d0 + d1*d1) + d2*d2 ...) in fp32 with every operation rounded (no FMA contraction);
// q itself is dropped afterwards. Output is a fixed-shape neighbour table nbr[N, k+1] (-1 padded,
// ascending distance) + deg[N], so no host synchronisation is needed to size a [2,E] tensor.
//
// Layout / mapping: one thread per query, 128 consecutive queries per CTA. The union of the
// events those queries belong to is a contiguous node range; it is streamed through shared memory
// in SoA chunks (coalesced fill, broadcast reads) and every thread scans only the part of each
// chunk that lies inside its own event, in ascending index order, so a strict '>' insertion keeps
// the lower index on ties exactly like the reference scan.
#include "common.cuh"

#define KNN_THREADS 128
#define KNN_MAX_K1 101   // torch_cluster asserts k <= 100
#define KNN_MAX_D 512

namespace {

__device__ __forceinline__ int find_segment(const int64_t* __restrict__ ptr, int nseg, int64_t q) {
    // largest b with ptr[b] <= q  (empty segments are skipped naturally)
    int lo = 0, hi = nseg;  // invariant: ptr[lo] <= q < ptr[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (ptr[mid] <= q) lo = mid; else hi = mid;
    }
    return lo;
}

template <int K1>
__device__ __forceinline__ void insert_static(float (&bd)[K1], int (&bi)[K1], float d, int j) {
    if (bd[K1 - 1] > d) {
#pragma unroll
        for (int e = K1 - 1; e >= 0; --e) {
            const bool gt = bd[e] > d;
            const bool prev_gt = (e > 0) ? (bd[e > 0 ? e - 1 : 0] > d) : false;
            if (gt) {
                bd[e] = prev_gt ? bd[e > 0 ? e - 1 : 0] : d;
                bi[e] = prev_gt ? bi[e > 0 ? e - 1 : 0] : j;
            }
        }
    }
}

template <int K1, int D>
__global__ void __launch_bounds__(KNN_THREADS)
knn_table_kernel(const float* __restrict__ x, int64_t ld, const int* __restrict__ cols, int d_rt,
                 const int64_t* __restrict__ ptr, int nseg, int64_t n, int k1_rt, int chunk,
                 int* __restrict__ nbr, int* __restrict__ deg) {
    gnb_pdl_begin();
    extern __shared__ float s_c[];            // [d][chunk]
    __shared__ int s_cols[KNN_MAX_D];
    __shared__ long long s_range[2];
    const int d = D ? D : d_rt;
    const int k1 = K1 ? K1 : k1_rt;
    const int tid = threadIdx.x;
    const int64_t q0 = (int64_t)blockIdx.x * KNN_THREADS;
    const int64_t q = q0 + tid;
    const bool active = q < n;
    for (int j = tid; j < d; j += KNN_THREADS) s_cols[j] = cols[j];

    int64_t lo = 0, hi = 0;
    if (active) {
        const int b = find_segment(ptr, nseg, q);
        lo = ptr[b];
        hi = ptr[b + 1];
    }
    if (tid == 0) s_range[0] = lo;
    const int64_t q_last = (q0 + KNN_THREADS < n ? q0 + KNN_THREADS : n) - 1;
    if (q == q_last) s_range[1] = hi;
    __syncthreads();
    const int64_t r_lo = s_range[0], r_hi = s_range[1];

    float qf[D ? D : 1];
    if (D && active) {
#pragma unroll
        for (int j = 0; j < (D ? D : 1); ++j) qf[j] = x[q * ld + s_cols[j]];
    }

    constexpr int KA = K1 ? K1 : KNN_MAX_K1;
    float bd[KA];
    int bi[KA];
#pragma unroll
    for (int e = 0; e < KA; ++e) { bd[e] = 1e10f; bi[e] = -1; }

    for (int64_t c0 = r_lo; c0 < r_hi; c0 += chunk) {
        const int cnt = (int)((r_hi - c0) < chunk ? (r_hi - c0) : chunk);
        __syncthreads();   // previous chunk fully consumed
        for (int idx = tid; idx < cnt * d; idx += KNN_THREADS) {
            const int j = idx / cnt, c = idx - j * cnt;
            s_c[j * chunk + c] = x[(c0 + c) * ld + s_cols[j]];
        }
        __syncthreads();
        if (active) {
            const int a = (int)((lo > c0 ? lo : c0) - c0);
            const int b = (int)((hi < c0 + cnt ? hi : c0 + cnt) - c0);
            int jj = a;
            if (D && K1) {
                // 4 candidates per iteration: the distance evaluations are independent (shared-memory latency
                // overlaps), the insertions stay in ascending candidate order so ties keep the lower index
                for (; jj + 4 <= b; jj += 4) {
                    float acc4[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const float d0 = s_c[jj + u] - qf[0];
                        float acc = __fmul_rn(d0, d0);
#pragma unroll
                        for (int j = 1; j < (D ? D : 1); ++j) {
                            const float dj = s_c[j * chunk + jj + u] - qf[j];
                            acc = __fadd_rn(acc, __fmul_rn(dj, dj));
                        }
                        acc4[u] = acc;
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) insert_static<KA>(bd, bi, acc4[u], (int)(c0 + jj + u));
                }
            }
            for (; jj < b; ++jj) {
                float acc;
                if (D) {
                    const float d0 = s_c[jj] - qf[0];
                    acc = __fmul_rn(d0, d0);
#pragma unroll
                    for (int j = 1; j < (D ? D : 1); ++j) {
                        const float dj = s_c[j * chunk + jj] - qf[j];
                        acc = __fadd_rn(acc, __fmul_rn(dj, dj));
                    }
                } else {
                    const float d0 = s_c[jj] - x[q * ld + s_cols[0]];
                    acc = __fmul_rn(d0, d0);
                    for (int j = 1; j < d; ++j) {
                        const float dj = s_c[j * chunk + jj] - x[q * ld + s_cols[j]];
                        acc = __fadd_rn(acc, __fmul_rn(dj, dj));
                    }
                }
                const int cand = (int)(c0 + jj);
                if (K1) {
                    insert_static<KA>(bd, bi, acc, cand);
                } else {
                    // runtime k: same strict '>' insertion as the reference scan
                    if (bd[k1 - 1] > acc) {
                        int e1 = 0;
                        while (!(bd[e1] > acc)) ++e1;
                        for (int e2 = k1 - 1; e2 > e1; --e2) { bd[e2] = bd[e2 - 1]; bi[e2] = bi[e2 - 1]; }
                        bd[e1] = acc;
                        bi[e1] = cand;
                    }
                }
            }
        }
    }
    if (active) {
        int cntd = 0;
        int* row = nbr + q * k1;
        if (K1) {
#pragma unroll
            for (int e = 0; e < KA; ++e)
                if (bi[e] >= 0 && bi[e] != (int)q) row[cntd++] = bi[e];
        } else {
            for (int e = 0; e < k1; ++e)
                if (bi[e] >= 0 && bi[e] != (int)q) row[cntd++] = bi[e];
        }
        for (int e = cntd; e < k1; ++e) row[e] = -1;
        deg[q] = cntd;
    }
}

// Split variant for the k = 8, 3-column graphs of DynEdge: S consecutive lanes share one query and scan disjoint subsets of
// the query's event in ascending index order; every lane keeps a sorted list that contains every member of the global top-9
// that lies in its subset (strict '>' insertion under ascending indices = the total order (distance, index)), then the S
// sorted lists are merged with S-lane shuffle minima under the same total order. Result is identical bit for bit.
//
// The scan is built around what the first version measured (ncu, profiles/r01/p_knn_table_split.ncu-rep: issue-bound, 65 SASS
// instructions per candidate, 27..45 of them the 9-entry insertion that a warp pays whenever ANY of its 32 lanes accepts a
// candidate -- and sum n^2 is dominated by the events with thousands of pulses, where a lane accepts a few per cent):
//   * a lane takes groups of 4 consecutive candidates (lane s: groups s, s + S, ...), so the coordinates arrive as three
//     16-byte shared-memory loads per group instead of twelve 4-byte loads, with 32-bit index arithmetic only;
//   * accepted candidates (distance below the lane's current 9th best) are only APPENDED to a per-lane FIFO in shared memory
//     (predicated stores, no branch); the warp runs the insertion loop when some lane's FIFO could overflow on the next
//     group, so its cost is shared by every pending entry of the warp instead of being paid per accepting lane. The FIFO
//     keeps ascending candidate order per lane, and a stale 9th best only lets extra candidates through (the insertion
//     re-checks), so the lists are the same as with immediate insertion;
//   * the S lanes of a query share a cut: the smallest of their 9th-best distances (refreshed by three shuffles after every
//     insertion round). Some lane already holds 9 candidates at or below the cut, so a candidate strictly above it cannot
//     be among the query's 9 best and is dropped without touching any list (a tie with the cut is kept: it may win on its
//     index). With 8 young lists per query a lane accepted 13 .. 60 % of its candidates; against the shared cut the accept
//     rate is that of one list over the whole event (2 .. 14 %). A lane's list is then no longer the exact top-9 of its
//     subset, but it still contains every member of the query's top-9 that lies in the subset, which is all the merge needs.
constexpr int KNN_CH = 1024;      // candidates per shared-memory chunk
constexpr int KNN_QC = 8;         // FIFO entries per lane (a group appends at most 4)

template <int K1>
__device__ __forceinline__ void knn_flush(float (&bd)[K1], int (&bi)[K1], const float (*s_qd)[KNN_THREADS],
                                          const int (*s_qi)[KNN_THREADS], int tid, int& qn) {
    const int m = __reduce_max_sync(0xffffffffu, qn);
    for (int e = 0; e < m; ++e)
        if (e < qn) insert_static<K1>(bd, bi, s_qd[e][tid], s_qi[e][tid]);
    qn = 0;
}

template <int K1, int D, int S>
__global__ void __launch_bounds__(KNN_THREADS)
knn_table_split_kernel(const float* __restrict__ x, int64_t ld, const int* __restrict__ cols,
                       const int64_t* __restrict__ ptr, int nseg, int64_t n, int64_t size_lo, int64_t size_hi,
                       int* __restrict__ nbr, int* __restrict__ deg) {
    gnb_pdl_begin();
    static_assert(D == 3 && (S & (S - 1)) == 0 && S <= 32, "scan is written for 3 coordinates; lanes per query a power of two");
    __shared__ __align__(16) float s_c[D][KNN_CH];
    __shared__ float s_qd[KNN_QC][KNN_THREADS];
    __shared__ int s_qi[KNN_QC][KNN_THREADS];
    __shared__ int s_cols[D];
    __shared__ long long s_range[2];
    constexpr int QPC = KNN_THREADS / S;      // queries per CTA
    const int tid = threadIdx.x;
    const int sub = tid % S;
    const int64_t q0 = (int64_t)blockIdx.x * QPC;
    const int64_t q = q0 + tid / S;
    bool active = q < n;
    if (tid < D) s_cols[tid] = cols[tid];
    if (tid == 0) { s_range[0] = 0x7fffffffffffffffLL; s_range[1] = 0; }
    __syncthreads();

    // this launch handles the queries whose event has size_lo <= pulses < size_hi (the other size class runs with another S)
    int64_t lo = 0, hi = 0;
    if (active) {
        const int b = find_segment(ptr, nseg, q);
        lo = ptr[b];
        hi = ptr[b + 1];
        active = hi - lo >= size_lo && hi - lo < size_hi;
        if (active && sub == 0) {
            atomicMin(reinterpret_cast<unsigned long long*>(&s_range[0]), (unsigned long long)lo);
            atomicMax(reinterpret_cast<unsigned long long*>(&s_range[1]), (unsigned long long)hi);
        }
    }
    __syncthreads();
    const int64_t r_lo = s_range[0], r_hi = s_range[1];     // union of the active queries' events (empty: r_lo > r_hi)
    if (r_lo >= r_hi) return;

    float qf[D];
#pragma unroll
    for (int j = 0; j < D; ++j) qf[j] = active ? x[q * ld + s_cols[j]] : 0.f;

    float bd[K1];
    int bi[K1];
#pragma unroll
    for (int e = 0; e < K1; ++e) { bd[e] = 1e10f; bi[e] = -1; }
    int qn = 0;
    float cut = 1e10f;        // smallest 9th best of the query's S lanes (as of the last flush)

    for (int64_t c0 = r_lo; c0 < r_hi; c0 += KNN_CH) {
        const int cnt = (int)((r_hi - c0) < KNN_CH ? (r_hi - c0) : KNN_CH);
        __syncthreads();   // previous chunk fully consumed
        for (int c = tid; c < cnt; c += KNN_THREADS) {       // one node per thread: its D coordinates share a sector
            const float* row = x + (c0 + c) * ld;
#pragma unroll
            for (int j = 0; j < D; ++j) s_c[j][c] = row[s_cols[j]];
        }
        __syncthreads();
        // the lane's part of the chunk, as positions relative to c0: [a, b) (empty when the event does not touch the chunk)
        int a = 0, b = 0;
        if (active) {
            a = (int)((lo > c0 ? lo : c0) - c0);
            b = (int)((hi < c0 + cnt ? hi : c0 + cnt) - c0);
            if (b <= a) a = b = 0;     // the event does not touch this chunk (positions beyond the chunk must never be formed)
        }
        const unsigned span = (unsigned)(b - a);
        const int gb = (b + 3) >> 2;                       // groups [a >> 2, gb)
        int g = (a >> 2) + sub;
        const int mine = gb > g ? (gb - g + S - 1) / S : 0;
        const int steps = __reduce_max_sync(0xffffffffu, mine);      // warp-uniform trip count: the flush votes with all lanes
        const int base = (int)c0;
        for (int t = 0; t < steps; ++t, g += S) {
            const bool gin = g < gb;
            const int p0 = gin ? (g << 2) : 0;
            const float4 cx = *reinterpret_cast<const float4*>(&s_c[0][p0]);
            const float4 cy = *reinterpret_cast<const float4*>(&s_c[1][p0]);
            const float4 cz = *reinterpret_cast<const float4*>(&s_c[2][p0]);
            const float vx[4] = {cx.x, cx.y, cx.z, cx.w}, vy[4] = {cy.x, cy.y, cy.z, cy.w}, vz[4] = {cz.x, cz.y, cz.z, cz.w};
            const float lim = bd[K1 - 1];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float d0 = vx[u] - qf[0], d1 = vy[u] - qf[1], d2 = vz[u] - qf[2];
                const float acc = __fadd_rn(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1)), __fmul_rn(d2, d2));
                const bool hit = gin && (unsigned)(p0 + u - a) < span && lim > acc && acc <= cut;
                if (hit) {
                    s_qd[qn][tid] = acc;
                    s_qi[qn][tid] = base + p0 + u;
                    ++qn;
                }
            }
            if (__any_sync(0xffffffffu, qn > KNN_QC - 4)) {
                knn_flush<K1>(bd, bi, s_qd, s_qi, tid, qn);
                cut = bd[K1 - 1];
#pragma unroll
                for (int off = 1; off < S; off <<= 1) cut = fminf(cut, __shfl_xor_sync(0xffffffffu, cut, off));
            }
        }
    }
    knn_flush<K1>(bd, bi, s_qd, s_qi, tid, qn);
    // merge the S sorted partial lists: K1 rounds of "smallest head under (distance, index)"; the owner pops its head
    int res[K1];
#pragma unroll
    for (int r = 0; r < K1; ++r) {
        float md = bd[0];
        int mi = bi[0];
#pragma unroll
        for (int off = 1; off < S; off <<= 1) {
            const float od = __shfl_xor_sync(0xffffffffu, md, off);
            const int oi = __shfl_xor_sync(0xffffffffu, mi, off);
            // empty entries are (1e10, -1): as unsigned, -1 loses every index tie
            if (od < md || (od == md && (unsigned)oi < (unsigned)mi)) { md = od; mi = oi; }
        }
        res[r] = mi;
        if (mi >= 0 && bi[0] == mi) {
#pragma unroll
            for (int e = 0; e + 1 < K1; ++e) { bd[e] = bd[e + 1]; bi[e] = bi[e + 1]; }
            bd[K1 - 1] = 1e10f;
            bi[K1 - 1] = -1;
        }
    }
    if (active && sub == 0) {
        int cntd = 0;
        int* row = nbr + q * K1;
#pragma unroll
        for (int e = 0; e < K1; ++e)
            if (res[e] >= 0 && res[e] != (int)q) row[cntd++] = res[e];
        for (int e = cntd; e < K1; ++e) row[e] = -1;
        deg[q] = cntd;
    }
}

// ptr[b] = first i with batch[i] >= b  (batch sorted ascending; ptr has nseg+1 entries)
__global__ void batch_to_ptr_kernel(const int64_t* __restrict__ batch, int64_t n, int64_t nseg,
                                    int64_t* __restrict__ ptr) {
    gnb_pdl_begin();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    const int64_t prev = (i == 0) ? -1 : batch[i - 1];
    const int64_t cur = (i == n) ? nseg : batch[i];
    for (int64_t b = prev + 1; b <= cur && b <= nseg; ++b) ptr[b] = i;
}

// edge_index[0, rowptr[q]+s] = nbr[q,s]; edge_index[1, ...] = q
__global__ void table_to_edge_index_kernel(const int* __restrict__ nbr, const int* __restrict__ deg,
                                           const int64_t* __restrict__ rowptr, int64_t n, int width,
                                           int64_t n_edges, int64_t* __restrict__ edge_index) {
    gnb_pdl_begin();
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t q = t / width;
    const int s = (int)(t - q * width);
    if (q >= n || s >= deg[q]) return;
    const int64_t e = rowptr[q] + s;
    edge_index[e] = nbr[q * width + s];
    edge_index[n_edges + e] = q;
}

template <int K1, int D>
int launch_knn(const float* x, int64_t ld, const int* cols, int d, const int64_t* ptr, int nseg, int64_t n,
               int k1, int* nbr, int* deg, cudaStream_t st) {
    int chunk = 8192 / d;
    if (chunk > 1024) chunk = 1024;
    if (chunk < 32) chunk = 32;
    const size_t smem = (size_t)chunk * d * sizeof(float);
    auto kern = knn_table_kernel<K1, D>;
    if (smem > 48 * 1024) GNB_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gnb_launch(kern, gnb_div_up(n, KNN_THREADS), KNN_THREADS, smem, st)(x, ld, cols, d, ptr, nseg, n, k1, chunk, nbr, deg);
    GNB_RETURN_LAUNCH();
}

// One launch with 8 lanes per query for every event. (Two size classes -- 2 lanes per query for events below 512 pulses, 8
// above, each launch skipping the other's queries -- were measured at 162 + 226 us per step against 345 us: no gain; the
// size window arguments stay for experiments.)
template <int K1, int D>
int launch_knn_split(const float* x, int64_t ld, const int* cols, const int64_t* ptr, int nseg, int64_t n, int* nbr, int* deg,
                     cudaStream_t st) {
    gnb_launch(knn_table_split_kernel<K1, D, 8>, gnb_div_up(n, KNN_THREADS / 8), KNN_THREADS, 0, st)(x, ld, cols, ptr, nseg, n, 0,
                                                                                              (int64_t)1 << 62, nbr, deg);
    GNB_RETURN_LAUNCH();
}

int g_knn_variant = 0;   // 0 auto (split kernel for k = 8, 3 columns), 1 one thread per query

}  // namespace

// Kernel selection for gnb_knn_table (tests pin both): 0 auto, 1 one thread per query, 2 split (k = 8, 3 columns only).
GNB_EXPORT int gnb_knn_set_variant(int32_t v) {
    if (v < 0 || v > 2) return GNB_ERR_ARG;
    g_knn_variant = v;
    return GNB_OK;
}

GNB_EXPORT int gnb_batch_to_ptr(const int64_t* batch, int64_t n, int64_t nseg, int64_t* ptr, void* stream) {
    if (n < 0 || nseg < 0) return GNB_ERR_ARG;
    gnb_launch(batch_to_ptr_kernel, gnb_div_up(n + 1, 256), 256, 0, (cudaStream_t)stream)(batch, n, nseg, ptr);
    GNB_RETURN_LAUNCH();
}

GNB_EXPORT int gnb_knn_table(const float* x, int64_t ld, const int32_t* cols, int32_t d, const int64_t* ptr,
                             int64_t nseg, int64_t n, int32_t k, int32_t* nbr, int32_t* deg, void* stream) {
    if (k < 1 || k + 1 > KNN_MAX_K1 || d < 1 || d > KNN_MAX_D || n < 0 || nseg < 0) return GNB_ERR_ARG;
    if (n == 0) return GNB_OK;
    if (n >= (int64_t)1 << 31) return GNB_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const int k1 = k + 1;
    if (d == 3) {
        if (k1 == 9 && g_knn_variant != 1) return launch_knn_split<9, 3>(x, ld, cols, ptr, (int)nseg, n, nbr, deg, st);
        if (k1 == 9) return launch_knn<9, 3>(x, ld, cols, d, ptr, (int)nseg, n, k1, nbr, deg, st);
        if (k1 == 5) return launch_knn<5, 3>(x, ld, cols, d, ptr, (int)nseg, n, k1, nbr, deg, st);
        if (k1 == 17) return launch_knn<17, 3>(x, ld, cols, d, ptr, (int)nseg, n, k1, nbr, deg, st);
        return launch_knn<0, 3>(x, ld, cols, d, ptr, (int)nseg, n, k1, nbr, deg, st);
    }
    if (k1 == 9) return launch_knn<9, 0>(x, ld, cols, d, ptr, (int)nseg, n, k1, nbr, deg, st);
    return launch_knn<0, 0>(x, ld, cols, d, ptr, (int)nseg, n, k1, nbr, deg, st);
}

GNB_EXPORT int gnb_table_to_edge_index(const int32_t* nbr, const int32_t* deg, const int64_t* rowptr, int64_t n,
                                       int32_t width, int64_t n_edges, int64_t* edge_index, void* stream) {
    if (n == 0 || n_edges == 0) return GNB_OK;
    gnb_launch(table_to_edge_index_kernel, gnb_div_up(n * width, 256), 256, 0, (cudaStream_t)stream)(
        nbr, deg, rowptr, n, width, n_edges, edge_index);
    GNB_RETURN_LAUNCH();
}
