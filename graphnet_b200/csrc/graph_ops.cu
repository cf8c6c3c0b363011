// HBM-bound graph operators of the DynEdge path: global variables, neighbour gathers, k-neighbour
// aggregation, global pooling and their backward passes. All are single-pass, coalesced (a warp
// always walks consecutive channels of one row) and use the fixed-width neighbour table
// nbr[N, W] / deg[N] written by knn.cu.
//
// Reference call sites replaced (paths relative to /root/reference):
//   global variables  src/graphnet/models/gnn/dynedge.py:266-293, 300-319 + models/utils.py:13-29
//   EdgeConv gather / aggregate  models/components/layers.py:60 (PyG EdgeConv.propagate)
//   global pooling    src/graphnet/models/gnn/dynedge.py:251-264 (torch_scatter.scatter_*)
#include "common.cuh"
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <float.h>

namespace {

constexpr int GV_THREADS = 512;
constexpr int GV_MAX_F = 32;

// One CTA per event: feature means, x/y/z/t homophily over the event's edges, log10(n_pulses);
// then the same CTA writes x0 = [x | g[event]] (zero padded to ld0) for the event's nodes, which
// replaces the reference's dense [N,B] "distribute" mask (dynedge.py:308-319) by a gather.
__global__ void __launch_bounds__(GV_THREADS)
global_vars_kernel(const float* __restrict__ x, int64_t ldx, int nf, const int* __restrict__ nbr,
                   const int* __restrict__ deg, int width, const int64_t* __restrict__ ptr,
                   const float* __restrict__ n_pulses, float* __restrict__ g, float* __restrict__ x0,
                   int64_t ld0, int x0_cols) {
    gnb_pdl_begin();
    const int b = blockIdx.x;
    const int64_t lo = ptr[b], hi = ptr[b + 1];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = GV_THREADS / 32;
    __shared__ float s_part[NW][GV_MAX_F + 5];
    __shared__ float s_g[GV_MAX_F + 5];
    float acc[GV_MAX_F];
#pragma unroll
    for (int f = 0; f < GV_MAX_F; ++f) acc[f] = 0.f;
    float same[4] = {0.f, 0.f, 0.f, 0.f};
    float edges = 0.f;
    for (int64_t i = lo + tid; i < hi; i += GV_THREADS) {
        const float* xi = x + i * ldx;
#pragma unroll
        for (int f = 0; f < GV_MAX_F; ++f)
            if (f < nf) acc[f] += xi[f];
        const int dg = deg[i];
        const float xi0 = xi[0], xi1 = xi[1], xi2 = xi[2], xi3 = xi[3];
        // the neighbour indices first, then all gathers: 4 slots at a time are independent loads in flight (one slot
        // per iteration made every gather wait for the previous compare: the largest event set the kernel's time)
        for (int s0 = 0; s0 < dg; s0 += 4) {
            int nb[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) nb[u] = s0 + u < dg ? nbr[i * width + s0 + u] : -1;
            float v[4][4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float* xj = x + (int64_t)(nb[u] < 0 ? i : nb[u]) * ldx;
#pragma unroll
                for (int c = 0; c < 4; ++c) v[u][c] = xj[c];
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (nb[u] >= 0) {
                    same[0] += (v[u][0] == xi0) ? 1.f : 0.f; same[1] += (v[u][1] == xi1) ? 1.f : 0.f;
                    same[2] += (v[u][2] == xi2) ? 1.f : 0.f; same[3] += (v[u][3] == xi3) ? 1.f : 0.f;
                }
            }
        }
        edges += (float)dg;
    }
    // block reduction: warp shuffles then one smem row per warp
#pragma unroll
    for (int f = 0; f < GV_MAX_F; ++f) {
        if (f < nf) {
            const float v = gnb_warp_sum(acc[f]);
            if (lane == 0) s_part[warp][f] = v;
        }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const float v = gnb_warp_sum(same[c]);
        if (lane == 0) s_part[warp][GV_MAX_F + c] = v;
    }
    {
        const float v = gnb_warp_sum(edges);
        if (lane == 0) s_part[warp][GV_MAX_F + 4] = v;
    }
    __syncthreads();
    const int ng = nf + 5;
    if (tid < nf) {
        float v = 0.f;
        for (int w = 0; w < NW; ++w) v += s_part[w][tid];
        const float cnt = (float)(hi - lo);
        s_g[tid] = v / (cnt < 1.f ? 1.f : cnt);
    } else if (tid < nf + 4) {
        const int c = tid - nf;
        float v = 0.f, e = 0.f;
        for (int w = 0; w < NW; ++w) { v += s_part[w][GV_MAX_F + c]; e += s_part[w][GV_MAX_F + 4]; }
        s_g[tid] = v / (e < 1.f ? 1.f : e);
    } else if (tid == nf + 4) {
        s_g[tid] = log10f(n_pulses[b]);
    }
    __syncthreads();
    if (tid < ng) g[(int64_t)b * ng + tid] = s_g[tid];
    if (x0 != nullptr && !(ld0 & 3) && ld0 <= 4 * GV_THREADS && !(reinterpret_cast<uintptr_t>(x0) & 15u)) {
        // one float4 of a row per thread and step: no integer division in the loop, 16-byte coalesced stores
        const int chunks = (int)(ld0 >> 2), rows_per_step = GV_THREADS / chunks;
        const int ch = tid % chunks, r0 = tid / chunks;
        if (r0 < rows_per_step) {
            // 4 rows per thread and step: the loads of x are in flight together (one row per step made the largest event's
            // load -> store chain the kernel's time)
            for (int64_t ib = lo + r0; ib < hi; ib += 4 * rows_per_step) {
                float v[4][4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int64_t i = ib + (int64_t)u * rows_per_step;
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int c = 4 * ch + e;
                        v[u][e] = c < nf ? (i < hi ? x[i * ldx + c] : 0.f) : (c < x0_cols ? s_g[c - nf] : 0.f);
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int64_t i = ib + (int64_t)u * rows_per_step;
                    if (i < hi) reinterpret_cast<float4*>(x0 + i * ld0)[ch] = make_float4(v[u][0], v[u][1], v[u][2], v[u][3]);
                }
            }
        }
    } else if (x0 != nullptr) {
        const int64_t total = (hi - lo) * ld0;
        for (int64_t t = tid; t < total; t += GV_THREADS) {
            const int64_t i = lo + t / ld0;
            const int c = (int)(t % ld0);
            float v = 0.f;
            if (c < nf) v = x[i * ldx + c];
            else if (c < x0_cols) v = s_g[c - nf];
            x0[i * ld0 + c] = v;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// hoisted EdgeConv hidden layer: row r = (i, s) of the padded edge list,
//   h[r, :] = act(P[i, :] + Q[nbr[i,s], :])   (zero row when s >= deg[i])
// PQ is [N, 2H]: P = columns [0,H) (already includes the bias), Q = columns [H, 2H).
// One warp per row, float4 along the channel dimension.
__global__ void edge_hidden_fwd_kernel(const float* __restrict__ pq, int64_t ldpq, int hdim,
                                       const int* __restrict__ nbr, const int* __restrict__ deg, int width,
                                       int64_t n, int act, float* __restrict__ h, int64_t ldh,
                                       unsigned* __restrict__ hmask, int mask_ld) {
    gnb_pdl_begin();
    const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= n * width) return;
    const int lane = threadIdx.x & 31;
    const int64_t i = r / width;
    const int s = (int)(r - i * width);
    float4* out = reinterpret_cast<float4*>(h + r * ldh);
    const int h4 = hdim >> 2;
    // optional bit mask of the activation, mask_ld words per row: channel c = 128 it + 4 L + k (float4 chunk L of
    // iteration it, component k) is bit L of word 4 it + k -- one __ballot_sync per component, no shuffles
    unsigned* mrow = hmask != nullptr ? hmask + r * mask_ld : nullptr;
    if (s >= deg[i]) {
        for (int c = lane; c < h4; c += 32) out[c] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (mrow != nullptr)
            for (int w = lane; w < mask_ld; w += 32) mrow[w] = 0u;
        return;
    }
    const int64_t j = nbr[i * width + s];
    const float4* p = reinterpret_cast<const float4*>(pq + i * ldpq);
    const float4* q = reinterpret_cast<const float4*>(pq + j * ldpq + hdim);
    const bool relu = (act & 0xff) == GNB_ACT_RELU, leaky = (act & 0xff) == GNB_ACT_LEAKY, rnd = (act & GNB_FLAG_ROUND_TF32) != 0;
    if (mrow == nullptr) {
        for (int c = lane; c < h4; c += 32) {
            const float4 a = p[c], b = q[c];
            float4 v = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
            if (relu) {
                v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
            }
            if (leaky) {
                v.x = v.x > 0.f ? v.x : GNB_LEAKY_SLOPE * v.x; v.y = v.y > 0.f ? v.y : GNB_LEAKY_SLOPE * v.y;
                v.z = v.z > 0.f ? v.z : GNB_LEAKY_SLOPE * v.z; v.w = v.w > 0.f ? v.w : GNB_LEAKY_SLOPE * v.w;
            }
            if (rnd) {
                v.x = gnb_round_tf32(v.x); v.y = gnb_round_tf32(v.y); v.z = gnb_round_tf32(v.z); v.w = gnb_round_tf32(v.w);
            }
            out[c] = v;
        }
        return;
    }
    // all loads first (independent), then the arithmetic, the stores and the nibble -> word shuffles
    float4 va[4], vb[4];
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        const int c = it * 32 + lane;
        if (it * 4 < mask_ld && c < h4) { va[it] = p[c]; vb[it] = q[c]; }
    }
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        if (it * 4 < mask_ld) {                               // warp-uniform
            const int c = it * 32 + lane;
            bool px = false, py = false, pz = false, pw = false;
            if (c < h4) {
                float4 v = make_float4(va[it].x + vb[it].x, va[it].y + vb[it].y, va[it].z + vb[it].z, va[it].w + vb[it].w);
                if (relu) {
                    v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
                }
                if (leaky) {
                    v.x = v.x > 0.f ? v.x : GNB_LEAKY_SLOPE * v.x; v.y = v.y > 0.f ? v.y : GNB_LEAKY_SLOPE * v.y;
                    v.z = v.z > 0.f ? v.z : GNB_LEAKY_SLOPE * v.z; v.w = v.w > 0.f ? v.w : GNB_LEAKY_SLOPE * v.w;
                }
                if (rnd) {
                    v.x = gnb_round_tf32(v.x); v.y = gnb_round_tf32(v.y); v.z = gnb_round_tf32(v.z); v.w = gnb_round_tf32(v.w);
                }
                out[c] = v;
                px = v.x > 0.f; py = v.y > 0.f; pz = v.z > 0.f; pw = v.w > 0.f;
            }
            const unsigned bx = __ballot_sync(0xffffffffu, px), by = __ballot_sync(0xffffffffu, py);
            const unsigned bz = __ballot_sync(0xffffffffu, pz), bw = __ballot_sync(0xffffffffu, pw);
            if (lane == 0) *reinterpret_cast<uint4*>(mrow + it * 4) = make_uint4(bx, by, bz, bw);
        }
    }
}

// Same operation, one warp per TARGET NODE: P_i, deg and the neighbour row are read once, the Q gathers of three slots
// are in flight together, and every slot row is written with 512-byte warp stores. NIT = ceil(hdim / 128) float4
// iterations per row. MASK adds the activation bits (layout above).
template <int NIT, bool MASK>
__global__ void __launch_bounds__(256)
edge_hidden_fwd_node_kernel(const float* __restrict__ pq, int64_t ldpq, int hdim, const int* __restrict__ nbr,
                            const int* __restrict__ deg, int width, int64_t n, int act, float* __restrict__ h, int64_t ldh,
                            unsigned* __restrict__ hmask, int mask_ld) {
    gnb_pdl_begin();
    const int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= n) return;
    const int lane = threadIdx.x & 31;
    const int h4 = hdim >> 2;
    const int dg = deg[i];
    const int nb = lane < width ? nbr[i * width + lane] : -1;
    const bool relu = (act & 0xff) == GNB_ACT_RELU, leaky = (act & 0xff) == GNB_ACT_LEAKY, rnd = (act & GNB_FLAG_ROUND_TF32) != 0;
    const float4* p = reinterpret_cast<const float4*>(pq + i * ldpq);
    float4 pa[NIT];
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
        const int c = it * 32 + lane;
        pa[it] = c < h4 ? p[c] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int s0 = 0; s0 < width; s0 += 3) {
        float4 qv[3][NIT];
#pragma unroll
        for (int u = 0; u < 3; ++u) {
            const int s = s0 + u;
            const int j = __shfl_sync(0xffffffffu, nb, s < 32 ? s : 0);
            const bool valid = s < width && s < dg && j >= 0;
            const float4* q = reinterpret_cast<const float4*>(pq + (int64_t)(valid ? j : 0) * ldpq + hdim);
#pragma unroll
            for (int it = 0; it < NIT; ++it) {
                const int c = it * 32 + lane;
                qv[u][it] = (valid && c < h4) ? q[c] : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
            }
        }
#pragma unroll
        for (int u = 0; u < 3; ++u) {
            const int s = s0 + u;
            if (s < width) {                                  // warp-uniform
                const int64_t r = i * width + s;
                float4* out = reinterpret_cast<float4*>(h + r * ldh);
                const bool valid = s < dg && __shfl_sync(0xffffffffu, nb, s < 32 ? s : 0) >= 0;
#pragma unroll
                for (int it = 0; it < NIT; ++it) {
                    const int c = it * 32 + lane;
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (valid) {
                        v = make_float4(pa[it].x + qv[u][it].x, pa[it].y + qv[u][it].y, pa[it].z + qv[u][it].z,
                                        pa[it].w + qv[u][it].w);
                        if (relu) {
                            v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
                        }
                        if (leaky) {
                            v.x = v.x > 0.f ? v.x : GNB_LEAKY_SLOPE * v.x; v.y = v.y > 0.f ? v.y : GNB_LEAKY_SLOPE * v.y;
                            v.z = v.z > 0.f ? v.z : GNB_LEAKY_SLOPE * v.z; v.w = v.w > 0.f ? v.w : GNB_LEAKY_SLOPE * v.w;
                        }
                        if (rnd) {
                            v.x = gnb_round_tf32(v.x); v.y = gnb_round_tf32(v.y);
                            v.z = gnb_round_tf32(v.z); v.w = gnb_round_tf32(v.w);
                        }
                    }
                    if (c < h4) out[c] = v;
                    if (MASK) {
                        const bool in = c < h4;
                        const unsigned bx = __ballot_sync(0xffffffffu, in && v.x > 0.f), by = __ballot_sync(0xffffffffu, in && v.y > 0.f);
                        const unsigned bz = __ballot_sync(0xffffffffu, in && v.z > 0.f), bw = __ballot_sync(0xffffffffu, in && v.w > 0.f);
                        if (lane == 0 && it * 4 < mask_ld)
                            *reinterpret_cast<uint4*>(hmask + r * mask_ld + it * 4) = make_uint4(bx, by, bz, bw);
                    }
                }
                if (MASK)                                     // words beyond the last iteration (mask_ld > 4 NIT never happens
                    for (int w = 4 * NIT + lane; w < mask_ld; w += 32) hmask[r * mask_ld + w] = 0u;   // for mask_ld = 4 ceil(hdim/128))
            }
        }
    }
}

template <bool MASK>
static void launch_hidden_node(int nit, dim3 grid, cudaStream_t st, const float* pq, int64_t ldpq, int hdim, const int* nbr,
                               const int* deg, int width, int64_t n, int act, float* h, int64_t ldh, unsigned* hmask,
                               int mask_ld) {
    switch (nit) {
        case 1: gnb_launch(edge_hidden_fwd_node_kernel<1, MASK>, grid, 256, 0, st)(pq, ldpq, hdim, nbr, deg, width, n, act, h, ldh, hmask, mask_ld); break;
        case 2: gnb_launch(edge_hidden_fwd_node_kernel<2, MASK>, grid, 256, 0, st)(pq, ldpq, hdim, nbr, deg, width, n, act, h, ldh, hmask, mask_ld); break;
        case 3: gnb_launch(edge_hidden_fwd_node_kernel<3, MASK>, grid, 256, 0, st)(pq, ldpq, hdim, nbr, deg, width, n, act, h, ldh, hmask, mask_ld); break;
        default: gnb_launch(edge_hidden_fwd_node_kernel<4, MASK>, grid, 256, 0, st)(pq, ldpq, hdim, nbr, deg, width, n, act, h, ldh, hmask, mask_ld); break;
    }
}


// The same hidden layer written as bf16 PLANES for the bf16 / bf16x3 precision modes (csrc/gemm_tc.cu, MODE 2 / 3):
// h0 = bf16(h), h1 = bf16(h - h0) (NP = 2), each [n * width, ldh] bf16. The activation bits are taken from the fp32 value.
__device__ __forceinline__ uint2 pack_bf16x4(float a, float b, float c, float d) {
    const __nv_bfloat162 lo = __floats2bfloat162_rn(a, b), hi = __floats2bfloat162_rn(c, d);
    return make_uint2(*reinterpret_cast<const unsigned*>(&lo), *reinterpret_cast<const unsigned*>(&hi));
}
__device__ __forceinline__ float bf16_lo_f(unsigned w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi_f(unsigned w) { return __uint_as_float(w & 0xFFFF0000u); }

// fp16 planes of four scaled values and of their remainders (mode mixed16): p0 = fp16(v s), p1 = fp16(v s - p0)
__device__ __forceinline__ uint2 pack_f16x4(float a, float b, float c, float d) {
    const __half2 lo = __floats2half2_rn(a, b), hi = __floats2half2_rn(c, d);
    return make_uint2(*reinterpret_cast<const unsigned*>(&lo), *reinterpret_cast<const unsigned*>(&hi));
}
__device__ __forceinline__ float2 unpack_f16x2(unsigned w) { return __half22float2(*reinterpret_cast<const __half2*>(&w)); }

// F16: the planes are fp16 of h * scale, scale = gnb_pow2_scale(*scale_bits).x with *scale_bits >= fp32 bits of max h.
template <int NIT, int NP, bool F16>
__global__ void __launch_bounds__(256)
edge_hidden_fwd_node_bf16_kernel(const float* __restrict__ pq, int64_t ldpq, int hdim, const int* __restrict__ nbr,
                                 const int* __restrict__ deg, int width, int64_t n, __nv_bfloat16* __restrict__ h0,
                                 __nv_bfloat16* __restrict__ h1, int64_t ldh, unsigned* __restrict__ hmask, int mask_ld,
                                 const unsigned* __restrict__ scale_bits) {
    gnb_pdl_begin();
    const float scale = F16 ? gnb_pow2_scale(*scale_bits).x : 1.f;
    const int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= n) return;
    const int lane = threadIdx.x & 31;
    const int h4 = hdim >> 2;
    const int dg = deg[i];
    const int nb = lane < width ? nbr[i * width + lane] : -1;
    const float4* p = reinterpret_cast<const float4*>(pq + i * ldpq);
    float4 pa[NIT];
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
        const int c = it * 32 + lane;
        pa[it] = c < h4 ? p[c] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int s0 = 0; s0 < width; s0 += 3) {
        float4 qv[3][NIT];
#pragma unroll
        for (int u = 0; u < 3; ++u) {
            const int s = s0 + u;
            const int j = __shfl_sync(0xffffffffu, nb, s < 32 ? s : 0);
            const bool valid = s < width && s < dg && j >= 0;
            const float4* q = reinterpret_cast<const float4*>(pq + (int64_t)(valid ? j : 0) * ldpq + hdim);
#pragma unroll
            for (int it = 0; it < NIT; ++it) {
                const int c = it * 32 + lane;
                qv[u][it] = (valid && c < h4) ? q[c] : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
            }
        }
#pragma unroll
        for (int u = 0; u < 3; ++u) {
            const int s = s0 + u;
            if (s < width) {                                  // warp-uniform
                const int64_t r = i * width + s;
                uint2* o0 = reinterpret_cast<uint2*>(h0 + r * ldh);
                uint2* o1 = NP == 2 ? reinterpret_cast<uint2*>(h1 + r * ldh) : nullptr;
                const bool valid = s < dg && __shfl_sync(0xffffffffu, nb, s < 32 ? s : 0) >= 0;
#pragma unroll
                for (int it = 0; it < NIT; ++it) {
                    const int c = it * 32 + lane;
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (valid)
                        v = make_float4(fmaxf(pa[it].x + qv[u][it].x, 0.f), fmaxf(pa[it].y + qv[u][it].y, 0.f),
                                        fmaxf(pa[it].z + qv[u][it].z, 0.f), fmaxf(pa[it].w + qv[u][it].w, 0.f));
                    if (F16) {
                        const float4 sv = make_float4(v.x * scale, v.y * scale, v.z * scale, v.w * scale);
                        const uint2 b0 = pack_f16x4(sv.x, sv.y, sv.z, sv.w);
                        if (c < h4) {
                            o0[c] = b0;
                            if (NP == 2) {
                                const float2 f01 = unpack_f16x2(b0.x), f23 = unpack_f16x2(b0.y);
                                o1[c] = pack_f16x4(sv.x - f01.x, sv.y - f01.y, sv.z - f23.x, sv.w - f23.y);
                            }
                        }
                    } else {
                        const uint2 b0 = pack_bf16x4(v.x, v.y, v.z, v.w);
                        if (c < h4) {
                            o0[c] = b0;
                            if (NP == 2)
                                o1[c] = pack_bf16x4(v.x - bf16_lo_f(b0.x), v.y - bf16_hi_f(b0.x), v.z - bf16_lo_f(b0.y), v.w - bf16_hi_f(b0.y));
                        }
                    }
                    if (hmask != nullptr) {
                        const bool in = c < h4;
                        const unsigned bx = __ballot_sync(0xffffffffu, in && v.x > 0.f), by = __ballot_sync(0xffffffffu, in && v.y > 0.f);
                        const unsigned bz = __ballot_sync(0xffffffffu, in && v.z > 0.f), bw = __ballot_sync(0xffffffffu, in && v.w > 0.f);
                        if (lane == 0 && it * 4 < mask_ld)
                            *reinterpret_cast<uint4*>(hmask + r * mask_ld + it * 4) = make_uint4(bx, by, bz, bw);
                    }
                }
                if (hmask != nullptr)
                    for (int w = 4 * NIT + lane; w < mask_ld; w += 32) hmask[r * mask_ld + w] = 0u;
            }
        }
    }
}

template <int NP, bool F16>
static void launch_hidden_node_bf16(int nit, dim3 grid, cudaStream_t st, const float* pq, int64_t ldpq, int hdim, const int* nbr,
                                    const int* deg, int width, int64_t n, __nv_bfloat16* h0, __nv_bfloat16* h1, int64_t ldh,
                                    unsigned* hmask, int mask_ld, const unsigned* sb) {
    switch (nit) {
        case 1: gnb_launch(edge_hidden_fwd_node_bf16_kernel<1, NP, F16>, grid, 256, 0, st)(pq, ldpq, hdim, nbr, deg, width, n, h0, h1, ldh, hmask, mask_ld, sb); break;
        case 2: gnb_launch(edge_hidden_fwd_node_bf16_kernel<2, NP, F16>, grid, 256, 0, st)(pq, ldpq, hdim, nbr, deg, width, n, h0, h1, ldh, hmask, mask_ld, sb); break;
        case 3: gnb_launch(edge_hidden_fwd_node_bf16_kernel<3, NP, F16>, grid, 256, 0, st)(pq, ldpq, hdim, nbr, deg, width, n, h0, h1, ldh, hmask, mask_ld, sb); break;
        default: gnb_launch(edge_hidden_fwd_node_bf16_kernel<4, NP, F16>, grid, 256, 0, st)(pq, ldpq, hdim, nbr, deg, width, n, h0, h1, ldh, hmask, mask_ld, sb); break;
    }
}

// backward of the above: da1 = gh * act'(h); dPQ[i, 0:H] = sum_s da1[(i,s)] (own rows, no atomics);
// dPQ[j, H:2H] += da1[(i,s)] (scatter to the source node: vector atomics, 32 consecutive channels
// per warp instruction). dPQ's Q half must be zero on entry. One warp per target node.
__global__ void edge_hidden_bwd_kernel(const float* __restrict__ gh, int64_t ldg, const float* __restrict__ h,
                                       int64_t ldh, int hdim, const int* __restrict__ nbr,
                                       const int* __restrict__ deg, int width, int64_t n, int act,
                                       float* __restrict__ dpq, int64_t ldpq) {
    gnb_pdl_begin();
    const int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= n) return;
    const int lane = threadIdx.x & 31;
    const int dg = deg[i];
    const int h4 = hdim >> 2;
    float4* dp = reinterpret_cast<float4*>(dpq + i * ldpq);
    for (int c = lane; c < h4; c += 32) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int s = 0; s < dg; ++s) {
            const int64_t r = i * width + s;
            float4 g = reinterpret_cast<const float4*>(gh + r * ldg)[c];
            if ((act & 0xff) == GNB_ACT_RELU) {
                const float4 hv = reinterpret_cast<const float4*>(h + r * ldh)[c];
                g.x = hv.x > 0.f ? g.x : 0.f; g.y = hv.y > 0.f ? g.y : 0.f;
                g.z = hv.z > 0.f ? g.z : 0.f; g.w = hv.w > 0.f ? g.w : 0.f;
            }
            if ((act & 0xff) == GNB_ACT_LEAKY) {       // h = leaky(a): h > 0 <=> a > 0
                const float4 hv = reinterpret_cast<const float4*>(h + r * ldh)[c];
                g.x = hv.x > 0.f ? g.x : GNB_LEAKY_SLOPE * g.x; g.y = hv.y > 0.f ? g.y : GNB_LEAKY_SLOPE * g.y;
                g.z = hv.z > 0.f ? g.z : GNB_LEAKY_SLOPE * g.z; g.w = hv.w > 0.f ? g.w : GNB_LEAKY_SLOPE * g.w;
            }
            acc.x += g.x; acc.y += g.y; acc.z += g.z; acc.w += g.w;
            const int64_t j = nbr[r];
            atomicAdd(reinterpret_cast<float4*>(dpq + j * ldpq + hdim) + c, g);
        }
        dp[c] = acc;
    }
}

// ---------------------------------------------------------------------------------------------
// generic EdgeConv message input: u[(i,s), :] = [x_i | x_j - x_i]  (zero row when s >= deg[i])
__global__ void edge_cat_fwd_kernel(const float* __restrict__ x, int64_t ldx, int c_in,
                                    const int* __restrict__ nbr, const int* __restrict__ deg, int width,
                                    int64_t n, float* __restrict__ u, int64_t ldu) {
    gnb_pdl_begin();
    const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= n * width) return;
    const int lane = threadIdx.x & 31;
    const int64_t i = r / width;
    const int s = (int)(r - i * width);
    float* out = u + r * ldu;
    if (s >= deg[i]) {
        for (int c = lane; c < 2 * c_in; c += 32) out[c] = 0.f;
        return;
    }
    const int64_t j = nbr[i * width + s];
    for (int c = lane; c < c_in; c += 32) {
        const float a = x[i * ldx + c], b = x[j * ldx + c];
        out[c] = a;
        out[c_in + c] = b - a;
    }
}

// backward: dx[i] += sum_s (du_a - du_b); dx[j] += du_b.  dx must be zero on entry.
__global__ void edge_cat_bwd_kernel(const float* __restrict__ du, int64_t ldu, int c_in,
                                    const int* __restrict__ nbr, const int* __restrict__ deg, int width,
                                    int64_t n, float* __restrict__ dx, int64_t ldx) {
    gnb_pdl_begin();
    const int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= n) return;
    const int lane = threadIdx.x & 31;
    const int dg = deg[i];
    for (int c = lane; c < c_in; c += 32) {
        float acc = 0.f;
        for (int s = 0; s < dg; ++s) {
            const int64_t r = i * width + s;
            const float ga = du[r * ldu + c], gb = du[r * ldu + c_in + c];
            acc += ga - gb;
            atomicAdd(dx + (int64_t)nbr[r] * ldx + c, gb);
        }
        if (dg > 0) atomicAdd(dx + i * ldx + c, acc);
    }
}

// ---------------------------------------------------------------------------------------------
// k-neighbour aggregation over the padded edge list: y[i,:] = AGG_{s<deg[i]} m[(i,s),:]
// add / mean / max (max also writes the arg slot, first occurrence; empty neighbourhood -> 0).
__global__ void edge_aggregate_fwd_kernel(const float* __restrict__ m, int64_t ldm, int c_out,
                                          const int* __restrict__ deg, int width, int64_t n, int aggr,
                                          float* __restrict__ y, int64_t ldy, int8_t* __restrict__ arg) {
    gnb_pdl_begin();
    const int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= n) return;
    const int lane = threadIdx.x & 31;
    const int dg = deg[i];
    const bool rnd = (aggr & GNB_FLAG_ROUND_TF32) != 0;
    aggr &= 0xff;
    for (int c = lane; c < c_out; c += 32) {
        float acc = 0.f;
        int best = -1;
        if (aggr == GNB_AGGR_MAX) {
            if (dg > 0) { acc = m[(i * width) * ldm + c]; best = 0; }
            for (int s = 1; s < dg; ++s) {
                const float v = m[(i * width + s) * ldm + c];
                if (v > acc) { acc = v; best = s; }
            }
            if (arg) arg[i * c_out + c] = (int8_t)best;
        } else {
            for (int s = 0; s < dg; ++s) acc += m[(i * width + s) * ldm + c];
            if (aggr == GNB_AGGR_MEAN && dg > 0) acc = acc / (float)dg;
        }
        y[i * ldy + c] = rnd ? gnb_round_tf32(acc) : acc;
    }
}

__global__ void edge_aggregate_bwd_kernel(const float* __restrict__ gy, int64_t ldy, int c_out,
                                          const int* __restrict__ deg, int width, int64_t n, int aggr,
                                          const int8_t* __restrict__ arg, float* __restrict__ gm, int64_t ldm) {
    gnb_pdl_begin();
    const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= n * width) return;
    const int lane = threadIdx.x & 31;
    const int64_t i = r / width;
    const int s = (int)(r - i * width);
    const int dg = deg[i];
    for (int c = lane; c < c_out; c += 32) {
        float v = 0.f;
        if (s < dg) {
            const float g = gy[i * ldy + c];
            if (aggr == GNB_AGGR_ADD) v = g;
            else if (aggr == GNB_AGGR_MEAN) v = g / (float)dg;
            else v = (arg[i * c_out + c] == s) ? g : 0.f;
        }
        gm[r * ldm + c] = v;
    }
}

// Backward of the max-aggregating GEMM epilogue (gnb_edge_linear_aggmax_fwd_*): y[i, c] = max_s act(pre[(i,s), c]) sends its
// gradient to ONE slot, the arg-max (torch_scatter's scatter_max rule; first slot on ties), through the activation's slope
// there: dz[(i,s), c] = (s == slot(arg[i, c])) ? gy[i, c] * (pos(arg[i, c]) ? 1 : slope) : 0, where arg = slot | 0x40 if the
// winning pre-activation was > 0, -1 for a node without neighbours. db[c] += column sums of dz (may be NULL). One warp per
// node; a CTA reduces its bias contributions in shared memory before the global atomics.
constexpr int AMB_NODES = 8;      // nodes per warp
__global__ void __launch_bounds__(256)
edge_argmax_bwd_kernel(const float* __restrict__ gy, int64_t ldy, const int8_t* __restrict__ arg, int64_t ldarg, int c_out,
                       int width, int64_t n, float slope, int rnd, float* __restrict__ dz, int64_t ldz, float* __restrict__ db) {
    gnb_pdl_begin();
    __shared__ float s_db[512];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (db != nullptr) {
        for (int c = threadIdx.x; c < 512; c += 256) s_db[c] = 0.f;
        __syncthreads();
    }
    const int64_t i0 = ((int64_t)blockIdx.x * 8 + warp) * AMB_NODES;
    for (int c = lane; c < c_out; c += 32) {
        float colacc = 0.f;
        for (int f = 0; f < AMB_NODES; ++f) {
            const int64_t i = i0 + f;
            if (i >= n) break;
            const int a = arg[i * ldarg + c];
            float d = 0.f;
            if (a >= 0) {
                d = gy[i * ldy + c] * ((a & 0x40) ? 1.f : slope);
                colacc += d;
                if (rnd) d = gnb_round_tf32(d);
            }
            const int slot = a >= 0 ? (a & 0x3f) : -1;
            for (int s = 0; s < width; ++s) dz[(i * width + s) * ldz + c] = s == slot ? d : 0.f;
        }
        if (db != nullptr) atomicAdd(&s_db[c & 511], colacc);
    }
    if (db != nullptr) {
        __syncthreads();
        for (int c = threadIdx.x; c < c_out && c < 512; c += 256) atomicAdd(db + c, s_db[c]);
    }
}

// ---------------------------------------------------------------------------------------------
// global pooling: [N, C] -> [B, P*C] for up to 4 schemes in caller order, single read of x.
// CTA = (64 channels) x (16 row lanes); grid = (B, C/64). arg (int32 node index, -1 for sum/mean or an
// empty event) is kept for the min/max backward; ties resolve to the lowest node index.
constexpr int POOL_CX = 64, POOL_RY = 16;   // 16 row lanes: the kernel's time is the largest event's row count / POOL_RY

__global__ void __launch_bounds__(POOL_CX * POOL_RY)
segment_pool_fwd_kernel(const float* __restrict__ x, int64_t ldx, int c_tot, const int64_t* __restrict__ ptr,
                        int np, int s0, int s1, int s2, int s3, float* __restrict__ out,
                        int* __restrict__ arg) {
    gnb_pdl_begin();
    const int b = blockIdx.x;
    const int c = blockIdx.y * POOL_CX + threadIdx.x;
    const int ty = threadIdx.y;
    const int64_t lo = ptr[b], hi = ptr[b + 1];
    __shared__ float s_min[POOL_RY][POOL_CX], s_max[POOL_RY][POOL_CX], s_sum[POOL_RY][POOL_CX];
    __shared__ int s_amin[POOL_RY][POOL_CX], s_amax[POOL_RY][POOL_CX];
    float vmin = FLT_MAX, vmax = -FLT_MAX, vsum = 0.f;
    int amin = -1, amax = -1;
    if (c < c_tot) {
        // 4 rows per step: the loads are issued together, the (ordered) compare chain follows -- with one load per step
        // every load waited for the previous row's compares and the largest event set the kernel's time
        for (int64_t i = lo + ty; i < hi; i += 4 * POOL_RY) {
            float v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = i + u * POOL_RY < hi ? x[(i + u * POOL_RY) * ldx + c] : 0.f;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (i + u * POOL_RY < hi) {
                    const int idx = (int)(i + u * POOL_RY);
                    vsum += v[u];
                    if (amin < 0 || v[u] < vmin) { vmin = v[u]; amin = idx; }
                    if (amax < 0 || v[u] > vmax) { vmax = v[u]; amax = idx; }
                }
            }
        }
    }
    s_min[ty][threadIdx.x] = vmin; s_max[ty][threadIdx.x] = vmax; s_sum[ty][threadIdx.x] = vsum;
    s_amin[ty][threadIdx.x] = amin; s_amax[ty][threadIdx.x] = amax;
    __syncthreads();
    if (ty != 0 || c >= c_tot) return;
    for (int t = 1; t < POOL_RY; ++t) {
        const int a1 = s_amin[t][threadIdx.x], a2 = s_amax[t][threadIdx.x];
        const float m1 = s_min[t][threadIdx.x], m2 = s_max[t][threadIdx.x];
        if (a1 >= 0 && (amin < 0 || m1 < vmin || (m1 == vmin && a1 < amin))) { vmin = m1; amin = a1; }
        if (a2 >= 0 && (amax < 0 || m2 > vmax || (m2 == vmax && a2 < amax))) { vmax = m2; amax = a2; }
        vsum += s_sum[t][threadIdx.x];
    }
    const int schemes[4] = {s0, s1, s2, s3};
    const float cnt = (float)(hi - lo);
    for (int p = 0; p < np; ++p) {
        float v; int a = -1;
        switch (schemes[p]) {
            case GNB_POOL_MIN: v = amin < 0 ? 0.f : vmin; a = amin; break;
            case GNB_POOL_MAX: v = amax < 0 ? 0.f : vmax; a = amax; break;
            case GNB_POOL_SUM: v = vsum; break;
            default: v = vsum / (cnt < 1.f ? 1.f : cnt); break;
        }
        const int64_t o = (int64_t)b * np * c_tot + (int64_t)p * c_tot + c;
        out[o] = v;
        if (arg) arg[o] = a;
    }
}

// Vectorised form for c % 4 == 0 and 16-byte aligned rows (every DynEdge configuration): thread = (4 channels, row lane),
// CTA = 16 float4 lanes x 16 row lanes per (event, 64-channel block); 4 rows per thread and step in flight, pointer bumps
// instead of 64-bit index arithmetic (the scalar kernel above spends 44 thread-instructions per element, ncu capture
// profiles/r02/z_tail_kernels_knn_pool_globalvars_dzprep.ncu-rep).
struct Pool4 { float mn[4], mx[4], sm[4]; int amn[4], amx[4]; };
__device__ __forceinline__ void pool4_take(Pool4& p, const float4& v, int idx) {
    const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        p.sm[k] += e[k];
        if (e[k] < p.mn[k]) { p.mn[k] = e[k]; p.amn[k] = idx; }       // strict: the first occurrence wins inside a thread
        if (e[k] > p.mx[k]) { p.mx[k] = e[k]; p.amx[k] = idx; }
    }
}
constexpr int POOL4_RY = 16;          // row lanes per CTA
__device__ __forceinline__ void pool4_merge(Pool4& p, const Pool4& o) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        p.sm[k] += o.sm[k];
        if (o.amn[k] >= 0 && (p.amn[k] < 0 || o.mn[k] < p.mn[k] || (o.mn[k] == p.mn[k] && o.amn[k] < p.amn[k]))) { p.mn[k] = o.mn[k]; p.amn[k] = o.amn[k]; }
        if (o.amx[k] >= 0 && (p.amx[k] < 0 || o.mx[k] > p.mx[k] || (o.mx[k] == p.mx[k] && o.amx[k] < p.amx[k]))) { p.mx[k] = o.mx[k]; p.amx[k] = o.amx[k]; }
    }
}
// The launch's time is its largest event (one CTA per (event, 64 channels)): with 16 row lanes and 4 rows in flight a
// 2 566-pulse event was 40 dependent DRAM round trips = the 62 us both the scalar and the first float4 version took, whatever
// their instruction counts. 8 rows in flight: 20 round trips (32 row lanes would halve that again but leave one CTA per SM
// for the many small events).
__global__ void __launch_bounds__(16 * POOL4_RY, 4)
segment_pool_fwd4_kernel(const float* __restrict__ x, int64_t ldx, int c_tot, const int64_t* __restrict__ ptr,
                         int np, int s0, int s1, int s2, int s3, float* __restrict__ out, int* __restrict__ arg) {
    gnb_pdl_begin();
    const int b = blockIdx.x;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;          // float4 lane, row lane
    const int c4 = blockIdx.y * 16 + tx;
    const bool col_ok = 4 * c4 < c_tot;
    const int64_t lo = ptr[b], hi = ptr[b + 1];
    Pool4 p;
#pragma unroll
    for (int k = 0; k < 4; ++k) { p.mn[k] = FLT_MAX; p.mx[k] = -FLT_MAX; p.sm[k] = 0.f; p.amn[k] = -1; p.amx[k] = -1; }
    if (col_ok) {
        const int cnt = (int)(hi - lo);
        const float4* row = reinterpret_cast<const float4*>(x + (lo + ty) * ldx) + c4;
        const int64_t step4 = POOL4_RY * (ldx >> 2);                   // POOL4_RY rows further, in float4 units
        int r = ty;
        for (; r + 7 * POOL4_RY < cnt; r += 8 * POOL4_RY) {
            float4 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = row[u * step4];
            row += 8 * step4;
#pragma unroll
            for (int u = 0; u < 8; ++u) pool4_take(p, v[u], (int)lo + r + u * POOL4_RY);
        }
        {   // tail: up to 7 rows, again all in flight together
            float4 v[7];
#pragma unroll
            for (int u = 0; u < 7; ++u) v[u] = r + u * POOL4_RY < cnt ? row[u * step4] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int u = 0; u < 7; ++u) if (r + u * POOL4_RY < cnt) pool4_take(p, v[u], (int)lo + r + u * POOL4_RY);
        }
        // (all values of the lane equal to +-FLT_MAX: the strict compares above never fired; any seen row is a valid arg)
        if (cnt > ty) {
#pragma unroll
            for (int k = 0; k < 4; ++k) { if (p.amn[k] < 0) p.amn[k] = (int)lo + ty; if (p.amx[k] < 0) p.amx[k] = (int)lo + ty; }
        }
    }
    __shared__ Pool4 s_p[POOL4_RY][16];
    s_p[ty][tx] = p;
    __syncthreads();
#pragma unroll
    for (int half = POOL4_RY / 2; half > 0; half >>= 1) {              // tree over the row lanes (lower lane = lower node indices first)
        if (ty < half) { pool4_merge(p, s_p[ty + half][tx]); s_p[ty][tx] = p; }
        __syncthreads();
    }
    if (ty != 0 || !col_ok) return;
    const int schemes[4] = {s0, s1, s2, s3};
    const float cntf = (float)(hi - lo);
    for (int q = 0; q < np; ++q) {
        float v[4]; int a[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            a[k] = -1;
            switch (schemes[q]) {
                case GNB_POOL_MIN: v[k] = p.amn[k] < 0 ? 0.f : p.mn[k]; a[k] = p.amn[k]; break;
                case GNB_POOL_MAX: v[k] = p.amx[k] < 0 ? 0.f : p.mx[k]; a[k] = p.amx[k]; break;
                case GNB_POOL_SUM: v[k] = p.sm[k]; break;
                default: v[k] = p.sm[k] / (cntf < 1.f ? 1.f : cntf); break;
            }
        }
        const int64_t o = (int64_t)b * np * c_tot + (int64_t)q * c_tot + 4 * c4;
        *reinterpret_cast<float4*>(out + o) = make_float4(v[0], v[1], v[2], v[3]);
        if (arg) *reinterpret_cast<int4*>(arg + o) = make_int4(a[0], a[1], a[2], a[3]);
    }
}

// backward: one warp per PB_NODES consecutive nodes (the event of the first node is found by a binary search over ptr, all
// lanes on the same cached addresses; the following nodes only compare against the event's end), lanes stride over the
// channels; the pooled gradients / arg tables (B x P x C) stay L2-resident. (One node per warp spent most of its time in
// the 9 dependent loads of the search: 79 us for 81 MB.)
constexpr int PB_NODES = 8;
__global__ void __launch_bounds__(256)
segment_pool_bwd_kernel(const float* __restrict__ gout, int64_t ldg, const int* __restrict__ arg, int c_tot,
                        const int64_t* __restrict__ ptr, int nseg, int64_t n, int np, int s0,
                        int s1, int s2, int s3, float* __restrict__ gx, int64_t ldx, int vec4) {
    gnb_pdl_begin();
    const int64_t i0 = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * PB_NODES;
    if (i0 >= n) return;
    const int lane = threadIdx.x & 31;
    int lo = 0, hi = nseg;
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (ptr[mid] <= i0) lo = mid; else hi = mid; }
    int b = lo;
    int64_t next = ptr[b + 1];
    for (int k = 0; k < PB_NODES && i0 + k < n; ++k) {
        const int64_t i = i0 + k;
        while (i >= next && b + 1 < nseg) { ++b; next = ptr[b + 1]; }          // also steps over empty events
        const float cnt = (float)(next - ptr[b]);
        const float* gb = gout + (int64_t)b * ldg;
        const int* ab = arg != nullptr ? arg + (int64_t)b * np * c_tot : nullptr;
        const int ii = (int)i;
        auto term = [&](int scheme, int p, int c) -> float {
            if (p >= np) return 0.f;
            const float g = gb[(int64_t)p * c_tot + c];
            if (scheme == GNB_POOL_SUM) return g;
            if (scheme == GNB_POOL_MEAN) return g / cnt;
            return ab[(int64_t)p * c_tot + c] == ii ? g : 0.f;
        };
        if (vec4) {   // c_tot % 4 == 0, 16-byte aligned rows: four channels per lane, 512-byte warp stores
            const int c4n = c_tot >> 2;
            auto term4 = [&](int scheme, int p, int c4) -> float4 {
                if (p >= np) return make_float4(0.f, 0.f, 0.f, 0.f);
                float4 g = reinterpret_cast<const float4*>(gb + (int64_t)p * c_tot)[c4];
                if (scheme == GNB_POOL_SUM) return g;
                if (scheme == GNB_POOL_MEAN) return make_float4(g.x / cnt, g.y / cnt, g.z / cnt, g.w / cnt);
                const int4 a = reinterpret_cast<const int4*>(ab + (int64_t)p * c_tot)[c4];
                return make_float4(a.x == ii ? g.x : 0.f, a.y == ii ? g.y : 0.f, a.z == ii ? g.z : 0.f, a.w == ii ? g.w : 0.f);
            };
            for (int c4 = lane; c4 < c4n; c4 += 32) {
                float4 acc = term4(s0, 0, c4);
                const float4 t1 = term4(s1, 1, c4), t2 = term4(s2, 2, c4), t3 = term4(s3, 3, c4);
                acc.x += t1.x; acc.y += t1.y; acc.z += t1.z; acc.w += t1.w;
                acc.x += t2.x; acc.y += t2.y; acc.z += t2.z; acc.w += t2.w;
                acc.x += t3.x; acc.y += t3.y; acc.z += t3.z; acc.w += t3.w;
                reinterpret_cast<float4*>(gx + i * ldx)[c4] = acc;
            }
            continue;
        }
        for (int c = lane; c < c_tot; c += 32) {
            // same summation order as the schemes are listed
            float acc = term(s0, 0, c);
            acc += term(s1, 1, c);
            acc += term(s2, 2, c);
            acc += term(s3, 3, c);
            gx[i * ldx + c] = acc;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Device-side graph definition: detector standardisation of a raw pulse batch and the
// collate arithmetic, so that raw [N, F] pulse arrays become the DynEdge input on the GPU.
// Reference: Detector._standardize (src/graphnet/models/detector/detector.py:63-77) applies one callable per named
// column; for IceCube86 (detector/icecube.py:21-48) these are x / 500, (t - 1e4) / 3e4, log10(q), (rde - 1.25) / 0.25,
// area / 0.05, identity -- i.e. every column is one of {identity, (x - a) / b, log10(x)}. Same fp32 operations in the
// same order (IEEE subtraction and division, log10f), one pass over the tensor.
struct StdTable { int kind[GNB_STD_MAX_F]; float sub[GNB_STD_MAX_F]; float div[GNB_STD_MAX_F]; };

__global__ void standardize_kernel(const float* __restrict__ x, int64_t ldx, int64_t n, int f, const StdTable t,
                                   float* __restrict__ out, int64_t ldo) {
    gnb_pdl_begin();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * f) return;
    const int64_t r = i / f;
    const int c = (int)(i - r * f);
    const float v = x[r * ldx + c];
    float o = v;
    if (t.kind[c] == 1) o = __fdiv_rn(__fsub_rn(v, t.sub[c]), t.div[c]);
    else if (t.kind[c] == 2) o = log10f(v);
    out[r * ldo + c] = o;
}

// batch[i] = b for ptr[b] <= i < ptr[b+1] (the `batch` vector Batch.from_data_list builds, dataloader.py:12-18)
__global__ void ptr_to_batch_kernel(const int64_t* __restrict__ ptr, int nseg, int64_t n, int64_t* __restrict__ batch) {
    gnb_pdl_begin();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int lo = 0, hi = nseg;
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (ptr[mid] <= i) lo = mid; else hi = mid; }
    batch[i] = lo;
}

// ---------------------------------------------------------------------------------------------
// small dense helpers
__global__ void relu_bwd_kernel(const float* __restrict__ g, int64_t ldg, const float* __restrict__ y, int64_t ldy,
                                int64_t rows, int cols, float* __restrict__ dz, int64_t ldz, int flags) {
    gnb_pdl_begin();
    const int c4 = cols >> 2;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= rows * c4) return;
    const int64_t r = t / c4;
    const int c = (int)(t - r * c4);
    float4 gv = reinterpret_cast<const float4*>(g + r * ldg)[c];
    const float4 yv = reinterpret_cast<const float4*>(y + r * ldy)[c];
    gv.x = yv.x > 0.f ? gv.x : 0.f; gv.y = yv.y > 0.f ? gv.y : 0.f;
    gv.z = yv.z > 0.f ? gv.z : 0.f; gv.w = yv.w > 0.f ? gv.w : 0.f;
    if (flags & GNB_FLAG_ROUND_TF32) {
        gv.x = gnb_round_tf32(gv.x); gv.y = gnb_round_tf32(gv.y); gv.z = gnb_round_tf32(gv.z); gv.w = gnb_round_tf32(gv.w);
    }
    reinterpret_cast<float4*>(dz + r * ldz)[c] = gv;
}


// Fused backward of "activation + bias" for a Linear output, optionally combined with the backward of the
// k-neighbour aggregation in front of it:
//   dz[r, :] = g_row(r) * act'(y[r, :])       db[:] += sum_r dz[r, :]
// with g_row(r) = g[r] (plain) or, when deg != nullptr, the broadcast of the aggregated gradient
// g[i] (add) / g[i]/deg[i] (mean) to the valid slots of node i = r / width (zero for padding slots).
// One pass over the E x C tensors instead of three (aggregate-bwd, relu-bwd, colsum).
// CTA = bx float4 columns x by row lanes (bx = the row's float4 count rounded up to a warp when that is <= 128, else 64 with
// grid.y over 256-column chunks); `rpc` rows per CTA, chosen by the launcher so that small tensors still fill the GPU (the
// read-out's 512 x 128 tensor used to run on 2 CTAs of 64 dependent steps: 33 us); 4 rows per thread and step in flight.
constexpr int ACT_ROWS = 256;

__global__ void __launch_bounds__(256)
act_bwd_colsum_kernel(const float* __restrict__ g, int64_t ldg, const float* __restrict__ y, int64_t ldy,
                      int64_t rows, int cols4, float* __restrict__ dz, int64_t ldz, float* __restrict__ db, int flags,
                      const int* __restrict__ deg, int width, int aggr, int rpc) {
    gnb_pdl_begin();
    const int bx = blockDim.x, by = blockDim.y;
    const int c4 = blockIdx.y * bx + threadIdx.x;
    const int ty = threadIdx.y;
    const bool col_ok = c4 < cols4;
    const bool relu = (flags & 0xff) == GNB_ACT_RELU, rnd = (flags & GNB_FLAG_ROUND_TF32) != 0;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const int64_t r0 = (int64_t)blockIdx.x * rpc;
    const int64_t r1 = r0 + rpc < rows ? r0 + rpc : rows;
    if (col_ok) {
        for (int64_t rb = r0 + ty; rb < r1; rb += 4 * by) {
            float4 gv[4], yv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int64_t r = rb + (int64_t)u * by;
                gv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                yv[u] = make_float4(1.f, 1.f, 1.f, 1.f);
                if (r < r1) {
                    if (deg != nullptr) {
                        const int64_t i = r / width;
                        const int s = (int)(r - i * width);
                        const int dg = deg[i];
                        if (s < dg) {
                            gv[u] = reinterpret_cast<const float4*>(g + i * ldg)[c4];
                            if (aggr == GNB_AGGR_MEAN) {
                                const float sc = 1.f / (float)dg;
                                gv[u].x *= sc; gv[u].y *= sc; gv[u].z *= sc; gv[u].w *= sc;
                            }
                        }
                    } else {
                        gv[u] = reinterpret_cast<const float4*>(g + r * ldg)[c4];
                    }
                    if (relu) yv[u] = reinterpret_cast<const float4*>(y + r * ldy)[c4];
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int64_t r = rb + (int64_t)u * by;
                if (r < r1) {
                    float4 v = gv[u];
                    if (relu) {
                        v.x = yv[u].x > 0.f ? v.x : 0.f; v.y = yv[u].y > 0.f ? v.y : 0.f;
                        v.z = yv[u].z > 0.f ? v.z : 0.f; v.w = yv[u].w > 0.f ? v.w : 0.f;
                    }
                    if (rnd) {
                        v.x = gnb_round_tf32(v.x); v.y = gnb_round_tf32(v.y);
                        v.z = gnb_round_tf32(v.z); v.w = gnb_round_tf32(v.w);
                    }
                    reinterpret_cast<float4*>(dz + r * ldz)[c4] = v;
                    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
                }
            }
        }
    }
    if (db == nullptr) return;
    __shared__ float4 s_acc[256];
    s_acc[ty * bx + threadIdx.x] = acc;
    __syncthreads();
    if (ty == 0 && col_ok) {
        for (int t = 1; t < by; ++t) {
            const float4 o = s_acc[t * bx + threadIdx.x];
            acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
        }
        atomicAdd(db + 4 * c4 + 0, acc.x); atomicAdd(db + 4 * c4 + 1, acc.y);
        atomicAdd(db + 4 * c4 + 2, acc.z); atomicAdd(db + 4 * c4 + 3, acc.w);
    }
}

// Backward of "ReLU + k-sum" from the bit mask of the aggregating GEMM epilogue (gemm_tc.cu: one uint4 = 126 bits per
// (14-node tile, channel); bit = slot valid && pre-activation > 0):  dz[(i, s), c] = bit ? g[i, c] : 0, db[c] += sum.
// One CTA pass per tile: the tile's 14 gradient rows and its mask words are staged in shared memory (coalesced),
// then 126 rows x 256 channels are written as coalesced float4 stores. CTA = 64 float4 columns x 4 row lanes.
constexpr int EM_W = 9, EM_NPT = 14, EM_ROWS = EM_W * EM_NPT;

__global__ void __launch_bounds__(256)
edge_mask_bwd_kernel(const float* __restrict__ g, int64_t ldg, const uint4* __restrict__ mask4, int64_t n, int cols,
                     float* __restrict__ dz, int64_t ldz, float* __restrict__ db, int rnd, int64_t n_tiles) {
    gnb_pdl_begin();
    __shared__ float4 s_g[EM_NPT][64];
    __shared__ __align__(16) unsigned s_m[4][256];
    __shared__ float4 s_acc[4][64];
    const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * 64 + tx;
    const int cols4 = cols >> 2;
    const int c4 = blockIdx.y * 64 + tx;
    const bool col_ok = c4 < cols4;
    const int64_t rows = n * EM_W;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        __syncthreads();
        for (int e = tid; e < EM_NPT * 64; e += 256) {
            const int f = e >> 6, cc = blockIdx.y * 64 + (e & 63);
            const int64_t node = tile * EM_NPT + f;
            s_g[f][e & 63] = (node < n && cc < cols4) ? reinterpret_cast<const float4*>(g + node * ldg)[cc]
                                                      : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        {
            const int ch = blockIdx.y * 256 + tid;
            const uint4 m = ch < cols ? mask4[tile * cols + ch] : make_uint4(0u, 0u, 0u, 0u);
            s_m[0][tid] = m.x; s_m[1][tid] = m.y; s_m[2][tid] = m.z; s_m[3][tid] = m.w;
        }
        __syncthreads();
        if (col_ok) {
            const int64_t row_base = tile * EM_ROWS;
            const int r_end = rows - row_base < EM_ROWS ? (int)(rows - row_base) : EM_ROWS;
#pragma unroll 4
            for (int r = ty; r < r_end; r += 4) {
                float4 gv = s_g[r / EM_W][tx];
                const uint4 w = reinterpret_cast<const uint4*>(s_m[r >> 5])[tx];
                const unsigned sh = r & 31;
                gv.x = ((w.x >> sh) & 1u) ? gv.x : 0.f; gv.y = ((w.y >> sh) & 1u) ? gv.y : 0.f;
                gv.z = ((w.z >> sh) & 1u) ? gv.z : 0.f; gv.w = ((w.w >> sh) & 1u) ? gv.w : 0.f;
                if (rnd) {
                    gv.x = gnb_round_tf32(gv.x); gv.y = gnb_round_tf32(gv.y);
                    gv.z = gnb_round_tf32(gv.z); gv.w = gnb_round_tf32(gv.w);
                }
                reinterpret_cast<float4*>(dz + (row_base + r) * ldz)[c4] = gv;
                acc.x += gv.x; acc.y += gv.y; acc.z += gv.z; acc.w += gv.w;
            }
        }
    }
    if (db == nullptr) return;
    s_acc[ty][tx] = acc;
    __syncthreads();
    if (ty == 0 && col_ok) {
        for (int t = 1; t < 4; ++t) {
            const float4 o = s_acc[t][tx];
            acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
        }
        atomicAdd(db + 4 * c4 + 0, acc.x); atomicAdd(db + 4 * c4 + 1, acc.y);
        atomicAdd(db + 4 * c4 + 2, acc.z); atomicAdd(db + 4 * c4 + 3, acc.w);
    }
}


// The mask backward written as bf16 planes (precision modes bf16 / bf16x3): dz0 = bf16(dz), dz1 = bf16(dz - dz0) (NP = 2);
// the bias gradient db sums the fp32 values.
// NP = 0: ONE fp16 plane of dz * scale, scale = gnb_pow2_scale(*scale_bits).x (mode mixed16).
template <int NP>
__global__ void __launch_bounds__(256)
edge_mask_bwd_bf16_kernel(const float* __restrict__ g, int64_t ldg, const uint4* __restrict__ mask4, int64_t n, int cols,
                          __nv_bfloat16* __restrict__ dz0, __nv_bfloat16* __restrict__ dz1, int64_t ldz, float* __restrict__ db,
                          int64_t n_tiles, const unsigned* __restrict__ scale_bits) {
    gnb_pdl_begin();
    const float scale = (NP == 0) ? gnb_pow2_scale(*scale_bits).x : 1.f;
    __shared__ float4 s_g[EM_NPT][64];
    __shared__ __align__(16) unsigned s_m[4][256];
    __shared__ float4 s_acc[4][64];
    const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * 64 + tx;
    const int cols4 = cols >> 2;
    const int c4 = blockIdx.y * 64 + tx;
    const bool col_ok = c4 < cols4;
    const int64_t rows = n * EM_W;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        __syncthreads();
        for (int e = tid; e < EM_NPT * 64; e += 256) {
            const int f = e >> 6, cc = blockIdx.y * 64 + (e & 63);
            const int64_t node = tile * EM_NPT + f;
            s_g[f][e & 63] = (node < n && cc < cols4) ? reinterpret_cast<const float4*>(g + node * ldg)[cc]
                                                      : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        {
            const int ch = blockIdx.y * 256 + tid;
            const uint4 m = ch < cols ? mask4[tile * cols + ch] : make_uint4(0u, 0u, 0u, 0u);
            s_m[0][tid] = m.x; s_m[1][tid] = m.y; s_m[2][tid] = m.z; s_m[3][tid] = m.w;
        }
        __syncthreads();
        if (col_ok) {
            const int64_t row_base = tile * EM_ROWS;
            const int r_end = rows - row_base < EM_ROWS ? (int)(rows - row_base) : EM_ROWS;
#pragma unroll 4
            for (int r = ty; r < r_end; r += 4) {
                float4 gv = s_g[r / EM_W][tx];
                const uint4 w = reinterpret_cast<const uint4*>(s_m[r >> 5])[tx];
                const unsigned sh = r & 31;
                gv.x = ((w.x >> sh) & 1u) ? gv.x : 0.f; gv.y = ((w.y >> sh) & 1u) ? gv.y : 0.f;
                gv.z = ((w.z >> sh) & 1u) ? gv.z : 0.f; gv.w = ((w.w >> sh) & 1u) ? gv.w : 0.f;
                uint2 b0;
                if (NP == 0) {
                    const __half2 lo = __floats2half2_rn(gv.x * scale, gv.y * scale), hi = __floats2half2_rn(gv.z * scale, gv.w * scale);
                    b0 = make_uint2(*reinterpret_cast<const unsigned*>(&lo), *reinterpret_cast<const unsigned*>(&hi));
                } else {
                    b0 = pack_bf16x4(gv.x, gv.y, gv.z, gv.w);
                }
                reinterpret_cast<uint2*>(dz0 + (row_base + r) * ldz)[c4] = b0;
                if (NP == 2)
                    reinterpret_cast<uint2*>(dz1 + (row_base + r) * ldz)[c4] =
                        pack_bf16x4(gv.x - bf16_lo_f(b0.x), gv.y - bf16_hi_f(b0.x), gv.z - bf16_lo_f(b0.y), gv.w - bf16_hi_f(b0.y));
                acc.x += gv.x; acc.y += gv.y; acc.z += gv.z; acc.w += gv.w;
            }
        }
    }
    if (db == nullptr) return;
    s_acc[ty][tx] = acc;
    __syncthreads();
    if (ty == 0 && col_ok) {
        for (int t = 1; t < 4; ++t) {
            const float4 o = s_acc[t][tx];
            acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
        }
        atomicAdd(db + 4 * c4 + 0, acc.x); atomicAdd(db + 4 * c4 + 1, acc.y);
        atomicAdd(db + 4 * c4 + 2, acc.z); atomicAdd(db + 4 * c4 + 3, acc.w);
    }
}

// out[c] += sum_r a[r, c]; out must be zero on entry. CTA = 32 x 8, 256 rows per CTA.
__global__ void colsum_kernel(const float* __restrict__ a, int64_t lda, int64_t rows, int cols,
                              float* __restrict__ out) {
    gnb_pdl_begin();
    __shared__ float s[8][33];
    const int c = blockIdx.x * 32 + threadIdx.x;
    const int64_t r0 = (int64_t)blockIdx.y * 256;
    float acc = 0.f;
    if (c < cols)
        for (int64_t r = r0 + threadIdx.y; r < r0 + 256 && r < rows; r += 8) acc += a[r * lda + c];
    s[threadIdx.y][threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.y == 0 && c < cols) {
        for (int t = 1; t < 8; ++t) acc += s[t][threadIdx.x];
        atomicAdd(out + c, acc);
    }
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace

GNB_EXPORT int gnb_global_vars(const float* x, int64_t ldx, int32_t nf, const int32_t* nbr, const int32_t* deg,
                               int32_t width, const int64_t* ptr, int64_t nseg, const float* n_pulses, float* g,
                               float* x0, int64_t ld0, void* stream) {
    if (nf < 4 || nf > GV_MAX_F || nseg < 0) return GNB_ERR_ARG;
    if (x0 != nullptr && ld0 < 2 * nf + 5) return GNB_ERR_ARG;
    if (nseg == 0) return GNB_OK;
    gnb_launch(global_vars_kernel, (unsigned)nseg, GV_THREADS, 0, (cudaStream_t)stream)(
        x, ldx, nf, nbr, deg, width, ptr, n_pulses, g, x0, ld0, 2 * nf + 5);
    GNB_RETURN_LAUNCH();
}

GNB_EXPORT int gnb_edge_hidden_fwd(const float* pq, int64_t ldpq, int32_t hdim, const int32_t* nbr,
                                   const int32_t* deg, int32_t width, int64_t n, int32_t act, float* h, int64_t ldh,
                                   void* stream) {
    if ((hdim & 3) || (ldpq & 3) || (ldh & 3) || !aligned16(pq) || !aligned16(h)) return GNB_ERR_ARG;
    if (n == 0) return GNB_OK;
    if (hdim <= 512 && width <= 32) {
        launch_hidden_node<false>((hdim + 127) / 128, dim3((unsigned)gnb_div_up(n, 8)), (cudaStream_t)stream, pq, ldpq, hdim, nbr,
                                  deg, width, n, act, h, ldh, nullptr, 0);
        GNB_RETURN_LAUNCH();
    }
    gnb_launch(edge_hidden_fwd_kernel, gnb_div_up(n * width, 8), 256, 0, (cudaStream_t)stream)(pq, ldpq, hdim, nbr, deg, width,
                                                                                       n, act, h, ldh, nullptr, 0);
    GNB_RETURN_LAUNCH();
}

// Same, additionally writing the activation bit mask hmask[r, mask_ld] (bit c of row r = h[r, c] > 0; mask_ld words per
// row, mask_ld * 32 >= hdim, bits beyond hdim zero) consumed by gnb_edge_hidden_dgrad_scatter_tf32.
GNB_EXPORT int gnb_edge_hidden_fwd_mask(const float* pq, int64_t ldpq, int32_t hdim, const int32_t* nbr,
                                        const int32_t* deg, int32_t width, int64_t n, int32_t act, float* h, int64_t ldh,
                                        uint32_t* hmask, int32_t mask_ld, void* stream) {
    if ((hdim & 3) || (ldpq & 3) || (ldh & 3) || !aligned16(pq) || !aligned16(h)) return GNB_ERR_ARG;
    if (hmask == nullptr || (int64_t)mask_ld * 32 < hdim || (mask_ld & 3) || mask_ld > 16 || !aligned16(hmask)) return GNB_ERR_ARG;
    if (n == 0) return GNB_OK;
    if (hdim <= 512 && width <= 32 && mask_ld == 4 * ((hdim + 127) / 128)) {
        launch_hidden_node<true>((hdim + 127) / 128, dim3((unsigned)gnb_div_up(n, 8)), (cudaStream_t)stream, pq, ldpq, hdim, nbr,
                                 deg, width, n, act, h, ldh, hmask, mask_ld);
        GNB_RETURN_LAUNCH();
    }
    gnb_launch(edge_hidden_fwd_kernel, gnb_div_up(n * width, 8), 256, 0, (cudaStream_t)stream)(pq, ldpq, hdim, nbr, deg, width,
                                                                                       n, act, h, ldh, hmask, mask_ld);
    GNB_RETURN_LAUNCH();
}

GNB_EXPORT int gnb_edge_hidden_bwd(const float* gh, int64_t ldg, const float* h, int64_t ldh, int32_t hdim,
                                   const int32_t* nbr, const int32_t* deg, int32_t width, int64_t n, int32_t act,
                                   float* dpq, int64_t ldpq, void* stream) {
    if ((hdim & 3) || (ldpq & 3) || (ldh & 3) || (ldg & 3) || !aligned16(gh) || !aligned16(h) || !aligned16(dpq))
        return GNB_ERR_ARG;
    if (n == 0) return GNB_OK;
    gnb_launch(edge_hidden_bwd_kernel, gnb_div_up(n, 8), 256, 0, (cudaStream_t)stream)(gh, ldg, h, ldh, hdim, nbr, deg, width,
                                                                               n, act, dpq, ldpq);
    GNB_RETURN_LAUNCH();
}

GNB_EXPORT int gnb_edge_cat_fwd(const float* x, int64_t ldx, int32_t c_in, const int32_t* nbr, const int32_t* deg,
                                int32_t width, int64_t n, float* u, int64_t ldu, void* stream) {
    if (n == 0) return GNB_OK;
    gnb_launch(edge_cat_fwd_kernel, gnb_div_up(n * width, 8), 256, 0, (cudaStream_t)stream)(x, ldx, c_in, nbr, deg, width, n, u,
                                                                                    ldu);
    GNB_RETURN_LAUNCH();
}

GNB_EXPORT int gnb_edge_cat_bwd(const float* du, int64_t ldu, int32_t c_in, const int32_t* nbr, const int32_t* deg,
                                int32_t width, int64_t n, float* dx, int64_t ldx, void* stream) {
    if (n == 0) return GNB_OK;
    gnb_launch(edge_cat_bwd_kernel, gnb_div_up(n, 8), 256, 0, (cudaStream_t)stream)(du, ldu, c_in, nbr, deg, width, n, dx, ldx);
    GNB_RETURN_LAUNCH();
}

GNB_EXPORT int gnb_edge_aggregate_fwd(const float* m, int64_t ldm, int32_t c_out, const int32_t* deg, int32_t width,
                                      int64_t n, int32_t aggr, float* y, int64_t ldy, int8_t* arg, void* stream) {
    if ((aggr & 0xff) > 2 || width > 127) return GNB_ERR_ARG;
    if (n == 0) return GNB_OK;
    gnb_launch(edge_aggregate_fwd_kernel, gnb_div_up(n, 8), 256, 0, (cudaStream_t)stream)(m, ldm, c_out, deg, width, n, aggr, y,
                                                                                  ldy, arg);
    GNB_RETURN_LAUNCH();
}

GNB_EXPORT int gnb_edge_aggregate_bwd(const float* gy, int64_t ldy, int32_t c_out, const int32_t* deg, int32_t width,
                                      int64_t n, int32_t aggr, const int8_t* arg, float* gm, int64_t ldm, void* stream) {
    if (aggr < 0 || aggr > 2 || (aggr == GNB_AGGR_MAX && arg == nullptr)) return GNB_ERR_ARG;
    if (n == 0) return GNB_OK;
    gnb_launch(edge_aggregate_bwd_kernel, gnb_div_up(n * width, 8), 256, 0, (cudaStream_t)stream)(gy, ldy, c_out, deg, width, n,
                                                                                          aggr, arg, gm, ldm);
    GNB_RETURN_LAUNCH();
}

GNB_EXPORT int gnb_edge_argmax_bwd(const float* gy, int64_t ldy, const int8_t* arg, int64_t ldarg, int32_t c_out, int32_t width,
                                   int64_t n, int32_t act, int32_t flags, float* dz, int64_t ldz, float* db, void* stream) {
    if (gy == nullptr || arg == nullptr || dz == nullptr || c_out < 1 || c_out > 512 || width < 1 || width > 64 || act < 0 || act > 2)
        return GNB_ERR_ARG;
    if (n == 0) return GNB_OK;
    const float slope = act == GNB_ACT_RELU ? 0.f : (act == GNB_ACT_LEAKY ? GNB_LEAKY_SLOPE : 1.f);
    gnb_launch(edge_argmax_bwd_kernel, gnb_div_up(n, 8 * AMB_NODES), 256, 0, (cudaStream_t)stream)(
        gy, ldy, arg, ldarg, c_out, width, n, slope, (flags & GNB_FLAG_ROUND_TF32) ? 1 : 0, dz, ldz, db);
    GNB_RETURN_LAUNCH();
}

GNB_EXPORT int gnb_segment_pool_fwd(const float* x, int64_t ldx, int32_t c, const int64_t* ptr, int64_t nseg,
                                    const int32_t* schemes, int32_t np, float* out, int32_t* arg, void* stream) {
    if (np < 1 || np > 4 || c < 1) return GNB_ERR_ARG;
    for (int p = 0; p < np; ++p) if (schemes[p] < 0 || schemes[p] > 3) return GNB_ERR_ARG;
    if (nseg == 0) return GNB_OK;
    int s[4] = {0, 0, 0, 0};
    for (int p = 0; p < np; ++p) s[p] = schemes[p];
    if (!(c & 3) && !(ldx & 3) && aligned16(x) && aligned16(out) && (arg == nullptr || aligned16(arg))) {
        gnb_launch(segment_pool_fwd4_kernel, dim3((unsigned)nseg, (unsigned)gnb_div_up(c, 64)), 16 * POOL4_RY, 0, (cudaStream_t)stream)(
            x, ldx, c, ptr, np, s[0], s[1], s[2], s[3], out, arg);
        GNB_RETURN_LAUNCH();
    }
    dim3 grid((unsigned)nseg, (unsigned)gnb_div_up(c, POOL_CX)), block(POOL_CX, POOL_RY);
    gnb_launch(segment_pool_fwd_kernel, grid, block, 0, (cudaStream_t)stream)(x, ldx, c, ptr, np, s[0], s[1], s[2], s[3], out,
                                                                      arg);
    GNB_RETURN_LAUNCH();
}

GNB_EXPORT int gnb_segment_pool_bwd(const float* gout, int64_t ldg, const int32_t* arg, int32_t c, const int64_t* ptr, int64_t nseg,
                                    int64_t n, const int32_t* schemes, int32_t np, float* gx, int64_t ldx,
                                    void* stream) {
    if (np < 1 || np > 4 || c < 1) return GNB_ERR_ARG;
    if (n == 0) return GNB_OK;
    int s[4] = {0, 0, 0, 0};
    for (int p = 0; p < np; ++p) s[p] = schemes[p];
    const int vec4 = !(c & 3) && !(ldg & 3) && !(ldx & 3) && aligned16(gout) && aligned16(gx) && (arg == nullptr || aligned16(arg));
    gnb_launch(segment_pool_bwd_kernel, gnb_div_up(n, 8 * PB_NODES), 256, 0, (cudaStream_t)stream)(gout, ldg, arg, c, ptr, (int)nseg, n, np,
                                                                                      s[0], s[1], s[2], s[3], gx, ldx, vec4);
    GNB_RETURN_LAUNCH();
}

GNB_EXPORT int gnb_relu_bwd(const float* g, int64_t ldg, const float* y, int64_t ldy, int64_t rows, int32_t cols,
                            float* dz, int64_t ldz, int32_t flags, void* stream) {
    if ((cols & 3) || (ldg & 3) || (ldy & 3) || (ldz & 3) || !aligned16(g) || !aligned16(y) || !aligned16(dz))
        return GNB_ERR_ARG;
    if (rows == 0) return GNB_OK;
    gnb_launch(relu_bwd_kernel, gnb_div_up(rows * (cols >> 2), 256), 256, 0, (cudaStream_t)stream)(g, ldg, y, ldy, rows, cols,
                                                                                           dz, ldz, flags);
    GNB_RETURN_LAUNCH();
}


// GNB_FLAG_ZERO_SRC form (no activation, no aggregation, no column sums): dz = maybe_round(g), then g = 0. `g` is an
// accumulation buffer that its next user expects empty; it is read and written through the same non-restrict pointer,
// four independent rows per thread and step so that the loads still overlap.
__global__ void __launch_bounds__(256)
round_move_zero_kernel(float* g, int64_t ldg, int64_t rows, int cols4, float* __restrict__ dz, int64_t ldz, int rnd) {
    gnb_pdl_begin();
    const int c4 = blockIdx.y * 64 + threadIdx.x;
    if (c4 >= cols4) return;
    const int64_t r0 = (int64_t)blockIdx.x * ACT_ROWS;
    const int64_t r1 = r0 + ACT_ROWS < rows ? r0 + ACT_ROWS : rows;
    for (int64_t r = r0 + threadIdx.y; r < r1; r += 16) {
        float4 gv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (r + 4 * u < r1) gv[u] = reinterpret_cast<const float4*>(g + (r + 4 * u) * ldg)[c4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (r + 4 * u < r1) {
                reinterpret_cast<float4*>(g + (r + 4 * u) * ldg)[c4] = make_float4(0.f, 0.f, 0.f, 0.f);
                float4 v = gv[u];
                if (rnd) { v.x = gnb_round_tf32(v.x); v.y = gnb_round_tf32(v.y); v.z = gnb_round_tf32(v.z); v.w = gnb_round_tf32(v.w); }
                reinterpret_cast<float4*>(dz + (r + 4 * u) * ldz)[c4] = v;
            }
    }
}

GNB_EXPORT int gnb_act_bwd_colsum(const float* g, int64_t ldg, const float* y, int64_t ldy, int64_t rows, int32_t cols,
                                  float* dz, int64_t ldz, float* db, int32_t flags, const int32_t* deg, int32_t width,
                                  int32_t aggr, void* stream) {
    if ((cols & 3) || (ldg & 3) || (ldz & 3) || !aligned16(g) || !aligned16(dz)) return GNB_ERR_ARG;
    if ((flags & 0xff) == GNB_ACT_RELU && ((ldy & 3) || !aligned16(y))) return GNB_ERR_ARG;
    if (deg != nullptr && (width < 1 || aggr < 0 || aggr > 1)) return GNB_ERR_ARG;
    if (rows == 0) return GNB_OK;
    dim3 grid((unsigned)gnb_div_up(rows, ACT_ROWS), (unsigned)gnb_div_up(cols >> 2, 64)), block(64, 4);
    if (flags & GNB_FLAG_ZERO_SRC) {
        if ((flags & 0xff) != GNB_ACT_NONE || deg != nullptr || db != nullptr) return GNB_ERR_ARG;
        gnb_launch(round_move_zero_kernel, grid, block, 0, (cudaStream_t)stream)(const_cast<float*>(g), ldg, rows, cols >> 2, dz, ldz,
                                                                         (flags & GNB_FLAG_ROUND_TF32) ? 1 : 0);
        GNB_RETURN_LAUNCH();
    }
    {
        const int cols4 = cols >> 2;
        const int bx = cols4 <= 128 ? (cols4 + 31) / 32 * 32 : 64, by = 256 / bx < 1 ? 1 : 256 / bx;
        int64_t rpc = gnb_div_up(rows, 4 * 148);                      // ~4 CTAs per SM
        rpc = rpc > ACT_ROWS ? ACT_ROWS : rpc;
        rpc = gnb_div_up(rpc, by) * by;
        dim3 grid2((unsigned)gnb_div_up(rows, rpc), (unsigned)gnb_div_up(cols4, bx)), block2(bx, by);
        gnb_launch(act_bwd_colsum_kernel, grid2, block2, 0, (cudaStream_t)stream)(g, ldg, y, ldy, rows, cols4, dz, ldz, db, flags,
                                                                          deg, width, aggr, (int)rpc);
    }
    GNB_RETURN_LAUNCH();
}

// Same, with the ReLU mask taken from the bit mask written by gnb_edge_linear_agg_fwd_tf32 (k = 8 tables: width 9,
// 14 nodes per tile) instead of the message tensor. g = gradient of the aggregated node tensor [n, cols].
GNB_EXPORT int gnb_edge_mask_bwd_colsum(const float* g, int64_t ldg, const uint32_t* maskbits, int64_t n, int32_t cols,
                                        const int32_t* deg, float* dz, int64_t ldz, float* db, int32_t flags, void* stream) {
    if (maskbits == nullptr || deg == nullptr) return GNB_ERR_ARG;
    if ((cols & 3) || (ldg & 3) || (ldz & 3) || !aligned16(g) || !aligned16(dz) || !aligned16(maskbits)) return GNB_ERR_ARG;
    if (n == 0) return GNB_OK;
    const int64_t n_tiles = (n + EM_NPT - 1) / EM_NPT;
    const int64_t max_ctas = 148 * 8;
    dim3 grid((unsigned)(n_tiles < max_ctas ? n_tiles : max_ctas), (unsigned)gnb_div_up(cols >> 2, 64)), block(64, 4);
    gnb_launch(edge_mask_bwd_kernel, grid, block, 0, (cudaStream_t)stream)(g, ldg, reinterpret_cast<const uint4*>(maskbits), n, cols,
                                                                    dz, ldz, db, (flags & GNB_FLAG_ROUND_TF32) ? 1 : 0,
                                                                    n_tiles);
    GNB_RETURN_LAUNCH();
}

GNB_EXPORT int gnb_colsum(const float* a, int64_t lda, int64_t rows, int32_t cols, float* out, void* stream) {
    if (rows == 0) return GNB_OK;
    dim3 grid((unsigned)gnb_div_up(cols, 32), (unsigned)gnb_div_up(rows, 256)), block(32, 8);
    gnb_launch(colsum_kernel, grid, block, 0, (cudaStream_t)stream)(a, lda, rows, cols, out);
    GNB_RETURN_LAUNCH();
}

// out[r, c] = kind[c] == 0 ? x : kind[c] == 1 ? (x - sub[c]) / div[c] : log10(x); kind / sub / div are HOST arrays (f <= 32).
GNB_EXPORT int gnb_standardize(const float* x, int64_t ldx, int64_t n, int32_t f, const int32_t* kind, const float* sub,
                               const float* div, float* out, int64_t ldo, void* stream) {
    if (n < 0 || f < 1 || f > GNB_STD_MAX_F || ldx < f || ldo < f || kind == nullptr || sub == nullptr || div == nullptr)
        return GNB_ERR_ARG;
    StdTable t;
    for (int c = 0; c < GNB_STD_MAX_F; ++c) { t.kind[c] = 0; t.sub[c] = 0.f; t.div[c] = 1.f; }
    for (int c = 0; c < f; ++c) {
        if (kind[c] < 0 || kind[c] > 2) return GNB_ERR_ARG;
        t.kind[c] = kind[c]; t.sub[c] = sub[c]; t.div[c] = div[c];
    }
    if (n == 0) return GNB_OK;
    gnb_launch(standardize_kernel, gnb_div_up(n * f, 256), 256, 0, (cudaStream_t)stream)(x, ldx, n, f, t, out, ldo);
    GNB_RETURN_LAUNCH();
}

GNB_EXPORT int gnb_ptr_to_batch(const int64_t* ptr, int64_t nseg, int64_t n, int64_t* batch, void* stream) {
    if (n < 0 || nseg < 1 || nseg >= ((int64_t)1 << 31)) return GNB_ERR_ARG;
    if (n == 0) return GNB_OK;
    gnb_launch(ptr_to_batch_kernel, gnb_div_up(n, 256), 256, 0, (cudaStream_t)stream)(ptr, (int)nseg, n, batch);
    GNB_RETURN_LAUNCH();
}

// ---- bf16-plane producers (precision modes bf16 / bf16x3; pitches in bf16 ELEMENTS, multiples of 8) -----------------------------
// h planes of the hoisted EdgeConv hidden layer: h0 = bf16(relu(P_i + Q_j)), h1 = bf16(h - h0) (h1 may be NULL), plus the
// activation bits of gnb_edge_hidden_fwd_mask (hmask may be NULL: inference). hdim <= 512, width <= 32.
static int hidden_fwd16_impl(const float* pq, int64_t ldpq, int32_t hdim, const int32_t* nbr, const int32_t* deg,
                             int32_t width, int64_t n, void* h0, void* h1, int64_t ldh, uint32_t* hmask,
                             int32_t mask_ld, const uint32_t* scale_bits, void* stream) {
    if ((hdim & 7) || (ldpq & 3) || (ldh & 7) || ldh < hdim || !aligned16(pq) || !aligned16(h0) || !aligned16(h1) || h0 == nullptr)
        return GNB_ERR_ARG;
    if (hdim > 512 || width > 32 || width < 1) return GNB_ERR_UNSUPPORTED;
    if (hmask != nullptr && (mask_ld != 4 * ((hdim + 127) / 128) || !aligned16(hmask))) return GNB_ERR_ARG;
    if (n == 0) return GNB_OK;
    const dim3 grid((unsigned)gnb_div_up(n, 8));
    const int nit = (hdim + 127) / 128;
    cudaStream_t st = (cudaStream_t)stream;
    __nv_bfloat16 *a = (__nv_bfloat16*)h0, *b = (__nv_bfloat16*)h1;
    if (scale_bits != nullptr) {
        if (h1 != nullptr) launch_hidden_node_bf16<2, true>(nit, grid, st, pq, ldpq, hdim, nbr, deg, width, n, a, b, ldh, hmask, mask_ld, scale_bits);
        else launch_hidden_node_bf16<1, true>(nit, grid, st, pq, ldpq, hdim, nbr, deg, width, n, a, nullptr, ldh, hmask, mask_ld, scale_bits);
    } else {
        if (h1 != nullptr) launch_hidden_node_bf16<2, false>(nit, grid, st, pq, ldpq, hdim, nbr, deg, width, n, a, b, ldh, hmask, mask_ld, nullptr);
        else launch_hidden_node_bf16<1, false>(nit, grid, st, pq, ldpq, hdim, nbr, deg, width, n, a, nullptr, ldh, hmask, mask_ld, nullptr);
    }
    GNB_RETURN_LAUNCH();
}
GNB_EXPORT int gnb_edge_hidden_fwd_bf16(const float* pq, int64_t ldpq, int32_t hdim, const int32_t* nbr, const int32_t* deg,
                                        int32_t width, int64_t n, void* h0, void* h1, int64_t ldh, uint32_t* hmask,
                                        int32_t mask_ld, void* stream) {
    return hidden_fwd16_impl(pq, ldpq, hdim, nbr, deg, width, n, h0, h1, ldh, hmask, mask_ld, nullptr, stream);
}
// The same as fp16 planes of h * 2^s (mode mixed16): 2^s = gnb_pow2_scale(*scale_bits).x, *scale_bits >= fp32 bits of max h
// (gnb_absmax_bits over PQ with shift = 1: h = relu(P_i + Q_j) <= 2 max|PQ|).
GNB_EXPORT int gnb_edge_hidden_fwd_f16(const float* pq, int64_t ldpq, int32_t hdim, const int32_t* nbr, const int32_t* deg,
                                       int32_t width, int64_t n, void* h0, void* h1, int64_t ldh, uint32_t* hmask,
                                       int32_t mask_ld, const uint32_t* scale_bits, void* stream) {
    if (scale_bits == nullptr) return GNB_ERR_ARG;
    return hidden_fwd16_impl(pq, ldpq, hdim, nbr, deg, width, n, h0, h1, ldh, hmask, mask_ld, scale_bits, stream);
}
// dz planes of the "ReLU + k-sum" backward from the aggregating epilogue's bit mask (see gnb_edge_mask_bwd_colsum):
// dz0 = bf16(dz), dz1 = bf16(dz - dz0) (may be NULL); db[c] += column sums of the fp32 values.
GNB_EXPORT int gnb_edge_mask_bwd_colsum_bf16(const float* g, int64_t ldg, const uint32_t* maskbits, int64_t n, int32_t cols,
                                             void* dz0, void* dz1, int64_t ldz, float* db, void* stream) {
    if (maskbits == nullptr || dz0 == nullptr) return GNB_ERR_ARG;
    if ((cols & 7) || (ldg & 3) || (ldz & 7) || ldz < cols || !aligned16(g) || !aligned16(dz0) || !aligned16(dz1) || !aligned16(maskbits))
        return GNB_ERR_ARG;
    if (n == 0) return GNB_OK;
    const int64_t n_tiles = (n + EM_NPT - 1) / EM_NPT;
    const int64_t max_ctas = 148 * 8;
    dim3 grid((unsigned)(n_tiles < max_ctas ? n_tiles : max_ctas), (unsigned)gnb_div_up(cols >> 2, 64)), block(64, 4);
    if (dz1 != nullptr)
        gnb_launch(edge_mask_bwd_bf16_kernel<2>, grid, block, 0, (cudaStream_t)stream)(g, ldg, reinterpret_cast<const uint4*>(maskbits), n, cols,
                                                                             (__nv_bfloat16*)dz0, (__nv_bfloat16*)dz1, ldz, db, n_tiles, nullptr);
    else
        gnb_launch(edge_mask_bwd_bf16_kernel<1>, grid, block, 0, (cudaStream_t)stream)(g, ldg, reinterpret_cast<const uint4*>(maskbits), n, cols,
                                                                             (__nv_bfloat16*)dz0, nullptr, ldz, db, n_tiles, nullptr);
    GNB_RETURN_LAUNCH();
}
// The same as ONE fp16 plane of dz * 2^s, 2^s = gnb_pow2_scale(*scale_bits).x with *scale_bits = fp32 bits of max|g|
// (gnb_absmax_bits on g): fp16 carries tf32's significand in half the bytes once the tensor's range is centred.
GNB_EXPORT int gnb_edge_mask_bwd_colsum_f16(const float* g, int64_t ldg, const uint32_t* maskbits, int64_t n, int32_t cols,
                                            void* dz, int64_t ldz, float* db, const uint32_t* scale_bits, void* stream) {
    if (maskbits == nullptr || dz == nullptr || scale_bits == nullptr) return GNB_ERR_ARG;
    if ((cols & 7) || (ldg & 3) || (ldz & 7) || ldz < cols || !aligned16(g) || !aligned16(dz) || !aligned16(maskbits)) return GNB_ERR_ARG;
    if (n == 0) return GNB_OK;
    const int64_t n_tiles = (n + EM_NPT - 1) / EM_NPT;
    const int64_t max_ctas = 148 * 8;
    dim3 grid((unsigned)(n_tiles < max_ctas ? n_tiles : max_ctas), (unsigned)gnb_div_up(cols >> 2, 64)), block(64, 4);
    gnb_launch(edge_mask_bwd_bf16_kernel<0>, grid, block, 0, (cudaStream_t)stream)(g, ldg, reinterpret_cast<const uint4*>(maskbits), n, cols,
                                                                         (__nv_bfloat16*)dz, nullptr, ldz, db, n_tiles, scale_bits);
    GNB_RETURN_LAUNCH();
}

// *out_bits = max(*out_bits, fp32 bits of 2^shift max |a[r, c]|) over a [rows, cols] matrix (cols % 4 == 0); *out_bits must be
// initialised (0) by the caller. Non-negative floats order like their bit patterns, so one atomicMax per CTA suffices.
__global__ void __launch_bounds__(256) absmax_bits_kernel(const float* __restrict__ a, int64_t lda, int64_t rows, int cols4,
                                                           unsigned* __restrict__ out_bits, unsigned shift) {
    gnb_pdl_begin();
    unsigned m = 0u;
    const int64_t total = rows * cols4;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    auto upd = [&](const float4 v) {
        m = max(max(m, __float_as_uint(fabsf(v.x))), max(__float_as_uint(fabsf(v.y)), max(__float_as_uint(fabsf(v.z)), __float_as_uint(fabsf(v.w)))));
    };
    if (lda == 4 * (int64_t)cols4) {              // contiguous: flat float4 stream, four independent loads in flight per thread
        const float4* p = reinterpret_cast<const float4*>(a);
        int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        for (; t + 3 * stride < total; t += 4 * stride) {
            const float4 v0 = p[t], v1 = p[t + stride], v2 = p[t + 2 * stride], v3 = p[t + 3 * stride];
            upd(v0); upd(v1); upd(v2); upd(v3);
        }
        for (; t < total; t += stride) upd(p[t]);
    } else {
        for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
            const int64_t r = t / cols4;
            const int c = (int)(t - r * cols4);
            upd(reinterpret_cast<const float4*>(a + r * lda)[c]);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    __shared__ unsigned s_m[8];
    if ((threadIdx.x & 31) == 0) s_m[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) m = max(m, s_m[w]);
        if (m < 0x7f800000u && m != 0u) atomicMax(out_bits, min(m + (shift << 23), 0x7f000000u));   // NaN / Inf do not set the scale
    }
}
GNB_EXPORT int gnb_absmax_bits(const float* a, int64_t lda, int64_t rows, int32_t cols, int32_t shift, uint32_t* out_bits,
                               void* stream) {
    if ((cols & 3) || (lda & 3) || !aligned16(a) || out_bits == nullptr || rows < 0 || shift < 0 || shift > 8) return GNB_ERR_ARG;
    if (rows == 0 || cols == 0) return GNB_OK;
    const int64_t total = rows * (cols >> 2);
    int64_t ctas = (total + 256 * 8 - 1) / (256 * 8);
    if (ctas > 148 * 8) ctas = 148 * 8;
    gnb_launch(absmax_bits_kernel, (unsigned)ctas, 256, 0, (cudaStream_t)stream)(a, lda, rows, cols >> 2, out_bits, (unsigned)shift);
    GNB_RETURN_LAUNCH();
}


// a[r, 0:cols] = 0 for a [rows, cols] block of pitch lda (cols % 4 == 0, 16-byte aligned): the scatter target of the fused
// data-gradient kernels (the Q half of dPQ). cudaMemset2DAsync moves this block at ~2 TB/s; plain float4 stores at HBM speed.
__global__ void __launch_bounds__(256) zero_block_kernel(float* __restrict__ a, int64_t lda, int64_t rows, int cols4) {
    gnb_pdl_begin();
    const int64_t total = rows * cols4;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = t / cols4;
        const int c = (int)(t - r * cols4);
        reinterpret_cast<float4*>(a + r * lda)[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}
GNB_EXPORT int gnb_zero_block(float* a, int64_t lda, int64_t rows, int32_t cols, void* stream) {
    if ((cols & 3) || (lda & 3) || !aligned16(a) || rows < 0 || cols < 0) return GNB_ERR_ARG;
    if (rows == 0 || cols == 0) return GNB_OK;
    const int64_t total = rows * (cols >> 2);
    int64_t ctas = (total + 255) / 256;
    if (ctas > 148 * 16) ctas = 148 * 16;
    gnb_launch(zero_block_kernel, (unsigned)ctas, 256, 0, (cudaStream_t)stream)(a, lda, rows, cols >> 2);
    GNB_RETURN_LAUNCH();
}


// Preparation of the dz-free backward of an aggregating Linear (fp16-plane modes): from the node-level gradient g [n, cols] and
// the tile-major ReLU bits of the aggregating epilogue (maskbits[(i / 14) * cols + c] = uint4, bit 9 (i % 14) + s) it writes
//   g16[i, c]   = fp16(g[i, c] * 2^s),  2^s = gnb_pow2_scale(*scale_bits).x                       (operand values of dz)
//   rowmask[(i * 9 + s) * (cols / 32) + c / 32] bit c % 32 = the same bits, ROW-major         (what the GEMM builders read)
//   db[c]      += sum_i g[i, c] * popcount(bits of node i, channel c)                           (bias gradient = column sums of dz)
// in one pass over 0.1 GB instead of the 0.37 GB write of a stored dz. CTA = 256 channels of one 14-node tile per iteration;
// the 32 x 32 bit transposes are warp ballots. cols % 32 == 0.
// W = 9: the layout above. W = 8 (graphs in which no node keeps k + 1 = 9 neighbours; the caller's device flag says so): 16 nodes per
// tile, bit 8 (i % 16) + s of maskbits[(i / 16) * cols + c], rows (i * 8 + s) of rowmask -- 128 rows per tile, no padding rows.
template <int W>
__device__ __forceinline__ void edge_dz_prep_body(const float* __restrict__ g, int64_t ldg, const uint4* __restrict__ mask4,
                                                  int64_t n, int cols, const unsigned* __restrict__ scale_bits,
                                                  __half* __restrict__ g16, unsigned* __restrict__ rowmask, float* __restrict__ db) {
    constexpr int NPT = W == 8 ? 16 : EM_NPT, ROWS = W * NPT;
    const int64_t n_tiles = (n + NPT - 1) / NPT;
    const int c = blockIdx.y * 256 + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const bool c_on = c < cols;                               // warp-uniform (cols % 32 == 0)
    if (!c_on) return;
    const float scale = gnb_pow2_scale(*scale_bits).x;
    const int cw = cols >> 5;
    float acc = 0.f;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t node0 = tile * NPT;
        float gv[NPT];
#pragma unroll
        for (int f = 0; f < NPT; ++f) {
            const int64_t nd = node0 + f < n ? node0 + f : n - 1;
            gv[f] = g[nd * ldg + c];
        }
        const uint4 m = mask4[tile * cols + c];
        const unsigned w[5] = {m.x, m.y, m.z, m.w, 0u};
#pragma unroll
        for (int f = 0; f < NPT; ++f) {
            const int bp = W * f;                                                   // compile-time after unrolling
            const unsigned bw = __funnelshift_r(w[bp >> 5], w[(bp >> 5) + 1], bp & 31) & ((1u << W) - 1u);
            if (node0 + f < n) {
                acc += gv[f] * (float)__popc(bw);
                g16[(node0 + f) * cols + c] = __float2half_rn(gv[f] * scale);
            }
        }
        // row-major words: 32 x 32 bit transposes across the warp (lane = channel, bit = row  ->  lane = row, bit = channel) by
        // five rounds of block swaps with the lane's partner (recursive transpose: 5 shuffles instead of 32 ballots per word)
        unsigned* rm = rowmask + tile * ROWS * cw + (c >> 5);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            unsigned a = w[k];
#pragma unroll
            for (int jb = 16; jb > 0; jb >>= 1) {
                const unsigned msk = jb == 16 ? 0x0000FFFFu : jb == 8 ? 0x00FF00FFu : jb == 4 ? 0x0F0F0F0Fu : jb == 2 ? 0x33333333u : 0x55555555u;
                const unsigned p = __shfl_xor_sync(0xffffffffu, a, jb);
                // lanes with bit jb clear keep their low blocks and take the partner's low blocks into their high blocks
                a = (lane & jb) ? ((a & ~msk) | ((p >> jb) & msk)) : ((a & msk) | ((p << jb) & ~msk));
            }
            if (32 * k + lane < ROWS) rm[(int64_t)(32 * k + lane) * cw] = a;
        }
    }
    atomicAdd(db + c, acc);
}
// Zero job riding on the launch (the executor's backward pass): a [rows, 4 cols4] block of pitch lda (the Q half of dPQ, the scatter
// target of the data-gradient kernel that follows) and nb floats at b (the bias-gradient scratch) -- a zero_block launch, a memset
// and their two launch gaps (~24 us per layer) folded into a kernel that has store bandwidth to spare. a / b may be NULL.
struct DzPrepZero { float* a; int64_t lda, rows; int cols4; float* b; int nb; };
__global__ void __launch_bounds__(256) edge_dz_prep_kernel(const float* __restrict__ g, int64_t ldg, const uint4* __restrict__ mask4,
                                                            int64_t n, int cols, const unsigned* __restrict__ scale_bits,
                                                            __half* __restrict__ g16, unsigned* __restrict__ rowmask,
                                                            float* __restrict__ db, const int* __restrict__ full9, const DzPrepZero zj) {
    gnb_pdl_begin();
    {
        const int64_t nthreads = (int64_t)gridDim.x * gridDim.y * 256;
        const int64_t t0 = ((int64_t)blockIdx.y * gridDim.x + blockIdx.x) * 256 + threadIdx.x;
        if (zj.a != nullptr) {
            const int64_t total = zj.rows * zj.cols4;
            for (int64_t t = t0; t < total; t += nthreads) {
                const int64_t r = t / zj.cols4;
                reinterpret_cast<float4*>(zj.a + r * zj.lda)[(int)(t - r * zj.cols4)] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        if (zj.b != nullptr)
            for (int64_t t = t0; t < zj.nb; t += nthreads) zj.b[t] = 0.f;
    }
    if (full9 == nullptr || *full9 != 0) edge_dz_prep_body<9>(g, ldg, mask4, n, cols, scale_bits, g16, rowmask, db);
    else edge_dz_prep_body<8>(g, ldg, mask4, n, cols, scale_bits, g16, rowmask, db);
}
// *flag = 1 when some node keeps more than k neighbours (the k + 1 duplicate quirk of the table, edges.py:72-80 -> torch_cluster), else
// 0: the device-side switch between the 9-slot and the 8-slot layout of the per-edge kernels (no host synchronisation). One CTA.
__global__ void __launch_bounds__(1024) edge_slot_flag_kernel(const int* __restrict__ deg, int64_t n, int k, int* __restrict__ flag) {
    gnb_pdl_begin();
    int any = 0;
    const int64_t n4 = ((reinterpret_cast<uintptr_t>(deg) & 15u) == 0) ? n >> 2 : 0;       // 16-byte loads, four in flight per thread
    const int4* d4 = reinterpret_cast<const int4*>(deg);
    for (int64_t i = threadIdx.x; i < n4; i += 4096) {
        int4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = i + 1024 * u < n4 ? __ldg(d4 + i + 1024 * u) : make_int4(0, 0, 0, 0);
#pragma unroll
        for (int u = 0; u < 4; ++u) any |= (v[u].x > k) | (v[u].y > k) | (v[u].z > k) | (v[u].w > k);
    }
    for (int64_t i = 4 * n4 + threadIdx.x; i < n; i += 1024) any |= deg[i] > k;
    any = __syncthreads_or(any);
    if (threadIdx.x == 0) *flag = any ? 1 : 0;
}
// The same test spread over the GPU for a flag word the caller has zeroed: CTAs that see such a node store 1 (a benign race), the
// others store nothing. (One CTA over 158 k nodes took 10 us per launch in the 1024-event inference step.)
__global__ void __launch_bounds__(256) edge_slot_flag_or_kernel(const int* __restrict__ deg, int64_t n, int k, int* __restrict__ flag) {
    gnb_pdl_begin();
    int any = 0;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) any |= deg[i] > k;
    any = __syncthreads_or(any);
    if (any && threadIdx.x == 0) *flag = 1;
}
GNB_EXPORT int gnb_edge_slot_flag_or(const int32_t* deg, int64_t n, int32_t k, int32_t* flag, void* stream) {
    if (deg == nullptr || flag == nullptr || n < 0) return GNB_ERR_ARG;
    if (n == 0) return GNB_OK;
    int64_t ctas = (n + 1023) / 1024;                 // four nodes per thread
    if (ctas > 148 * 4) ctas = 148 * 4;
    gnb_launch(edge_slot_flag_or_kernel, (unsigned)ctas, 256, 0, (cudaStream_t)stream)(deg, n, k, flag);
    GNB_RETURN_LAUNCH();
}
GNB_EXPORT int gnb_edge_slot_flag(const int32_t* deg, int64_t n, int32_t k, int32_t* flag, void* stream) {
    if (deg == nullptr || flag == nullptr || n < 0) return GNB_ERR_ARG;
    gnb_launch(edge_slot_flag_kernel, 1, 1024, 0, (cudaStream_t)stream)(deg, n, k, flag);
    GNB_RETURN_LAUNCH();
}
GNB_EXPORT int gnb_edge_dz_prep_wz(const float* g, int64_t ldg, const uint32_t* maskbits, int64_t n, int32_t cols,
                                   const uint32_t* scale_bits, void* g16, uint32_t* rowmask, float* db, const int32_t* full9,
                                   float* zero_a, int64_t lda, int64_t zrows, int32_t zcols, float* zero_b, int32_t zb_count,
                                   void* stream) {
    if (zero_a != nullptr && ((zcols & 3) || (lda & 3) || !aligned16(zero_a) || zrows < 0 || zcols < 0)) return GNB_ERR_ARG;
    if (zero_b != nullptr && zb_count < 0) return GNB_ERR_ARG;
    if (maskbits == nullptr || g == nullptr || db == nullptr || g16 == nullptr || rowmask == nullptr || scale_bits == nullptr ||
        cols < 32 || (cols & 31) || !aligned16(maskbits))
        return GNB_ERR_ARG;
    if (n == 0) return GNB_OK;
    const int64_t n_tiles = (n + EM_NPT - 1) / EM_NPT;
    const int64_t max_ctas = 148 * 8;
    dim3 grid((unsigned)(n_tiles < max_ctas ? n_tiles : max_ctas), (unsigned)gnb_div_up(cols, 256));
    gnb_launch(edge_dz_prep_kernel, grid, 256, 0, (cudaStream_t)stream)(g, ldg, reinterpret_cast<const uint4*>(maskbits), n, cols, scale_bits,
                                                                (__half*)g16, rowmask, db, full9,
                                                                DzPrepZero{zero_a, lda, zrows, zcols >> 2, zero_b, zb_count});
    GNB_RETURN_LAUNCH();
}
GNB_EXPORT int gnb_edge_dz_prep_w(const float* g, int64_t ldg, const uint32_t* maskbits, int64_t n, int32_t cols,
                                  const uint32_t* scale_bits, void* g16, uint32_t* rowmask, float* db, const int32_t* full9,
                                  void* stream) {
    return gnb_edge_dz_prep_wz(g, ldg, maskbits, n, cols, scale_bits, g16, rowmask, db, full9, nullptr, 0, 0, 0, nullptr, 0, stream);
}
GNB_EXPORT int gnb_edge_dz_prep(const float* g, int64_t ldg, const uint32_t* maskbits, int64_t n, int32_t cols,
                                const uint32_t* scale_bits, void* g16, uint32_t* rowmask, float* db, void* stream) {
    return gnb_edge_dz_prep_w(g, ldg, maskbits, n, cols, scale_bits, g16, rowmask, db, nullptr, stream);
}
