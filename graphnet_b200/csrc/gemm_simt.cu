// fp32 SIMT GEMM family for the "fp32" precision mode of the DynEdge path (exact fp32 products,
// fp32 accumulation) and for shapes the tcgen05 kernels do not cover.
//
//   C[m,n] = epilogue( sum_k Aop(m,k) * Bop(n,k) )
//     TA = 0: A is [M,K] row-major (K contiguous)     TA = 1: A is [K,M] row-major (M contiguous)
//     TB = 0: B is [N,K] row-major (K contiguous)     TB = 1: B is [K,N] row-major (N contiguous)
//   forward Linear      y  = act(x W^T + b)        : TA=0, TB=0   (torch.nn.Linear at dynedge.py:200-247)
//   backward data       dx = dz W                  : TA=0, TB=1
//   backward weights    dW = dz^T x  (split-K)     : TA=1, TB=1   (reduction over rows, fp32 atomics)
//
// 128x128x16 CTA tile, 256 threads, 8x8 register tile per thread (split 4+4 so every shared-memory
// read is a conflict-free float4), register-staged double buffering of the global loads.
#include "common.cuh"

namespace {

constexpr int BM = 128, BN = 128, BK = 16, PADS = 4, NT = 256;

template <bool T, bool VEC>
__device__ __forceinline__ void load_tile(const float* __restrict__ p, int64_t ld, int64_t r0, int64_t rmax, int64_t k0,
                                          int64_t kmax, int tid, float4 (&reg)[2]) {
    // T = 0: operand is [rows, K] (K contiguous). thread -> row = tid/4 + {0,64}, k quad = tid%4
    // T = 1: operand is [K, rows] (rows contiguous). thread -> k = tid/32 + {0,8}, row quad = tid%32
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (!T) {
            const int64_t r = r0 + (tid >> 2) + h * 64;
            const int64_t k = k0 + (tid & 3) * 4;
            if (r < rmax) {
                const float* src = p + r * ld + k;
                if (VEC && k + 3 < kmax) {
                    v = *reinterpret_cast<const float4*>(src);
                } else {
                    if (k + 0 < kmax) v.x = src[0];
                    if (k + 1 < kmax) v.y = src[1];
                    if (k + 2 < kmax) v.z = src[2];
                    if (k + 3 < kmax) v.w = src[3];
                }
            }
        } else {
            const int64_t k = k0 + (tid >> 5) + h * 8;
            const int64_t r = r0 + (tid & 31) * 4;
            if (k < kmax) {
                const float* src = p + k * ld + r;
                if (VEC && r + 3 < rmax) {
                    v = *reinterpret_cast<const float4*>(src);
                } else {
                    if (r + 0 < rmax) v.x = src[0];
                    if (r + 1 < rmax) v.y = src[1];
                    if (r + 2 < rmax) v.z = src[2];
                    if (r + 3 < rmax) v.w = src[3];
                }
            }
        }
        reg[h] = v;
    }
}

template <bool T>
__device__ __forceinline__ void store_tile(float (*s)[BM + PADS], int tid, const float4 (&reg)[2]) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        if (!T) {
            const int r = (tid >> 2) + h * 64;
            const int k = (tid & 3) * 4;
            s[k + 0][r] = reg[h].x; s[k + 1][r] = reg[h].y; s[k + 2][r] = reg[h].z; s[k + 3][r] = reg[h].w;
        } else {
            const int k = (tid >> 5) + h * 8;
            const int r = (tid & 31) * 4;
            *reinterpret_cast<float4*>(&s[k][r]) = reg[h];
        }
    }
}

template <bool TA, bool TB, bool VEC>
__global__ void __launch_bounds__(NT)
gemm_f32_kernel(const float* __restrict__ A, int64_t lda, const float* __restrict__ B, int64_t ldb,
                float* __restrict__ C, int64_t ldc, int64_t M, int64_t N, int64_t K,
                const float* __restrict__ bias, int act, int accumulate, int64_t k_per_split, int atomic_flag) {
    gnb_pdl_begin();
    __shared__ __align__(16) float As[2][BK][BM + PADS];
    __shared__ __align__(16) float Bs[2][BK][BN + PADS];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int64_t m0 = (int64_t)blockIdx.y * BM, n0 = (int64_t)blockIdx.x * BN;
    const int64_t kbeg = (int64_t)blockIdx.z * k_per_split;
    const int64_t kend = (kbeg + k_per_split < K) ? kbeg + k_per_split : K;
    const bool atomic_out = atomic_flag != 0;

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    float4 ra[2], rb[2];
    load_tile<TA, VEC>(A, lda, m0, M, kbeg, kend, tid, ra);
    load_tile<TB, VEC>(B, ldb, n0, N, kbeg, kend, tid, rb);
    store_tile<TA>(As[0], tid, ra);
    store_tile<TB>(Bs[0], tid, rb);
    __syncthreads();

    int buf = 0;
    for (int64_t k0 = kbeg; k0 < kend; k0 += BK) {
        const bool has_next = k0 + BK < kend;
        if (has_next) {
            load_tile<TA, VEC>(A, lda, m0, M, k0 + BK, kend, tid, ra);
            load_tile<TB, VEC>(B, ldb, n0, N, k0 + BK, kend, tid, rb);
        }
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][64 + tx * 4]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (has_next) {
            store_tile<TA>(As[buf ^ 1], tid, ra);
            store_tile<TB>(Bs[buf ^ 1], tid, rb);
            __syncthreads();
            buf ^= 1;
        }
    }

    // epilogue
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (m >= M) continue;
#pragma unroll
        for (int jh = 0; jh < 2; ++jh) {
            const int64_t n = n0 + jh * 64 + tx * 4;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (n + j >= N) continue;
                float v = acc[i][jh * 4 + j];
                float* dst = C + m * ldc + n + j;
                if (atomic_out) {
                    if (bias != nullptr && blockIdx.z == 0) v += bias[n + j];
                    atomicAdd(dst, v);
                } else {
                    if (bias != nullptr) v += bias[n + j];
                    if (accumulate) v += *dst;
                    if (act == GNB_ACT_RELU) v = fmaxf(v, 0.f);
                    *dst = v;
                }
            }
        }
    }
}

inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

template <bool TA, bool TB>
int launch(const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc, int64_t M, int64_t N,
           int64_t K, const float* bias, int act, int accumulate, int splits, bool atomic, cudaStream_t st) {
    if (M <= 0 || N <= 0) return GNB_OK;
    if (splits < 1) splits = 1;
    int64_t kps = (K + splits - 1) / splits;
    kps = ((kps + BK - 1) / BK) * BK;
    if (kps < BK) kps = BK;
    splits = (int)((K + kps - 1) / kps);
    if (splits < 1) splits = 1;
    if (splits > 1) atomic = true;
    if (atomic && (act != GNB_ACT_NONE)) return GNB_ERR_ARG;
    dim3 grid((unsigned)gnb_div_up(N, BN), (unsigned)gnb_div_up(M, BM), (unsigned)splits);
    if (grid.y > 65535u || grid.z > 65535u) return GNB_ERR_ARG;
    const bool vec = al16(A) && al16(B) && (lda % 4 == 0) && (ldb % 4 == 0);
    if (vec)
        gnb_launch(gemm_f32_kernel<TA, TB, true>, grid, NT, 0, st)(A, lda, B, ldb, C, ldc, M, N, K, bias, act, accumulate, kps, atomic ? 1 : 0);
    else
        gnb_launch(gemm_f32_kernel<TA, TB, false>, grid, NT, 0, st)(A, lda, B, ldb, C, ldc, M, N, K, bias, act, accumulate, kps, atomic ? 1 : 0);
    GNB_RETURN_LAUNCH();
}

}  // namespace

// y[M,N] = act(x[M,K] W[N,K]^T + bias (+ y if accumulate))
GNB_EXPORT int gnb_linear_fwd_f32(const float* x, int64_t ldx, const float* w, int64_t ldw, const float* bias, float* y,
                                  int64_t ldy, int64_t m, int64_t n, int64_t k, int32_t act, int32_t accumulate,
                                  void* stream) {
    return launch<false, false>(x, ldx, w, ldw, y, ldy, m, n, k, bias, act, accumulate, 1, false, (cudaStream_t)stream);
}

// dx[M,K] = dz[M,N] W[N,K] (+ dx if accumulate)
GNB_EXPORT int gnb_linear_bwd_data_f32(const float* dz, int64_t lddz, const float* w, int64_t ldw, float* dx,
                                       int64_t lddx, int64_t m, int64_t n, int64_t k, int32_t accumulate, void* stream) {
    return launch<false, true>(dz, lddz, w, ldw, dx, lddx, m, k, n, nullptr, GNB_ACT_NONE, accumulate, 1, false,
                               (cudaStream_t)stream);
}

// dW[N,K] += dz[M,N]^T x[M,K]   (dW must hold the value to accumulate onto, e.g. zeros; split over M)
GNB_EXPORT int gnb_linear_bwd_weight_f32(const float* dz, int64_t lddz, const float* x, int64_t ldx, float* dw,
                                         int64_t lddw, int64_t m, int64_t n, int64_t k, void* stream) {
    // enough K-splits to fill the machine: tiles(N) * tiles(K) * splits ~ 4 waves of 148 SMs
    const int tiles = gnb_div_up(n, BM) * gnb_div_up(k, BN);
    int splits = (4 * 148 + tiles - 1) / tiles;
    const int max_splits = (int)((m + 4 * BK - 1) / (4 * BK));
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    return launch<true, true>(dz, lddz, x, ldx, dw, lddw, n, k, m, nullptr, GNB_ACT_NONE, 0, splits, true,
                              (cudaStream_t)stream);
}
