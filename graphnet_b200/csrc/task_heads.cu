// Task heads + losses of the training benchmark (BASELINE config #3) fused into two kernels:
//   EnergyReconstruction  (src/graphnet/models/task/reconstruction.py:101-112): Linear(H -> 1), softplus(beta=0.05) + eps
//   LogCoshLoss on log10   (src/graphnet/training/loss_functions.py:93-112; examples/04_training/01_train_dynedge.py:113-124)
//   DirectionReconstructionWithKappa (reconstruction.py:49-70): Linear(H -> 3), kappa = |z| + eps, (z / kappa, kappa)
//   VonMisesFisher3DLoss   (loss_functions.py:281-353, 424-447) with log C_3(kappa) in closed form, i.e. what the
//                           reference's scipy-Bessel `LogCMK` evaluates for m = 3 (tests/training/test_loss_functions.py:66-95)
//                           without its device -> host round trip every step.
// Forward: one warp per event computes both affine heads, both predictions, both per-event loss terms and the
// derivatives of the summed loss w.r.t. the four affine outputs; backward: dfeat / dW / db from those derivatives
// (one warp per event, shared-memory reduction over the CTA's events, one fp32 atomic per weight and CTA).
#include "common.cuh"

namespace {

constexpr int TH_WARPS = 8;
constexpr float TH_EPS = 1.1920928955078125e-07f;       // torch.finfo(torch.float32).eps (graphnet eps_like)
constexpr float TH_LN10 = 2.302585092994046f, TH_LN2 = 0.6931471805599453f, TH_LOG_2PI = 1.8378770664093453f;

__device__ __forceinline__ float softplus1(float x) {    // torch softplus(beta = 1, threshold = 20)
    return x > 20.f ? x : log1pf(expf(x));
}

// VonMisesFisherLoss.log_cmk (loss_functions.py:307-326) for m = 3: exact below kappa_switch = 100, from there on the
// approximation -sqrt(4 + k^2) shifted by approx(100) - exact(100) for continuity.
constexpr float TH_KAPPA_SWITCH = 100.f, TH_LOG_C3_OFFSET = -2.78729248046875f;   // the reference's float32 offset
__device__ __forceinline__ float log_c3(float k) {
    if (k >= TH_KAPPA_SWITCH) return -sqrtf(4.f + k * k) - TH_LOG_C3_OFFSET;
    return logf(k) - TH_LOG_2PI - k - logf(-expm1f(-2.f * k));
}
// d/dk log C_3(k) = 1/k - coth(k); series below 0.1 (the closed form cancels catastrophically there)
__device__ __forceinline__ float dlog_c3(float k) {
    if (k >= TH_KAPPA_SWITCH) return -k / sqrtf(4.f + k * k);
    if (k < 0.1f) { const float k2 = k * k; return k * (-1.f / 3.f + k2 * (1.f / 45.f - k2 * (2.f / 945.f))); }
    const float e = expf(-2.f * k);
    return 1.f / k - (1.f + e) / (1.f - e);
}

__global__ void __launch_bounds__(TH_WARPS * 32)
task_heads_fwd_kernel(const float* __restrict__ feat, int64_t ldf, int hdim, const float* __restrict__ we,
                      const float* __restrict__ be, const float* __restrict__ wd, const float* __restrict__ bd,
                      const float* __restrict__ energy, const float* __restrict__ direction, int64_t nev, float inv_n,
                      float* __restrict__ pred_e, float* __restrict__ pred_d, float* __restrict__ dz,
                      float* __restrict__ loss) {
    gnb_pdl_begin();
    __shared__ float s_loss[TH_WARPS][2];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t ev = (int64_t)blockIdx.x * TH_WARPS + warp;
    float le = 0.f, ld = 0.f;
    if (ev < nev) {
        const float* f = feat + ev * ldf;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        for (int h = lane; h < hdim; h += 32) {
            const float v = f[h];
            a0 += v * we[h]; a1 += v * wd[h]; a2 += v * wd[hdim + h]; a3 += v * wd[2 * hdim + h];
        }
        a0 = gnb_warp_sum(a0); a1 = gnb_warp_sum(a1); a2 = gnb_warp_sum(a2); a3 = gnb_warp_sum(a3);
        if (lane == 0) {
            // ---- energy: softplus(z, beta = 0.05) + eps; log-cosh of the log10 residual
            const float ze = a0 + be[0];
            const float bz = 0.05f * ze;
            const float sp = bz > 20.f ? ze : log1pf(expf(bz)) / 0.05f;
            const float pe = sp + TH_EPS;
            const float diff = log10f(pe) - log10f(energy[ev]);
            le = diff + softplus1(-2.f * diff) - TH_LN2;
            const float dsp = bz > 20.f ? 1.f : 1.f / (1.f + expf(-bz));          // d softplus / dz
            const float dze = tanhf(diff) / (pe * TH_LN10) * dsp;
            // ---- direction: kappa = |z| + eps, prediction (z / kappa, kappa); vMF-3D negative log-likelihood
            const float zx = a1 + bd[0], zy = a2 + bd[1], zz = a3 + bd[2];
            const float nz = sqrtf(zx * zx + zy * zy + zz * zz);
            const float kappa = nz + TH_EPS;
            const float ux = zx / kappa, uy = zy / kappa, uz = zz / kappa;
            const float px = kappa * ux, py = kappa * uy, pz = kappa * uz;       // what the loss rebuilds from the prediction
            const float kn = sqrtf(px * px + py * py + pz * pz);
            const float tx = direction[ev * 3], ty = direction[ev * 3 + 1], tz = direction[ev * 3 + 2];
            const float logc = log_c3(kn);
            ld = -logc - (px * tx + py * ty + pz * tz);
            const float s = kn > 1e-30f ? -dlog_c3(kn) / kn : 0.f;
            pred_e[ev] = pe;
            pred_d[ev * 4] = ux; pred_d[ev * 4 + 1] = uy; pred_d[ev * 4 + 2] = uz; pred_d[ev * 4 + 3] = kappa;
            dz[ev * 4] = dze * inv_n;
            dz[ev * 4 + 1] = (s * px - tx) * inv_n; dz[ev * 4 + 2] = (s * py - ty) * inv_n; dz[ev * 4 + 3] = (s * pz - tz) * inv_n;
        }
    }
    if (lane == 0) { s_loss[warp][0] = le; s_loss[warp][1] = ld; }
    __syncthreads();
    if (threadIdx.x < 2) {
        float v = 0.f;
        for (int w = 0; w < TH_WARPS; ++w) v += s_loss[w][threadIdx.x];
        atomicAdd(loss + threadIdx.x, v * inv_n);
    }
}

// dfeat[ev, :] = g * (dz_e We + sum_c dz_c Wd[c]);  dWe += g * sum_ev dz_e feat[ev], ... (g = upstream scalar gradient)
__global__ void __launch_bounds__(TH_WARPS * 32)
task_heads_bwd_kernel(const float* __restrict__ feat, int64_t ldf, int hdim, const float* __restrict__ we,
                      const float* __restrict__ wd, const float* __restrict__ dz, const float* __restrict__ gout,
                      int64_t nev, float* __restrict__ dfeat, int64_t lddf, float* __restrict__ dwe,
                      float* __restrict__ dbe, float* __restrict__ dwd, float* __restrict__ dbd) {
    gnb_pdl_begin();
    extern __shared__ float s_acc[];            // [4][hdim] weight-gradient partial sums of this CTA + [4] bias sums
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float g = gout != nullptr ? gout[0] : 1.f;
    for (int t = threadIdx.x; t < 4 * hdim + 4; t += blockDim.x) s_acc[t] = 0.f;
    __syncthreads();
    const int64_t ev = (int64_t)blockIdx.x * TH_WARPS + warp;
    if (ev < nev) {
        const float d0 = dz[ev * 4] * g, d1 = dz[ev * 4 + 1] * g, d2 = dz[ev * 4 + 2] * g, d3 = dz[ev * 4 + 3] * g;
        const float* f = feat + ev * ldf;
        for (int h = lane; h < hdim; h += 32) {
            const float v = f[h];
            if (dfeat != nullptr) dfeat[ev * lddf + h] = d0 * we[h] + d1 * wd[h] + d2 * wd[hdim + h] + d3 * wd[2 * hdim + h];
            atomicAdd(&s_acc[h], d0 * v); atomicAdd(&s_acc[hdim + h], d1 * v);
            atomicAdd(&s_acc[2 * hdim + h], d2 * v); atomicAdd(&s_acc[3 * hdim + h], d3 * v);
        }
        if (lane == 0) {
            atomicAdd(&s_acc[4 * hdim], d0); atomicAdd(&s_acc[4 * hdim + 1], d1);
            atomicAdd(&s_acc[4 * hdim + 2], d2); atomicAdd(&s_acc[4 * hdim + 3], d3);
        }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < hdim; t += blockDim.x) {
        atomicAdd(dwe + t, s_acc[t]);
        atomicAdd(dwd + t, s_acc[hdim + t]); atomicAdd(dwd + hdim + t, s_acc[2 * hdim + t]);
        atomicAdd(dwd + 2 * hdim + t, s_acc[3 * hdim + t]);
    }
    if (threadIdx.x == 0) atomicAdd(dbe, s_acc[4 * hdim]);
    if (threadIdx.x < 3) atomicAdd(dbd + threadIdx.x, s_acc[4 * hdim + 1 + threadIdx.x]);
}

}  // namespace

// feat [nev, hdim] (DynEdge output); we [hdim], be [1]: energy head; wd [3, hdim], bd [3]: direction head;
// energy [nev] (> 0), direction [nev, 3] (unit vectors). Outputs: pred_e [nev], pred_d [nev, 4] = (unit vector, kappa),
// dz [nev, 4] = d(loss_e + loss_d)/d(affine outputs) already divided by nev, loss[2] += (mean log-cosh, mean vMF NLL)
// (loss must be zero on entry).
GNB_EXPORT int gnb_task_heads_fwd(const float* feat, int64_t ldf, int32_t hdim, const float* we, const float* be,
                                  const float* wd, const float* bd, const float* energy, const float* direction,
                                  int64_t nev, float* pred_e, float* pred_d, float* dz, float* loss, void* stream) {
    if (hdim < 1 || nev < 0 || ldf < hdim) return GNB_ERR_ARG;
    if (nev == 0) return GNB_OK;
    gnb_launch(task_heads_fwd_kernel, gnb_div_up(nev, TH_WARPS), TH_WARPS * 32, 0, (cudaStream_t)stream)(
        feat, ldf, hdim, we, be, wd, bd, energy, direction, nev, 1.f / (float)nev, pred_e, pred_d, dz, loss);
    GNB_RETURN_LAUNCH();
}

// Backward of the above for an upstream scalar gradient gout[0] (NULL = 1): dfeat [nev, hdim] is written (may be NULL),
// dwe [hdim], dbe [1], dwd [3, hdim], dbd [3] are ACCUMULATED (+=), so they can be the parameters' .grad buffers.
GNB_EXPORT int gnb_task_heads_bwd(const float* feat, int64_t ldf, int32_t hdim, const float* we, const float* wd,
                                  const float* dz, const float* gout, int64_t nev, float* dfeat, int64_t lddf, float* dwe,
                                  float* dbe, float* dwd, float* dbd, void* stream) {
    if (hdim < 1 || hdim > 2048 || nev < 0 || ldf < hdim) return GNB_ERR_ARG;
    if (nev == 0) return GNB_OK;
    const size_t smem = (size_t)(4 * hdim + 4) * sizeof(float);
    gnb_launch(task_heads_bwd_kernel, gnb_div_up(nev, TH_WARPS), TH_WARPS * 32, smem, (cudaStream_t)stream)(
        feat, ldf, hdim, we, wd, dz, gout, nev, dfeat, lddf, dwe, dbe, dwd, dbd);
    GNB_RETURN_LAUNCH();
}
