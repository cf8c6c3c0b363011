// Adam step over ONE flat fp32 parameter buffer (the step that follows the hot path in a training iteration).
//
// Replaces torch.optim.Adam as configured by the reference (src/graphnet/models/easy_model.py:215-219 with
// examples/04_training/01_train_dynedge.py:128-129: optimizer_class=Adam, lr=1e-3, eps=1e-3): same update, same fp32
// state, but one launch over the flat buffer that already holds every gradient for the single NCCL all-reduce
// (graphnet_b200/distributed.py) instead of a multi-tensor launch over 26 small tensors, and the gradient buffer is
// zeroed behind the read so the next step needs no separate fill.
//   g' = g + wd p;  m = m + (1 - b1)(g' - m);  v = b2 v + (1 - b2) g'^2;  p -= (lr / (1 - b1^t)) m / (sqrt(v) / sqrt(1 - b2^t) + eps)
#include "common.cuh"

namespace {

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, float step_size, float b1, float b2, float eps,
                                         float inv_sqrt_bc2, float wd) {
    g = wd != 0.f ? g + wd * p : g;
    m = m + (1.f - b1) * (g - m);
    v = b2 * v + (1.f - b2) * g * g;
    const float denom = sqrtf(v) * inv_sqrt_bc2 + eps;
    p -= step_size * (m / denom);
}

__global__ void __launch_bounds__(256)
adam_flat_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, int64_t n,
                 float step_size, float b1, float b2, float eps, float inv_sqrt_bc2, float wd, float gs, int zero_grad) {
    gnb_pdl_begin();
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i >= n) return;
    if (i + 4 <= n) {
        float4 pv = *reinterpret_cast<float4*>(p + i), gv = *reinterpret_cast<float4*>(g + i);
        gv.x *= gs; gv.y *= gs; gv.z *= gs; gv.w *= gs;
        float4 mv = *reinterpret_cast<float4*>(m + i), vv = *reinterpret_cast<float4*>(v + i);
        adam_one(pv.x, gv.x, mv.x, vv.x, step_size, b1, b2, eps, inv_sqrt_bc2, wd);
        adam_one(pv.y, gv.y, mv.y, vv.y, step_size, b1, b2, eps, inv_sqrt_bc2, wd);
        adam_one(pv.z, gv.z, mv.z, vv.z, step_size, b1, b2, eps, inv_sqrt_bc2, wd);
        adam_one(pv.w, gv.w, mv.w, vv.w, step_size, b1, b2, eps, inv_sqrt_bc2, wd);
        *reinterpret_cast<float4*>(p + i) = pv;
        *reinterpret_cast<float4*>(m + i) = mv;
        *reinterpret_cast<float4*>(v + i) = vv;
        if (zero_grad) *reinterpret_cast<float4*>(g + i) = make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
        for (int64_t j = i; j < n; ++j) {
            adam_one(p[j], g[j] * gs, m[j], v[j], step_size, b1, b2, eps, inv_sqrt_bc2, wd);
            if (zero_grad) g[j] = 0.f;
        }
    }
}

}  // namespace

// step_size = lr / (1 - beta1^t), inv_sqrt_bc2 = 1 / sqrt(1 - beta2^t) (computed by the caller in double precision).
// p, g, m, v: [n] fp32, 16-byte aligned. zero_grad != 0 leaves g zeroed. grad_scale multiplies g on the way in: 1 / world_size
// turns the SUM all-reduce of the flat gradient buffer into DDP's mean (easy_model.py:90-110) without a separate division pass.
GNB_EXPORT int gnb_adam_flat(float* p, float* g, float* m, float* v, int64_t n, float step_size, float beta1, float beta2,
                             float eps, float inv_sqrt_bc2, float weight_decay, float grad_scale, int32_t zero_grad, void* stream) {
    if (n < 0 || !(beta1 >= 0.f && beta1 < 1.f) || !(beta2 >= 0.f && beta2 < 1.f) || !(eps >= 0.f)) return GNB_ERR_ARG;
    if (((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
          reinterpret_cast<uintptr_t>(v)) & 15u) != 0)
        return GNB_ERR_ARG;
    if (n == 0) return GNB_OK;
    gnb_launch(adam_flat_kernel, gnb_div_up(gnb_div_up(n, 4), 256), 256, 0, (cudaStream_t)stream)(p, g, m, v, n, step_size, beta1, beta2, eps,
                                                                                           inv_sqrt_bc2, weight_decay, grad_scale, zero_grad);
    GNB_RETURN_LAUNCH();
}
