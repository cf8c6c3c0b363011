// Shared helpers for the graphnet_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define GNB_OK 0
#define GNB_ERR_ARG (-1)          // bad argument (shape / alignment / unsupported size)
#define GNB_ERR_UNSUPPORTED (-2)  // valid request this build does not implement

#define GNB_EXPORT extern "C" __attribute__((visibility("default")))

// Number of kernels this library has launched (bench.py reports the per-step delta as `gpu_launches`).
extern "C" long long gnb_launch_counter;

// Launch-error check: returns the cudaError_t (>0) to the C-ABI caller; never throws.
#define GNB_RETURN_LAUNCH()                          \
    do {                                             \
        ++gnb_launch_counter;                        \
        cudaError_t e__ = cudaGetLastError();        \
        return e__ == cudaSuccess ? GNB_OK : (int)e__; \
    } while (0)

#define GNB_CHECK(call)                                  \
    do {                                                 \
        cudaError_t e__ = (call);                        \
        if (e__ != cudaSuccess) return (int)e__;         \
    } while (0)

static inline int gnb_div_up(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// ---- programmatic dependent launch (experimental, OFF unless GNB_PDL=1) ---------------------------------------------------
// A training step is ~110 kernels on one stream, each a full dependency of the next; measured (scripts/r02/gaps.py) the
// stream is idle 1.8 us (median) between two kernels, ~230 us = 3.5 % of the step. With GNB_PDL=1 every kernel of this library
// is launched with cudaLaunchAttributeProgrammaticStreamSerialization and starts with gnb_pdl_begin(): `launch_dependents` lets
// the NEXT kernel's CTAs be scheduled as soon as every CTA of this one has started, `griddepcontrol.wait` blocks them until
// this grid has completed and its writes are visible. Measured: 6.682 -> 6.599 ms per training step (+1.2 %), but
// tests/test_gpu_bf16.py::test_mixed16_unfused_forward_matches_the_fused_forward then FAILS (first-layer outputs differ): the
// kernels read their predecessor's output through the non-coherent path (__ldg / const __restrict__ -> LDG.CONSTANT), whose
// "read-only for the kernel's lifetime" contract an overlapping predecessor breaks. Not worth 1.2 %: the attribute is off by
// default, and without it the two device-side instructions are no-ops.
__device__ __forceinline__ void gnb_pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void gnb_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void gnb_pdl_begin() { gnb_pdl_trigger(); gnb_pdl_wait(); }
#ifdef __CUDACC__
#include <cstdlib>
#include <utility>
static inline bool gnb_pdl_enabled() {
    static const bool on = [] { const char* e = getenv("GNB_PDL"); return e != nullptr && e[0] == '1'; }();
    return on;
}
template <class K>
struct GnbLaunch {
    K kern; dim3 grid, block; size_t smem; cudaStream_t st;
    template <class... A> void operator()(A&&... a) const {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = gnb_pdl_enabled() ? 1 : 0;
        (void)cudaLaunchKernelEx(&cfg, kern, std::forward<A>(a)...);      // errors surface through cudaGetLastError() like <<< >>>
    }
};
template <class K>
static inline GnbLaunch<K> gnb_launch(K kern, dim3 grid, dim3 block, size_t smem, cudaStream_t st) { return GnbLaunch<K>{kern, grid, block, smem, st}; }
#endif

__device__ __forceinline__ float gnb_warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// activation codes shared by every epilogue
enum : int { GNB_ACT_NONE = 0, GNB_ACT_RELU = 1, GNB_ACT_LEAKY = 2 };      // LEAKY: torch.nn.LeakyReLU() default slope
#define GNB_LEAKY_SLOPE 0.01f
// OR-ed into an `act` / `aggr` argument: round the stored result to tf32 (cvt.rna) so that a following
// tcgen05 kind::tf32 GEMM, which truncates its fp32 operands, sees exactly representable values.
#define GNB_STD_MAX_F 32
enum : int { GNB_FLAG_ROUND_TF32 = 0x100, GNB_FLAG_ACCUMULATE = 0x200, GNB_FLAG_ZERO_SRC = 0x400, GNB_FLAG_HMASK_ROWMAJOR = 0x800 };   // 0x200 in `act` of the tf32 Linear: y += result

__device__ __forceinline__ float gnb_round_tf32(float v) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return __uint_as_float(r);
}

// Lane-interleaved order of a 64-column block of PQ for the fused EdgeConv forward (gnb_edgeconv_fused_fwd_f16, pq_layout 1):
// stored 16-byte piece j (j < 8) holds hidden units 8 j .. 8 j + 3, piece 8 + j holds units 8 j + 4 .. 8 j + 7.
__host__ __device__ __forceinline__ int gnb_pq_unit_of_stored(int s) { return s < 32 ? 8 * (s >> 2) + (s & 3) : 8 * ((s - 32) >> 2) + 4 + (s & 3); }
__host__ __device__ __forceinline__ int gnb_pq_stored_of_unit(int q) { return 4 * (q >> 3) + (q & 3) + 32 * ((q >> 2) & 1); }
// row r of [P half | Q half] (each hid rows): the same row with the permutation applied inside every FULL 64-row block of its half
__host__ __device__ __forceinline__ int gnb_pq_map_row(int r, int hid, bool to_unit) {
    const int half = r >= hid ? hid : 0, rr = r - half, blk = rr & ~63;
    if (blk + 64 > hid) return r;
    return half + blk + (to_unit ? gnb_pq_unit_of_stored(rr & 63) : gnb_pq_stored_of_unit(rr & 63));
}

// aggregation codes
enum : int { GNB_AGGR_ADD = 0, GNB_AGGR_MEAN = 1, GNB_AGGR_MAX = 2 };

// pooling scheme codes
enum : int { GNB_POOL_MIN = 0, GNB_POOL_MAX = 1, GNB_POOL_SUM = 2, GNB_POOL_MEAN = 3 };

// Power-of-two scaling of a gradient tensor stored as ONE fp16 plane (precision mode mixed16): `maxbits` = fp32 bits of
// max|g| (gnb_absmax_bits). Returns {scale, 1 / scale} with scale = 2^(14 - floor(log2 max)), so that max * scale lies in
// [2^14, 2^15) (fp16's largest finite value is 65504) and elements down to 2^-29 of the maximum keep a normal fp16 encoding.
// Exact in both directions (powers of two); 1 for an all-zero tensor.
__host__ __device__ __forceinline__ float2 gnb_pow2_scale(unsigned maxbits) {
    const int e = (int)((maxbits >> 23) & 0xffu);
    if (e == 0 || e == 255) return float2{1.f, 1.f};
    int se = 268 - e;                        // biased exponent of the scale: 127 + 14 - (e - 127)
    se = se < 2 ? 2 : (se > 252 ? 252 : se);
    const unsigned sb = (unsigned)se << 23, ib = (unsigned)(254 - se) << 23;
#ifdef __CUDA_ARCH__
    return float2{__uint_as_float(sb), __uint_as_float(ib)};
#else
    float s, i;
    __builtin_memcpy(&s, &sb, 4);
    __builtin_memcpy(&i, &ib, 4);
    return float2{s, i};
#endif
}
