// fp32 reduction throughput into a [n, ld] matrix: scalar red (warp = 32 channels of one row) vs v4 red (warp = 4 rows x
// 32 channels), rows pseudo-random within +-64 of the warp's home row (kNN-like).  Same number of floats either way.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/probes/_build/red_probe scripts/probes/red_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t hash(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

template <int MODE>
__global__ void red_kernel(float* out, int n, int ld, int chans, int iters) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const int home = (int)((int64_t)warp * n / nwarps);
    for (int it = 0; it < iters; ++it) {
        for (int cb = 0; cb < chans; cb += 32) {
            if (MODE == 0) {
#pragma unroll 4
                for (int a = 0; a < 4; ++a) {
                    int row = home + (int)(hash(warp * 977 + it * 4 + a) % 129) - 64;
                    row = row < 0 ? 0 : (row >= n ? n - 1 : row);
                    atomicAdd(out + (int64_t)row * ld + cb + lane, 1.0f);
                }
            } else {
                int row = home + (int)(hash(warp * 977 + it * 4 + (lane & 3)) % 129) - 64;
                row = row < 0 ? 0 : (row >= n ? n - 1 : row);
                float4* p = reinterpret_cast<float4*>(out + (int64_t)row * ld + cb + (lane & ~3));
                atomicAdd(p, make_float4(1.f, 1.f, 1.f, 1.f));
            }
        }
    }
}

int main() {
    const int n = 79261, ld = 672, chans = 320;
    float* d;
    cudaMalloc(&d, (size_t)n * ld * 4);
    cudaMemset(d, 0, (size_t)n * ld * 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = 148 * 4, threads = 256, iters = 300;      // 4736 warps
    const double floats = (double)blocks * threads / 32 * iters * 4 * chans * 32 / 32;   // per warp-iter: 4 rows x chans
    for (int mode = 0; mode < 2; ++mode) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) red_kernel<0><<<blocks, threads>>>(d, n, ld, chans, iters);
            else red_kernel<1><<<blocks, threads>>>(d, n, ld, chans, iters);
            cudaEventRecord(e1);
            cudaError_t err = cudaDeviceSynchronize();
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            if (rep == 1)
                printf("%s: %s %.3f ms, %.1f G fp32 reductions/s, %.2f TB/s of reduced bytes\n", mode == 0 ? "scalar red.f32" : "vector red.v4.f32",
                       cudaGetErrorString(err), ms, floats / ms / 1e6, floats * 4 / ms / 1e9);
        }
    }
    return 0;
}
