// cta_group::2 MMA issue/throughput probe: cluster of two CTAs, the leader issues 4 x (M256 x N x K8, tf32) per iteration.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I graphnet_b200/csrc -o scripts/probes/_build/pair_mma_probe scripts/probes/pair_mma_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
#include "tc_common.cuh"

__global__ void __cluster_dims__(2, 1, 1) probe(int n, int iters, long long* out, int mode) {
    extern __shared__ uint8_t dyn[];
    uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(dyn) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar[2];
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = tc::cluster_ctarank();
    if (threadIdx.x == 0) { tc::mbar_init(&bar[0], 1); tc::mbar_init(&bar[1], 1); tc::fence_barrier_init(); }
    if (warp == 0) tc::tmem_alloc_2cta<512>(&slot);
    tc::tcgen05_fence_before();
    __syncthreads();
    tc::cluster_sync_all();
    tc::tcgen05_fence_after();
    long long t0 = clock64();
    if (rank == 0 && warp == 1) {
        const uint32_t idesc = tc::umma_idesc_tf32(256, (uint32_t)n);
        for (int i = 0; i < iters; ++i) {
            const uint32_t st = tc::smem_u32(sm + (i & 1) * 65536);
            const uint64_t adesc = tc::umma_desc_sw128_kmajor(st);
            const uint64_t bdesc = tc::umma_desc_sw128_kmajor(st + 16384);
            if (mode == 0) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (tc::elect_one()) tc::umma_tf32_2cta(slot, adesc + 2 * k, bdesc + 2 * k, idesc, 1u);
            } else {          // one election per K block: the four MMAs issued back to back by the elected lane
                if (tc::elect_one()) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) tc::umma_tf32_2cta(slot, adesc + 2 * k, bdesc + 2 * k, idesc, 1u);
                }
            }
            __syncwarp();
        }
        if (tc::elect_one()) tc::umma_commit_2cta(&bar[0], 3);
        __syncwarp();
    }
    if (warp == 1) tc::mbar_wait_warp(&bar[0], 0);
    long long t1 = clock64();
    if (lane == 0 && warp == 1) out[rank] = t1 - t0;
    tc::tcgen05_fence_before();
    __syncthreads();
    tc::cluster_sync_all();
    if (warp == 0) tc::tmem_dealloc_2cta<512>(slot);
}

int main() {
    long long* d;
    cudaMalloc(&d, 16);
    const int iters = 2000;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 65536 + 1024);
    for (int mode = 0; mode < 2; ++mode)
    for (int n : {256, 64}) {
        probe<<<2, 64, 2 * 65536 + 1024>>>(n, iters, d, mode);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[2] = {0, 0};
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("mode %d cta_group::2 M256 x N%d x K8: %s  %.1f cycles per 4 MMAs (leader), %.1f (peer)\n", mode, n, cudaGetErrorString(e),
               (double)h[0] / iters, (double)h[1] / iters);
    }
    return 0;
}
