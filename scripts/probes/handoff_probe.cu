// Hand-off latencies of the tcgen05 pipeline primitives (cycles, one CTA): build with
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I graphnet_b200/csrc -o gpurun_out/handoff_probe scripts/probes/handoff_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
#include "tc_common.cuh"

__device__ __forceinline__ uint32_t test_wait(uint32_t addr, uint32_t parity) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    return done;
}

// mode 0: commit -> completion seen by the same thread (test_wait spin)
// mode 1: ping-pong lane0 <-> lane0 of another warp with try_wait
// mode 2: ping-pong with test_wait spin
// mode 3: ping-pong, whole warps polling try_wait + vote
// mode 4: commit by warp 1 -> seen by warp 0 (try_wait), warp 0 arrives -> seen by warp 1 (try_wait): the pipeline loop
__global__ void probe(int mode, int iters, long long* out) {
    __shared__ uint64_t bar[2];
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { tc::mbar_init(&bar[0], 1); tc::mbar_init(&bar[1], 1); tc::fence_barrier_init(); }
    if (warp == 0) tc::tmem_alloc<256>(&slot);
    tc::tcgen05_fence_before();
    __syncthreads();
    tc::tcgen05_fence_after();
    const uint32_t a0 = tc::smem_u32(&bar[0]), a1 = tc::smem_u32(&bar[1]);
    long long t0 = clock64();
    if (mode == 0) {
        if (threadIdx.x == 0)
            for (int i = 0; i < iters; ++i) { tc::umma_commit(&bar[0]); while (!test_wait(a0, i & 1)) {} }
    } else if (mode == 1 || mode == 2) {
        if (lane == 0 && warp < 2)
            for (int i = 0; i < iters; ++i) {
                if (warp == 0) {
                    tc::mbar_arrive(&bar[0]);
                    if (mode == 1) while (!tc::mbar_try_wait(a1, i & 1)) {} else while (!test_wait(a1, i & 1)) {}
                } else {
                    if (mode == 1) while (!tc::mbar_try_wait(a0, i & 1)) {} else while (!test_wait(a0, i & 1)) {}
                    tc::mbar_arrive(&bar[1]);
                }
            }
    } else if (mode == 3) {
        if (warp < 2)
            for (int i = 0; i < iters; ++i) {
                if (warp == 0) {
                    if (tc::elect_one()) tc::mbar_arrive(&bar[0]);
                    __syncwarp();
                    tc::mbar_wait_warp(&bar[1], i & 1);
                } else {
                    tc::mbar_wait_warp(&bar[0], i & 1);
                    if (tc::elect_one()) tc::mbar_arrive(&bar[1]);
                    __syncwarp();
                }
            }
    } else if (mode == 4) {
        if (warp < 2)
            for (int i = 0; i < iters; ++i) {
                if (warp == 0) {          // "producer": waits for the commit, then arrives
                    tc::mbar_wait_warp(&bar[0], i & 1);
                    if (tc::elect_one()) tc::mbar_arrive(&bar[1]);
                    __syncwarp();
                } else {                  // "MMA": commit, then waits for the producer
                    if (tc::elect_one()) tc::umma_commit(&bar[0]);
                    __syncwarp();
                    tc::mbar_wait_warp(&bar[1], i & 1);
                    tc::tcgen05_fence_after();
                }
            }
    }
    else if (mode == 5) {          // back-to-back commits, no MMAs, no waits: commit issue cost
        if (warp == 1)
            for (int i = 0; i < iters; ++i) { if (tc::elect_one()) tc::umma_commit(&bar[i & 1]); __syncwarp(); }
    } else if (mode >= 6 && mode <= 8) {        // 8 MMAs (128x128x8 tf32) + 1 commit per iteration, operands = whatever is in smem
        extern __shared__ uint8_t dyn[];
        uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(dyn) + 1023) & ~uintptr_t(1023));
        if (warp == 1) {
            constexpr uint32_t idesc = tc::umma_idesc_tf32(128, 128);
            for (int i = 0; i < iters; ++i) {
                if (mode == 7) { tc::mbar_wait_warp(&bar[1], 1); tc::tcgen05_fence_after(); }   // always already complete
                const uint32_t st = tc::smem_u32(sm + (i & 3) * 49152);
                const uint64_t bdesc = tc::umma_desc_sw128_kmajor(st + 32768);
                for (int m = 0; m < 2; ++m) {
                    const uint64_t adesc = tc::umma_desc_sw128_kmajor(st + m * 16384);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (tc::elect_one()) tc::umma_tf32(slot + m * 128, adesc + 2 * k, bdesc + 2 * k, idesc, 1u);
                }
                if (mode != 8) { if (tc::elect_one()) tc::umma_commit(&bar[0]); }
                __syncwarp();
            }
            if (tc::elect_one()) tc::umma_commit(&bar[0]);
            __syncwarp();
            // wait until everything retired: parity of the last completed phase
            const int nph = (mode != 8 ? iters : 0) + 1;
            tc::mbar_wait_warp(&bar[0], (nph - 1) & 1);
        }
    }
    if (mode >= 9 && mode <= 12) {     // shape sweep: 8 MMAs of 128 x N x 8 per iteration into one accumulator
        extern __shared__ uint8_t dyn[];
        uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(dyn) + 1023) & ~uintptr_t(1023));
        if (warp == 1) {
            const uint32_t n = mode == 9 ? 256u : (mode == 10 ? 64u : (mode == 11 ? 32u : 128u));
            const uint32_t idesc = tc::umma_idesc_tf32(128, n);
            for (int i = 0; i < iters; ++i) {
                const uint32_t st = tc::smem_u32(sm + (i & 1) * 65536);
                const uint64_t adesc = tc::umma_desc_sw128_kmajor(st);
                const uint64_t bdesc = tc::umma_desc_sw128_kmajor(st + 16384);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    // mode 12: N = 128 but the SAME K slice every time (is it the smem fetch?)
                    const int kk = mode == 12 ? 0 : (k & 3);
                    if (tc::elect_one()) tc::umma_tf32(slot, adesc + 2 * kk, bdesc + 2 * kk, idesc, 1u);
                }
                __syncwarp();
            }
            if (tc::elect_one()) tc::umma_commit(&bar[0]);
            __syncwarp();
            tc::mbar_wait_warp(&bar[0], 0);
        }
    }
    long long t1 = clock64();
    if (lane == 0 && warp < 2) out[warp] = t1 - t0;
    tc::tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<256>(slot);
}

int main() {
    long long* d;
    cudaMalloc(&d, 16);
    const int iters = 2000;
    const char* names[] = {"commit -> own test_wait", "ping-pong try_wait (1 lane)", "ping-pong test_wait (1 lane)",
                           "ping-pong try_wait + vote (warp)", "commit -> warp wait -> arrive -> warp wait",
                           "back-to-back commits", "8 MMA + commit per iter", "wait + fence + 8 MMA + commit per iter",
                           "8 MMA per iter, one commit at the end", "8 MMA 128x256x8 per iter", "8 MMA 128x64x8 per iter",
                           "8 MMA 128x32x8 per iter", "8 MMA 128x128x8, same K slice"};
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 49152 + 1024);
    for (int mode = 0; mode < 13; ++mode) {
        probe<<<1, 64, 4 * 49152 + 1024>>>(mode, iters, d);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[2] = {0, 0};
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("mode %d %-45s: %s  %.1f cycles per round trip (warp0), %.1f (warp1)\n", mode, names[mode], cudaGetErrorString(e),
               (double)h[0] / iters, (double)h[1] / iters);
    }
    return 0;
}
