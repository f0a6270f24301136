// Micro-benchmark: issue rate of small tcgen05.mma instructions from one elected thread (the attention kernels' shapes).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I tunevlseg_b200/csrc -I include tools/microbench/umma_rate.cu -o tools/microbench/umma_rate
// For M = 128, K = 16 (bf16) and N = 64 / 128 / 256, operands A from shared memory (SS) or from TMEM (TS): clk per instruction
// when `iters` x 4 MMAs are issued back to back and the whole stream is committed once.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include "sm100_ptx.cuh"

using namespace tvs::ptx;

template <int N, bool TS, bool MN_B>
__global__ void __launch_bounds__(128) umma_rate_kernel(int iters, unsigned long long* out) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~static_cast<uintptr_t>(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;      // finite bf16 pairs
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(&slot, 512);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = slot;
    if (warp == 0) {
        const uint64_t a0 = umma_desc_sw128(smem_u32(smem));
        const uint64_t b0 = umma_desc_sw128(smem_u32(smem + 16384));
        const uint32_t idesc = umma_idesc_bf16(128, N, 0, MN_B ? 1 : 0);
        long long t0 = 0, t1 = 0;
        if (elect_one()) {
            t0 = clock64();
            for (int it = 0; it < iters; ++it) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (TS) umma_ts(tb, tb + 256 + 8 * k, b0 + (MN_B ? 128 * k : 2 * k), idesc, (it | k) ? 1u : 0u);
                    else umma_ss(tb, a0 + 2 * k, b0 + 2 * k, idesc, (it | k) ? 1u : 0u);
                }
            }
            t1 = clock64();
            umma_commit(&bar);
        }
        __syncwarp();
        mbar_wait(&bar, 0);
        const long long t2 = clock64();
        if (threadIdx.x == 0) out[blockIdx.x * 2 + 1] = t2;
        if (t0) { out[blockIdx.x * 2] = static_cast<unsigned long long>(t1 - t0); out[blockIdx.x * 2 + 1] = static_cast<unsigned long long>(t2 - t0); }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tb, 512);
    }
}

template <int N, bool TS, bool MN_B>
static void run(const char* what, int iters) {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    unsigned long long* d;
    cudaMalloc(&d, sizeof(unsigned long long) * 2 * sms);
    cudaMemset(d, 0, sizeof(unsigned long long) * 2 * sms);
    const int dsm = 16384 + 32768 + 2048;
    cudaFuncSetAttribute(umma_rate_kernel<N, TS, MN_B>, cudaFuncAttributeMaxDynamicSharedMemorySize, dsm);
    for (int rep = 0; rep < 2; ++rep) umma_rate_kernel<N, TS, MN_B><<<sms, 128, dsm>>>(iters, d);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { printf("%s: %s\n", what, cudaGetErrorString(err)); exit(1); }
    unsigned long long* h = (unsigned long long*)malloc(sizeof(unsigned long long) * 2 * sms);
    cudaMemcpy(h, d, sizeof(unsigned long long) * 2 * sms, cudaMemcpyDeviceToHost);
    double issue = 0, total = 0;
    for (int i = 0; i < sms; ++i) { issue += h[2 * i]; total += h[2 * i + 1]; }
    const double n = 4.0 * iters;
    printf("%-44s N=%3d: issue %6.1f clk / MMA, issue + drain %6.1f clk / MMA (ideal %3d)\n", what, N, issue / sms / n, total / sms / n, N / 2);
    free(h);
    cudaFree(d);
}

int main() {
    const int iters = 2000;
    run<64, false, false>("SS  A smem K-major, B smem K-major", iters);
    run<128, false, false>("SS  A smem K-major, B smem K-major", iters);
    run<256, false, false>("SS  A smem K-major, B smem K-major", iters);
    run<64, true, true>("TS  A tmem, B smem MN-major", iters);
    run<128, true, true>("TS  A tmem, B smem MN-major", iters);
    run<64, true, false>("TS  A tmem, B smem K-major", iters);
    return 0;
}
