// Micro-benchmark: how fast can the element-wise warps of the attention kernels read TMEM, and what else limits them?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I tunevlseg_b200/csrc -I include tools/microbench/tmem_bw.cu -o tools/microbench/tmem_bw
// Prints bytes / clk / SM for tcgen05.ld (32x32b.x32) with 4 / 8 / 16 warps per CTA and 1 / 2 CTAs per SM, the same with the
// loads followed by one MUFU.EX2 + pack per element (the softmax stage), and tcgen05.st.  Measured numbers go to DESIGN.md.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include "sm100_ptx.cuh"

using namespace tvs::ptx;

__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
}

// MODE 0: loads only (two x32 loads in flight per wait); 1: loads + exp2 + pack per element; 2: + store of the packed half;
// 3: stores only; 4: exp2 + pack only (no TMEM)
template <int MODE>
__global__ void tmem_bw_kernel(int iters, int cols, unsigned long long* clk_out, uint32_t* sink) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) tmem_alloc(&slot, cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = slot;
    const uint32_t tl = tb + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    const int grp = warp >> 2;                 // warps of one lane quarter use different column blocks
    const int nblk = cols / 64;
    uint32_t acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        const uint32_t c = 64u * ((grp + it) % nblk);
        uint32_t a[32], b[32];
        if (MODE <= 2) {
            tmem_ld32(tl + c, a);
            tmem_ld32(tl + c + 32, b);
            tmem_ld_wait();
        } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) { a[i] = (it + i) ^ acc; b[i] = (it - i) ^ acc; }
        }
        if (MODE == 0) {
            acc ^= a[0] ^ b[31] ^ a[17];
        } else if (MODE == 1 || MODE == 2 || MODE == 4) {
            uint32_t pa[16], pb[16];
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
                pa[i / 2] = pack2(ex2(__uint_as_float(a[i]) * 1.44f - 3.f), ex2(__uint_as_float(a[i + 1]) * 1.44f - 3.f));
                pb[i / 2] = pack2(ex2(__uint_as_float(b[i]) * 1.44f - 3.f), ex2(__uint_as_float(b[i + 1]) * 1.44f - 3.f));
            }
            if (MODE == 2) {
                tmem_st16(tl + c, pa);
                tmem_st16(tl + c + 32, pb);
                tmem_st_wait();
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) acc ^= pa[i] ^ pb[i];
            }
        } else if (MODE == 3) {
            tmem_st32(tl + c, a);
            tmem_st32(tl + c + 32, b);
            tmem_st_wait();
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x % 32 == 0) clk_out[blockIdx.x * (blockDim.x / 32) + warp] = static_cast<unsigned long long>(t1 - t0);
    if (acc == 0x12345678u) sink[0] = acc;
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tb, cols);
    }
}

template <int MODE>
static void run(const char* what, int warps, int ctas_per_sm, int iters) {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int cols = ctas_per_sm == 1 ? 512 : 256;
    const int grid = sms * ctas_per_sm;
    unsigned long long* clk;
    uint32_t* sink;
    cudaMalloc(&clk, sizeof(unsigned long long) * grid * warps);
    cudaMalloc(&sink, 4);
    // dynamic shared memory pads the CTA so that exactly ctas_per_sm fit
    const int dsm = ctas_per_sm == 1 ? 120 * 1024 : 60 * 1024;
    cudaFuncSetAttribute(tmem_bw_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, dsm);
    tmem_bw_kernel<MODE><<<grid, warps * 32, dsm>>>(iters, cols, clk, sink);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    tmem_bw_kernel<MODE><<<grid, warps * 32, dsm>>>(iters, cols, clk, sink);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) {
        printf("%s: %s\n", what, cudaGetErrorString(err));
        exit(1);
    }
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    unsigned long long* h = (unsigned long long*)malloc(sizeof(unsigned long long) * grid * warps);
    cudaMemcpy(h, clk, sizeof(unsigned long long) * grid * warps, cudaMemcpyDeviceToHost);
    unsigned long long mx = 0;
    for (int i = 0; i < grid * warps; ++i) mx = h[i] > mx ? h[i] : mx;
    const double elems_per_sm = double(ctas_per_sm) * warps * iters * 2.0 * 32 * 32;        // 32-bit TMEM words (or exps) per SM
    printf("%-34s warps/CTA %2d CTAs/SM %d: %7.1f B/clk/SM (%5.2f elements/clk/SM), %8.1f us, %llu clk\n", what, warps, ctas_per_sm,
           4.0 * elems_per_sm / double(mx), elems_per_sm / double(mx), ms * 1e3, mx);
    free(h);
    cudaFree(clk);
    cudaFree(sink);
}

int main() {
    const int iters = 4000;
    for (int cps = 1; cps <= 2; ++cps)
        for (int w : {4, 8, 16}) {
            run<0>("tcgen05.ld x32 only", w, cps, iters);
            run<1>("ld + ex2 + pack", w, cps, iters);
            run<2>("ld + ex2 + pack + st x16", w, cps, iters);
            run<3>("tcgen05.st x32 only", w, cps, iters);
            run<4>("ex2 + pack only (no TMEM)", w, cps, iters);
        }
    return 0;
}
