// Shared device/host helpers for the tunevlseg_b200 CUDA library (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace tvs {

// ---- error plumbing: every C-ABI entry returns 0 on success, <0 on error; message via tvs_last_error() ----
void set_error(const char* fmt, ...);
int check_launch(const char* what);

#define TVS_REQUIRE(cond, ...)                 \
    do {                                       \
        if (!(cond)) {                         \
            ::tvs::set_error(__VA_ARGS__);     \
            return -1;                         \
        }                                      \
    } while (0)

#define TVS_CUDA(expr)                                                                   \
    do {                                                                                 \
        cudaError_t _e = (expr);                                                         \
        if (_e != cudaSuccess) {                                                         \
            ::tvs::set_error("%s failed: %s", #expr, cudaGetErrorString(_e));            \
            return -2;                                                                   \
        }                                                                                \
    } while (0)

int sm_count();

// ---- programmatic dependent launch ----
// A step is ~500 kernels of 5-100 us inside one CUDA graph; the drain -> launch -> prologue gap between two of them is
// ~2 us.  Kernels launched through launch_pdl() may become resident while their predecessor in the stream is still
// running: they do their private prologue (barrier init, TMEM allocation, descriptor prefetch) and then block in
// pdl_wait() until the predecessor has completed and its writes are visible.  Only kernels that call pdl_wait() before
// their first global-memory access may be launched this way.
bool pdl_enabled();
int pdl_max_ctas();
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, unsigned cluster,
                                     Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    unsigned n = 0;
    if (cluster > 1) {
        attr[n].id = cudaLaunchAttributeClusterDimension;
        attr[n].val.clusterDim.x = cluster;
        attr[n].val.clusterDim.y = 1;
        attr[n].val.clusterDim.z = 1;
        ++n;
    }
    if (pdl_enabled() && (pdl_max_ctas() == 0 || static_cast<long long>(grid.x) * grid.y * grid.z <= pdl_max_ctas())) {
        attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[n].val.programmaticStreamSerializationAllowed = 1;
        ++n;
    }
    cfg.attrs = attr;
    cfg.numAttrs = n;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// ---- small device helpers ----
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&t);
}
// IEEE fp16 pairs: the forward operands of the vision tower (11-bit significand instead of bf16's 8; the values are
// LayerNorm outputs, attention outputs, GELU outputs and frozen weights, all far inside fp16's range)
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
    __half2 t = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float2 unpack_f16x2(uint32_t v) {
    __half2 t = *reinterpret_cast<__half2*>(&v);
    return __half22float2(t);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t v) {
    __nv_bfloat162 t = *reinterpret_cast<__nv_bfloat162*>(&v);
    return __bfloat1622float2(t);
}
// fp32 -> nearest tf32 (10-bit mantissa, ties away).  The kind::tf32 MMA TRUNCATES its fp32 operands (a systematic
// -3.4e-4 relative bias per GEMM), so every producer of a tf32 GEMM operand rounds with this on the way out.
__device__ __forceinline__ float round_tf32_rn(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
__device__ __forceinline__ float sigmoidf_fast(float x) { return 1.0f / (1.0f + __expf(-x)); }
// sigmoid through ONE special-function op: sigma(z) = 0.5 * tanh(z / 2) + 0.5 (MUFU.TANH, ~2^-11 relative error).
// An IEEE division in a GEMM epilogue costs a slow-path subroutine per element (measured: 3x the whole GEMM time).
__device__ __forceinline__ float tanh_approx(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float sigmoid_tanh(float x) { return fmaf(0.5f, tanh_approx(0.5f * x), 0.5f); }
// quick_gelu(x) = x * sigmoid(1.702 x)   (transformers ACT2FN["quick_gelu"])
__device__ __forceinline__ float quick_gelu(float x) { return x * sigmoid_tanh(1.702f * x); }
__device__ __forceinline__ float quick_gelu_grad(float x) {
    float s = sigmoid_tanh(1.702f * x);
    return s * fmaf(1.702f * x, 1.0f - s, 1.0f);
}

}  // namespace tvs
