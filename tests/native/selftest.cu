// Native self-test of libtvs_b200.so through its C ABI (no torch): GEMM (tcgen05) against a naive CUDA-core GEMM,
// attention forward/backward against a double-precision CPU reference, LayerNorm against the CPU, the fused
// loss/metric kernel against oracle/loss_metrics.c (bit-exact counters).  Run on the B200 box:
//     tests/native/selftest [gemm|attn|ln|loss|all]
// Exit code 0 = all selected cases pass.  A watchdog (alarm) aborts a hung kernel after 180 s.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <signal.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include <algorithm>
#include <random>
#include <string>
#include <vector>

#include "tvs_b200.h"

extern "C" void oracle_dicebce_metrics(const float* logits, const float* mask, long long B, long long N, float thr, double* parts,
                                       int64_t* counts, int64_t* conf);

#define CK(x)                                                                              \
    do {                                                                                   \
        cudaError_t e_ = (x);                                                              \
        if (e_ != cudaSuccess) {                                                           \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            exit(2);                                                                       \
        }                                                                                  \
    } while (0)
#define TV(x)                                                                 \
    do {                                                                      \
        int r_ = (x);                                                         \
        if (r_ != 0) {                                                        \
            printf("tvs error %d: %s  (%s:%d)\n", r_, tvs_last_error(), __FILE__, __LINE__); \
            exit(3);                                                          \
        }                                                                     \
    } while (0)

static int g_fail = 0;
static std::mt19937 rng(1234);
static float frand(float s = 1.f) { return std::normal_distribution<float>(0.f, s)(rng); }

static std::vector<__nv_bfloat16> to_bf16(const std::vector<float>& v) {
    std::vector<__nv_bfloat16> o(v.size());
    for (size_t i = 0; i < v.size(); ++i) o[i] = __float2bfloat16(v[i]);
    return o;
}
static float bf(const __nv_bfloat16& x) { return __bfloat162float(x); }
template <class T>
static T* dev(const std::vector<T>& h) {
    T* p;
    CK(cudaMalloc(&p, h.size() * sizeof(T) + 16));
    CK(cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
    return p;
}
template <class T>
static T* dev_zero(size_t n) {
    T* p;
    CK(cudaMalloc(&p, n * sizeof(T) + 16));
    CK(cudaMemset(p, 0, n * sizeof(T)));
    return p;
}
template <class T>
static std::vector<T> host(const T* d, size_t n) {
    std::vector<T> h(n);
    CK(cudaMemcpy(h.data(), d, n * sizeof(T), cudaMemcpyDeviceToHost));
    return h;
}
static void report(const char* name, double err, double tol) {
    bool ok = err <= tol && err == err;
    printf("%-64s max_err %.3e (tol %.1e) %s\n", name, err, tol, ok ? "PASS" : "FAIL");
    if (!ok) g_fail++;
}

// ------------------------------------------------------------------------------------------------
// GEMM
// ------------------------------------------------------------------------------------------------
__device__ float qgelu_d(float x) { return x / (1.f + expf(-1.702f * x)); }
__device__ float qgelu_grad_d(float x) {
    float s = 1.f / (1.f + expf(-1.702f * x));
    return s * (1.f + 1.702f * x * (1.f - s));
}
__global__ void ref_gemm_kernel(const __nv_bfloat16* A, const __nv_bfloat16* W, int M, int N, int K, const float* bias, const float* residual,
                                const __nv_bfloat16* aux, int act, float* out, float* pre) {
    int n = blockIdx.x * blockDim.x + threadIdx.x, m = blockIdx.y;
    if (n >= N || m >= M) return;
    float acc = 0.f;
    for (int k = 0; k < K; ++k) acc += __bfloat162float(A[(size_t)m * K + k]) * __bfloat162float(W[(size_t)n * K + k]);
    if (bias) acc += bias[n];
    if (pre) pre[(size_t)m * N + n] = acc;
    if (act == TVS_ACT_QGELU) acc = qgelu_d(acc);
    else if (act == TVS_ACT_RELU) acc = fmaxf(acc, 0.f);
    else if (act == TVS_ACT_DQGELU) acc *= qgelu_grad_d(__bfloat162float(aux[(size_t)m * N + n]));
    else if (act == TVS_ACT_DRELU) acc = __bfloat162float(aux[(size_t)m * N + n]) > 0.f ? acc : 0.f;
    if (residual) acc += residual[(size_t)m * N + n];
    out[(size_t)m * N + n] = acc;
}

static void gemm_case(int M, int N, int K, int act, bool use_bias, bool use_res, bool out32, bool out16, bool pre, int tile_n, bool timeit = false, bool f16out = false) {
    std::vector<float> a((size_t)M * K), w((size_t)N * K), bias(N), res((size_t)M * N), aux((size_t)M * N);
    for (auto& x : a) x = frand();
    for (auto& x : w) x = frand(1.f / sqrtf((float)K));
    for (auto& x : bias) x = frand(0.5f);
    for (auto& x : res) x = frand();
    for (auto& x : aux) x = frand();
    auto a16 = to_bf16(a), w16 = to_bf16(w), aux16 = to_bf16(aux);
    __nv_bfloat16 *dA = dev(a16), *dW = dev(w16), *dAux = dev(aux16);
    float *dBias = dev(bias), *dRes = dev(res);
    float* dOut32 = dev_zero<float>((size_t)M * N);
    __nv_bfloat16* dOut16 = dev_zero<__nv_bfloat16>((size_t)M * N);
    __nv_bfloat16* dPre = dev_zero<__nv_bfloat16>((size_t)M * N);
    float* dRef = dev_zero<float>((size_t)M * N);
    float* dRefPre = dev_zero<float>((size_t)M * N);
    bool deriv = act == TVS_ACT_DQGELU || act == TVS_ACT_DRELU;

    tvs_gemm_args g;
    memset(&g, 0, sizeof(g));
    g.A = dA; g.lda = K; g.W = dW; g.ldw = K; g.M = M; g.N = N; g.K = K;
    g.bias = use_bias ? dBias : nullptr;
    g.residual = use_res ? dRes : nullptr; g.ldr = N;
    g.out_f32 = out32 ? dOut32 : nullptr; g.ldo32 = N;
    g.out_bf16 = out16 ? dOut16 : nullptr; g.ldo16 = N;
    g.pre_bf16 = pre ? dPre : nullptr; g.ldpre = N;
    g.aux_bf16 = deriv ? dAux : nullptr; g.ldaux = N;
    g.act = act; g.tile_n = tile_n;
    if (f16out) g.reserved |= TVS_GEMM_OUT16_F16;        // the 16-bit activation output as IEEE fp16 (fc1 of the vision tower)
    TV(tvs_gemm_bf16(&g, nullptr));
    CK(cudaDeviceSynchronize());
    dim3 grid((N + 127) / 128, M);
    ref_gemm_kernel<<<grid, 128>>>(dA, dW, M, N, K, g.bias, g.residual, deriv ? dAux : nullptr, act, dRef, dRefPre);
    CK(cudaDeviceSynchronize());
    auto ref = host(dRef, (size_t)M * N), refpre = host(dRefPre, (size_t)M * N);
    char name[160];
    double scale = 0;
    for (auto v : ref) scale = std::max(scale, (double)fabsf(v));
    if (out32) {
        auto o = host(dOut32, (size_t)M * N);
        double err = 0;
        for (size_t i = 0; i < o.size(); ++i) err = std::max(err, (double)fabsf(o[i] - ref[i]));
        snprintf(name, sizeof name, "gemm M=%d N=%d K=%d act=%d b=%d r=%d bn=%d f32", M, N, K, act, use_bias, use_res, tile_n);
        report(name, err / std::max(1.0, scale), 2e-3);
    }
    if (out16) {
        auto o = host(dOut16, (size_t)M * N);
        double err = 0;
        for (size_t i = 0; i < o.size(); ++i)
            err = std::max(err, (double)fabsf((f16out ? __half2float(*reinterpret_cast<const __half*>(&o[i])) : bf(o[i])) - ref[i]));
        snprintf(name, sizeof name, "gemm M=%d N=%d K=%d act=%d b=%d r=%d bn=%d %s", M, N, K, act, use_bias, use_res, tile_n, f16out ? "fp16" : "bf16");
        report(name, err / std::max(1.0, scale), f16out ? 2e-3 : 1e-2);
    }
    if (pre) {
        auto o = host(dPre, (size_t)M * N);
        double err = 0, sc = 0;
        for (size_t i = 0; i < o.size(); ++i) { err = std::max(err, (double)fabsf(bf(o[i]) - refpre[i])); sc = std::max(sc, (double)fabsf(refpre[i])); }
        snprintf(name, sizeof name, "gemm M=%d N=%d K=%d pre-activation bf16", M, N, K);
        report(name, err / std::max(1.0, sc), 1e-2);
    }
    if (timeit) {
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        for (int i = 0; i < 3; ++i) TV(tvs_gemm_bf16(&g, nullptr));
        CK(cudaEventRecord(e0));
        const int it = 20;
        for (int i = 0; i < it; ++i) TV(tvs_gemm_bf16(&g, nullptr));
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("    timing: %.1f us/launch, %.1f TFLOP/s (bn=%d)\n", ms / it * 1e3, 2.0 * M * N * K / (ms / it * 1e-3) / 1e12, tile_n);
    }
    cudaFree(dA); cudaFree(dW); cudaFree(dAux); cudaFree(dBias); cudaFree(dRes); cudaFree(dOut32); cudaFree(dOut16); cudaFree(dPre); cudaFree(dRef); cudaFree(dRefPre);
}

// fp32 operands through the kind::tf32 MMA, against a double-precision CPU reference of the fp32 inputs
static void gemm_tf32_case(int M, int N, int K, int tile_n) {
    std::vector<float> a((size_t)M * K), w((size_t)N * K), bias(N);
    for (auto& x : a) x = frand();
    for (auto& x : w) x = frand(1.f / sqrtf((float)K));
    for (auto& x : bias) x = frand(0.5f);
    float *dA = dev(a), *dW = dev(w), *dBias = dev(bias);
    float* dOut = dev_zero<float>((size_t)M * N);
    tvs_gemm_args g;
    memset(&g, 0, sizeof(g));
    g.A = dA; g.lda = K; g.W = dW; g.ldw = K; g.M = M; g.N = N; g.K = K;
    g.bias = dBias; g.out_f32 = dOut; g.ldo32 = N; g.tile_n = tile_n; g.ab_dtype = TVS_AB_TF32;
    TV(tvs_gemm_bf16(&g, nullptr));
    CK(cudaDeviceSynchronize());
    auto o = host(dOut, (size_t)M * N);
    double err = 0, scale = 0;
    for (int m = 0; m < M; m += 7)
        for (int n = 0; n < N; ++n) {
            double acc = bias[n];
            for (int k = 0; k < K; ++k) acc += (double)a[(size_t)m * K + k] * w[(size_t)n * K + k];
            err = std::max(err, fabs(acc - o[(size_t)m * N + n]));
            scale = std::max(scale, fabs(acc));
        }
    char name[128];
    snprintf(name, sizeof name, "gemm tf32 M=%d N=%d K=%d bn=%d (rel to max)", M, N, K, tile_n);
    report(name, err / std::max(1.0, scale), 1.5e-3);
    cudaFree(dA); cudaFree(dW); cudaFree(dBias); cudaFree(dOut);
}

// the shapes of the B = 32 training step with the automatic tile choice (dynamic tile scheduler on every multi-wave one)
static void test_gemm_step_shapes() {
#define STEP_CASE(...) do { printf("start %s\n", #__VA_ARGS__); __VA_ARGS__; } while (0)
    STEP_CASE(gemm_case(15488, 768, 768, TVS_ACT_NONE, false, false, true, false, false, 0));
    STEP_CASE(gemm_case(15648, 2304, 768, TVS_ACT_NONE, true, false, false, true, false, 0));
    STEP_CASE(gemm_case(15648, 768, 768, TVS_ACT_NONE, true, true, true, false, false, 0));
    STEP_CASE(gemm_case(15648, 3072, 768, TVS_ACT_QGELU, true, false, false, true, true, 0));
    STEP_CASE(gemm_case(15648, 3072, 768, TVS_ACT_QGELU, true, false, false, true, true, 0, false, true));      // fp16 activation out (EPI_FC1)
    STEP_CASE(gemm_case(15648, 768, 3072, TVS_ACT_NONE, true, true, true, false, false, 0));
    STEP_CASE(gemm_case(15648, 3072, 768, TVS_ACT_DQGELU, false, false, false, true, false, 0));
    STEP_CASE(gemm_case(15648, 768, 3072, TVS_ACT_NONE, false, false, false, true, false, 0));
    STEP_CASE(gemm_case(15648, 768, 768, TVS_ACT_NONE, false, false, false, true, false, 0));
    STEP_CASE(gemm_case(15648, 768, 2304, TVS_ACT_NONE, false, false, false, true, false, 0));
    STEP_CASE(gemm_case(15648, 768, 64, TVS_ACT_NONE, false, false, false, true, false, 0));
    STEP_CASE(gemm_case(15648, 64, 192, TVS_ACT_NONE, false, true, true, false, false, 0));
    STEP_CASE(gemm_tf32_case(15648, 64, 768, 0));
    STEP_CASE(gemm_tf32_case(15648, 192, 64, 0));
    STEP_CASE(gemm_tf32_case(15648, 64, 64, 0));
    STEP_CASE(gemm_tf32_case(15488, 256, 64, 0));
    STEP_CASE(gemm_tf32_case(15488, 25, 64, 0));
    STEP_CASE(gemm_tf32_case(384, 2048, 512, 0));
#undef STEP_CASE
}
static void test_gemm() {
    gemm_tf32_case(128, 64, 32, 64);
    gemm_tf32_case(300, 192, 64, 0);
    gemm_tf32_case(978, 64, 768, 0);
    gemm_tf32_case(978, 2048, 64, 0);
    gemm_tf32_case(978, 64, 2048, 0);
    gemm_tf32_case(968, 25, 64, 0);
    gemm_tf32_case(77 * 4, 1536, 512, 0);
    gemm_tf32_case(3, 512, 512, 0);
    // tiny first: one tile, one k-block
    gemm_case(128, 128, 64, TVS_ACT_NONE, false, false, true, false, false, 128);
    gemm_case(128, 64, 64, TVS_ACT_NONE, false, false, true, false, false, 64);
    gemm_case(128, 256, 64, TVS_ACT_NONE, false, false, true, false, false, 256);
    gemm_case(128, 128, 256, TVS_ACT_NONE, true, false, true, true, false, 128);
    // ragged M, several tiles / k-blocks, ring wrap-around, persistent loop
    gemm_case(489 * 2, 768, 768, TVS_ACT_NONE, true, true, true, true, false, 128);
    gemm_case(489 * 2, 768, 768, TVS_ACT_NONE, true, true, true, true, false, 256);
    gemm_case(1000, 2304, 768, TVS_ACT_NONE, true, false, false, true, false, 0);
    gemm_case(1000, 3072, 768, TVS_ACT_QGELU, true, false, false, true, true, 0);
    gemm_case(1000, 768, 3072, TVS_ACT_NONE, true, true, true, false, false, 0);
    gemm_case(1000, 3072, 768, TVS_ACT_DQGELU, false, false, false, true, false, 0);
    gemm_case(700, 2048, 64, TVS_ACT_RELU, true, false, false, true, false, 0);
    gemm_case(700, 64, 2048, TVS_ACT_DRELU, false, false, true, true, false, 0);
    gemm_case(700, 64, 768, TVS_ACT_NONE, true, true, true, true, false, 0);
    gemm_case(700, 192, 64, TVS_ACT_NONE, true, false, false, true, false, 0);
    gemm_case(968, 25, 64, TVS_ACT_NONE, false, false, true, false, false, 0);    // ragged N (additive 5x5 map)
    gemm_case(968, 256, 64, TVS_ACT_NONE, false, false, true, false, false, 0);
    gemm_case(77 * 4, 1536, 512, TVS_ACT_NONE, true, false, false, true, false, 0);
    gemm_case(300, 512, 40, TVS_ACT_NONE, true, false, true, false, false, 0);     // K not a multiple of 64
    // the north-star shapes (B=32, S=489), with timing
    gemm_case(15648, 768, 768, TVS_ACT_NONE, true, true, true, false, false, 128, true);
    gemm_case(15648, 768, 768, TVS_ACT_NONE, true, true, true, false, false, 256, true);
    gemm_case(15648, 2304, 768, TVS_ACT_NONE, true, false, false, true, false, 128, true);
    gemm_case(15648, 2304, 768, TVS_ACT_NONE, true, false, false, true, false, 256, true);
    gemm_case(15648, 3072, 768, TVS_ACT_QGELU, true, false, false, true, true, 128, true);
    gemm_case(15648, 3072, 768, TVS_ACT_QGELU, true, false, false, true, true, 256, true);
    gemm_case(15648, 768, 3072, TVS_ACT_NONE, true, true, true, false, false, 128, true);
    gemm_case(15648, 768, 3072, TVS_ACT_NONE, true, true, true, false, false, 256, true);
}

// ------------------------------------------------------------------------------------------------
// attention
// ------------------------------------------------------------------------------------------------
static void attn_case(int B, int S, int H, int hd, int causal, bool use_mask, bool timeit = false) {
    const int E = H * hd;
    std::vector<float> qkv((size_t)B * S * 3 * E), dout((size_t)B * S * E);
    for (auto& x : qkv) x = frand(1.0f);
    for (size_t i = 0; i < qkv.size(); ++i)
        if ((i % (3 * E)) < (size_t)E) qkv[i] *= 0.35f;  // pre-scaled q
    for (auto& x : dout) x = frand(1.0f);
    std::vector<uint8_t> mask((size_t)B * S, 1);
    if (use_mask)
        for (int b = 0; b < B; ++b)
            for (int s = 0; s < S; ++s) mask[(size_t)b * S + s] = s < std::max(2, S - 3 - 5 * b);
    auto qkv16 = to_bf16(qkv), dout16 = to_bf16(dout);
    __nv_bfloat16 *dQKV = dev(qkv16), *dDO = dev(dout16);
    uint8_t* dMask = dev(mask);
    __nv_bfloat16* dOut = dev_zero<__nv_bfloat16>((size_t)B * S * E);
    __nv_bfloat16* dDQKV = dev_zero<__nv_bfloat16>((size_t)B * S * 3 * E);
    float* dLse = dev_zero<float>((size_t)B * H * S);
    float* dDelta = dev_zero<float>((size_t)B * H * S);
    TV(tvs_attn_fwd(dQKV, B, S, H, hd, causal, use_mask ? dMask : nullptr, dOut, nullptr, dLse, 0, nullptr));
    CK(cudaDeviceSynchronize());
    TV(tvs_attn_bwd(dQKV, dOut, dDO, dLse, B, S, H, hd, causal, use_mask ? dMask : nullptr, dDelta, dDQKV, 0, nullptr));
    CK(cudaDeviceSynchronize());
    auto out = host(dOut, (size_t)B * S * E);
    auto dqkv = host(dDQKV, (size_t)B * S * 3 * E);
    auto lse = host(dLse, (size_t)B * H * S);

    // CPU reference in double from the bf16-rounded inputs
    double err_o = 0, err_l = 0, err_dq = 0, err_dk = 0, err_dv = 0, sc_dq = 1e-9, sc_dk = 1e-9, sc_dv = 1e-9, sc_o = 0;
    std::vector<double> P((size_t)S * S), dP((size_t)S * S);
    for (int b = 0; b < B; ++b)
        for (int h = 0; h < H; ++h) {
            auto q = [&](int s, int d) { return (double)bf(qkv16[((size_t)b * S + s) * 3 * E + h * hd + d]); };
            auto k = [&](int s, int d) { return (double)bf(qkv16[((size_t)b * S + s) * 3 * E + E + h * hd + d]); };
            auto v = [&](int s, int d) { return (double)bf(qkv16[((size_t)b * S + s) * 3 * E + 2 * E + h * hd + d]); };
            auto go = [&](int s, int d) { return (double)bf(dout16[((size_t)b * S + s) * E + h * hd + d]); };
            std::vector<double> O((size_t)S * hd, 0.0);
            for (int i = 0; i < S; ++i) {
                double mx = -1e300;
                for (int j = 0; j < S; ++j) {
                    bool ok = (!causal || j <= i) && mask[(size_t)b * S + j];
                    double s = 0;
                    if (ok) for (int d = 0; d < hd; ++d) s += q(i, d) * k(j, d);
                    P[(size_t)i * S + j] = ok ? s : -1e300;
                    if (ok) mx = std::max(mx, s);
                }
                double sum = 0;
                for (int j = 0; j < S; ++j) {
                    double p = P[(size_t)i * S + j] <= -1e299 ? 0.0 : exp(P[(size_t)i * S + j] - mx);
                    P[(size_t)i * S + j] = p;
                    sum += p;
                }
                for (int j = 0; j < S; ++j) P[(size_t)i * S + j] /= sum;
                err_l = std::max(err_l, fabs((mx + log(sum)) - (double)lse[((size_t)b * H + h) * S + i]));
                for (int d = 0; d < hd; ++d) {
                    double o = 0;
                    for (int j = 0; j < S; ++j) o += P[(size_t)i * S + j] * v(j, d);
                    O[(size_t)i * hd + d] = o;
                    err_o = std::max(err_o, fabs(o - (double)bf(out[((size_t)b * S + i) * E + h * hd + d])));
                    sc_o = std::max(sc_o, fabs(o));
                }
            }
            // backward
            std::vector<double> dQ((size_t)S * hd, 0.0), dK((size_t)S * hd, 0.0), dV((size_t)S * hd, 0.0);
            for (int i = 0; i < S; ++i) {
                double delta = 0;
                for (int d = 0; d < hd; ++d) delta += go(i, d) * O[(size_t)i * hd + d];
                for (int j = 0; j < S; ++j) {
                    double p = P[(size_t)i * S + j];
                    if (p == 0.0) continue;
                    double dp = 0;
                    for (int d = 0; d < hd; ++d) dp += go(i, d) * v(j, d);
                    double ds = p * (dp - delta);
                    for (int d = 0; d < hd; ++d) {
                        dQ[(size_t)i * hd + d] += ds * k(j, d);
                        dK[(size_t)j * hd + d] += ds * q(i, d);
                        dV[(size_t)j * hd + d] += p * go(i, d);
                    }
                }
            }
            for (int s = 0; s < S; ++s)
                for (int d = 0; d < hd; ++d) {
                    size_t base = ((size_t)b * S + s) * 3 * E + h * hd + d;
                    err_dq = std::max(err_dq, fabs(dQ[(size_t)s * hd + d] - (double)bf(dqkv[base])));
                    err_dk = std::max(err_dk, fabs(dK[(size_t)s * hd + d] - (double)bf(dqkv[base + E])));
                    err_dv = std::max(err_dv, fabs(dV[(size_t)s * hd + d] - (double)bf(dqkv[base + 2 * E])));
                    sc_dq = std::max(sc_dq, fabs(dQ[(size_t)s * hd + d]));
                    sc_dk = std::max(sc_dk, fabs(dK[(size_t)s * hd + d]));
                    sc_dv = std::max(sc_dv, fabs(dV[(size_t)s * hd + d]));
                }
        }
    char name[160];
    snprintf(name, sizeof name, "attn B=%d S=%d H=%d hd=%d causal=%d mask=%d  out", B, S, H, hd, causal, use_mask);
    // bf16 output: half an ulp is 2^-9 of the value; probabilities are bf16 too.  2e-2 absolute covers |out| <= ~3, beyond that
    // the bound scales with the largest reference output (the inputs are random: their range moves with the random stream)
    report(name, err_o, std::max(2e-2, 6e-3 * sc_o));
    snprintf(name, sizeof name, "attn B=%d S=%d H=%d hd=%d causal=%d mask=%d  lse", B, S, H, hd, causal, use_mask);
    report(name, err_l, 1e-3);
    snprintf(name, sizeof name, "attn B=%d S=%d H=%d hd=%d causal=%d mask=%d  dq (rel)", B, S, H, hd, causal, use_mask);
    report(name, err_dq / sc_dq, 2e-2);
    snprintf(name, sizeof name, "attn B=%d S=%d H=%d hd=%d causal=%d mask=%d  dk (rel)", B, S, H, hd, causal, use_mask);
    report(name, err_dk / sc_dk, 2e-2);
    snprintf(name, sizeof name, "attn B=%d S=%d H=%d hd=%d causal=%d mask=%d  dv (rel)", B, S, H, hd, causal, use_mask);
    report(name, err_dv / sc_dv, 2e-2);
    cudaFree(dQKV); cudaFree(dDO); cudaFree(dMask); cudaFree(dOut); cudaFree(dDQKV); cudaFree(dLse); cudaFree(dDelta);
    (void)timeit;
}

static void attn_timing(int B, int S, int H, int hd) {
    const int E = H * hd;
    __nv_bfloat16* dQKV = dev_zero<__nv_bfloat16>((size_t)B * S * 3 * E);
    __nv_bfloat16* dDO = dev_zero<__nv_bfloat16>((size_t)B * S * E);
    __nv_bfloat16* dOut = dev_zero<__nv_bfloat16>((size_t)B * S * E);
    __nv_bfloat16* dDQKV = dev_zero<__nv_bfloat16>((size_t)B * S * 3 * E);
    float* dLse = dev_zero<float>((size_t)B * H * S);
    float* dDelta = dev_zero<float>((size_t)B * H * S);
    cudaEvent_t e0, e1, e2;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); CK(cudaEventCreate(&e2));
    const int it = 10;
    for (int w = 0; w < 2; ++w) {
        CK(cudaEventRecord(e0));
        for (int i = 0; i < it; ++i) TV(tvs_attn_fwd(dQKV, B, S, H, hd, 0, nullptr, dOut, nullptr, dLse, 0, nullptr));
        CK(cudaEventRecord(e1));
        for (int i = 0; i < it; ++i) TV(tvs_attn_bwd(dQKV, dOut, dDO, dLse, B, S, H, hd, 0, nullptr, dDelta, dDQKV, 0, nullptr));
        CK(cudaEventRecord(e2));
        CK(cudaEventSynchronize(e2));
    }
    float f, b;
    CK(cudaEventElapsedTime(&f, e0, e1));
    CK(cudaEventElapsedTime(&b, e1, e2));
    double fl = 4.0 * B * H * (double)S * S * hd;
    printf("    attn timing B=%d S=%d H=%d hd=%d: fwd %.1f us (%.1f TFLOP/s), bwd %.1f us (%.1f TFLOP/s algorithmic 2.5x)\n", B, S, H, hd, f / it * 1e3,
           fl / (f / it * 1e-3) / 1e12, b / it * 1e3, 2.5 * fl / (b / it * 1e-3) / 1e12);
    cudaFree(dQKV); cudaFree(dDO); cudaFree(dOut); cudaFree(dDQKV); cudaFree(dLse); cudaFree(dDelta);
}

static void test_attn(bool timing = true) {
    attn_case(1, 128, 1, 64, 0, false);
    attn_case(1, 64, 1, 64, 0, false);
    attn_case(2, 493, 12, 64, 0, false);
    attn_case(2, 100, 2, 64, 0, false);
    attn_case(2, 489, 2, 64, 0, false);
    attn_case(3, 77, 2, 64, 1, true);
    attn_case(2, 12, 8, 64, 1, true);
    attn_case(2, 130, 4, 16, 0, false);
    attn_case(1, 493, 4, 16, 0, false);
    // (appended after the older cases so that those keep their random streams)
    attn_case(10, 300, 12, 64, 0, false);      // 360 (tile, head, sample) items > 2 x 148 resident CTAs: the persistent kernels' multi-item path
    attn_case(40, 100, 8, 64, 0, false);       // 320 items of TWO inner steps each (<= ring depth): the next item's outer tiles are requested after the loop
    if (!timing) return;
    attn_timing(32, 489, 12, 64);
    attn_timing(32, 489, 4, 16);
}

// ------------------------------------------------------------------------------------------------
// fused decoder FFN (D = 64)
// ------------------------------------------------------------------------------------------------
static void ffn_case(int M, int F, bool split, bool timeit = false) {
    const int D = 64;
    std::vector<float> x((size_t)M * D), g((size_t)M * D), w1((size_t)F * D), w2t((size_t)F * D), b1(F), b2(D);
    for (auto& v : x) v = frand();
    for (auto& v : g) v = frand();
    for (auto& v : w1) v = frand(0.125f);
    for (auto& v : w2t) v = frand(0.03f);
    for (auto& v : b1) v = frand(0.3f);
    for (auto& v : b2) v = frand(0.3f);
    auto hi1 = to_bf16(w1), hi2 = to_bf16(w2t);
    std::vector<float> r1(w1.size()), r2(w2t.size());
    for (size_t i = 0; i < w1.size(); ++i) { r1[i] = w1[i] - bf(hi1[i]); r2[i] = w2t[i] - bf(hi2[i]); }
    auto lo1 = to_bf16(r1), lo2 = to_bf16(r2);
    float *dx = dev(x), *dg = dev(g), *db1 = dev(b1), *db2 = dev(b2);
    __nv_bfloat16 *dh1 = dev(hi1), *dh2 = dev(hi2), *dl1 = dev(lo1), *dl2 = dev(lo2);
    float *dout = dev_zero<float>((size_t)M * D), *ddx = dev_zero<float>((size_t)M * D);
    TV(tvs_ffn64_fwd(dx, dh1, split ? dl1 : nullptr, dh2, split ? dl2 : nullptr, db1, db2, M, D, F, dout, nullptr));
    TV(tvs_ffn64_bwd(dx, dg, dh1, split ? dl1 : nullptr, dh2, split ? dl2 : nullptr, db1, M, D, F, ddx, nullptr));
    CK(cudaDeviceSynchronize());
    auto out = host(dout, (size_t)M * D);
    auto gx = host(ddx, (size_t)M * D);
    double e_f = 0, e_b = 0, n_f = 0, n_b = 0;
    const int step = M > 2000 ? 97 : 1;       // sample rows of the big case
    std::vector<double> h(F), dh(F);
    for (int m = 0; m < M; m += step) {
        for (int f = 0; f < F; ++f) {
            double s = b1[f], d = 0;
            for (int k = 0; k < D; ++k) { s += (double)x[(size_t)m * D + k] * w1[(size_t)f * D + k]; d += (double)g[(size_t)m * D + k] * w2t[(size_t)f * D + k]; }
            h[f] = s > 0 ? s : 0;
            dh[f] = fabs(s) < (split ? 1e-4 : 6e-2) ? NAN : (s > 0 ? d : 0);       // mask decided within rounding distance of 0: skip the row
        }
        bool skip = false;
        for (int f = 0; f < F; ++f) if (std::isnan(dh[f])) skip = true;
        for (int k = 0; k < D; ++k) {
            double o = b2[k] + x[(size_t)m * D + k], dd = g[(size_t)m * D + k];
            for (int f = 0; f < F; ++f) { o += h[f] * w2t[(size_t)f * D + k]; if (!skip) dd += dh[f] * w1[(size_t)f * D + k]; }
            e_f = std::max(e_f, fabs(o - out[(size_t)m * D + k])); n_f = std::max(n_f, fabs(o));
            if (!skip) { e_b = std::max(e_b, fabs(dd - gx[(size_t)m * D + k])); n_b = std::max(n_b, fabs(dd)); }
        }
    }
    char name[128];
    snprintf(name, sizeof name, "ffn64 M=%d F=%d %s fwd (max |ref| %.2f)", M, F, split ? "split" : "bf16", n_f); report(name, e_f, split ? 2e-4 : 5e-2);
    snprintf(name, sizeof name, "ffn64 M=%d F=%d %s bwd (max |ref| %.2f)", M, F, split ? "split" : "bf16", n_b); report(name, e_b, split ? 2e-4 : 5e-2);
    if (timeit) {
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        float ms;
        const int reps = 20;
        CK(cudaEventRecord(e0));
        for (int it = 0; it < reps; ++it) TV(tvs_ffn64_fwd(dx, dh1, split ? dl1 : nullptr, dh2, split ? dl2 : nullptr, db1, db2, M, D, F, dout, nullptr));
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("    timing fwd: %.1f us/launch\n", ms * 1e3 / reps);
        CK(cudaEventRecord(e0));
        for (int it = 0; it < reps; ++it) TV(tvs_ffn64_bwd(dx, dg, dh1, split ? dl1 : nullptr, dh2, split ? dl2 : nullptr, db1, M, D, F, ddx, nullptr));
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("    timing bwd: %.1f us/launch\n", ms * 1e3 / reps);
    }
    cudaFree(dx); cudaFree(dg); cudaFree(db1); cudaFree(db2); cudaFree(dh1); cudaFree(dh2); cudaFree(dl1); cudaFree(dl2); cudaFree(dout); cudaFree(ddx);
}
static void test_ffn(bool timeit) {
    ffn_case(128, 64, true);
    ffn_case(128, 64, false);
    ffn_case(300, 256, true);
    ffn_case(300, 256, false);
    ffn_case(15648, 2048, true, timeit);
    ffn_case(15648, 2048, false, timeit);
}

// ------------------------------------------------------------------------------------------------
// layernorm
// ------------------------------------------------------------------------------------------------
static void ln_case(int M, int D) {
    std::vector<float> x((size_t)M * D), g(D), b(D), dy((size_t)M * D), add((size_t)M * D);
    for (auto& v : x) v = frand(2.f) + 0.3f;
    for (auto& v : g) v = 1.f + frand(0.2f);
    for (auto& v : b) v = frand(0.2f);
    for (auto& v : dy) v = frand();
    for (auto& v : add) v = frand();
    float *dx = dev(x), *dg = dev(g), *db = dev(b), *ddy = dev(dy), *dadd = dev(add);
    float* dy32 = dev_zero<float>((size_t)M * D);
    __nv_bfloat16* dy16 = dev_zero<__nv_bfloat16>((size_t)M * D);
    float *dmean = dev_zero<float>(M), *drstd = dev_zero<float>(M);
    float* ddx = dev_zero<float>((size_t)M * D);
    __nv_bfloat16* ddx16 = dev_zero<__nv_bfloat16>((size_t)M * D);
    TV(tvs_layernorm_fwd(dx, dg, db, 1e-5f, M, D, dy32, dy16, dmean, drstd, 0, nullptr));
    TV(tvs_layernorm_bwd(nullptr, ddy, dx, dg, dmean, drstd, dadd, M, D, ddx, ddx16, nullptr));
    CK(cudaDeviceSynchronize());
    auto y = host(dy32, (size_t)M * D);
    auto y16 = host(dy16, (size_t)M * D);
    auto gx = host(ddx, (size_t)M * D);
    double e1 = 0, e2 = 0, e3 = 0;
    for (int m = 0; m < M; ++m) {
        double mu = 0, var = 0;
        for (int d = 0; d < D; ++d) mu += x[(size_t)m * D + d];
        mu /= D;
        for (int d = 0; d < D; ++d) var += (x[(size_t)m * D + d] - mu) * (x[(size_t)m * D + d] - mu);
        var /= D;
        double rs = 1.0 / sqrt(var + 1e-5);
        double c1 = 0, c2 = 0;
        for (int d = 0; d < D; ++d) {
            double xh = (x[(size_t)m * D + d] - mu) * rs, gg = dy[(size_t)m * D + d] * g[d];
            c1 += gg; c2 += gg * xh;
            double yy = xh * g[d] + b[d];
            e1 = std::max(e1, fabs(yy - y[(size_t)m * D + d]));
            e2 = std::max(e2, fabs(yy - bf(y16[(size_t)m * D + d])));
        }
        c1 /= D; c2 /= D;
        for (int d = 0; d < D; ++d) {
            double xh = (x[(size_t)m * D + d] - mu) * rs, gg = dy[(size_t)m * D + d] * g[d];
            double r = rs * (gg - c1 - xh * c2) + add[(size_t)m * D + d];
            e3 = std::max(e3, fabs(r - gx[(size_t)m * D + d]));
        }
    }
    char name[128];
    snprintf(name, sizeof name, "layernorm M=%d D=%d fwd f32", M, D); report(name, e1, 1e-4);
    snprintf(name, sizeof name, "layernorm M=%d D=%d fwd bf16", M, D); report(name, e2, 4e-2);
    snprintf(name, sizeof name, "layernorm M=%d D=%d bwd f32", M, D); report(name, e3, 1e-4);
    cudaFree(dx); cudaFree(dg); cudaFree(db); cudaFree(ddy); cudaFree(dadd); cudaFree(dy32); cudaFree(dy16); cudaFree(dmean); cudaFree(drstd); cudaFree(ddx); cudaFree(ddx16);
}
// timing of the vision-tower LayerNorm shape over rotating buffer sets (each set 192 MB; three sets exceed the L2)
static void ln_timing(int M, int D) {
    const int NSET = 3;
    const size_t n = (size_t)M * D;
    float *x[NSET], *add[NSET], *o32[NSET], *mean, *rstd, *g, *b;
    __nv_bfloat16 *dy[NSET], *o16[NSET];
    std::vector<float> hx(n), hg(D, 1.f);
    for (auto& v : hx) v = frand(2.f);
    for (int s = 0; s < NSET; ++s) {
        x[s] = dev(hx); add[s] = dev(hx); o32[s] = dev_zero<float>(n);
        dy[s] = dev(to_bf16(hx)); o16[s] = dev_zero<__nv_bfloat16>(n);
    }
    g = dev(hg); b = dev(hg); mean = dev_zero<float>(M); rstd = dev_zero<float>(M);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float ms;
    const int reps = 30;
    for (int it = 0; it < 3; ++it) TV(tvs_layernorm_fwd(x[it], g, b, 1e-5f, M, D, nullptr, o16[it], mean, rstd, 0, nullptr));
    CK(cudaEventRecord(e0));
    for (int it = 0; it < reps; ++it) TV(tvs_layernorm_fwd(x[it % NSET], g, b, 1e-5f, M, D, nullptr, o16[it % NSET], mean, rstd, 0, nullptr));
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("    layernorm fwd  M=%d D=%d: %.1f us/launch, %.0f GB/s (6 B/elem)\n", M, D, ms * 1e3 / reps, 6.0 * n / (ms * 1e-3 / reps) * 1e-9);
    for (int it = 0; it < 3; ++it) TV(tvs_layernorm_bwd(dy[it], nullptr, x[it], g, mean, rstd, add[it], M, D, o32[it], o16[it], nullptr));
    CK(cudaEventRecord(e0));
    for (int it = 0; it < reps; ++it)
        TV(tvs_layernorm_bwd(dy[it % NSET], nullptr, x[it % NSET], g, mean, rstd, add[it % NSET], M, D, o32[it % NSET], o16[it % NSET], nullptr));
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("    layernorm bwd  M=%d D=%d: %.1f us/launch, %.0f GB/s (16 B/elem)\n", M, D, ms * 1e3 / reps, 16.0 * n / (ms * 1e-3 / reps) * 1e-9);
    for (int s = 0; s < NSET; ++s) { cudaFree(x[s]); cudaFree(add[s]); cudaFree(o32[s]); cudaFree(dy[s]); cudaFree(o16[s]); }
    cudaFree(g); cudaFree(b); cudaFree(mean); cudaFree(rstd);
}
static void test_ln() {
    ln_case(37, 768);
    ln_case(5, 512);
    ln_case(1003, 64);
    ln_case(9, 1024);
    ln_case(2051, 768);       // two rows per warp (M >= 2048), odd row count
    ln_case(15648, 64);       // half-warp-per-row kernels at the decoder shape
    ln_case(9001, 768);       // M >= 8192: the persistent pipelined forward when TVS_LN_FWD=4 selects it
}

// ------------------------------------------------------------------------------------------------
// loss / metrics
// ------------------------------------------------------------------------------------------------
static void loss_case(int B, long long N, bool adversarial) {
    std::vector<float> x((size_t)B * N), y((size_t)B * N);
    for (auto& v : x) v = frand(2.f);
    for (auto& v : y) v = (rng() % 100) < 30 ? 1.f : 0.f;
    if (adversarial) {
        // logits straddling the p == 0.5 decision: exact zeros, denormals, +-k * 2^-25 ... and soft masks
        for (size_t i = 0; i < x.size(); i += 3) {
            int k = (int)(rng() % 64) - 32;
            x[i] = ldexpf((float)k, -26 - (int)(rng() % 4));
        }
        for (size_t i = 1; i < y.size(); i += 7) y[i] = (rng() % 1000) / 1000.f;
        if (B > 1) for (long long i = 0; i < N; ++i) { x[(size_t)N + i] = -5.f - fabsf(x[(size_t)N + i]); y[(size_t)N + i] = 0.f; }  // empty sample
    }
    float *dx = dev(x), *dy = dev(y);
    double* dparts = dev_zero<double>((size_t)B * 4);
    int64_t* dcounts = dev_zero<int64_t>((size_t)B * 3);
    int64_t* dconf = dev_zero<int64_t>(4);
    float* dloss = dev_zero<float>(1);
    void* scratch;
    CK(cudaMalloc(&scratch, tvs_dicebce_scratch_bytes(B, N)));
    TV(tvs_dicebce_metrics_fwd(dx, dy, B, N, 0.5f, 1.0f, 0.2f, dparts, dcounts, dconf, dloss, scratch, nullptr));
    CK(cudaDeviceSynchronize());
    auto parts = host(dparts, (size_t)B * 4);
    auto counts = host(dcounts, (size_t)B * 3);
    auto conf = host(dconf, 4);
    auto loss = host(dloss, 1);
    std::vector<double> rp((size_t)B * 4);
    std::vector<int64_t> rc((size_t)B * 3), rconf(4);
    oracle_dicebce_metrics(x.data(), y.data(), B, N, 0.5f, rp.data(), rc.data(), rconf.data());
    double perr = 0, dice = 0, bce = 0;
    long long cerr = 0;
    for (int b = 0; b < B; ++b) {
        for (int j = 0; j < 4; ++j) perr = std::max(perr, fabs(parts[b * 4 + j] - rp[b * 4 + j]) / std::max(1.0, fabs(rp[b * 4 + j])));
        for (int j = 0; j < 3; ++j) cerr += llabs((long long)(counts[b * 3 + j] - rc[b * 3 + j]));
        dice += 1.0 - (2 * rp[b * 4] + 1e-5) / (rp[b * 4 + 1] + rp[b * 4 + 2] + 1e-5);
        bce += rp[b * 4 + 3];
    }
    for (int j = 0; j < 4; ++j) cerr += llabs((long long)(conf[j] - rconf[j]));
    double rl = dice / B + 0.2 * bce / ((double)B * N);
    char name[128];
    snprintf(name, sizeof name, "dicebce B=%d N=%lld adv=%d partial sums (rel)", B, N, adversarial); report(name, perr, 1e-5);
    snprintf(name, sizeof name, "dicebce B=%d N=%lld adv=%d integer counters (abs diff, bit-exact)", B, N, adversarial); report(name, (double)cerr, 0.0);
    snprintf(name, sizeof name, "dicebce B=%d N=%lld adv=%d loss", B, N, adversarial); report(name, fabs(rl - loss[0]), 1e-5);
    // backward against a double-precision closed form
    float* dg = dev_zero<float>((size_t)B * N);
    std::vector<float> gs(1, 0.7f);
    float* dgs = dev(gs);
    TV(tvs_dicebce_bwd(dx, dy, dparts, dgs, B, N, 1.0f, 0.2f, dg, nullptr));
    CK(cudaDeviceSynchronize());
    auto g = host(dg, (size_t)B * N);
    double gerr = 0, gsc = 1e-30;
    for (int b = 0; b < B; ++b) {
        double I = rp[b * 4], P = rp[b * 4 + 1], G = rp[b * 4 + 2], den = P + G + 1e-5;
        for (long long i = 0; i < N; i += 97) {
            double xv = x[(size_t)b * N + i], yv = y[(size_t)b * N + i], p = 1 / (1 + exp(-xv));
            double d = (-(2 * yv * den - (2 * I + 1e-5)) / (den * den)) / B * p * (1 - p) + 0.2 * (p - yv) / ((double)B * N);
            d *= 0.7;
            gerr = std::max(gerr, fabs(d - g[(size_t)b * N + i]));
            gsc = std::max(gsc, fabs(d));
        }
    }
    snprintf(name, sizeof name, "dicebce B=%d N=%lld adv=%d dlogits (rel)", B, N, adversarial); report(name, gerr / gsc, 1e-3);
    cudaFree(dx); cudaFree(dy); cudaFree(dparts); cudaFree(dcounts); cudaFree(dconf); cudaFree(dloss); cudaFree(scratch); cudaFree(dg); cudaFree(dgs);
}
static void test_loss() {
    loss_case(2, 64 * 64, false);
    loss_case(3, 352 * 352, true);
    loss_case(1, 1001, true);     // unaligned / scalar path
    loss_case(32, 352 * 352, true);
    loss_case(256, 416 * 416, false);
}

static void on_alarm(int) {
    printf("WATCHDOG: self-test exceeded its time budget (hung kernel?)\n");
    fflush(stdout);
    _exit(9);
}

int main(int argc, char** argv) {
    signal(SIGALRM, on_alarm);
    alarm(argc > 2 ? atoi(argv[2]) : 180);
    setvbuf(stdout, nullptr, _IOLBF, 0);
    std::string what = argc > 1 ? argv[1] : "all";
    if (tvs_device_check() != 0) {
        printf("device check failed: %s\n", tvs_last_error());
        return 4;
    }
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    printf("device: %s, %d SMs, abi %d\n", prop.name, prop.multiProcessorCount, tvs_version());
    // every group starts from its own seed: its data do not depend on which groups ran before it ("all" vs a single group)
    if (what == "loss" || what == "all") { rng.seed(1234); test_loss(); }
    if (what == "ln" || what == "all") { rng.seed(2345); test_ln(); }
    if (what == "attn" || what == "all") { rng.seed(3456); test_attn(); }
    if (what == "attncheck") { rng.seed(3456); test_attn(false); }      // correctness cases only (compute-sanitizer runs)
    if (what == "gemm" || what == "all") { rng.seed(4567); test_gemm(); }
    if (what == "ffn" || what == "all") { rng.seed(5678); test_ffn(what == "ffn"); }
    if (what == "gemmstep" || what == "all") { rng.seed(6789); test_gemm_step_shapes(); }
    if (what == "lnprof") { ln_case(9001, 768); ln_case(37, 768); ln_case(5, 512); ln_case(33, 256); ln_case(1003, 64); ln_case(9, 1024); ln_case(7, 2048); ln_timing(15648, 768); }
    if (what == "gemmprof") {   // epilogue cost isolation on the fc1 shape (for timing / ncu)
        const int bn = argc > 3 ? atoi(argv[3]) : 256;
        printf("-- N=3072: bf16 out only\n");
        gemm_case(15648, 3072, 768, TVS_ACT_NONE, true, false, false, true, false, bn, true);
        printf("-- N=3072: bf16 out + pre (two stores, no gelu)\n");
        gemm_case(15648, 3072, 768, TVS_ACT_NONE, true, false, false, true, true, bn, true);
        printf("-- N=3072: gelu, one store\n");
        gemm_case(15648, 3072, 768, TVS_ACT_QGELU, true, false, false, true, false, bn, true);
        printf("-- N=3072: gelu + pre (fc1 forward)\n");
        gemm_case(15648, 3072, 768, TVS_ACT_QGELU, true, false, false, true, true, bn, true);
        printf("-- N=3072: gelu + pre, fp16 activation out (fc1 forward as the engine runs it)\n");
        gemm_case(15648, 3072, 768, TVS_ACT_QGELU, true, false, false, true, true, bn, true, true);
        printf("-- N=3072: dgelu (aux read, one store)\n");
        gemm_case(15648, 3072, 768, TVS_ACT_DQGELU, false, false, false, true, false, bn, true);
        printf("-- N=768 K=768: residual f32 in/out\n");
        gemm_case(15648, 768, 768, TVS_ACT_NONE, true, true, true, false, false, 128, true);
        printf("-- N=768 K=768: bf16 out only\n");
        gemm_case(15648, 768, 768, TVS_ACT_NONE, true, false, false, true, false, 128, true);
        printf("-- N=768 K=3072: residual f32 in/out (fc2 forward)\n");
        gemm_case(15648, 768, 3072, TVS_ACT_NONE, true, true, true, false, false, 128, true);
        printf("-- N=2304 K=768: bias, bf16 out (QKV)\n");
        gemm_case(15648, 2304, 768, TVS_ACT_NONE, true, false, false, true, false, 128, true);
        printf("-- tile width 256 on the N = 768 / 2304 shapes (pairs: 256 x 256 tiles)\n");
        gemm_case(15648, 768, 3072, TVS_ACT_NONE, true, true, true, false, false, 256, true);
        gemm_case(15648, 768, 3072, TVS_ACT_NONE, false, false, false, true, false, 256, true);
        gemm_case(15648, 768, 3072, TVS_ACT_NONE, false, false, false, true, false, 128, true);
        gemm_case(15648, 2304, 768, TVS_ACT_NONE, true, false, false, true, false, 256, true);
        gemm_case(15648, 768, 2304, TVS_ACT_NONE, false, false, false, true, false, 256, true);
        gemm_case(15648, 768, 2304, TVS_ACT_NONE, false, false, false, true, false, 128, true);
        gemm_case(15648, 768, 768, TVS_ACT_NONE, true, true, true, false, false, 256, true);
    }
    printf("launches: %lld\n", (long long)tvs_launch_count());
    printf(g_fail ? "SELFTEST FAILED (%d cases)\n" : "SELFTEST PASSED\n", g_fail);
    return g_fail ? 1 : 0;
}
