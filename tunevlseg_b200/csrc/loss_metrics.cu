// Fused Dice+BCE loss and Dice / IoU integer counters: ONE read of logits and mask (8 B/pixel), HBM-bound.
//
// Replaces, on the reference's hot path (src/models/image_text_mask_module.py:87-119):
//   monai DiceCELoss(sigmoid=True, lambda_dice=1, lambda_ce=0.2)  (configs/model/maple_clipseg.yaml:29-33)
//   torch.sigmoid(logits), mask.long()
//   torchmetrics Dice(threshold, average="samples").update  -> per-sample tp/fp/fn with p >= thr
//   torchmetrics JaccardIndex(task="binary").update         -> global [[tn,fp],[fn,tp]] with p >  thr
// Formulas: oracle/loss_metrics.py.  Integer counters are bit-exact against oracle/loss_metrics.c: the
// thresholded probability is fl32(1 / fl32(1 + fl32(exp(-x)))) with a correctly rounded exp; the kernel takes
// the fast __expf path and re-evaluates in double only when p is within 1e-4 of the threshold.
//
// Two kernels: (1) grid (nblk, B) - every block reduces a slice of one sample to 4 doubles + 7 counters in
// `scratch`; (2) one block - fixed-order reduction over blocks and samples (deterministic), writes
// parts / counts / loss and accumulates the confusion matrix.
#include "common.cuh"
#include "tvs_b200.h"

namespace tvs {

constexpr int LOSS_THREADS = 256;

static int loss_blocks_per_sample(int B, long long N) {
    long long want = (16LL * sm_count() + B - 1) / B;     // ~8 resident blocks per SM: the kernels are latency bound on their two loads per pixel quad
    long long cap = (N + 4 * LOSS_THREADS - 1) / (4 * LOSS_THREADS);
    long long n = want < cap ? want : cap;
    return static_cast<int>(n < 1 ? 1 : n);
}

struct PixelStats {
    float I, P, G, bce;
    unsigned tp, fp, fn;          // p >= thr
    unsigned c00, c01, c10, c11;  // [t][p > thr]
};

template <bool PROBS>
__device__ __forceinline__ void pixel(float x, float y, float thr, PixelStats& s) {
    float p;
    if (PROBS) {
        p = x;   // the caller already holds probabilities (metric(preds, target) API): threshold them as they are
    } else {
        // fast path: one MUFU.EX2, one MUFU.RCP, one MUFU.LG2 per pixel (an IEEE division plus log1pf had made this HBM-sized
        // kernel instruction bound).  t = exp(-|x|) in (0, 1]; sigmoid(x) = 1 / (1 + t) for x >= 0, t / (1 + t) otherwise.
        const float t = __expf(-fabsf(x));
        const float r = __fdividef(1.0f, 1.0f + t);
        p = x >= 0.f ? r : t * r;
        if (fabsf(p - thr) < 1e-4f) {      // the integer counters must match fl32(1 / fl32(1 + fl32(exp(-x)))) exactly
            const float e = static_cast<float>(exp(-static_cast<double>(x)));
            p = __fdiv_rn(1.0f, __fadd_rn(1.0f, e));
        }
        s.bce += fmaxf(x, 0.f) - x * y + __logf(1.0f + t);
    }
    s.I += p * y;
    s.P += p;
    s.G += y;
    const unsigned t = static_cast<unsigned>(static_cast<long long>(y)) & 1u;
    const unsigned ge = p >= thr, gt = p > thr;
    s.tp += ge & t;
    s.fp += ge & (t ^ 1u);
    s.fn += (ge ^ 1u) & t;
    s.c00 += (t ^ 1u) & (gt ^ 1u);
    s.c01 += (t ^ 1u) & gt;
    s.c10 += t & (gt ^ 1u);
    s.c11 += t & gt;
}

template <bool PROBS>
__global__ void __launch_bounds__(LOSS_THREADS)
dicebce_partial_kernel(const float* __restrict__ logits, const float* __restrict__ mask, long long N, float thr, double* __restrict__ part_out,
                       long long* __restrict__ cnt_out) {
    const int b = blockIdx.y, nblk = gridDim.x;
    const float* x = logits + static_cast<long long>(b) * N;
    const float* y = mask + static_cast<long long>(b) * N;
    PixelStats s = {0.f, 0.f, 0.f, 0.f, 0, 0, 0, 0, 0, 0, 0};
    const bool vec = (N % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0) && ((reinterpret_cast<uintptr_t>(y) & 15) == 0);
    if (vec) {
        const long long n4 = N >> 2;
        for (long long i = static_cast<long long>(blockIdx.x) * LOSS_THREADS + threadIdx.x; i < n4; i += static_cast<long long>(nblk) * LOSS_THREADS) {
            const float4 xv = __ldg(reinterpret_cast<const float4*>(x) + i);
            const float4 yv = __ldg(reinterpret_cast<const float4*>(y) + i);
            pixel<PROBS>(xv.x, yv.x, thr, s);
            pixel<PROBS>(xv.y, yv.y, thr, s);
            pixel<PROBS>(xv.z, yv.z, thr, s);
            pixel<PROBS>(xv.w, yv.w, thr, s);
        }
    } else {
        for (long long i = static_cast<long long>(blockIdx.x) * LOSS_THREADS + threadIdx.x; i < N; i += static_cast<long long>(nblk) * LOSS_THREADS)
            pixel<PROBS>(x[i], y[i], thr, s);
    }
    // block reduction: floats in double, counters as 64-bit
    __shared__ double sd[LOSS_THREADS / 32][4];
    __shared__ unsigned long long sc[LOSS_THREADS / 32][7];
    double d[4] = {s.I, s.P, s.G, s.bce};
    unsigned long long c[7] = {s.tp, s.fp, s.fn, s.c00, s.c01, s.c10, s.c11};
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int k = 0; k < 4; ++k) d[k] += __shfl_xor_sync(0xffffffffu, d[k], o);
#pragma unroll
        for (int k = 0; k < 7; ++k) c[k] += __shfl_xor_sync(0xffffffffu, c[k], o);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 4; ++k) sd[warp][k] = d[k];
#pragma unroll
        for (int k = 0; k < 7; ++k) sc[warp][k] = c[k];
    }
    __syncthreads();
    if (threadIdx.x < 11) {
        const long long slot = static_cast<long long>(b) * nblk + blockIdx.x;
        if (threadIdx.x < 4) {
            double t = 0;
            for (int w = 0; w < LOSS_THREADS / 32; ++w) t += sd[w][threadIdx.x];
            part_out[slot * 4 + threadIdx.x] = t;
        } else {
            unsigned long long t = 0;
            for (int w = 0; w < LOSS_THREADS / 32; ++w) t += sc[w][threadIdx.x - 4];
            cnt_out[slot * 7 + (threadIdx.x - 4)] = static_cast<long long>(t);
        }
    }
}

// one warp per sample: lanes stride over the blocks of that sample, then a fixed-order butterfly - deterministic, and the
// nblk * 11 loads of a sample are in flight together (the first version walked them serially in one thread per sample: 33 us)
__global__ void __launch_bounds__(256)
dicebce_finalize_kernel(const double* __restrict__ part_in, const long long* __restrict__ cnt_in, int B, int nblk, long long N,
                        float lambda_dice, float lambda_ce, double* __restrict__ parts, long long* __restrict__ counts,
                        long long* __restrict__ confmat, float* __restrict__ loss) {
    extern __shared__ double sh[];  // [B] dice term, [B] bce sum, then 4*B int64 conf
    double* s_dice = sh;
    double* s_bce = sh + B;
    long long* s_conf = reinterpret_cast<long long*>(sh + 2 * B);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
    for (int b = warp; b < B; b += nwarp) {
        double p[4] = {0, 0, 0, 0};
        long long c[7] = {0, 0, 0, 0, 0, 0, 0};
        for (int k = lane; k < nblk; k += 32) {
            const long long slot = static_cast<long long>(b) * nblk + k;
#pragma unroll
            for (int j = 0; j < 4; ++j) p[j] += part_in[slot * 4 + j];
#pragma unroll
            for (int j = 0; j < 7; ++j) c[j] += cnt_in[slot * 7 + j];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int j = 0; j < 4; ++j) p[j] += __shfl_xor_sync(0xffffffffu, p[j], o);
#pragma unroll
            for (int j = 0; j < 7; ++j) c[j] += __shfl_xor_sync(0xffffffffu, c[j], o);
        }
        if (lane == 0) {
            if (parts)
                for (int j = 0; j < 4; ++j) parts[b * 4 + j] = p[j];
            if (counts)
                for (int j = 0; j < 3; ++j) counts[b * 3 + j] = c[j];
            for (int j = 0; j < 4; ++j) s_conf[b * 4 + j] = c[3 + j];
            s_dice[b] = 1.0 - (2.0 * p[0] + 1e-5) / (p[1] + p[2] + 1e-5);
            s_bce[b] = p[3];
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double dice = 0, bce = 0;
        long long cf[4] = {0, 0, 0, 0};
        for (int b = 0; b < B; ++b) {
            dice += s_dice[b];
            bce += s_bce[b];
            for (int j = 0; j < 4; ++j) cf[j] += s_conf[b * 4 + j];
        }
        if (loss) *loss = static_cast<float>(lambda_dice * dice / B + lambda_ce * bce / (static_cast<double>(B) * static_cast<double>(N)));
        if (confmat)
            for (int j = 0; j < 4; ++j) confmat[j] += cf[j];
    }
}

__global__ void __launch_bounds__(LOSS_THREADS)
dicebce_bwd_kernel(const float* __restrict__ logits, const float* __restrict__ mask, const double* __restrict__ parts, const float* __restrict__ gscale,
                   int B, long long N, float lambda_dice, float lambda_ce, float* __restrict__ dlogits) {
    const int b = blockIdx.y;
    const double I = parts[b * 4 + 0], P = parts[b * 4 + 1], G = parts[b * 4 + 2];
    const double den = P + G + 1e-5;
    const float gs = gscale ? *gscale : 1.0f;
    // d dice_b / d p_i = -(2 y_i den - (2I + s)) / den^2  = a * y_i + c
    const float a = static_cast<float>(-2.0 / den) * (lambda_dice / B) * gs;
    const float c = static_cast<float>((2.0 * I + 1e-5) / (den * den)) * (lambda_dice / B) * gs;
    const float w = lambda_ce / (static_cast<float>(B) * static_cast<float>(N)) * gs;
    const float* x = logits + static_cast<long long>(b) * N;
    const float* y = mask + static_cast<long long>(b) * N;
    float* o = dlogits + static_cast<long long>(b) * N;
    const bool vec = (N % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0) && ((reinterpret_cast<uintptr_t>(y) & 15) == 0) &&
                     ((reinterpret_cast<uintptr_t>(o) & 15) == 0);
    auto f = [&](float xv, float yv) {
        const float t = __expf(-fabsf(xv));
        const float r = __fdividef(1.0f, 1.0f + t);
        const float p = xv >= 0.f ? r : t * r;
        return (a * yv + c) * p * (1.0f - p) + w * (p - yv);
    };
    if (vec) {
        const long long n4 = N >> 2;
        for (long long i = static_cast<long long>(blockIdx.x) * LOSS_THREADS + threadIdx.x; i < n4; i += static_cast<long long>(gridDim.x) * LOSS_THREADS) {
            const float4 xv = __ldg(reinterpret_cast<const float4*>(x) + i);
            const float4 yv = __ldg(reinterpret_cast<const float4*>(y) + i);
            reinterpret_cast<float4*>(o)[i] = make_float4(f(xv.x, yv.x), f(xv.y, yv.y), f(xv.z, yv.z), f(xv.w, yv.w));
        }
    } else {
        for (long long i = static_cast<long long>(blockIdx.x) * LOSS_THREADS + threadIdx.x; i < N; i += static_cast<long long>(gridDim.x) * LOSS_THREADS)
            o[i] = f(x[i], y[i]);
    }
}

// B * 48 bytes of dynamic shared memory: above 48 KB (B > 1024, up to the 4096 the entry points accept = 192 KB) the kernel
// needs the opt-in attribute, otherwise the launch fails
static int finalize_smem_opt_in(size_t sh) {
    static size_t granted = 48 * 1024;
    if (sh <= granted) return 0;
    cudaError_t e = cudaFuncSetAttribute(dicebce_finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sh));
    if (e != cudaSuccess) {
        set_error("dicebce_finalize_kernel: cannot opt in to %zu bytes of shared memory: %s", sh, cudaGetErrorString(e));
        return -2;
    }
    granted = sh;
    return 0;
}

}  // namespace tvs

extern "C" __attribute__((visibility("default"))) int64_t tvs_dicebce_scratch_bytes(int32_t B, int64_t N) {
    const int nblk = tvs::loss_blocks_per_sample(B, N);
    return static_cast<int64_t>(B) * nblk * (4 * sizeof(double) + 7 * sizeof(long long));
}

extern "C" __attribute__((visibility("default"))) int tvs_dicebce_metrics_fwd(const float* logits, const float* mask, int32_t B, int64_t N, float threshold, float lambda_dice,
                                       float lambda_ce, double* parts, int64_t* counts, int64_t* confmat, float* loss, void* scratch,
                                       void* stream) {
    using namespace tvs;
    TVS_REQUIRE(logits && mask && scratch, "tvs_dicebce_metrics_fwd: null pointer");
    TVS_REQUIRE(B > 0 && B <= 4096 && N > 0, "tvs_dicebce_metrics_fwd: bad shape B=%d N=%lld", B, (long long)N);
    const int nblk = loss_blocks_per_sample(B, N);
    double* part = static_cast<double*>(scratch);
    long long* cnt = reinterpret_cast<long long*>(part + static_cast<long long>(B) * nblk * 4);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    dicebce_partial_kernel<false><<<dim3(nblk, B), LOSS_THREADS, 0, st>>>(logits, mask, N, threshold, part, cnt);
    if (int rc = check_launch("dicebce_partial_kernel")) return rc;
    const size_t sh = static_cast<size_t>(B) * (2 * sizeof(double) + 4 * sizeof(long long));
    if (int rc = finalize_smem_opt_in(sh)) return rc;
    dicebce_finalize_kernel<<<1, 256, sh, st>>>(part, cnt, B, nblk, N, lambda_dice, lambda_ce, parts, reinterpret_cast<long long*>(counts),
                                                reinterpret_cast<long long*>(confmat), loss);
    return check_launch("dicebce_finalize_kernel");
}

extern "C" __attribute__((visibility("default"))) int tvs_dicebce_bwd(const float* logits, const float* mask, const double* parts, const float* gscale, int32_t B, int64_t N,
                               float lambda_dice, float lambda_ce, float* dlogits, void* stream) {
    using namespace tvs;
    TVS_REQUIRE(logits && mask && parts && dlogits, "tvs_dicebce_bwd: null pointer");
    TVS_REQUIRE(B > 0 && N > 0, "tvs_dicebce_bwd: bad shape");
    const int nblk = loss_blocks_per_sample(B, N);
    dicebce_bwd_kernel<<<dim3(nblk, B), LOSS_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(logits, mask, parts, gscale, B, N, lambda_dice,
                                                                                              lambda_ce, dlogits);
    return check_launch("dicebce_bwd_kernel");
}

// metric(preds, target) entry for callers that only hold probabilities (torchmetrics-style API): same counters, no loss.
extern "C" __attribute__((visibility("default"))) int tvs_metrics_from_probs(const float* preds, const float* mask, int32_t B, int64_t N, float threshold,
                                                                              int64_t* counts, int64_t* confmat, void* scratch, void* stream) {
    using namespace tvs;
    TVS_REQUIRE(preds && mask && scratch, "tvs_metrics_from_probs: null pointer");
    TVS_REQUIRE(B > 0 && B <= 4096 && N > 0, "tvs_metrics_from_probs: bad shape B=%d N=%lld", B, (long long)N);
    const int nblk = loss_blocks_per_sample(B, N);
    double* part = static_cast<double*>(scratch);
    long long* cnt = reinterpret_cast<long long*>(part + static_cast<long long>(B) * nblk * 4);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    dicebce_partial_kernel<true><<<dim3(nblk, B), LOSS_THREADS, 0, st>>>(preds, mask, N, threshold, part, cnt);
    if (int rc = check_launch("dicebce_partial_kernel")) return rc;
    const size_t sh = static_cast<size_t>(B) * (2 * sizeof(double) + 4 * sizeof(long long));
    if (int rc = finalize_smem_opt_in(sh)) return rc;
    dicebce_finalize_kernel<<<1, 256, sh, st>>>(part, cnt, B, nblk, N, 0.f, 0.f, nullptr, reinterpret_cast<long long*>(counts),
                                                reinterpret_cast<long long*>(confmat), nullptr);
    return check_launch("dicebce_finalize_kernel");
}
