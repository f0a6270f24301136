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
#include <stdlib.h>

#include "common.cuh"
#include "tvs_b200.h"

namespace tvs {

constexpr int LOSS_THREADS = 256;
constexpr int NSLOT = 9;          // per block: I, P, G, bce (double) + tp, n_ge, n_t, tp_gt, n_gt (int64), 9 x 8 bytes per slot

static int loss_blocks_per_sample(int B, long long N, int per_sm = 8) {
    static const int env = [] { const char* e = getenv("TVS_LOSS_BLOCKS_PER_SM"); return e ? atoi(e) : 0; }();     // tuning switch
    if (env > 0) per_sm = env;
    long long want = (static_cast<long long>(per_sm) * sm_count() + B - 1) / B;      // blocks of 256 threads per SM, over all samples
    long long cap = (N + 8 * LOSS_THREADS - 1) / (8 * LOSS_THREADS);      // at least two quads per thread
    long long n = want < cap ? want : cap;
    return static_cast<int>(n < 1 ? 1 : n);
}

// Per-pixel work, kept off the integer pipe: the first version spent 58 % of the ALU pipe on 11 integer counters per pixel and
// a 64-bit float->int conversion of the mask, and ran at 58 % of the HBM roofline (ncu, B = 256 @ 416^2).  The counters are
// kept as FLOATS (0/1 increments, exact below 2^24 per thread and per block) on the FMA pipe, and only five are needed:
//   tp = #(p >= thr, t), n_ge = #(p >= thr), n_t = #(t), tp_gt = #(p > thr, t), n_gt = #(p > thr)
//   fp = n_ge - tp, fn = n_t - tp;  confusion matrix: c11 = tp_gt, c01 = n_gt - tp_gt, c10 = n_t - tp_gt, c00 = n - n_t - c01
// t = mask.long() of a {0, 1} mask (image_text_mask_module.py:107): t = (y >= 1).
struct PixelStats {
    float I, P, G, bce;
    float tp, nge, nt, tpgt, ngt;
};

template <bool PROBS>
__device__ __forceinline__ void pixel(float x, float y, float thr, PixelStats& s) {
    float p;
    if (PROBS) {
        p = x;   // the caller already holds probabilities (metric(preds, target) API): threshold them as they are
    } else {
        // fast path: one MUFU.EX2, one MUFU.RCP, one MUFU.LG2 per pixel.  t = exp(-|x|) in (0, 1];
        // sigmoid(x) = 1 / (1 + t) for x >= 0, t / (1 + t) otherwise.
        const float t = __expf(-fabsf(x));
        const float r = __fdividef(1.0f, 1.0f + t);
        p = x >= 0.f ? r : t * r;
        if (fabsf(p - thr) < 1e-4f) {      // the integer counters must match fl32(1 / fl32(1 + fl32(exp(-x)))) exactly
            const float e = static_cast<float>(exp(-static_cast<double>(x)));
            p = __fdiv_rn(1.0f, __fadd_rn(1.0f, e));
        }
        s.bce += fmaxf(x, 0.f) - x * y + __logf(1.0f + t);
    }
    s.I = fmaf(p, y, s.I);
    s.P += p;
    s.G += y;
    const float tf = y >= 1.0f ? 1.0f : 0.0f;
    const float gef = p >= thr ? 1.0f : 0.0f, gtf = p > thr ? 1.0f : 0.0f;
    s.tp = fmaf(gef, tf, s.tp);
    s.nge += gef;
    s.nt += tf;
    s.tpgt = fmaf(gtf, tf, s.tpgt);
    s.ngt += gtf;
}

template <bool PROBS>
__global__ void __launch_bounds__(LOSS_THREADS)
dicebce_partial_kernel(const float* __restrict__ logits, const float* __restrict__ mask, long long N, float thr, double* __restrict__ slot_out) {
    const int b = blockIdx.y, nblk = gridDim.x;
    const float* x = logits + static_cast<long long>(b) * N;
    const float* y = mask + static_cast<long long>(b) * N;
    PixelStats s = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const bool vec = (N % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0) && ((reinterpret_cast<uintptr_t>(y) & 15) == 0);
    if (vec) {
        const long long n4 = N >> 2, stride = static_cast<long long>(nblk) * LOSS_THREADS;
        long long i = static_cast<long long>(blockIdx.x) * LOSS_THREADS + threadIdx.x;
        // two quads per trip: four 128-bit loads in flight per thread before the first use
        for (; i + stride < n4; i += 2 * stride) {
            const float4 xa = __ldcs(reinterpret_cast<const float4*>(x) + i), ya = __ldcs(reinterpret_cast<const float4*>(y) + i);
            const float4 xb = __ldcs(reinterpret_cast<const float4*>(x) + i + stride), yb = __ldcs(reinterpret_cast<const float4*>(y) + i + stride);
            pixel<PROBS>(xa.x, ya.x, thr, s); pixel<PROBS>(xa.y, ya.y, thr, s); pixel<PROBS>(xa.z, ya.z, thr, s); pixel<PROBS>(xa.w, ya.w, thr, s);
            pixel<PROBS>(xb.x, yb.x, thr, s); pixel<PROBS>(xb.y, yb.y, thr, s); pixel<PROBS>(xb.z, yb.z, thr, s); pixel<PROBS>(xb.w, yb.w, thr, s);
        }
        if (i < n4) {
            const float4 xa = __ldcs(reinterpret_cast<const float4*>(x) + i), ya = __ldcs(reinterpret_cast<const float4*>(y) + i);
            pixel<PROBS>(xa.x, ya.x, thr, s); pixel<PROBS>(xa.y, ya.y, thr, s); pixel<PROBS>(xa.z, ya.z, thr, s); pixel<PROBS>(xa.w, ya.w, thr, s);
        }
    } else {
        for (long long i = static_cast<long long>(blockIdx.x) * LOSS_THREADS + threadIdx.x; i < N; i += static_cast<long long>(nblk) * LOSS_THREADS)
            pixel<PROBS>(x[i], y[i], thr, s);
    }
    // block reduction.  Loss sums: float per thread, double across threads.  Counters: exact small integers in float, summed in
    // float inside the warp (<= 32 * 2^24 / 32 ...: a warp never sees 2^24 pixels), in double across warps.
    __shared__ double sd[LOSS_THREADS / 32][NSLOT];
    double d[4] = {s.I, s.P, s.G, s.bce};
    float c[5] = {s.tp, s.nge, s.nt, s.tpgt, s.ngt};
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int k = 0; k < 4; ++k) d[k] += __shfl_xor_sync(0xffffffffu, d[k], o);
#pragma unroll
        for (int k = 0; k < 5; ++k) c[k] += __shfl_xor_sync(0xffffffffu, c[k], o);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 4; ++k) sd[warp][k] = d[k];
#pragma unroll
        for (int k = 0; k < 5; ++k) sd[warp][4 + k] = static_cast<double>(c[k]);
    }
    __syncthreads();
    if (threadIdx.x < NSLOT) {
        double t = 0;
        for (int w = 0; w < LOSS_THREADS / 32; ++w) t += sd[w][threadIdx.x];
        slot_out[(static_cast<long long>(b) * nblk + blockIdx.x) * NSLOT + threadIdx.x] = t;      // counters: exact integers in double
    }
}

// One block, one warp per sample at a time: lanes stride over the blocks of that sample (all loads of a sample in flight
// together), fixed-order butterfly - deterministic.  Then warp 0 reduces over the samples, again in a fixed order.
__global__ void __launch_bounds__(1024)
dicebce_finalize_kernel(const double* __restrict__ slot_in, int B, int nblk, long long N, float lambda_dice, float lambda_ce,
                        double* __restrict__ parts, long long* __restrict__ counts, long long* __restrict__ confmat, float* __restrict__ loss) {
    extern __shared__ double sh[];  // [B] dice term, [B] bce sum, then 4*B conf counts (exact integers in double)
    double* s_dice = sh;
    double* s_bce = sh + B;
    double* s_conf = sh + 2 * B;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
    for (int b = warp; b < B; b += nwarp) {
        double v[NSLOT];
#pragma unroll
        for (int j = 0; j < NSLOT; ++j) v[j] = 0;
        for (int k = lane; k < nblk; k += 32) {
            const double* sl = slot_in + (static_cast<long long>(b) * nblk + k) * NSLOT;
#pragma unroll
            for (int j = 0; j < NSLOT; ++j) v[j] += sl[j];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int j = 0; j < NSLOT; ++j) v[j] += __shfl_xor_sync(0xffffffffu, v[j], o);
        }
        if (lane == 0) {
            const double tp = v[4], nge = v[5], nt = v[6], tpgt = v[7], ngt = v[8];
            if (parts)
                for (int j = 0; j < 4; ++j) parts[b * 4 + j] = v[j];
            if (counts) {
                counts[b * 3 + 0] = static_cast<long long>(tp);
                counts[b * 3 + 1] = static_cast<long long>(nge - tp);
                counts[b * 3 + 2] = static_cast<long long>(nt - tp);
            }
            const double c01 = ngt - tpgt;
            s_conf[b * 4 + 0] = static_cast<double>(N) - nt - c01;     // tn
            s_conf[b * 4 + 1] = c01;                                   // fp
            s_conf[b * 4 + 2] = nt - tpgt;                             // fn
            s_conf[b * 4 + 3] = tpgt;                                  // tp
            s_dice[b] = 1.0 - (2.0 * v[0] + 1e-5) / (v[1] + v[2] + 1e-5);
            s_bce[b] = v[3];
        }
    }
    __syncthreads();
    if (warp == 0) {
        double acc[6] = {0, 0, 0, 0, 0, 0};
        for (int b = lane; b < B; b += 32) {
            acc[0] += s_dice[b];
            acc[1] += s_bce[b];
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[2 + j] += s_conf[b * 4 + j];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int j = 0; j < 6; ++j) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
        }
        if (lane == 0) {
            if (loss) *loss = static_cast<float>(lambda_dice * acc[0] / B + lambda_ce * acc[1] / (static_cast<double>(B) * static_cast<double>(N)));
            if (confmat)
                for (int j = 0; j < 4; ++j) confmat[j] += static_cast<long long>(acc[2 + j]);
        }
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
    return static_cast<int64_t>(B) * nblk * tvs::NSLOT * sizeof(double);
}

extern "C" __attribute__((visibility("default"))) int tvs_dicebce_metrics_fwd(const float* logits, const float* mask, int32_t B, int64_t N, float threshold, float lambda_dice,
                                       float lambda_ce, double* parts, int64_t* counts, int64_t* confmat, float* loss, void* scratch,
                                       void* stream) {
    using namespace tvs;
    TVS_REQUIRE(logits && mask && scratch, "tvs_dicebce_metrics_fwd: null pointer");
    TVS_REQUIRE(B > 0 && B <= 4096 && N > 0, "tvs_dicebce_metrics_fwd: bad shape B=%d N=%lld", B, (long long)N);
    const int nblk = loss_blocks_per_sample(B, N);
    double* slots = static_cast<double*>(scratch);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    dicebce_partial_kernel<false><<<dim3(nblk, B), LOSS_THREADS, 0, st>>>(logits, mask, N, threshold, slots);
    if (int rc = check_launch("dicebce_partial_kernel")) return rc;
    const size_t sh = static_cast<size_t>(B) * 6 * sizeof(double);
    if (int rc = finalize_smem_opt_in(sh)) return rc;
    dicebce_finalize_kernel<<<1, 1024, sh, st>>>(slots, B, nblk, N, lambda_dice, lambda_ce, parts, reinterpret_cast<long long*>(counts),
                                                 reinterpret_cast<long long*>(confmat), loss);
    return check_launch("dicebce_finalize_kernel");
}

extern "C" __attribute__((visibility("default"))) int tvs_dicebce_bwd(const float* logits, const float* mask, const double* parts, const float* gscale, int32_t B, int64_t N,
                               float lambda_dice, float lambda_ce, float* dlogits, void* stream) {
    using namespace tvs;
    TVS_REQUIRE(logits && mask && parts && dlogits, "tvs_dicebce_bwd: null pointer");
    TVS_REQUIRE(B > 0 && N > 0, "tvs_dicebce_bwd: bad shape");
    const int nblk = loss_blocks_per_sample(B, N, 16);
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
    double* slots = static_cast<double*>(scratch);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    dicebce_partial_kernel<true><<<dim3(nblk, B), LOSS_THREADS, 0, st>>>(preds, mask, N, threshold, slots);
    if (int rc = check_launch("dicebce_partial_kernel")) return rc;
    const size_t sh = static_cast<size_t>(B) * 6 * sizeof(double);
    if (int rc = finalize_smem_opt_in(sh)) return rc;
    dicebce_finalize_kernel<<<1, 1024, sh, st>>>(slots, B, nblk, N, 0.f, 0.f, nullptr, reinterpret_cast<long long*>(counts),
                                                 reinterpret_cast<long long*>(confmat), nullptr);
    return check_launch("dicebce_finalize_kernel");
}
