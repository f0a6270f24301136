// HBM-bound glue kernels of the CLIPSeg prompt-tuning path: patch im2col, embedding assembly, deep-prompt row
// replacement (+ gradient), FiLM, the fused decoder head (transposed-conv pixel shuffle + additive upsample/5x5
// stencil + blend), casts and the flat AdamW.  All coalesced / vectorised, no shared-memory staging needed except
// the head, which keeps its low-resolution neighbourhood in shared memory.
#include "common.cuh"
#include "tvs_b200.h"

namespace tvs {

static inline unsigned blocks_for(long long n, int threads, int max_blocks = 1 << 20) {
    long long b = (n + threads - 1) / threads;
    if (b > max_blocks) b = max_blocks;
    return static_cast<unsigned>(b < 1 ? 1 : b);
}

// ------------------------------------------------------------------------------------------------
// im2col for the stride-P patch embedding (transformers modeling_clipseg.py:141-147, :202-203)
// out[(b*G + gy)*G + gx][c*P*P + py*P + px] = image[b][c][gy*P+py][gx*P+px]
// ------------------------------------------------------------------------------------------------
__global__ void im2col_kernel(const float* __restrict__ img, int B, int C, int H, int W, int P, __nv_bfloat16* __restrict__ out, int f16) {
    const int G = W / P, GH = H / P;
    const long long total4 = static_cast<long long>(B) * C * H * W / 4;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total4; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long e = i * 4;
        const int xw = static_cast<int>(e % W);
        const int yh = static_cast<int>((e / W) % H);
        const int c = static_cast<int>((e / (static_cast<long long>(W) * H)) % C);
        const int b = static_cast<int>(e / (static_cast<long long>(W) * H * C));
        const float4 v = __ldg(reinterpret_cast<const float4*>(img) + i);
        const int gy = yh / P, py = yh % P, gx = xw / P, px = xw % P;   // P % 4 == 0 -> the 4 pixels stay in one patch row
        const long long row = (static_cast<long long>(b) * GH + gy) * G + gx;
        const long long col = (static_cast<long long>(c) * P + py) * P + px;
        *reinterpret_cast<uint2*>(out + row * (static_cast<long long>(C) * P * P) + col) =
            f16 ? make_uint2(pack_f16x2(v.x, v.y), pack_f16x2(v.z, v.w)) : make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
    }
}

// h[b,0] = cls + pos[0]; h[b,1+p] = patches[b*G2+p] + pos[1+p]; h[b,1+G2+j] = ctx[(b),j]
__global__ void vision_assemble_kernel(const float* __restrict__ patches, const float* __restrict__ cls, const float* __restrict__ pos,
                                       const float* __restrict__ ctx, long long ctx_bs, int B, int G2, int n, int D, float* __restrict__ h) {
    const int S = 1 + G2 + n;
    const int d4 = D / 4;
    const long long total = static_cast<long long>(B) * S * d4;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(i % d4);
        const int s = static_cast<int>((i / d4) % S);
        const int b = static_cast<int>(i / (static_cast<long long>(d4) * S));
        float4 v;
        if (s == 0) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(cls) + c), p = __ldg(reinterpret_cast<const float4*>(pos) + c);
            v = make_float4(a.x + p.x, a.y + p.y, a.z + p.z, a.w + p.w);
        } else if (s <= G2) {
            const float4 a = reinterpret_cast<const float4*>(patches + (static_cast<long long>(b) * G2 + (s - 1)) * D)[c];
            const float4 p = __ldg(reinterpret_cast<const float4*>(pos + static_cast<long long>(s) * D) + c);
            v = make_float4(a.x + p.x, a.y + p.y, a.z + p.z, a.w + p.w);
        } else {
            v = __ldg(reinterpret_cast<const float4*>(ctx + b * ctx_bs + static_cast<long long>(s - 1 - G2) * D) + c);
        }
        reinterpret_cast<float4*>(h)[i] = v;
    }
}

// x[b, row0+j, :] = ctx[(b), j, :]
__global__ void prompt_overwrite_kernel(float* __restrict__ x, __nv_bfloat16* __restrict__ x16, int B, int S, int D, int row0, int n,
                                        const float* __restrict__ ctx, long long ctx_bs) {
    const int d4 = D / 4;
    const long long total = static_cast<long long>(B) * n * d4;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(i % d4);
        const int j = static_cast<int>((i / d4) % n);
        const int b = static_cast<int>(i / (static_cast<long long>(d4) * n));
        const float4 v = __ldg(reinterpret_cast<const float4*>(ctx + b * ctx_bs + static_cast<long long>(j) * D) + c);
        const long long off = (static_cast<long long>(b) * S + row0 + j) * D;
        reinterpret_cast<float4*>(x + off)[c] = v;
        if (x16) reinterpret_cast<uint2*>(x16 + off)[c] = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
    }
}

// dctx[(b), j, d] (+)= sum_b dx[b, row0+j, d]; optionally zero those rows of dx (and of its bf16 mirror).
// Batch-reduced form: grid (n, ceil(D/128)), block (128 channels, 8 batch groups); the batch loop is split over
// threadIdx.y and finished through shared memory, so no thread walks the whole batch serially.
__global__ void __launch_bounds__(1024)
prompt_grad_reduce_kernel(float* __restrict__ dx, __nv_bfloat16* __restrict__ dx16, int B, int S, int D, int row0, float* __restrict__ dctx,
                          int zero_rows) {
    __shared__ float part[8][128];
    const int j = blockIdx.x, d = blockIdx.y * 128 + threadIdx.x, g = threadIdx.y;
    float acc = 0.f;
    if (d < D) {
        for (int b = g; b < B; b += 8) {
            const long long off = (static_cast<long long>(b) * S + row0 + j) * D + d;
            acc += dx[off];
            if (zero_rows) {
                dx[off] = 0.f;
                if (dx16) dx16[off] = __float2bfloat16(0.f);
            }
        }
    }
    part[g][threadIdx.x] = acc;
    __syncthreads();
    if (g == 0 && d < D) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += part[k][threadIdx.x];
        dctx[static_cast<long long>(j) * D + d] += t;
    }
}

// per-sample form (CoCoOp): dctx[b, j, d] += dx[b, row0+j, d]
__global__ void prompt_grad_kernel(float* __restrict__ dx, __nv_bfloat16* __restrict__ dx16, int B, int S, int D, int row0, int n,
                                   float* __restrict__ dctx, long long ctx_bs, int zero_rows) {
    const long long per = static_cast<long long>(n) * D;
    const long long total = per * B;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int d = static_cast<int>(i % D);
        const int j = static_cast<int>((i / D) % n);
        const int b = static_cast<int>(i / per);
        const long long off = (static_cast<long long>(b) * S + row0 + j) * D + d;
        dctx[b * ctx_bs + static_cast<long long>(j) * D + d] += dx[off];
        if (zero_rows) {
            dx[off] = 0.f;
            if (dx16) dx16[off] = __float2bfloat16(0.f);
        }
    }
}

__global__ void cast_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, long long n4) {
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
        reinterpret_cast<uint2*>(y)[i] = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
    }
}

__global__ void add_f32_kernel(float* __restrict__ y, const float* __restrict__ x, long long n4) {
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        float4 a = reinterpret_cast<float4*>(y)[i];
        const float4 b = __ldg(reinterpret_cast<const float4*>(x) + i);
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
        reinterpret_cast<float4*>(y)[i] = a;
    }
}


// y[b, r, :] = x[b, row0 + r, :]  for r < nrows  (strip CLS / prompt rows; base_clipseg.py:132-142)
__global__ void slice_rows_kernel(const float* __restrict__ x, int B, int S, int D, int row0, int nrows, float* __restrict__ y32,
                                  __nv_bfloat16* __restrict__ y16) {
    const int d4 = D / 4;
    const long long total = static_cast<long long>(B) * nrows * d4;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(i % d4);
        const int r = static_cast<int>((i / d4) % nrows);
        const int b = static_cast<int>(i / (static_cast<long long>(d4) * nrows));
        const float4 v = reinterpret_cast<const float4*>(x + (static_cast<long long>(b) * S + row0 + r) * D)[c];
        if (y32) reinterpret_cast<float4*>(y32)[i] = v;
        if (y16) reinterpret_cast<uint2*>(y16)[i] = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
    }
}
// dx[b, s, :] = (row0 <= s < row0 + nrows) ? dy[b, s - row0, :] : 0
__global__ void unslice_rows_kernel(const float* __restrict__ dy, int B, int S, int D, int row0, int nrows, float* __restrict__ dx) {
    const int d4 = D / 4;
    const long long total = static_cast<long long>(B) * S * d4;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(i % d4);
        const int s = static_cast<int>((i / d4) % S);
        const int b = static_cast<int>(i / (static_cast<long long>(d4) * S));
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (s >= row0 && s < row0 + nrows) v = reinterpret_cast<const float4*>(dy + (static_cast<long long>(b) * nrows + (s - row0)) * D)[c];
        reinterpret_cast<float4*>(dx)[i] = v;
    }
}

// dw[n, k] += sum_m dy[m, n] * x[m, k]   for a SMALL weight (N * K <= 4096): the trainable additive 5x5 conv of the
// decoder head (base_clipseg.py:63-70) contracted at low resolution.  Each block reduces a slab of rows through
// shared memory and issues one atomicAdd per output.
constexpr int WG_ROWS = 64;
__global__ void __launch_bounds__(256)
wgrad_small_kernel(const float* __restrict__ dy, long long ld_dy, const float* __restrict__ x, long long ld_x, long long M, int N, int K,
                   float* __restrict__ dw) {
    extern __shared__ float sm[];
    float* s_dy = sm;                 // [WG_ROWS][N]
    float* s_x = sm + WG_ROWS * N;    // [WG_ROWS][K]
    const int NK = N * K;
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = 0.f;
    for (long long m0 = static_cast<long long>(blockIdx.x) * WG_ROWS; m0 < M; m0 += static_cast<long long>(gridDim.x) * WG_ROWS) {
        const int rows = static_cast<int>(min(static_cast<long long>(WG_ROWS), M - m0));
        for (int i = threadIdx.x; i < rows * N; i += blockDim.x) s_dy[i] = dy[(m0 + i / N) * ld_dy + i % N];
        for (int i = threadIdx.x; i < rows * K; i += blockDim.x) s_x[i] = x[(m0 + i / K) * ld_x + i % K];
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int o = threadIdx.x + j * 256;
            if (o < NK) {
                const int n = o / K, k = o % K;
                float a = 0.f;
                for (int r = 0; r < rows; ++r) a += s_dy[r * N + n] * s_x[r * K + k];
                acc[j] += a;
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const int o = threadIdx.x + j * 256;
        if (o < NK) atomicAdd(dw + o, acc[j]);
    }
}

// ------------------------------------------------------------------------------------------------
// FiLM (base_clipseg.py:111-115)
// ------------------------------------------------------------------------------------------------
__global__ void film_fwd_kernel(const float* __restrict__ x, const float* __restrict__ mul, const float* __restrict__ add, int B, int S, int D,
                                float* __restrict__ y, __nv_bfloat16* __restrict__ y16) {
    const int d4 = D / 4;
    const long long total = static_cast<long long>(B) * S * d4;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(i % d4);
        const int b = static_cast<int>(i / (static_cast<long long>(d4) * S));
        const float4 v = reinterpret_cast<const float4*>(x)[i];
        const float4 m = __ldg(reinterpret_cast<const float4*>(mul + static_cast<long long>(b) * D) + c);
        const float4 a = __ldg(reinterpret_cast<const float4*>(add + static_cast<long long>(b) * D) + c);
        const float4 o = make_float4(m.x * v.x + a.x, m.y * v.y + a.y, m.z * v.z + a.z, m.w * v.w + a.w);
        if (y) reinterpret_cast<float4*>(y)[i] = o;
        if (y16) reinterpret_cast<uint2*>(y16)[i] = make_uint2(pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
    }
}

// one block (1024 threads) per sample: thread (grp, d) sums rows s = grp, grp + ngrp, ... of channel d, then a
// shared-memory reduction over the groups.  dmul[b,d] = sum_s dy*x ; dadd[b,d] = sum_s dy ; dx = mul * dy
__global__ void __launch_bounds__(1024)
film_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ mul, int S, int D,
                float* __restrict__ dx, float* __restrict__ dmul, float* __restrict__ dadd) {
    __shared__ float sm_m[1024], sm_a[1024];
    const int b = blockIdx.x;
    const int ngrp = max(1, static_cast<int>(blockDim.x) / D);
    for (int d0 = 0; d0 < D; d0 += blockDim.x) {
        const int d = d0 + (threadIdx.x % (D < static_cast<int>(blockDim.x) ? D : blockDim.x));
        const int grp = D < static_cast<int>(blockDim.x) ? threadIdx.x / D : 0;
        float sm = 0.f, sa = 0.f;
        if (d < D && grp < ngrp) {
            const float m = mul[static_cast<long long>(b) * D + d];
            // four rows per trip, all eight loads issued before the first use: with one row per trip the ~30 dependent DRAM round
            // trips of a thread were the kernel (24.6 us for 12 MB at B = 32); same summation order per accumulator pair as before
            // is NOT kept - the partial sums are combined at the end (fp32, fixed order: deterministic)
            float sm1 = 0.f, sa1 = 0.f, sm2 = 0.f, sa2 = 0.f, sm3 = 0.f, sa3 = 0.f;
            int s = grp;
            for (; s + 3 * ngrp < S; s += 4 * ngrp) {
                const long long o0 = (static_cast<long long>(b) * S + s) * D + d, st = static_cast<long long>(ngrp) * D;
                const float g0 = __ldcs(dy + o0), g1 = __ldcs(dy + o0 + st), g2 = __ldcs(dy + o0 + 2 * st), g3 = __ldcs(dy + o0 + 3 * st);
                const float x0 = __ldcs(x + o0), x1 = __ldcs(x + o0 + st), x2 = __ldcs(x + o0 + 2 * st), x3 = __ldcs(x + o0 + 3 * st);
                sm = fmaf(g0, x0, sm); sa += g0;
                sm1 = fmaf(g1, x1, sm1); sa1 += g1;
                sm2 = fmaf(g2, x2, sm2); sa2 += g2;
                sm3 = fmaf(g3, x3, sm3); sa3 += g3;
                dx[o0] = m * g0; dx[o0 + st] = m * g1; dx[o0 + 2 * st] = m * g2; dx[o0 + 3 * st] = m * g3;
            }
            for (; s < S; s += ngrp) {
                const long long off = (static_cast<long long>(b) * S + s) * D + d;
                const float g = dy[off];
                sm = fmaf(g, x[off], sm);
                sa += g;
                dx[off] = m * g;
            }
            sm = (sm + sm1) + (sm2 + sm3);
            sa = (sa + sa1) + (sa2 + sa3);
        }
        sm_m[threadIdx.x] = sm;
        sm_a[threadIdx.x] = sa;
        __syncthreads();
        if (grp == 0 && d < D) {
            for (int k = 1; k < ngrp; ++k) {
                sm += sm_m[threadIdx.x + k * D];
                sa += sm_a[threadIdx.x + k * D];
            }
            dmul[static_cast<long long>(b) * D + d] = sm;
            dadd[static_cast<long long>(b) * D + d] = sa;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// decoder head.  Bilinear source coordinate of nn.Upsample(scale_factor=P, mode="bilinear", align_corners=False):
//   src = max((dst + 0.5) / P - 0.5, 0);  i0 = floor(src);  i1 = min(i0 + 1, G - 1);  w1 = src - i0
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void bilin(int dst, int P, int G, int& i0, int& i1, float& w1) {
    float src = (static_cast<float>(dst) + 0.5f) / static_cast<float>(P) - 0.5f;
    src = fmaxf(src, 0.f);
    i0 = static_cast<int>(src);
    if (i0 > G - 1) i0 = G - 1;
    i1 = min(i0 + 1, G - 1);
    w1 = src - static_cast<float>(i0);
}

constexpr int HEAD_MAXK = 7;   // ksize <= 7

// grid (G /*patch row gy*/, B); HEAD_THREADS threads; thread X owns one image column of the P pixel rows of this patch row.
// The additive branch is evaluated separably: the bilinear weights along x depend only on (X, kx), so the contraction
// over ky and the interpolation along y are done ONCE per pixel row into T[xi][kx] (G*ks values, cooperatively), and a
// pixel then needs 2*ks shared-memory reads instead of 4*ks*ks (the first version was LDS/ALU bound at 187 us; the
// kernel's traffic - tconv in, logits (+ add_out) out - is worth ~10 us).
constexpr int HEAD_THREADS = 512;     // upper bound; the launch uses the image width rounded up to a warp
__global__ void __launch_bounds__(HEAD_THREADS, 2)
head_fwd_kernel(const float* __restrict__ tconv, long long ld_t, const float* __restrict__ addmap, long long ld_a, const float* __restrict__ bias_t,
                const float* __restrict__ bias_a, const float* __restrict__ ratio, int blend, int G, int P, int ks, float* __restrict__ logits,
                float* __restrict__ add_out) {
    extern __shared__ float s_head[];
    const int gy = blockIdx.x, b = blockIdx.y;
    const int KK = ks * ks, W = G * P, half = (ks - 1) / 2;
    float* s_add = s_head;                          // [3][G][KK]  addmap rows gy-1 .. gy+1 (clamped)
    float* s_T = s_add + 3 * G * KK;                // [P][G][ks]  the y-interpolated, ky-contracted additive map of every pixel row
    int* s_y0 = reinterpret_cast<int*>(s_T + P * G * ks);       // [P][ks] row taps relative to gy-1
    int* s_y1 = s_y0 + P * ks;
    float* s_wy = reinterpret_cast<float*>(s_y1 + P * ks);
    float wa = 1.f, wb = 0.f;
    if (blend == 1) { const float r = *ratio; wa = 1.f - r; wb = r; }
    else if (blend == 2) { wa = 1.f; wb = 1.f; }
    if (blend != 0) {
        for (int i = threadIdx.x; i < 3 * G * KK; i += blockDim.x) {
            const int k = i % KK, xi = (i / KK) % G, ry = i / (KK * G);
            const int yi = min(max(gy - 1 + ry, 0), G - 1);
            s_add[i] = addmap[(static_cast<long long>(b) * G * G + yi * G + xi) * ld_a + k];
        }
        for (int i = threadIdx.x; i < P * ks; i += blockDim.x) {
            const int py = i / ks, ky = i - py * ks;
            const int Yc = min(max(gy * P + py + ky - half, 0), W - 1);       // square images: H == W
            int y0, y1;
            float wy;
            bilin(Yc, P, G, y0, y1, wy);
            s_y0[i] = min(max(y0 - (gy - 1), 0), 2);
            s_y1[i] = min(max(y1 - (gy - 1), 0), 2);
            s_wy[i] = wy;
        }
    }
    const float bt = bias_t ? *bias_t : 0.f;
    const float ba = (blend != 0 && bias_a) ? *bias_a : 0.f;
    const int X = threadIdx.x;
    int x0[HEAD_MAXK], x1[HEAD_MAXK];
    float wx[HEAD_MAXK];
    if (blend != 0 && X < W) {
#pragma unroll
        for (int kx = 0; kx < HEAD_MAXK; ++kx)
            if (kx < ks) bilin(min(max(X + kx - half, 0), W - 1), P, G, x0[kx], x1[kx], wx[kx]);
    }
    const int gx = X / P, px = X - gx * P;
    // this thread's column of the transposed-convolution patch row: all P loads in flight before anything waits on them
    constexpr int PMAX = 16;
    float tv[PMAX];
    const float* trow = tconv + (static_cast<long long>(b) * G * G + gy * G + gx) * ld_t + px;
    if (X < W && P <= PMAX) {
#pragma unroll
        for (int py = 0; py < PMAX; ++py)
            if (py < P) tv[py] = trow[py * P];
    }
    __syncthreads();
    if (blend != 0) {
        // T[py][xi][kx] for every pixel row of the patch row at once (one barrier instead of one per row)
        for (int i = threadIdx.x; i < P * G * ks; i += blockDim.x) {
            const int py = i / (G * ks), r = i - py * (G * ks);
            const int xi = r / ks, kx = r - xi * ks;
            float acc = 0.f;
            for (int ky = 0; ky < ks; ++ky) {
                const float lo = s_add[(s_y0[py * ks + ky] * G + xi) * KK + ky * ks + kx];
                const float hi = s_add[(s_y1[py * ks + ky] * G + xi) * KK + ky * ks + kx];
                acc += lo + s_wy[py * ks + ky] * (hi - lo);
            }
            s_T[i] = acc;
        }
        __syncthreads();
    }
    if (X < W) {
#pragma unroll
        for (int py = 0; py < PMAX; ++py) {
            if (py >= P) break;
            const int Y = gy * P + py;
            const float* T = s_T + py * G * ks;
            float v = wa * ((P <= PMAX ? tv[py] : trow[py * P]) + bt);
            if (blend != 0) {
                float acc = ba;
#pragma unroll
                for (int kx = 0; kx < HEAD_MAXK; ++kx)
                    if (kx < ks) {
                        const float lo = T[x0[kx] * ks + kx], hi = T[x1[kx] * ks + kx];
                        acc += lo + wx[kx] * (hi - lo);
                    }
                if (add_out) add_out[(static_cast<long long>(b) * W + Y) * W + X] = acc;
                v += wb * acc;
            }
            logits[(static_cast<long long>(b) * W + Y) * W + X] = v;
        }
        for (int py = PMAX; py < P; ++py) {        // patch sizes above 16 (not a CLIPSeg / CRIS geometry): plain loop
            const int Y = gy * P + py;
            const float* T = s_T + py * G * ks;
            float v = wa * (trow[py * P] + bt);
            if (blend != 0) {
                float acc = ba;
#pragma unroll
                for (int kx = 0; kx < HEAD_MAXK; ++kx)
                    if (kx < ks) {
                        const float lo = T[x0[kx] * ks + kx], hi = T[x1[kx] * ks + kx];
                        acc += lo + wx[kx] * (hi - lo);
                    }
                if (add_out) add_out[(static_cast<long long>(b) * W + Y) * W + X] = acc;
                v += wb * acc;
            }
            logits[(static_cast<long long>(b) * W + Y) * W + X] = v;
        }
    }
}

// Rows gy-1..gy+1 of the low-res grid are addressed relative to (gy-1); at the image border the clamp of yi in the
// loader and of the relative index above agree because bilin() never asks for a row outside [0, G-1].

// backward part 1: dtconv = wa * dlogits (pixel un-shuffle, bf16) and the two scalar gradients
__global__ void __launch_bounds__(256)
head_bwd_pix_kernel(const float* __restrict__ dlogits, const float* __restrict__ tconv, long long ld_t, const float* __restrict__ add_out,
                    const float* __restrict__ bias_t, const float* __restrict__ ratio, int blend, int G, int P, __nv_bfloat16* __restrict__ dtconv,
                    long long ld_dt, float* __restrict__ dbias_a, float* __restrict__ dratio) {
    const int gy = blockIdx.x, b = blockIdx.y;
    const int W = G * P;
    float wa = 1.f, wb = 0.f;
    if (blend == 1) { const float r = *ratio; wa = 1.f - r; wb = r; }
    else if (blend == 2) { wa = 1.f; wb = 1.f; }
    const float bt = bias_t ? *bias_t : 0.f;
    float s_g = 0.f, s_r = 0.f;
    for (int i = threadIdx.x; i < P * W; i += blockDim.x) {
        const int py = i / W, X = i % W;
        const int Y = gy * P + py;
        const int gx = X / P, px = X % P;
        const long long pix = (static_cast<long long>(b) * W + Y) * W + X;
        const float g = dlogits[pix];
        const long long t_off = (static_cast<long long>(b) * G * G + gy * G + gx);
        dtconv[t_off * ld_dt + py * P + px] = __float2bfloat16(wa * g);
        s_g += g;
        if (blend == 1 && add_out) s_r += g * (add_out[pix] - (tconv[t_off * ld_t + py * P + px] + bt));
    }
    s_g = warp_sum(s_g);
    s_r = warp_sum(s_r);
    __shared__ float sh[2][8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { sh[0][warp] = s_g; sh[1][warp] = s_r; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float tg = 0.f, tr = 0.f;
        for (int w = 0; w < (blockDim.x >> 5); ++w) { tg += sh[0][w]; tr += sh[1][w]; }
        if (dbias_a && blend != 0) atomicAdd(dbias_a, wb * tg);
        if (dratio && blend == 1) atomicAdd(dratio, tr);
    }
}

// backward part 2: daddmap[b, yi, xi, (ky,kx)] = wb * sum_{Y,X} dlogits[b,Y,X] * wy(Y;ky,yi) * wx(X;kx,xi), separable.
// grid (G /*yi*/, B), 256 threads.  Step 1: T[Yrel][xi][kx] = sum_X dl[Y][X] wx(X;kx,xi) for the <= 2P+ks rows that can
// touch yi; step 2: contract over Y with wy.
__device__ __forceinline__ float bilin_weight_on(int dst_clamped, int P, int G, int target) {
    int i0, i1;
    float w1;
    bilin(dst_clamped, P, G, i0, i1, w1);
    float w = 0.f;
    if (i0 == target) w += 1.f - w1;
    if (i1 == target) w += w1;
    return w;
}

// grid (G /*yi*/, B), 256 threads.  w(c; t) = bilinear weight with which clamped coordinate c reads low-res index t; it is
// tabulated once per block for every t over the 2P + 2 coordinates that can touch it (the image is square, so the same
// table serves x and y).  Step 1 walks each (pixel row, xi) once and feeds all ks column taps from that table (pixel rows
// staged through shared memory in coalesced chunks); step 2 contracts over the pixel rows.  The first version evaluated
// the weight (a float division) per (row, xi, kx, X) and read dlogits uncoalesced: 200 us for ~25 us of work.
constexpr int HB_CHUNK = 8;
__global__ void __launch_bounds__(256)
head_bwd_addmap_kernel(const float* __restrict__ dlogits, const float* __restrict__ ratio, int blend, int G, int P, int ks,
                       float* __restrict__ daddmap, long long ld_da) {
    extern __shared__ float s_hb[];
    const int yi = blockIdx.x, b = blockIdx.y;
    const int W = G * P, half = (ks - 1) / 2;
    const int TW = 2 * P + 2;                                   // coordinates base(t) .. base(t) + TW - 1 can read index t
    float* s_T = s_hb;                                          // [NRmax][G][ks]
    const int NRmax = 2 * P + ks + 2;
    float* s_w = s_T + NRmax * G * ks;                          // [G][TW]
    float* s_dl = s_w + G * TW;                                 // [HB_CHUNK][WP]
    const int WP = W + G;                                       // padded row: one extra word per patch
    float wb = 1.f;
    if (blend == 1) wb = *ratio;
    for (int i = threadIdx.x; i < G * TW; i += blockDim.x) {
        const int t = i / TW, c = t * P - P / 2 - 1 + (i - t * TW);
        s_w[i] = (c >= 0 && c < W) ? bilin_weight_on(c, P, G, t) : 0.f;
    }
    // rows Y whose clamped shifted coordinate can have bilinear support on yi
    const int Ylo = max(yi * P - P / 2 - half - 1, 0);
    const int Yhi = min(yi * P + P + P / 2 + half, W - 1);
    const int NR = Yhi - Ylo + 1;
    const float* dl = dlogits + static_cast<long long>(b) * W * W;
    for (int r0 = 0; r0 < NR; r0 += HB_CHUNK) {
        const int nr = min(HB_CHUNK, NR - r0);
        __syncthreads();                                         // table ready / previous chunk consumed
        // one padding word per P columns: the tasks of a warp read row[xi * P + j] for consecutive xi, a stride of P = 16
        // words (16-way bank conflicts); with stride P + 1 they fall into distinct banks
        for (int i = threadIdx.x; i < nr * W; i += blockDim.x) {
            const int yr = i / W, X = i - yr * W;
            s_dl[yr * WP + X + X / P] = dl[static_cast<long long>(Ylo + r0) * W + i];
        }
        __syncthreads();
        auto do_task = [&](int yr, int xi) {
            const int base = xi * P - P / 2 - 1;
            const int Xlo = max(base - half, 0), Xhi = min(base + TW - 1 + half, W - 1);
            float acc[HEAD_MAXK];
#pragma unroll
            for (int kx = 0; kx < HEAD_MAXK; ++kx) acc[kx] = 0.f;
            const float* row = s_dl + yr * WP;
            const float* wt = s_w + xi * TW;
            if (P == 16 && ks == 5 && base - 4 >= 0 && base + 33 + 4 <= W - 1) {
                // interior patch column of the CLIPSeg / CRIS geometry (no clamp is active): rel = u + kx - 4 for the window
                // u = X - (base - 2), so acc[kx] = sum_rel wt[rel] * win[rel - kx + 4]: 72 shared-memory reads and 170 FMAs
                // from registers instead of ~1700 instructions of index arithmetic in the general loop below
                float win[38];
#pragma unroll
                for (int u = 0; u < 38; ++u) {
                    const int X = base - 2 + u;
                    win[u] = row[X + (X >> 4)];
                }
#pragma unroll
                for (int rel = 0; rel < 34; ++rel) {
                    const float w = wt[rel];
#pragma unroll
                    for (int kx = 0; kx < 5; ++kx) acc[kx] = fmaf(w, win[rel - kx + 4], acc[kx]);
                }
#pragma unroll
                for (int kx = 0; kx < 5; ++kx) s_T[((r0 + yr) * G + xi) * ks + kx] = acc[kx];
                return;
            }
            int Xp = Xlo + Xlo / P, rem = Xlo % P;             // padded index of X, kept without a division per step
            for (int X = Xlo; X <= Xhi; ++X) {
                const float g = row[Xp];
                ++Xp;
                if (++rem == P) { rem = 0; ++Xp; }
#pragma unroll
                for (int kx = 0; kx < HEAD_MAXK; ++kx)
                    if (kx < ks) {
                        const int rel = min(max(X + kx - half, 0), W - 1) - base;
                        if (rel >= 0 && rel < TW) acc[kx] = fmaf(wt[rel], g, acc[kx]);
                    }
            }
#pragma unroll
            for (int kx = 0; kx < HEAD_MAXK; ++kx)
                if (kx < ks) s_T[((r0 + yr) * G + xi) * ks + kx] = acc[kx];
        };
        // interior patch columns first, the two border columns (whose clamps need the general loop) last: a warp that holds
        // both kinds of task pays for both paths, so the border tasks are packed into the last warp
        const int n_mid = G > 2 ? G - 2 : 0;
        for (int task = threadIdx.x; task < nr * n_mid; task += blockDim.x) do_task(task / n_mid, 1 + task % n_mid);
        const int n_edge = G > 1 ? 2 : 1;
        // ... of the block, which has no interior task (nr * n_mid <= 8 * 20 < 224), so both kinds run side by side
        for (int task = static_cast<int>(threadIdx.x) - (static_cast<int>(blockDim.x) - 32); task >= 0 && task < nr * n_edge; task += 32)
            do_task(task / n_edge, (task % n_edge) ? G - 1 : 0);
    }
    __syncthreads();
    const int ybase = yi * P - P / 2 - 1;
    for (int i = threadIdx.x; i < G * ks * ks; i += blockDim.x) {
        const int kx = i % ks, ky = (i / ks) % ks, xi = i / (ks * ks);
        float acc = 0.f;
        for (int yr = 0; yr < NR; ++yr) {
            const int rel = min(max(Ylo + yr + ky - half, 0), W - 1) - ybase;
            if (rel >= 0 && rel < TW) acc = fmaf(s_w[yi * TW + rel], s_T[(yr * G + xi) * ks + kx], acc);
        }
        daddmap[(static_cast<long long>(b) * G * G + yi * G + xi) * ld_da + ky * ks + kx] = wb * acc;
    }
}

// ------------------------------------------------------------------------------------------------
// AdamW (torch.optim.AdamW: decoupled weight decay, bias-corrected)
// ------------------------------------------------------------------------------------------------
__global__ void adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, long long n, float lr,
                             float b1, float b2, float eps, float wd, int step, float gscale, const int* __restrict__ step_dev,
                             const float* __restrict__ lr_dev) {
    if (step_dev) step = *step_dev;
    if (lr_dev) lr = *lr_dev;
    const float bc1 = 1.f - powf(b1, static_cast<float>(step));
    const float bc2 = 1.f - powf(b2, static_cast<float>(step));
    const float step_size = lr / bc1;
    const float inv_sqrt_bc2 = 1.0f / sqrtf(bc2);
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const float gr = g[i] * gscale;
        float pv = p[i] * (1.f - lr * wd);
        const float mv = b1 * m[i] + (1.f - b1) * gr;
        const float vv = b2 * v[i] + (1.f - b2) * gr * gr;
        m[i] = mv;
        v[i] = vv;
        pv -= step_size * mv / (sqrtf(vv) * inv_sqrt_bc2 + eps);
        p[i] = pv;
    }
}
__global__ void counter_inc_kernel(int* c) { *c += 1; }

}  // namespace tvs

using namespace tvs;

extern "C" __attribute__((visibility("default"))) int tvs_im2col_patches(const float* image, int32_t B, int32_t C, int32_t H, int32_t W, int32_t P, void* out_bf16, int32_t out_f16, void* stream) {
    TVS_REQUIRE(image && out_bf16, "tvs_im2col_patches: null pointer");
    TVS_REQUIRE(P % 4 == 0 && H % P == 0 && W % P == 0, "tvs_im2col_patches: P must divide H, W and be a multiple of 4");
    const long long n4 = static_cast<long long>(B) * C * H * W / 4;
    im2col_kernel<<<blocks_for(n4, 256, 148 * 16), 256, 0, static_cast<cudaStream_t>(stream)>>>(image, B, C, H, W, P, static_cast<__nv_bfloat16*>(out_bf16), out_f16);
    return check_launch("im2col_kernel");
}

extern "C" __attribute__((visibility("default"))) int tvs_vision_assemble(const float* patches, const float* cls, const float* pos, const float* ctx, int64_t ctx_batch_stride,
                                   int32_t B, int32_t G2, int32_t n, int32_t D, float* h, void* stream) {
    TVS_REQUIRE(patches && cls && pos && h && (n == 0 || ctx), "tvs_vision_assemble: null pointer");
    TVS_REQUIRE(D % 4 == 0, "tvs_vision_assemble: D %% 4");
    const long long total = static_cast<long long>(B) * (1 + G2 + n) * (D / 4);
    vision_assemble_kernel<<<blocks_for(total, 256, 148 * 16), 256, 0, static_cast<cudaStream_t>(stream)>>>(patches, cls, pos, ctx, ctx_batch_stride, B, G2, n, D, h);
    return check_launch("vision_assemble_kernel");
}

extern "C" __attribute__((visibility("default"))) int tvs_prompt_overwrite(float* x, void* x_bf16, int32_t B, int32_t S, int32_t D, int32_t row0, int32_t n, const float* ctx,
                                    int64_t ctx_batch_stride, void* stream) {
    TVS_REQUIRE(x && ctx, "tvs_prompt_overwrite: null pointer");
    TVS_REQUIRE(D % 4 == 0 && row0 >= 0 && n > 0 && row0 + n <= S, "tvs_prompt_overwrite: bad rows row0=%d n=%d S=%d", row0, n, S);
    const long long total = static_cast<long long>(B) * n * (D / 4);
    prompt_overwrite_kernel<<<blocks_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, static_cast<__nv_bfloat16*>(x_bf16), B, S, D, row0, n, ctx,
                                                                                                  ctx_batch_stride);
    return check_launch("prompt_overwrite_kernel");
}

extern "C" __attribute__((visibility("default"))) int tvs_prompt_grad(float* dx, void* dx_bf16, int32_t B, int32_t S, int32_t D, int32_t row0, int32_t n, float* dctx,
                               int64_t ctx_batch_stride, int32_t zero_rows, void* stream) {
    TVS_REQUIRE(dx && dctx, "tvs_prompt_grad: null pointer");
    TVS_REQUIRE(row0 >= 0 && n > 0 && row0 + n <= S, "tvs_prompt_grad: bad rows");
    if (ctx_batch_stride == 0) {
        prompt_grad_reduce_kernel<<<dim3(n, (D + 127) / 128), dim3(128, 8), 0, static_cast<cudaStream_t>(stream)>>>(
            dx, static_cast<__nv_bfloat16*>(dx_bf16), B, S, D, row0, dctx, zero_rows);
        return check_launch("prompt_grad_reduce_kernel");
    }
    const long long total = static_cast<long long>(n) * D * B;
    prompt_grad_kernel<<<blocks_for(total, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(dx, static_cast<__nv_bfloat16*>(dx_bf16), B, S, D, row0, n, dctx,
                                                                                            ctx_batch_stride, zero_rows);
    return check_launch("prompt_grad_kernel");
}

extern "C" __attribute__((visibility("default"))) int tvs_cast_bf16(const float* x, void* y_bf16, int64_t n, void* stream) {
    TVS_REQUIRE(x && y_bf16 && n % 4 == 0, "tvs_cast_bf16: bad arguments");
    cast_bf16_kernel<<<blocks_for(n / 4, 256, 148 * 16), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, static_cast<__nv_bfloat16*>(y_bf16), n / 4);
    return check_launch("cast_bf16_kernel");
}

extern "C" __attribute__((visibility("default"))) int tvs_add_f32(float* y, const float* x, int64_t n, void* stream) {
    TVS_REQUIRE(x && y && n % 4 == 0, "tvs_add_f32: bad arguments");
    add_f32_kernel<<<blocks_for(n / 4, 256, 148 * 16), 256, 0, static_cast<cudaStream_t>(stream)>>>(y, x, n / 4);
    return check_launch("add_f32_kernel");
}

extern "C" __attribute__((visibility("default"))) int tvs_film_fwd(const float* x, const float* mul, const float* add, int32_t B, int32_t S, int32_t D, float* y, void* y_bf16,
                            void* stream) {
    TVS_REQUIRE(x && mul && add && (y || y_bf16) && D % 4 == 0, "tvs_film_fwd: bad arguments");
    const long long total = static_cast<long long>(B) * S * (D / 4);
    film_fwd_kernel<<<blocks_for(total, 256, 148 * 16), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, mul, add, B, S, D, y, static_cast<__nv_bfloat16*>(y_bf16));
    return check_launch("film_fwd_kernel");
}

extern "C" __attribute__((visibility("default"))) int tvs_film_bwd(const float* dy, const float* x, const float* mul, int32_t B, int32_t S, int32_t D, float* dx, float* dmul,
                            float* dadd, void* stream) {
    TVS_REQUIRE(dy && x && mul && dx && dmul && dadd, "tvs_film_bwd: null pointer");
    film_bwd_kernel<<<B, 1024, 0, static_cast<cudaStream_t>(stream)>>>(dy, x, mul, S, D, dx, dmul, dadd);     // one block per sample: 1024 threads = 16 row groups at D = 64
    return check_launch("film_bwd_kernel");
}

extern "C" __attribute__((visibility("default"))) int tvs_head_fwd(const float* tconv, int64_t ld_tconv, const float* addmap, int64_t ld_addmap, const float* bias_t,
                            const float* bias_a, const float* ratio, int32_t blend, int32_t B, int32_t G, int32_t P, int32_t ksize,
                            float* logits, float* add_out, void* stream) {
    TVS_REQUIRE(tconv && logits, "tvs_head_fwd: null pointer");
    TVS_REQUIRE(blend >= 0 && blend <= 2, "tvs_head_fwd: blend must be 0, 1 or 2");
    TVS_REQUIRE(blend == 0 || (addmap && ksize >= 1 && ksize <= HEAD_MAXK && (ksize & 1)), "tvs_head_fwd: additive branch needs addmap and odd ksize <= %d", HEAD_MAXK);
    TVS_REQUIRE(blend != 1 || ratio, "tvs_head_fwd: ratio required for blend=1");
    TVS_REQUIRE(G * P <= HEAD_THREADS, "tvs_head_fwd: image width %d exceeds the %d columns one block covers", G * P, HEAD_THREADS);
    const size_t sh = blend ? (static_cast<size_t>(3) * G * ksize * ksize + static_cast<size_t>(P) * G * ksize + 3 * P * ksize) * sizeof(float) : 0;
    TVS_REQUIRE(sh <= 48 * 1024, "tvs_head_fwd: grid too large for the shared-memory neighbourhood");
    const int threads = max(128, (G * P + 31) / 32 * 32);
    head_fwd_kernel<<<dim3(G, B), threads, sh, static_cast<cudaStream_t>(stream)>>>(tconv, ld_tconv, addmap, ld_addmap, bias_t, bias_a, ratio, blend, G, P,
                                                                                   ksize, logits, add_out);
    return check_launch("head_fwd_kernel");
}

extern "C" __attribute__((visibility("default"))) int tvs_head_bwd(const float* dlogits, const float* tconv, int64_t ld_tconv, const float* add_out, const float* bias_t,
                            const float* ratio, int32_t blend, int32_t B, int32_t G, int32_t P, int32_t ksize, void* dtconv_bf16,
                            int64_t ld_dtconv, float* daddmap, int64_t ld_daddmap, float* dbias_a, float* dratio, void* stream) {
    TVS_REQUIRE(dlogits && dtconv_bf16, "tvs_head_bwd: null pointer");
    TVS_REQUIRE(blend >= 0 && blend <= 2, "tvs_head_bwd: blend must be 0, 1 or 2");
    TVS_REQUIRE(blend != 1 || (ratio && tconv && add_out), "tvs_head_bwd: blend=1 needs ratio, tconv and add_out");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    head_bwd_pix_kernel<<<dim3(G, B), 256, 0, st>>>(dlogits, tconv, ld_tconv, add_out, bias_t, ratio, blend, G, P,
                                                    static_cast<__nv_bfloat16*>(dtconv_bf16), ld_dtconv, dbias_a, dratio);
    if (int rc = check_launch("head_bwd_pix_kernel")) return rc;
    if (blend != 0 && daddmap) {
        TVS_REQUIRE(ksize >= 1 && ksize <= HEAD_MAXK && (ksize & 1), "tvs_head_bwd: odd ksize <= %d", HEAD_MAXK);
        const int nr = 2 * P + ksize + 2;
        const size_t sh = (static_cast<size_t>(nr) * G * ksize + static_cast<size_t>(G) * (2 * P + 2) + static_cast<size_t>(HB_CHUNK) * (G * P + G)) * sizeof(float);
        TVS_REQUIRE(sh <= 48 * 1024, "tvs_head_bwd: shared-memory tile too large");
        head_bwd_addmap_kernel<<<dim3(G, B), 256, sh, st>>>(dlogits, ratio, blend, G, P, ksize, daddmap, ld_daddmap);
        return check_launch("head_bwd_addmap_kernel");
    }
    return 0;
}

extern "C" __attribute__((visibility("default"))) int tvs_adamw_flat(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                              float beta2, float eps, float weight_decay, int32_t step, float grad_scale, const int32_t* step_dev,
                              const float* lr_dev, void* stream) {
    TVS_REQUIRE(param && grad && exp_avg && exp_avg_sq && n > 0, "tvs_adamw_flat: bad arguments");
    adamw_kernel<<<blocks_for(n, 256, 148 * 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps,
                                                                                           weight_decay, step, grad_scale, step_dev, lr_dev);
    return check_launch("adamw_kernel");
}

extern "C" __attribute__((visibility("default"))) int tvs_counter_inc(int32_t* counter_dev, void* stream) {
    TVS_REQUIRE(counter_dev, "tvs_counter_inc: null pointer");
    counter_inc_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(counter_dev);
    return check_launch("counter_inc_kernel");
}

extern "C" __attribute__((visibility("default"))) int tvs_slice_rows(const float* x, int32_t B, int32_t S, int32_t D, int32_t row0, int32_t nrows, float* y_f32, void* y_bf16,
                                                                      void* stream) {
    TVS_REQUIRE(x && (y_f32 || y_bf16) && D % 4 == 0 && row0 >= 0 && nrows > 0 && row0 + nrows <= S, "tvs_slice_rows: bad arguments");
    const long long total = static_cast<long long>(B) * nrows * (D / 4);
    slice_rows_kernel<<<blocks_for(total, 256, 148 * 16), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, B, S, D, row0, nrows, y_f32, static_cast<__nv_bfloat16*>(y_bf16));
    return check_launch("slice_rows_kernel");
}

extern "C" __attribute__((visibility("default"))) int tvs_unslice_rows(const float* dy, int32_t B, int32_t S, int32_t D, int32_t row0, int32_t nrows, float* dx, void* stream) {
    TVS_REQUIRE(dy && dx && D % 4 == 0 && row0 >= 0 && nrows > 0 && row0 + nrows <= S, "tvs_unslice_rows: bad arguments");
    const long long total = static_cast<long long>(B) * S * (D / 4);
    unslice_rows_kernel<<<blocks_for(total, 256, 148 * 16), 256, 0, static_cast<cudaStream_t>(stream)>>>(dy, B, S, D, row0, nrows, dx);
    return check_launch("unslice_rows_kernel");
}

extern "C" __attribute__((visibility("default"))) int tvs_wgrad_small(const float* dy, int64_t ld_dy, const float* x, int64_t ld_x, int64_t M, int32_t N, int32_t K, float* dw,
                                                                       void* stream) {
    TVS_REQUIRE(dy && x && dw && M > 0 && N > 0 && K > 0 && N * K <= 4096, "tvs_wgrad_small: N*K must be <= 4096");
    const size_t sh = static_cast<size_t>(WG_ROWS) * (N + K) * sizeof(float);
    TVS_REQUIRE(sh <= 48 * 1024, "tvs_wgrad_small: N + K too large");
    long long blocks = (M + WG_ROWS - 1) / WG_ROWS;
    if (blocks > 2 * sm_count()) blocks = 2 * sm_count();
    wgrad_small_kernel<<<static_cast<unsigned>(blocks), 256, sh, static_cast<cudaStream_t>(stream)>>>(dy, ld_dy, x, ld_x, M, N, K, dw);
    return check_launch("wgrad_small_kernel");
}
