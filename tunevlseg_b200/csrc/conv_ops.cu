// Channels-last (NHWC) convolution plumbing and the small CRIS-specific kernels.
//
// CRIS (src/models/components/cris_model) is convolutional: a CLIP-RN50 image encoder, an FPN neck, a 3-layer
// vision-language decoder and a dynamic-convolution projector.  Every activation lives in HBM as a row-major
// [B*H*W, C] matrix, so a 1x1 convolution IS a GEMM on the tcgen05 kernel (gemm_sm100.cu) and a k x k convolution is
// an im2col gather followed by that GEMM (BatchNorm folded into the weights, ReLU / residual in the GEMM epilogue).
// The backbone is frozen, so backward is dgrad only: dcol = dy * W (GEMM), dx = col2im(dcol).
//
// All kernels here are HBM-bound gathers / stencils: vectorised 16-byte accesses along the channel dimension, grids
// sized to a multiple of the SM count.
#include <algorithm>

#include <cooperative_groups.h>

#include "common.cuh"
#include "tvs_b200.h"

namespace cg = cooperative_groups;

namespace tvs {

static inline unsigned grid_for(long long total, int threads, int cap_blocks = 148 * 32) {
    long long b = (total + threads - 1) / threads;
    if (b < 1) b = 1;
    if (b > cap_blocks) b = cap_blocks;
    return static_cast<unsigned>(b);
}

// ---------------------------------------------------------------------------------------------------------------
// im2col: col[(b,oy,ox), (ky,kx,c)] = x[b, oy*s - p + ky, ox*s - p + kx, c]  (0 outside), tail columns zero.
// Elements are moved as opaque V-sized vectors (V = 16/8/4/2 bytes), so one kernel serves bf16 and f32 operands.
// ---------------------------------------------------------------------------------------------------------------
// kind::tf32 MMAs TRUNCATE their fp32 operands to 10 mantissa bits (measured: a systematic -3.4e-4 relative bias per
// GEMM, which compounds through ~60 layers); operands are therefore rounded to nearest on their way into a GEMM.
__device__ __forceinline__ uint32_t rn_tf32(uint32_t bits) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(__uint_as_float(bits)));
    return r;
}
__device__ __forceinline__ void round_words(uint4& v) { v.x = rn_tf32(v.x); v.y = rn_tf32(v.y); v.z = rn_tf32(v.z); v.w = rn_tf32(v.w); }
__device__ __forceinline__ void round_words(uint2& v) { v.x = rn_tf32(v.x); v.y = rn_tf32(v.y); }
__device__ __forceinline__ void round_words(uint32_t& v) { v = rn_tf32(v); }
__device__ __forceinline__ void round_words(uint16_t&) {}

// Thread layout: 2^tpr_shift threads walk one output row (consecutive lanes -> consecutive 16-byte vectors, so both the
// tap reads and the row writes are coalesced); the (tap, channel) position advances incrementally - no division in the
// inner loop (the first version divided five times per vector and was instruction-bound at 1/3 of HBM speed).
template <typename V, bool ROUND>
__global__ void __launch_bounds__(256) im2col_kernel(const V* __restrict__ x, V* __restrict__ col, int H, int W, int Cv, int Ho, int Wo, int ks,
                                                     int stride, int pad, int ldcol_v, int Kv, long long rows, int tpr_shift) {
    const int tpr = 1 << tpr_shift, rpb = blockDim.x >> tpr_shift;
    const int lane = threadIdx.x & (tpr - 1), rl = threadIdx.x >> tpr_shift;
    const int hw = Ho * Wo;
    const int tap0 = lane / Cv, cv0 = lane - tap0 * Cv, ky0 = tap0 / ks, kx0 = tap0 - ky0 * ks;
    for (long long row = static_cast<long long>(blockIdx.x) * rpb + rl; row < rows; row += static_cast<long long>(gridDim.x) * rpb) {
        const long long b = row / hw;
        const int r = static_cast<int>(row - b * hw);
        const int oy = r / Wo, ox = r - oy * Wo;
        const int iy0 = oy * stride - pad, ix0 = ox * stride - pad;
        const V* xb = x + b * H * W * Cv;
        V* crow = col + row * ldcol_v;
        int cv = cv0, ky = ky0, kx = kx0;
        for (int kv = lane; kv < ldcol_v; kv += tpr) {
            V v;
            memset(&v, 0, sizeof(V));
            if (kv < Kv) {
                const int iy = iy0 + ky, ix = ix0 + kx;
                if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = xb[(static_cast<long long>(iy) * W + ix) * Cv + cv];
                if (ROUND) round_words(v);
            }
            crow[kv] = v;
            cv += tpr;
            while (cv >= Cv) {
                cv -= Cv;
                if (++kx == ks) { kx = 0; ++ky; }
            }
        }
    }
}

// zero-bordered copy [B,H,W,C] -> [B,H+2,W+2,C] (the A operand of the implicit-GEMM 3x3 convolution), 16-byte vectors
template <bool ROUND>
__global__ void __launch_bounds__(256) pad_nhwc_kernel(const uint4* __restrict__ x, uint4* __restrict__ xp, int H, int W, int Cv, long long total) {
    const int Wp = W + 2, Hp = H + 2;
    for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long pix = idx / Cv;
        const int cv = static_cast<int>(idx - pix * Cv);
        const long long b = pix / (Hp * Wp);
        const int rem = static_cast<int>(pix - b * (Hp * Wp));
        const int yp = rem / Wp, xq = rem - yp * Wp;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (yp >= 1 && yp <= H && xq >= 1 && xq <= W) {
            v = x[((b * H + (yp - 1)) * W + (xq - 1)) * Cv + cv];
            if (ROUND) round_words(v);
        }
        xp[idx] = v;
    }
}

// col2im (stride 1, pad = ks/2): dx[b,y,x,c] = sum_{ky,kx} dcol[(b, y-ky+p, x-kx+p), (ky,kx,c)], c < Cx <= Ccol,
// optionally multiplied by the ReLU mask of the layer below (ymask > 0).  4 channels per thread.
__global__ void __launch_bounds__(256) col2im_kernel(const float* __restrict__ dcol, long long ldcol, int H, int W, int Ccol, int Cx, int ks,
                                                     const float* __restrict__ ymask, long long ld_mask, float* __restrict__ dx,
                                                     long long ld_dx, long long total) {
    const int c4n = Cx >> 2, pad = ks >> 1;
    for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long pix = idx / c4n;
        const int c = static_cast<int>(idx - pix * c4n) * 4;
        const int hw = H * W;
        const long long b = pix / hw;
        const int r = static_cast<int>(pix - b * hw);
        const int y = r / W, x = r - y * W;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int ky = 0; ky < ks; ++ky) {
            const int oy = y - ky + pad;
            if (oy < 0 || oy >= H) continue;
            for (int kx = 0; kx < ks; ++kx) {
                const int ox = x - kx + pad;
                if (ox < 0 || ox >= W) continue;
                const float* p = dcol + ((b * H + oy) * W + ox) * ldcol + static_cast<long long>(ky * ks + kx) * Ccol + c;
                // Ccol may be odd-sized (coordconv: 514), so the tap offset is only 8-byte aligned in general
                const float2 lo = *reinterpret_cast<const float2*>(p), hi = *reinterpret_cast<const float2*>(p + 2);
                acc.x += lo.x; acc.y += lo.y; acc.z += hi.x; acc.w += hi.y;
            }
        }
        if (ymask) {
            const float4 m = *reinterpret_cast<const float4*>(ymask + pix * ld_mask + c);
            acc.x = m.x > 0.f ? acc.x : 0.f; acc.y = m.y > 0.f ? acc.y : 0.f;
            acc.z = m.z > 0.f ? acc.z : 0.f; acc.w = m.w > 0.f ? acc.w : 0.f;
        }
        *reinterpret_cast<float4*>(dx + pix * ld_dx + c) = acc;
    }
}

// y = round-to-nearest tf32 of x, [M, C] views with row strides
__global__ void __launch_bounds__(256) round_tf32_kernel(const float* __restrict__ x, long long ld_x, float* __restrict__ y, long long ld_y, int C4,
                                                         long long total) {
    for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long m = idx / C4;
        const int c = static_cast<int>(idx - m * C4) * 4;
        uint4 v = *reinterpret_cast<const uint4*>(x + m * ld_x + c);
        round_words(v);
        *reinterpret_cast<uint4*>(y + m * ld_y + c) = v;
    }
}

// out[m, c] = y[m, c] > 0 ? dy[m, c] : 0   (row strides allow column slices of concatenated buffers)
__global__ void __launch_bounds__(256) relu_mask_kernel(const float* __restrict__ dy, long long ld_dy, const float* __restrict__ y, long long ld_y,
                                                        float* __restrict__ out, long long ld_out, int C4, long long total) {
    for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long m = idx / C4;
        const int c = static_cast<int>(idx - m * C4) * 4;
        const float4 g = *reinterpret_cast<const float4*>(dy + m * ld_dy + c);
        const float4 a = *reinterpret_cast<const float4*>(y + m * ld_y + c);
        *reinterpret_cast<float4*>(out + m * ld_out + c) =
            make_float4(a.x > 0.f ? g.x : 0.f, a.y > 0.f ? g.y : 0.f, a.z > 0.f ? g.z : 0.f, a.w > 0.f ? g.w : 0.f);
    }
}

// 2x2 / stride-2 average pooling, NHWC.  T = float (4 channels / thread) or bf16 (8 channels / thread).
template <bool BF16>
__global__ void __launch_bounds__(256) avgpool2_kernel(const void* __restrict__ xin, void* __restrict__ yout, int H, int W, int C, long long ld_out,
                                                       long long total, int round_out) {
    constexpr int VEC = BF16 ? 8 : 4;
    const int cvn = C / VEC, Ho = H >> 1, Wo = W >> 1;
    for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long opix = idx / cvn;
        const int c = static_cast<int>(idx - opix * cvn) * VEC;
        const long long b = opix / (Ho * Wo);
        const int r = static_cast<int>(opix - b * (Ho * Wo));
        const int oy = r / Wo, ox = r - oy * Wo;
        float acc[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc[i] = 0.f;
#pragma unroll
        for (int dy = 0; dy < 2; ++dy)
#pragma unroll
            for (int dx = 0; dx < 2; ++dx) {
                const long long ip = (b * H + 2 * oy + dy) * W + 2 * ox + dx;
                if (BF16) {
                    const uint4 q = *reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(xin) + ip * C + c);
                    const float2 a = unpack_bf16x2(q.x), bb = unpack_bf16x2(q.y), cc = unpack_bf16x2(q.z), d = unpack_bf16x2(q.w);
                    acc[0] += a.x; acc[1] += a.y; acc[2] += bb.x; acc[3] += bb.y;
                    acc[4 % VEC] += cc.x; acc[5 % VEC] += cc.y; acc[6 % VEC] += d.x; acc[7 % VEC] += d.y;
                } else {
                    const float4 q = *reinterpret_cast<const float4*>(static_cast<const float*>(xin) + ip * C + c);
                    acc[0] += q.x; acc[1] += q.y; acc[2] += q.z; acc[3] += q.w;
                }
            }
        if (BF16) {
            uint4 o;
            o.x = pack_bf16x2(acc[0] * 0.25f, acc[1] * 0.25f); o.y = pack_bf16x2(acc[2] * 0.25f, acc[3] * 0.25f);
            o.z = pack_bf16x2(acc[4 % VEC] * 0.25f, acc[5 % VEC] * 0.25f); o.w = pack_bf16x2(acc[6 % VEC] * 0.25f, acc[7 % VEC] * 0.25f);
            *reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(yout) + opix * ld_out + c) = o;
        } else {
            uint4 o = make_uint4(__float_as_uint(acc[0] * 0.25f), __float_as_uint(acc[1] * 0.25f), __float_as_uint(acc[2] * 0.25f),
                                 __float_as_uint(acc[3] * 0.25f));
            if (round_out) round_words(o);
            *reinterpret_cast<uint4*>(static_cast<float*>(yout) + opix * ld_out + c) = o;
        }
    }
}

// bilinear x2 upsampling, align_corners=False (torch: src = max((o + 0.5) / 2 - 0.5, 0), lo = floor, hi = min(lo+1, n-1))
__device__ __forceinline__ void up2_taps(int o, int n, int& lo, int& hi, float& whi) {
    float src = (o + 0.5f) * 0.5f - 0.5f;
    src = src < 0.f ? 0.f : src;
    lo = static_cast<int>(src);
    hi = lo + (lo < n - 1 ? 1 : 0);
    whi = src - lo;
}

__global__ void __launch_bounds__(256) upsample2x_fwd_kernel(const float* __restrict__ x, int H, int W, int C, float* __restrict__ y, long long ld_out,
                                                             long long total) {
    const int c4n = C >> 2, Ho = 2 * H, Wo = 2 * W;
    for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long opix = idx / c4n;
        const int c = static_cast<int>(idx - opix * c4n) * 4;
        const long long b = opix / (Ho * Wo);
        const int r = static_cast<int>(opix - b * (Ho * Wo));
        const int oy = r / Wo, ox = r - oy * Wo;
        int y0, y1, x0, x1;
        float wy, wx;
        up2_taps(oy, H, y0, y1, wy);
        up2_taps(ox, W, x0, x1, wx);
        const float* base = x + b * H * W * C + c;
        const float4 a = *reinterpret_cast<const float4*>(base + (static_cast<long long>(y0) * W + x0) * C);
        const float4 bq = *reinterpret_cast<const float4*>(base + (static_cast<long long>(y0) * W + x1) * C);
        const float4 cq = *reinterpret_cast<const float4*>(base + (static_cast<long long>(y1) * W + x0) * C);
        const float4 d = *reinterpret_cast<const float4*>(base + (static_cast<long long>(y1) * W + x1) * C);
        const float w00 = (1.f - wy) * (1.f - wx), w01 = (1.f - wy) * wx, w10 = wy * (1.f - wx), w11 = wy * wx;
        float4 o;
        o.x = w00 * a.x + w01 * bq.x + w10 * cq.x + w11 * d.x;
        o.y = w00 * a.y + w01 * bq.y + w10 * cq.y + w11 * d.y;
        o.z = w00 * a.z + w01 * bq.z + w10 * cq.z + w11 * d.z;
        o.w = w00 * a.w + w01 * bq.w + w10 * cq.w + w11 * d.w;
        *reinterpret_cast<float4*>(y + opix * ld_out + c) = o;
    }
}

// backward as a gather: every input pixel collects from the <= 4 x 4 outputs whose taps (after clamping) land on it
__global__ void __launch_bounds__(256) upsample2x_bwd_kernel(const float* __restrict__ dy, long long ld_dy, int H, int W, int C, float* __restrict__ dx,
                                                             long long total) {
    const int c4n = C >> 2, Ho = 2 * H, Wo = 2 * W;
    for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long ipix = idx / c4n;
        const int c = static_cast<int>(idx - ipix * c4n) * 4;
        const long long b = ipix / (H * W);
        const int r = static_cast<int>(ipix - b * (H * W));
        const int iy = r / W, ix = r - iy * W;
        float wys[4], wxs[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            int lo, hi;
            float wh;
            const int oy = 2 * iy - 1 + t;
            wys[t] = 0.f;
            if (oy >= 0 && oy < Ho) {
                up2_taps(oy, H, lo, hi, wh);
                wys[t] = (lo == iy ? 1.f - wh : 0.f) + (hi == iy ? wh : 0.f);
            }
            const int ox = 2 * ix - 1 + t;
            wxs[t] = 0.f;
            if (ox >= 0 && ox < Wo) {
                up2_taps(ox, W, lo, hi, wh);
                wxs[t] = (lo == ix ? 1.f - wh : 0.f) + (hi == ix ? wh : 0.f);
            }
        }
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int ty = 0; ty < 4; ++ty) {
            if (wys[ty] == 0.f) continue;
            const int oy = 2 * iy - 1 + ty;
#pragma unroll
            for (int tx = 0; tx < 4; ++tx) {
                if (wxs[tx] == 0.f) continue;
                const int ox = 2 * ix - 1 + tx;
                const float w = wys[ty] * wxs[tx];
                const float4 g = *reinterpret_cast<const float4*>(dy + ((b * Ho + oy) * Wo + ox) * ld_dy + c);
                acc.x += w * g.x; acc.y += w * g.y; acc.z += w * g.z; acc.w += w * g.w;
            }
        }
        *reinterpret_cast<float4*>(dx + ipix * C + c) = acc;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Cross attention of the CRIS decoder (layers.py:341-349): Sq = 676 vision queries against Sk <= 77 word keys with a
// key-padding mask, head dim 64, fp32.  The key/value set of one (batch, head) fits in shared memory, one thread
// owns one query row (q and the output accumulator live in registers), keys are broadcast reads.
// ---------------------------------------------------------------------------------------------------------------
constexpr int XA_HD = 64;
constexpr int XA_MAXK = 80;
constexpr int XA_THREADS = 128;

__device__ __forceinline__ void xa_load_kv(const float* __restrict__ k, const float* __restrict__ v, long long ld_kv, int Sk, int b, int h,
                                           float (*sk)[XA_HD], float (*sv)[XA_HD]) {
    for (int i = threadIdx.x; i < Sk * (XA_HD / 4); i += blockDim.x) {
        const int j = i / (XA_HD / 4), d = (i - j * (XA_HD / 4)) * 4;
        const long long off = (static_cast<long long>(b) * Sk + j) * ld_kv + h * XA_HD + d;
        *reinterpret_cast<float4*>(&sk[j][d]) = *reinterpret_cast<const float4*>(k + off);
        *reinterpret_cast<float4*>(&sv[j][d]) = *reinterpret_cast<const float4*>(v + off);
    }
}

// Four lanes (a quad) share one query row, 16 of the 64 head dims each: 48 live registers instead of 192 (the
// one-thread-per-row version spilled), dot products finish with two quad shuffles.
__device__ __forceinline__ float quad_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    return v;
}
constexpr int XA_QPB = XA_THREADS / 4;      // query rows per block

__global__ void __launch_bounds__(XA_THREADS) cross_attn_fwd_kernel(const float* __restrict__ q, long long ld_q, const float* __restrict__ k,
                                                                    const float* __restrict__ v, long long ld_kv, const uint8_t* __restrict__ key_mask,
                                                                    int Sq, int Sk, int causal, float* __restrict__ out, long long ld_o,
                                                                    float* __restrict__ lse) {
    __shared__ __align__(16) float sk[XA_MAXK][XA_HD];
    __shared__ __align__(16) float sv[XA_MAXK][XA_HD];
    __shared__ uint8_t sm[XA_MAXK];
    const int b = blockIdx.z, h = blockIdx.y, H = gridDim.y;
    xa_load_kv(k, v, ld_kv, Sk, b, h, sk, sv);
    for (int j = threadIdx.x; j < Sk; j += blockDim.x) sm[j] = key_mask ? key_mask[b * Sk + j] : 1;
    __syncthreads();
    const int part = threadIdx.x & 3;
    const int i = blockIdx.x * XA_QPB + (threadIdx.x >> 2);
    const bool live = i < Sq;                      // dead quads still take part in the shuffles
    const long long row = static_cast<long long>(b) * Sq + (live ? i : 0);
    float qr[16], acc[16];
#pragma unroll
    for (int d = 0; d < 16; d += 4) {
        const float4 t = *reinterpret_cast<const float4*>(q + row * ld_q + h * XA_HD + part * 16 + d);
        qr[d] = t.x; qr[d + 1] = t.y; qr[d + 2] = t.z; qr[d + 3] = t.w;
        acc[d] = acc[d + 1] = acc[d + 2] = acc[d + 3] = 0.f;
    }
    float mx = -INFINITY, sum = 0.f;
    const int jend = causal ? min(Sk, i + 1) : Sk;
    for (int j = 0; j < Sk; ++j) {
        float s = 0.f;
#pragma unroll
        for (int d = 0; d < 16; ++d) s = fmaf(qr[d], sk[j][part * 16 + d], s);
        s = quad_sum(s);
        if (j >= jend || !sm[j]) continue;          // uniform within the quad
        const float nm = fmaxf(mx, s);
        const float corr = __expf(mx - nm), p = __expf(s - nm);
        sum = sum * corr + p;
#pragma unroll
        for (int d = 0; d < 16; ++d) acc[d] = fmaf(acc[d], corr, p * sv[j][part * 16 + d]);
        mx = nm;
    }
    if (!live) return;
    const float inv = 1.f / sum;
#pragma unroll
    for (int d = 0; d < 16; d += 4)
        *reinterpret_cast<float4*>(out + row * ld_o + h * XA_HD + part * 16 + d) =
            make_float4(round_tf32_rn(acc[d] * inv), round_tf32_rn(acc[d + 1] * inv), round_tf32_rn(acc[d + 2] * inv),
                        round_tf32_rn(acc[d + 3] * inv));                  // the output only feeds the out-proj tf32 GEMM
    if (part == 0) lse[(static_cast<long long>(b) * H + h) * Sq + i] = mx + __logf(sum);
}

// dq[i] = sum_j ds_ij k_j,  ds_ij = p_ij (dO_i . v_j - delta_i),  delta_i = dO_i . O_i ; also stores delta
__global__ void __launch_bounds__(XA_THREADS) cross_attn_bwd_dq_kernel(const float* __restrict__ q, long long ld_q, const float* __restrict__ k,
                                                                       const float* __restrict__ v, long long ld_kv,
                                                                       const uint8_t* __restrict__ key_mask, const float* __restrict__ o,
                                                                       const float* __restrict__ dO, long long ld_o, const float* __restrict__ lse,
                                                                       int Sq, int Sk, int causal, float* __restrict__ dq, long long ld_dq,
                                                                       float* __restrict__ delta) {
    __shared__ __align__(16) float sk[XA_MAXK][XA_HD];
    __shared__ __align__(16) float sv[XA_MAXK][XA_HD];
    __shared__ uint8_t sm[XA_MAXK];
    const int b = blockIdx.z, h = blockIdx.y, H = gridDim.y;
    xa_load_kv(k, v, ld_kv, Sk, b, h, sk, sv);
    for (int j = threadIdx.x; j < Sk; j += blockDim.x) sm[j] = key_mask ? key_mask[b * Sk + j] : 1;
    __syncthreads();
    const int part = threadIdx.x & 3;
    const int i = blockIdx.x * XA_QPB + (threadIdx.x >> 2);
    const bool live = i < Sq;
    const long long row = static_cast<long long>(b) * Sq + (live ? i : 0);
    float qr[16], gr[16], acc[16];
    float dl = 0.f;
#pragma unroll
    for (int d = 0; d < 16; d += 4) {
        const long long off = h * XA_HD + part * 16 + d;
        const float4 t = *reinterpret_cast<const float4*>(q + row * ld_q + off);
        const float4 g = *reinterpret_cast<const float4*>(dO + row * ld_o + off);
        const float4 oo = *reinterpret_cast<const float4*>(o + row * ld_o + off);
        qr[d] = t.x; qr[d + 1] = t.y; qr[d + 2] = t.z; qr[d + 3] = t.w;
        gr[d] = g.x; gr[d + 1] = g.y; gr[d + 2] = g.z; gr[d + 3] = g.w;
        dl += g.x * oo.x + g.y * oo.y + g.z * oo.z + g.w * oo.w;
        acc[d] = acc[d + 1] = acc[d + 2] = acc[d + 3] = 0.f;
    }
    dl = quad_sum(dl);
    const long long li = (static_cast<long long>(b) * H + h) * Sq + (live ? i : 0);
    const float l = lse[li];
    const int jend = causal ? min(Sk, i + 1) : Sk;
    for (int j = 0; j < Sk; ++j) {
        float s = 0.f, dp = 0.f;
#pragma unroll
        for (int d = 0; d < 16; ++d) {
            s = fmaf(qr[d], sk[j][part * 16 + d], s);
            dp = fmaf(gr[d], sv[j][part * 16 + d], dp);
        }
        s = quad_sum(s);
        dp = quad_sum(dp);
        if (j >= jend || !sm[j]) continue;
        const float ds = __expf(s - l) * (dp - dl);
#pragma unroll
        for (int d = 0; d < 16; ++d) acc[d] = fmaf(ds, sk[j][part * 16 + d], acc[d]);
    }
    if (!live) return;
#pragma unroll
    for (int d = 0; d < 16; d += 4)
        *reinterpret_cast<float4*>(dq + row * ld_dq + h * XA_HD + part * 16 + d) = make_float4(acc[d], acc[d + 1], acc[d + 2], acc[d + 3]);
    if (part == 0) delta[li] = dl;
}

// dk_j = sum_i ds_ij q_i ; dv_j = sum_i p_ij dO_i.  Thread (j, part) owns 16 of the 64 dims of key j; q / dO rows are staged
// through shared memory 32 queries at a time.  The queries of one (b, h) are split over the CTAs of a thread-block cluster
// (grid.x = cluster size <= 8): with one CTA of 4 Sk threads per (b, h) the CRIS decoder's cross-attention (676 queries, 12
// words) ran 256 blocks of 48 threads for 430 us.  The partial sums meet in CTA 0 through distributed shared memory in
// rank order - deterministic, no scratch buffer, no atomics.
constexpr int XA_QT = 32;
__global__ void __launch_bounds__(XA_MAXK * 4) cross_attn_bwd_dkv_kernel(const float* __restrict__ q, long long ld_q, const float* __restrict__ k,
                                                                         const float* __restrict__ v, long long ld_kv,
                                                                         const uint8_t* __restrict__ key_mask, const float* __restrict__ dO,
                                                                         long long ld_o, const float* __restrict__ lse,
                                                                         const float* __restrict__ delta, int Sq, int Sk, int causal,
                                                                         float* __restrict__ dk, float* __restrict__ dv, long long ld_dkv) {
    __shared__ __align__(16) float sq[XA_QT][XA_HD];
    __shared__ __align__(16) float sg[XA_QT][XA_HD];
    __shared__ float sl[XA_QT], sd[XA_QT];
    __shared__ __align__(16) float sred[XA_MAXK * 4][16];
    pdl_wait();
    pdl_trigger();
    const int split = blockIdx.x, nsplit = gridDim.x;
    const int b = blockIdx.z, h = blockIdx.y, H = gridDim.y;
    const int chunk = ((Sq + nsplit - 1) / nsplit + XA_QT - 1) / XA_QT * XA_QT;
    const int q_begin = split * chunk, q_end = min(Sq, q_begin + chunk);
    const int j = threadIdx.x >> 2, part = threadIdx.x & 3;
    const bool live = j < Sk;
    const bool attend = live && (!key_mask || key_mask[b * Sk + j]);
    float kr[16], vr[16], ak[16], av[16];
#pragma unroll
    for (int d = 0; d < 16; ++d) {
        const long long off = (static_cast<long long>(b) * Sk + (live ? j : 0)) * ld_kv + h * XA_HD + part * 16 + d;
        kr[d] = k[off];
        vr[d] = v[off];
        ak[d] = av[d] = 0.f;
    }
    for (int i0 = q_begin; i0 < q_end; i0 += XA_QT) {
        __syncthreads();
        for (int t = threadIdx.x; t < XA_QT * (XA_HD / 4); t += blockDim.x) {
            const int ii = t / (XA_HD / 4), d = (t - ii * (XA_HD / 4)) * 4;
            const int i = i0 + ii;
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f), g = a;
            if (i < q_end) {
                const long long row = static_cast<long long>(b) * Sq + i;
                a = *reinterpret_cast<const float4*>(q + row * ld_q + h * XA_HD + d);
                g = *reinterpret_cast<const float4*>(dO + row * ld_o + h * XA_HD + d);
            }
            *reinterpret_cast<float4*>(&sq[ii][d]) = a;
            *reinterpret_cast<float4*>(&sg[ii][d]) = g;
        }
        for (int t = threadIdx.x; t < XA_QT; t += blockDim.x) {
            const int i = i0 + t;
            const long long li = (static_cast<long long>(b) * H + h) * Sq + i;
            sl[t] = i < q_end ? lse[li] : INFINITY;        // exp(s - inf) = 0: rows past the end contribute nothing
            sd[t] = i < q_end ? delta[li] : 0.f;
        }
        __syncthreads();
        for (int ii = 0; ii < XA_QT; ++ii) {
            // this lane's 16 dims of q_i / dO_i: four 128-bit broadcast reads each (scalar reads made the kernel
            // shared-memory-wavefront bound: 64 wavefronts per warp and query)
            float qv[16], gv[16];
#pragma unroll
            for (int d = 0; d < 16; d += 4) {
                const float4 a = *reinterpret_cast<const float4*>(&sq[ii][part * 16 + d]);
                const float4 g = *reinterpret_cast<const float4*>(&sg[ii][part * 16 + d]);
                qv[d] = a.x; qv[d + 1] = a.y; qv[d + 2] = a.z; qv[d + 3] = a.w;
                gv[d] = g.x; gv[d + 1] = g.y; gv[d + 2] = g.z; gv[d + 3] = g.w;
            }
            float s = 0.f, dp = 0.f;
            if (live) {
#pragma unroll
                for (int d = 0; d < 16; ++d) {
                    s = fmaf(qv[d], kr[d], s);
                    dp = fmaf(gv[d], vr[d], dp);
                }
            }
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            dp += __shfl_xor_sync(0xffffffffu, dp, 1);
            dp += __shfl_xor_sync(0xffffffffu, dp, 2);
            const float p = (attend && (!causal || j <= i0 + ii)) ? __expf(s - sl[ii]) : 0.f;
            const float ds = p * (dp - sd[ii]);
            if (live) {
#pragma unroll
                for (int d = 0; d < 16; ++d) {
                    ak[d] = fmaf(ds, qv[d], ak[d]);
                    av[d] = fmaf(p, gv[d], av[d]);
                }
            }
        }
    }
    if (nsplit > 1) {
        cg::cluster_group cluster = cg::this_cluster();
#pragma unroll
        for (int round = 0; round < 2; ++round) {
            float* acc = round == 0 ? ak : av;
#pragma unroll
            for (int d = 0; d < 16; d += 4) *reinterpret_cast<float4*>(&sred[threadIdx.x][d]) = make_float4(acc[d], acc[d + 1], acc[d + 2], acc[d + 3]);
            cluster.sync();
            if (split == 0) {
                for (int r = 1; r < nsplit; ++r) {
                    const float* peer = cluster.map_shared_rank(&sred[0][0], r) + threadIdx.x * 16;
#pragma unroll
                    for (int d = 0; d < 16; d += 4) {
                        const float4 t = *reinterpret_cast<const float4*>(peer + d);
                        acc[d] += t.x; acc[d + 1] += t.y; acc[d + 2] += t.z; acc[d + 3] += t.w;
                    }
                }
            }
            cluster.sync();       // nobody rewrites sred (or exits) while CTA 0 is still reading it
        }
    }
    if (split == 0 && live) {
#pragma unroll
        for (int d = 0; d < 16; ++d) {
            const long long off = (static_cast<long long>(b) * Sk + j) * ld_dkv + h * XA_HD + part * 16 + d;
            dk[off] = ak[d];
            dv[off] = av[d];
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Projector tail (layers.py:96-119): per-sample dynamic 3x3 convolution, C channels -> 1, weights produced by the
// text branch.  taps[b,p,t] = x[b,p,:] . w[b,:,t] is contracted once per pixel (one read of x), the 3x3 stencil over
// the 9 tap planes then gives out[b,p] = bias[b] + sum_t taps[b, p + off_t, t] (zero padding).
// ---------------------------------------------------------------------------------------------------------------
constexpr int DC_MAXC = 512;
__global__ void __launch_bounds__(256) dynconv_taps_kernel(const float* __restrict__ x, const float* __restrict__ w, long long ld_w, int HW, int C,
                                                           float* __restrict__ taps) {
    __shared__ __align__(16) float sw[9][DC_MAXC];        // w[b] is (C, 3, 3) flattened: index c * 9 + t
    const int b = blockIdx.y;
    for (int i = threadIdx.x; i < C * 9; i += blockDim.x) sw[i % 9][i / 9] = w[b * ld_w + i];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int p = blockIdx.x * 8 + warp; p < HW; p += gridDim.x * 8) {
        const float* xr = x + (static_cast<long long>(b) * HW + p) * C;
        float acc[9];
#pragma unroll
        for (int t = 0; t < 9; ++t) acc[t] = 0.f;
        for (int c = lane * 4; c < C; c += 128) {
            const float4 xv = *reinterpret_cast<const float4*>(xr + c);
#pragma unroll
            for (int t = 0; t < 9; ++t) {
                const float4 wv = *reinterpret_cast<const float4*>(&sw[t][c]);
                acc[t] += xv.x * wv.x + xv.y * wv.y + xv.z * wv.z + xv.w * wv.w;
            }
        }
#pragma unroll
        for (int t = 0; t < 9; ++t) acc[t] = warp_sum(acc[t]);
        if (lane < 9) {
            float vsel = acc[0];
#pragma unroll
            for (int t = 1; t < 9; ++t) vsel = lane == t ? acc[t] : vsel;
            taps[(static_cast<long long>(b) * HW + p) * 9 + lane] = vsel;
        }
    }
}

__global__ void __launch_bounds__(256) dynconv_stencil_kernel(const float* __restrict__ taps, const float* __restrict__ bias, long long ld_bias, int H,
                                                              int W, float* __restrict__ out, long long total) {
    for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long b = idx / (H * W);
        const int r = static_cast<int>(idx - b * (H * W));
        const int y = r / W, x = r - y * W;
        float acc = bias[b * ld_bias];
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const int yy = y + ky - 1;
            if (yy < 0 || yy >= H) continue;
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int xx = x + kx - 1;
                if (xx < 0 || xx >= W) continue;
                acc += taps[((b * H + yy) * W + xx) * 9 + ky * 3 + kx];
            }
        }
        out[idx] = acc;
    }
}

// dx[b,p,c] = sum_t dtap_t(p) w[b,c,t],  dtap_t(p) = dout[b, p - off_t]  (0 outside)
// dw partials: part[chunk, b, c*9+t] = sum_{p in chunk} x[b,p,c] dtap_t(p)
__device__ __forceinline__ void dc_dtaps(const float* __restrict__ dout, long long b, int H, int W, int p, float (&dt)[9]) {
    const int y = p / W, x = p - y * W;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
            const int yy = y - (ky - 1), xx = x - (kx - 1);
            dt[ky * 3 + kx] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? dout[(b * H + yy) * W + xx] : 0.f;
        }
}

__global__ void __launch_bounds__(256) dynconv_bwd_dx_kernel(const float* __restrict__ dout, const float* __restrict__ w, long long ld_w, int H, int W,
                                                             int C, float* __restrict__ dx) {
    __shared__ __align__(16) float sw[9][DC_MAXC];
    const int b = blockIdx.y, HW = H * W;
    for (int i = threadIdx.x; i < C * 9; i += blockDim.x) sw[i % 9][i / 9] = w[b * ld_w + i];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int p = blockIdx.x * 8 + warp; p < HW; p += gridDim.x * 8) {
        float dt[9];
        dc_dtaps(dout, b, H, W, p, dt);
        float* dr = dx + (static_cast<long long>(b) * HW + p) * C;
        for (int c = lane * 4; c < C; c += 128) {
            float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int t = 0; t < 9; ++t) {
                const float4 wv = *reinterpret_cast<const float4*>(&sw[t][c]);
                o.x = fmaf(dt[t], wv.x, o.x); o.y = fmaf(dt[t], wv.y, o.y);
                o.z = fmaf(dt[t], wv.z, o.z); o.w = fmaf(dt[t], wv.w, o.w);
            }
            *reinterpret_cast<float4*>(dr + c) = o;
        }
    }
}

__global__ void __launch_bounds__(256) dynconv_bwd_dw_kernel(const float* __restrict__ dout, const float* __restrict__ x, int H, int W, int C,
                                                             float* __restrict__ part) {
    __shared__ float sdt[64][9];
    const int b = blockIdx.y, HW = H * W, chunks = gridDim.x;
    const int per = (HW + chunks - 1) / chunks;
    const int p0 = blockIdx.x * per, p1 = min(HW, p0 + per);
    float acc[2][9];
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int t = 0; t < 9; ++t) acc[u][t] = 0.f;
    for (int pb = p0; pb < p1; pb += 64) {
        __syncthreads();
        if (threadIdx.x < 64) {
            float dt[9];
            const int p = pb + threadIdx.x;
            if (p < p1) dc_dtaps(dout, b, H, W, p, dt);
#pragma unroll
            for (int t = 0; t < 9; ++t) sdt[threadIdx.x][t] = p < p1 ? dt[t] : 0.f;
        }
        __syncthreads();
        const int n = min(64, p1 - pb);
        for (int ii = 0; ii < n; ++ii) {
            const float* xr = x + (static_cast<long long>(b) * HW + pb + ii) * C;
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int c = threadIdx.x + 256 * u;
                if (c < C) {
                    const float xv = xr[c];
#pragma unroll
                    for (int t = 0; t < 9; ++t) acc[u][t] = fmaf(xv, sdt[ii][t], acc[u][t]);
                }
            }
        }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const int c = threadIdx.x + 256 * u;
        if (c < C)
#pragma unroll
            for (int t = 0; t < 9; ++t) part[((static_cast<long long>(blockIdx.x) * gridDim.y + b) * C + c) * 9 + t] = acc[u][t];
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Table-driven separable resampling of single-channel maps (the bicubic, align_corners=True upsampling of the CRIS
// prediction, coop_cris.py:235).  The host supplies per-output-index taps (NT indices + weights per output) and, for
// the backward, the transposed table (per input index: up to MT (output index, weight) pairs, count in cnt).
// tile > 0 writes / reads the big map in the decoder-head layout [B*(Ho/tile)*(Wo/tile), tile*tile] (see tvs_head_fwd).
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ long long tiled_index(long long b, int y, int x, int Ho, int Wo, int tile) {
    if (tile <= 0) return (b * Ho + y) * Wo + x;
    const int gy = y / tile, gx = x / tile, py = y - gy * tile, px = x - gx * tile;
    return ((b * (Ho / tile) + gy) * (Wo / tile) + gx) * (static_cast<long long>(tile) * tile) + py * tile + px;
}

template <bool U8>
__global__ void __launch_bounds__(256) resample2d_fwd_kernel(const float* __restrict__ in, int Hi, int Wi, int Ho, int Wo, const int* __restrict__ iy,
                                                             const float* __restrict__ wy, const int* __restrict__ ix, const float* __restrict__ wx,
                                                             int NT, int tile, void* __restrict__ out, long long total) {
    for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long b = idx / (static_cast<long long>(Ho) * Wo);
        const int r = static_cast<int>(idx - b * Ho * Wo);
        const int y = r / Wo, x = r - y * Wo;
        float acc = 0.f;
        if (U8 && NT == 4) {
            // Byte-exact against the reference's host-side TF.resize (ATen UpSampleKernel.cpp Interpolate<2,...,4>, x86 build
            // with FMA contraction; restated and pinned in oracle/resize_u8.py): every 4-tap sum is
            // fma(v3, w3, fma(v2, w2, fma(v0, w0, fl(v1 * w1)))), x taps inside, y taps outside.
            float t[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const float* row = in + (b * Hi + iy[y * 4 + a]) * Wi;
                float s = __fmaf_rn(row[ix[x * 4 + 0]], wx[x * 4 + 0], __fmul_rn(row[ix[x * 4 + 1]], wx[x * 4 + 1]));
                s = __fmaf_rn(row[ix[x * 4 + 2]], wx[x * 4 + 2], s);
                t[a] = __fmaf_rn(row[ix[x * 4 + 3]], wx[x * 4 + 3], s);
            }
            acc = __fmaf_rn(t[0], wy[y * 4 + 0], __fmul_rn(t[1], wy[y * 4 + 1]));
            acc = __fmaf_rn(t[2], wy[y * 4 + 2], acc);
            acc = __fmaf_rn(t[3], wy[y * 4 + 3], acc);
        } else {
            for (int a = 0; a < NT; ++a) {
                const float* row = in + (b * Hi + iy[y * NT + a]) * Wi;
                float s = 0.f;
                for (int c = 0; c < NT; ++c) s = fmaf(wx[x * NT + c], row[ix[x * NT + c]], s);
                acc = fmaf(wy[y * NT + a], s, acc);
            }
        }
        if (U8) {
            // torchvision.utils.save_image: mul(255).add_(0.5).clamp_(0, 255).to(uint8) - two roundings, then truncation
            const float q = __fadd_rn(__fmul_rn(acc, 255.f), 0.5f);
            static_cast<uint8_t*>(out)[idx] = static_cast<uint8_t>(fminf(fmaxf(q, 0.f), 255.f));
        } else {
            static_cast<float*>(out)[tiled_index(b, y, x, Ho, Wo, tile)] = acc;
        }
    }
}

template <bool BF16>
__global__ void __launch_bounds__(256) resample2d_bwd_kernel(const void* __restrict__ dout, int Hi, int Wi, int Ho, int Wo, const int* __restrict__ ty,
                                                             const float* __restrict__ twy, const int* __restrict__ cy, const int* __restrict__ tx,
                                                             const float* __restrict__ twx, const int* __restrict__ cx, int MT, int tile,
                                                             float* __restrict__ din, long long total) {
    for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long b = idx / (static_cast<long long>(Hi) * Wi);
        const int r = static_cast<int>(idx - b * Hi * Wi);
        const int y = r / Wi, x = r - y * Wi;
        float acc = 0.f;
        for (int a = 0; a < cy[y]; ++a) {
            const int oy = ty[y * MT + a];
            float s = 0.f;
            for (int c = 0; c < cx[x]; ++c) {
                const long long o = tiled_index(b, oy, tx[x * MT + c], Ho, Wo, tile);
                const float g = BF16 ? __bfloat162float(static_cast<const __nv_bfloat16*>(dout)[o]) : static_cast<const float*>(dout)[o];
                s = fmaf(twx[x * MT + c], g, s);
            }
            acc = fmaf(twy[y * MT + a], s, acc);
        }
        din[idx] = acc;
    }
}

}  // namespace tvs

// ===================================================================================================================
// C ABI
// ===================================================================================================================
using namespace tvs;
#define TVS_API extern "C" __attribute__((visibility("default")))

TVS_API int tvs_im2col_nhwc(const void* x, int32_t elem_bytes, int32_t B, int32_t H, int32_t W, int32_t C, int32_t ksize, int32_t stride,
                            int32_t pad, void* col, int64_t ldcol, int32_t round_tf32, void* stream) {
    TVS_REQUIRE(x && col && (elem_bytes == 2 || elem_bytes == 4), "tvs_im2col_nhwc: bad pointers / elem_bytes");
    TVS_REQUIRE(!round_tf32 || elem_bytes == 4, "tvs_im2col_nhwc: round_tf32 needs f32 elements");
    TVS_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0 && ksize >= 1 && stride >= 1 && pad >= 0, "tvs_im2col_nhwc: bad geometry");
    const int Ho = (H + 2 * pad - ksize) / stride + 1, Wo = (W + 2 * pad - ksize) / stride + 1;
    const long long K = static_cast<long long>(ksize) * ksize * C;
    TVS_REQUIRE(Ho > 0 && Wo > 0 && ldcol >= K, "tvs_im2col_nhwc: ldcol %lld < k*k*C = %lld", static_cast<long long>(ldcol), K);
    const long long cb = static_cast<long long>(C) * elem_bytes, lb = ldcol * elem_bytes;
    const auto al = [&](int v) {
        return cb % v == 0 && lb % v == 0 && reinterpret_cast<uintptr_t>(x) % v == 0 && reinterpret_cast<uintptr_t>(col) % v == 0;
    };
    const long long rows = static_cast<long long>(B) * Ho * Wo;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
#define TVS_IM2COL(V)                                                                                                          \
    {                                                                                                                          \
        const int vb = static_cast<int>(sizeof(V));                                                                            \
        const int cv = static_cast<int>(cb / vb), ldv = static_cast<int>(lb / vb), kvn = static_cast<int>(K * elem_bytes / vb); \
        int shift = 5;                                                                                                         \
        while (shift < 8 && (1 << (shift + 1)) <= std::max(32, std::min(cv, ldv))) ++shift;                                    \
        const int rpb = 256 >> shift;                                                                                          \
        const unsigned grid = grid_for((rows + rpb - 1) / rpb, 1, 148 * 16);                                                   \
        if (round_tf32)                                                                                                        \
            im2col_kernel<V, true><<<grid, 256, 0, st>>>(static_cast<const V*>(x), static_cast<V*>(col), H, W, cv, Ho, Wo, ksize, \
                                                         stride, pad, ldv, kvn, rows, shift);                                  \
        else                                                                                                                   \
            im2col_kernel<V, false><<<grid, 256, 0, st>>>(static_cast<const V*>(x), static_cast<V*>(col), H, W, cv, Ho, Wo, ksize, \
                                                          stride, pad, ldv, kvn, rows, shift);                                 \
    }
    if (al(16)) TVS_IM2COL(uint4)
    else if (al(8)) TVS_IM2COL(uint2)
    else if (al(4)) TVS_IM2COL(uint32_t)
    else TVS_IM2COL(uint16_t)
#undef TVS_IM2COL
    return check_launch("im2col_kernel");
}

TVS_API int tvs_col2im_nhwc(const float* dcol, int64_t ldcol, int32_t B, int32_t H, int32_t W, int32_t Ccol, int32_t Cx, int32_t ksize,
                            const float* relu_mask, int64_t ld_mask, float* dx, int64_t ld_dx, void* stream) {
    TVS_REQUIRE(dcol && dx && (ksize & 1) && Cx > 0 && Cx <= Ccol && Cx % 4 == 0 && Ccol % 2 == 0 && ldcol % 2 == 0 && ld_dx % 4 == 0 &&
                    (!relu_mask || ld_mask % 4 == 0),
                "tvs_col2im_nhwc: needs odd ksize, Cx %% 4 == 0, even Ccol / ldcol, 16-byte aligned rows");
    const long long total = static_cast<long long>(B) * H * W * (Cx / 4);
    col2im_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(dcol, ldcol, H, W, Ccol, Cx, ksize, relu_mask, ld_mask, dx,
                                                                                       ld_dx, total);
    return check_launch("col2im_kernel");
}

TVS_API int tvs_pad_nhwc(const void* x, int32_t elem_bytes, int32_t B, int32_t H, int32_t W, int32_t C, void* xp, int32_t round_tf32,
                         void* stream) {
    TVS_REQUIRE(x && xp && (elem_bytes == 2 || elem_bytes == 4) && (static_cast<long long>(C) * elem_bytes) % 16 == 0,
                "tvs_pad_nhwc: rows of C elements must be a multiple of 16 bytes");
    TVS_REQUIRE(!round_tf32 || elem_bytes == 4, "tvs_pad_nhwc: round_tf32 needs f32 elements");
    const int cv = C * elem_bytes / 16;
    const long long total = static_cast<long long>(B) * (H + 2) * (W + 2) * cv;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (round_tf32)
        pad_nhwc_kernel<true><<<grid_for(total, 256), 256, 0, st>>>(static_cast<const uint4*>(x), static_cast<uint4*>(xp), H, W, cv, total);
    else
        pad_nhwc_kernel<false><<<grid_for(total, 256), 256, 0, st>>>(static_cast<const uint4*>(x), static_cast<uint4*>(xp), H, W, cv, total);
    return check_launch("pad_nhwc_kernel");
}

TVS_API int tvs_round_tf32(const float* x, int64_t ld_x, int64_t M, int32_t C, float* y, int64_t ld_y, void* stream) {
    TVS_REQUIRE(x && y && C % 4 == 0 && ld_x % 4 == 0 && ld_y % 4 == 0, "tvs_round_tf32: C and strides must be multiples of 4");
    const long long total = M * (C / 4);
    round_tf32_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, ld_x, y, ld_y, C / 4, total);
    return check_launch("round_tf32_kernel");
}

TVS_API int tvs_relu_mask(const float* dy, int64_t ld_dy, const float* y, int64_t ld_y, int64_t M, int32_t C, float* out, int64_t ld_out,
                          void* stream) {
    TVS_REQUIRE(dy && y && out && C % 4 == 0 && ld_dy % 4 == 0 && ld_y % 4 == 0 && ld_out % 4 == 0, "tvs_relu_mask: C and strides must be multiples of 4");
    const long long total = M * (C / 4);
    relu_mask_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(dy, ld_dy, y, ld_y, out, ld_out, C / 4, total);
    return check_launch("relu_mask_kernel");
}

TVS_API int tvs_avgpool2_nhwc(const void* x, int32_t is_bf16, int32_t B, int32_t H, int32_t W, int32_t C, void* y, int64_t ld_out, void* stream) {
    TVS_REQUIRE(x && y && H % 2 == 0 && W % 2 == 0, "tvs_avgpool2_nhwc: H and W must be even");
    const int round_out = (is_bf16 >> 1) & 1;
    is_bf16 &= 1;
    TVS_REQUIRE(!(round_out && is_bf16), "tvs_avgpool2_nhwc: tf32 rounding applies to f32 elements");
    const int vec = is_bf16 ? 8 : 4;
    TVS_REQUIRE(C % vec == 0 && ld_out % vec == 0, "tvs_avgpool2_nhwc: C and ld_out must be multiples of %d", vec);
    const long long total = static_cast<long long>(B) * (H / 2) * (W / 2) * (C / vec);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (is_bf16)
        avgpool2_kernel<true><<<grid_for(total, 256), 256, 0, st>>>(x, y, H, W, C, ld_out, total, 0);
    else
        avgpool2_kernel<false><<<grid_for(total, 256), 256, 0, st>>>(x, y, H, W, C, ld_out, total, round_out);
    return check_launch("avgpool2_kernel");
}

TVS_API int tvs_upsample2x_fwd(const float* x, int32_t B, int32_t H, int32_t W, int32_t C, float* y, int64_t ld_out, void* stream) {
    TVS_REQUIRE(x && y && C % 4 == 0 && ld_out % 4 == 0, "tvs_upsample2x_fwd: C and ld_out must be multiples of 4");
    const long long total = static_cast<long long>(B) * H * W * 4 * (C / 4);
    upsample2x_fwd_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, H, W, C, y, ld_out, total);
    return check_launch("upsample2x_fwd_kernel");
}

TVS_API int tvs_upsample2x_bwd(const float* dy, int64_t ld_dy, int32_t B, int32_t H, int32_t W, int32_t C, float* dx, void* stream) {
    TVS_REQUIRE(dy && dx && C % 4 == 0 && ld_dy % 4 == 0, "tvs_upsample2x_bwd: C and ld_dy must be multiples of 4");
    const long long total = static_cast<long long>(B) * H * W * (C / 4);
    upsample2x_bwd_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(dy, ld_dy, H, W, C, dx, total);
    return check_launch("upsample2x_bwd_kernel");
}

TVS_API int tvs_cross_attn_fwd(const float* q, int64_t ld_q, const float* k, const float* v, int64_t ld_kv, const uint8_t* key_mask, int32_t B,
                               int32_t Sq, int32_t Sk, int32_t H, int32_t hd, int32_t causal, float* out, int64_t ld_o, float* lse,
                               void* stream) {
    TVS_REQUIRE(!causal || Sq == Sk, "tvs_cross_attn_fwd: causal needs Sq == Sk");
    TVS_REQUIRE(q && k && v && out && lse, "tvs_cross_attn_fwd: null pointer");
    TVS_REQUIRE(hd == XA_HD && Sk >= 1 && Sk <= XA_MAXK, "tvs_cross_attn_fwd: head dim must be %d and 1 <= Sk <= %d (got %d, %d)", XA_HD, XA_MAXK, hd, Sk);
    TVS_REQUIRE(ld_q % 4 == 0 && ld_kv % 4 == 0 && ld_o % 4 == 0, "tvs_cross_attn_fwd: strides must be multiples of 4");
    cross_attn_fwd_kernel<<<dim3((Sq + XA_QPB - 1) / XA_QPB, H, B), XA_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
        q, ld_q, k, v, ld_kv, key_mask, Sq, Sk, causal, out, ld_o, lse);
    return check_launch("cross_attn_fwd_kernel");
}

TVS_API int tvs_cross_attn_bwd(const float* q, int64_t ld_q, const float* k, const float* v, int64_t ld_kv, const uint8_t* key_mask, const float* out,
                               const float* dout, int64_t ld_o, const float* lse, int32_t B, int32_t Sq, int32_t Sk, int32_t H, int32_t hd,
                               int32_t causal, float* dq, int64_t ld_dq, float* dk, float* dv, int64_t ld_dkv, float* delta, void* stream) {
    TVS_REQUIRE(!causal || Sq == Sk, "tvs_cross_attn_bwd: causal needs Sq == Sk");
    TVS_REQUIRE(q && k && v && out && dout && lse && dq && dk && dv && delta, "tvs_cross_attn_bwd: null pointer");
    TVS_REQUIRE(hd == XA_HD && Sk >= 1 && Sk <= XA_MAXK, "tvs_cross_attn_bwd: head dim must be %d and 1 <= Sk <= %d", XA_HD, XA_MAXK);
    TVS_REQUIRE(ld_q % 4 == 0 && ld_kv % 4 == 0 && ld_o % 4 == 0 && ld_dq % 4 == 0, "tvs_cross_attn_bwd: strides must be multiples of 4");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cross_attn_bwd_dq_kernel<<<dim3((Sq + XA_QPB - 1) / XA_QPB, H, B), XA_THREADS, 0, st>>>(q, ld_q, k, v, ld_kv, key_mask, out, dout, ld_o,
                                                                                                     lse, Sq, Sk, causal, dq, ld_dq, delta);
    if (check_launch("cross_attn_bwd_dq_kernel")) return -3;
    const int dkv_threads = (Sk * 4 + 31) / 32 * 32;       // 4 lanes per key; the kernel is issue bound, so no warp without a key is launched
    const int nsplit = std::min(8, std::max(1, Sq / 64));       // cluster of query splits; 1 for the text encoder's own attention (Sq <= 77)
    TVS_CUDA(launch_pdl(cross_attn_bwd_dkv_kernel, dim3(nsplit, H, B), dim3(dkv_threads), 0, st, nsplit, q, ld_q, k, v, ld_kv, key_mask, dout, ld_o, lse,
                        delta, Sq, Sk, causal, dk, dv, ld_dkv));
    return check_launch("cross_attn_bwd_dkv_kernel");
}

TVS_API int tvs_dynconv_fwd(const float* x, const float* w, int64_t ld_w, const float* bias, int64_t ld_bias, int32_t B, int32_t H, int32_t W,
                            int32_t C, float* taps, float* out, void* stream) {
    TVS_REQUIRE(x && w && bias && taps && out, "tvs_dynconv_fwd: null pointer");
    TVS_REQUIRE(C % 4 == 0 && C <= DC_MAXC, "tvs_dynconv_fwd: C must be a multiple of 4 and <= %d", DC_MAXC);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int HW = H * W;
    const int gx = std::max(1, std::min((HW + 7) / 8, (148 * 8 + B - 1) / B));
    dynconv_taps_kernel<<<dim3(gx, B), 256, 0, st>>>(x, w, ld_w, HW, C, taps);
    if (check_launch("dynconv_taps_kernel")) return -3;
    const long long total = static_cast<long long>(B) * HW;
    dynconv_stencil_kernel<<<grid_for(total, 256), 256, 0, st>>>(taps, bias, ld_bias, H, W, out, total);
    return check_launch("dynconv_stencil_kernel");
}

TVS_API int tvs_dynconv_bwd(const float* dout, const float* x, const float* w, int64_t ld_w, int32_t B, int32_t H, int32_t W, int32_t C, float* dx,
                            float* dw_part, int32_t chunks, void* stream) {
    TVS_REQUIRE(dout && x && w && dx && dw_part && chunks >= 1, "tvs_dynconv_bwd: null pointer");
    TVS_REQUIRE(C % 4 == 0 && C <= DC_MAXC, "tvs_dynconv_bwd: C must be a multiple of 4 and <= %d", DC_MAXC);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int HW = H * W;
    const int gx = std::max(1, std::min((HW + 7) / 8, (148 * 8 + B - 1) / B));
    dynconv_bwd_dx_kernel<<<dim3(gx, B), 256, 0, st>>>(dout, w, ld_w, H, W, C, dx);
    if (check_launch("dynconv_bwd_dx_kernel")) return -3;
    dynconv_bwd_dw_kernel<<<dim3(chunks, B), 256, 0, st>>>(dout, x, H, W, C, dw_part);
    return check_launch("dynconv_bwd_dw_kernel");
}

TVS_API int tvs_resample2d_fwd(const float* in, int32_t B, int32_t Hi, int32_t Wi, int32_t Ho, int32_t Wo, const int32_t* iy, const float* wy,
                               const int32_t* ix, const float* wx, int32_t ntaps, int32_t tile, float* out, void* stream) {
    TVS_REQUIRE(in && out && iy && wy && ix && wx && ntaps >= 1, "tvs_resample2d_fwd: null pointer");
    TVS_REQUIRE(tile == 0 || (Ho % tile == 0 && Wo % tile == 0), "tvs_resample2d_fwd: tile must divide the output size");
    const long long total = static_cast<long long>(B) * Ho * Wo;
    resample2d_fwd_kernel<false><<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(in, Hi, Wi, Ho, Wo, iy, wy, ix, wx, ntaps, tile,
                                                                                                      out, total);
    return check_launch("resample2d_fwd_kernel");
}

TVS_API int tvs_resample2d_u8(const float* in, int32_t B, int32_t Hi, int32_t Wi, int32_t Ho, int32_t Wo, const int32_t* iy, const float* wy,
                              const int32_t* ix, const float* wx, int32_t ntaps, uint8_t* out, void* stream) {
    TVS_REQUIRE(in && out && iy && wy && ix && wx && ntaps >= 1, "tvs_resample2d_u8: null pointer");
    const long long total = static_cast<long long>(B) * Ho * Wo;
    resample2d_fwd_kernel<true><<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(in, Hi, Wi, Ho, Wo, iy, wy, ix, wx, ntaps, 0, out,
                                                                                                     total);
    return check_launch("resample2d_u8_kernel");
}

TVS_API int tvs_resample2d_bwd(const void* dout, int32_t dout_is_bf16, int32_t B, int32_t Hi, int32_t Wi, int32_t Ho, int32_t Wo, const int32_t* ty,
                               const float* twy, const int32_t* cy, const int32_t* tx, const float* twx, const int32_t* cx, int32_t max_taps,
                               int32_t tile, float* din, void* stream) {
    TVS_REQUIRE(dout && din && ty && twy && cy && tx && twx && cx && max_taps >= 1, "tvs_resample2d_bwd: null pointer");
    TVS_REQUIRE(tile == 0 || (Ho % tile == 0 && Wo % tile == 0), "tvs_resample2d_bwd: tile must divide the output size");
    const long long total = static_cast<long long>(B) * Hi * Wi;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dout_is_bf16)
        resample2d_bwd_kernel<true><<<grid_for(total, 256), 256, 0, st>>>(dout, Hi, Wi, Ho, Wo, ty, twy, cy, tx, twx, cx, max_taps, tile, din, total);
    else
        resample2d_bwd_kernel<false><<<grid_for(total, 256), 256, 0, st>>>(dout, Hi, Wi, Ho, Wo, ty, twy, cy, tx, twx, cx, max_taps, tile, din, total);
    return check_launch("resample2d_bwd_kernel");
}
