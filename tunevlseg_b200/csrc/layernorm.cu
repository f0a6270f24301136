// LayerNorm forward / backward (dgrad only), one warp per row, float4 loads, warp-shuffle reductions.
// Replaces nn.LayerNorm on the path: transformers modeling_clipseg.py:362-365 (layer_norm1/2), :782 (pre_layrnorm),
// :636 (final_layer_norm), :784 (post_layernorm), :395-398 (decoder post-norms).  HBM-bound: reads the fp32 row once.
#include <stdlib.h>

#include "common.cuh"
#include "tvs_b200.h"

namespace tvs {

constexpr int LN_MAXC = 16;  // float4 chunks per lane: the kernels are instantiated for 8 (D <= 1024) and 16 (D <= 2048)
constexpr int LN_WARPS = 4;

template <int MAXC>
__global__ void __launch_bounds__(LN_WARPS * 32)
layernorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta, float eps, long long M,
                     int D, float* __restrict__ y32, __nv_bfloat16* __restrict__ y16, float* __restrict__ mean_out, float* __restrict__ rstd_out,
                     int round_out) {
    pdl_wait();
    pdl_trigger();
    const long long row = static_cast<long long>(blockIdx.x) * LN_WARPS + (threadIdx.x >> 5);
    if (row >= M) return;
    const int lane = threadIdx.x & 31;
    const int nch = D >> 2;
    const float4* xr = reinterpret_cast<const float4*>(x + row * D);
    float4 v[MAXC];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < MAXC; ++i) {
        const int c = lane + 32 * i;
        if (c < nch) {
            v[i] = xr[c];
            sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        }
    }
    const float mean = warp_sum(sum) / static_cast<float>(D);
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < MAXC; ++i) {
        const int c = lane + 32 * i;
        if (c < nch) {
            const float a = v[i].x - mean, b = v[i].y - mean, cc = v[i].z - mean, d = v[i].w - mean;
            sq += (a * a + b * b) + (cc * cc + d * d);
        }
    }
    const float rstd = 1.0f / sqrtf(warp_sum(sq) / static_cast<float>(D) + eps);
    if (lane == 0) {
        if (mean_out) mean_out[row] = mean;
        if (rstd_out) rstd_out[row] = rstd;
    }
    const float4* g4 = reinterpret_cast<const float4*>(gamma);
    const float4* b4 = reinterpret_cast<const float4*>(beta);
#pragma unroll
    for (int i = 0; i < MAXC; ++i) {
        const int c = lane + 32 * i;
        if (c < nch) {
            const float4 g = __ldg(g4 + c), b = __ldg(b4 + c);
            float4 o;
            o.x = (v[i].x - mean) * rstd * g.x + b.x;
            o.y = (v[i].y - mean) * rstd * g.y + b.y;
            o.z = (v[i].z - mean) * rstd * g.z + b.z;
            o.w = (v[i].w - mean) * rstd * g.w + b.w;
            if (y32)
                reinterpret_cast<float4*>(y32 + row * D)[c] =
                    (round_out & 1) ? make_float4(round_tf32_rn(o.x), round_tf32_rn(o.y), round_tf32_rn(o.z), round_tf32_rn(o.w)) : o;
            if (y16) reinterpret_cast<uint2*>(y16 + row * D)[c] = (round_out & 2) ? make_uint2(pack_f16x2(o.x, o.y), pack_f16x2(o.z, o.w)) : make_uint2(pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
        }
    }
}

// Two rows per warp (large M, D <= 768): all 2 x D / 128 loads of the pair are issued before the first reduction, the two
// reductions interleave, gamma / beta are fetched once for both rows.  The one-row kernel had ~3 KB in flight per warp and
// ran at 50 % of the copy bandwidth on [15 648, 768] (22.5 us cold; round 2).
constexpr int LN2_WARPS = 8;
template <int MAXC>
__global__ void __launch_bounds__(LN2_WARPS * 32)
layernorm_fwd2_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta, float eps, long long M,
                      int D, float* __restrict__ y32, __nv_bfloat16* __restrict__ y16, float* __restrict__ mean_out, float* __restrict__ rstd_out,
                      int round_out) {
    pdl_wait();
    pdl_trigger();
    const long long row0 = (static_cast<long long>(blockIdx.x) * LN2_WARPS + (threadIdx.x >> 5)) * 2;
    if (row0 >= M) return;
    const bool two = row0 + 1 < M;
    const int lane = threadIdx.x & 31;
    const int nch = D >> 2;
    const float4* xa = reinterpret_cast<const float4*>(x + row0 * D);
    const float4* xb = reinterpret_cast<const float4*>(x + (row0 + (two ? 1 : 0)) * D);
    float4 va[MAXC], vb[MAXC];
#pragma unroll
    for (int i = 0; i < MAXC; ++i) {
        const int c = lane + 32 * i;
        if (c < nch) {
            va[i] = __ldcs(xa + c);
            vb[i] = __ldcs(xb + c);
        }
    }
    float sa = 0.f, sb = 0.f;
#pragma unroll
    for (int i = 0; i < MAXC; ++i)
        if (lane + 32 * i < nch) {
            sa += (va[i].x + va[i].y) + (va[i].z + va[i].w);
            sb += (vb[i].x + vb[i].y) + (vb[i].z + vb[i].w);
        }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sa += __shfl_xor_sync(0xffffffffu, sa, o);
        sb += __shfl_xor_sync(0xffffffffu, sb, o);
    }
    const float ma = sa / static_cast<float>(D), mb = sb / static_cast<float>(D);
    float qa = 0.f, qb = 0.f;
#pragma unroll
    for (int i = 0; i < MAXC; ++i)
        if (lane + 32 * i < nch) {
            float a = va[i].x - ma, b = va[i].y - ma, c = va[i].z - ma, d = va[i].w - ma;
            qa += (a * a + b * b) + (c * c + d * d);
            a = vb[i].x - mb; b = vb[i].y - mb; c = vb[i].z - mb; d = vb[i].w - mb;
            qb += (a * a + b * b) + (c * c + d * d);
        }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        qa += __shfl_xor_sync(0xffffffffu, qa, o);
        qb += __shfl_xor_sync(0xffffffffu, qb, o);
    }
    const float ra = 1.0f / sqrtf(qa / static_cast<float>(D) + eps), rb = 1.0f / sqrtf(qb / static_cast<float>(D) + eps);
    if (lane == 0) {
        if (mean_out) { mean_out[row0] = ma; if (two) mean_out[row0 + 1] = mb; }
        if (rstd_out) { rstd_out[row0] = ra; if (two) rstd_out[row0 + 1] = rb; }
    }
    const float4* g4 = reinterpret_cast<const float4*>(gamma);
    const float4* b4 = reinterpret_cast<const float4*>(beta);
    auto emit = [&](const float4& v, float mean, float rstd, const float4& g, const float4& b, long long row, int c) {
        float4 o;
        o.x = (v.x - mean) * rstd * g.x + b.x;
        o.y = (v.y - mean) * rstd * g.y + b.y;
        o.z = (v.z - mean) * rstd * g.z + b.z;
        o.w = (v.w - mean) * rstd * g.w + b.w;
        if (y32)
            reinterpret_cast<float4*>(y32 + row * D)[c] =
                (round_out & 1) ? make_float4(round_tf32_rn(o.x), round_tf32_rn(o.y), round_tf32_rn(o.z), round_tf32_rn(o.w)) : o;
        if (y16)
            reinterpret_cast<uint2*>(y16 + row * D)[c] = (round_out & 2) ? make_uint2(pack_f16x2(o.x, o.y), pack_f16x2(o.z, o.w))
                                                                        : make_uint2(pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
    };
#pragma unroll
    for (int i = 0; i < MAXC; ++i) {
        const int c = lane + 32 * i;
        if (c < nch) {
            const float4 g = __ldg(g4 + c), b = __ldg(b4 + c);
            emit(va[i], ma, ra, g, b, row0, c);
            if (two) emit(vb[i], mb, rb, g, b, row0 + 1, c);
        }
    }
}

// Persistent, software-pipelined forward (TVS_LN_FWD=4): a fixed grid of warps walks the rows with a stride; the row a warp
// will normalise NEXT is already in flight (registers) while it reduces and stores the current one, so in steady state only
// bandwidth matters - the one-shot kernels above expose a DRAM round trip, the reductions and the store back to back per row
// and leave the tail of every wave of CTAs half empty.
template <int MAXC>
__global__ void __launch_bounds__(LN2_WARPS * 32, 3)
layernorm_fwd_pipe_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta, float eps, long long M,
                          int D, float* __restrict__ y32, __nv_bfloat16* __restrict__ y16, float* __restrict__ mean_out, float* __restrict__ rstd_out,
                          int round_out) {
    pdl_wait();
    pdl_trigger();
    const int lane = threadIdx.x & 31;
    const int nch = D >> 2;
    const long long stride = static_cast<long long>(gridDim.x) * LN2_WARPS;
    long long row = static_cast<long long>(blockIdx.x) * LN2_WARPS + (threadIdx.x >> 5);
    if (row >= M) return;
    const float4* g4 = reinterpret_cast<const float4*>(gamma);
    const float4* b4 = reinterpret_cast<const float4*>(beta);
    float4 cur[MAXC], nxt[MAXC];
    {
        const float4* xr = reinterpret_cast<const float4*>(x + row * D);
#pragma unroll
        for (int i = 0; i < MAXC; ++i)
            if (lane + 32 * i < nch) cur[i] = __ldcs(xr + lane + 32 * i);
    }
    const float inv_d = 1.0f / static_cast<float>(D);
    for (; row < M; row += stride) {
        const long long rn = row + stride;
        if (rn < M) {
            const float4* xr = reinterpret_cast<const float4*>(x + rn * D);
#pragma unroll
            for (int i = 0; i < MAXC; ++i)
                if (lane + 32 * i < nch) nxt[i] = __ldcs(xr + lane + 32 * i);
        }
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < MAXC; ++i)
            if (lane + 32 * i < nch) s += (cur[i].x + cur[i].y) + (cur[i].z + cur[i].w);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const float m = s / static_cast<float>(D);
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < MAXC; ++i)
            if (lane + 32 * i < nch) {
                const float a = cur[i].x - m, b = cur[i].y - m, c = cur[i].z - m, d = cur[i].w - m;
                q += (a * a + b * b) + (c * c + d * d);
            }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
        const float r = 1.0f / sqrtf(q / static_cast<float>(D) + eps);
        (void)inv_d;
        if (lane == 0) {
            if (mean_out) mean_out[row] = m;
            if (rstd_out) rstd_out[row] = r;
        }
#pragma unroll
        for (int i = 0; i < MAXC; ++i) {
            const int c = lane + 32 * i;
            if (c < nch) {
                const float4 g = __ldg(g4 + c), b = __ldg(b4 + c);
                float4 o;
                o.x = (cur[i].x - m) * r * g.x + b.x;
                o.y = (cur[i].y - m) * r * g.y + b.y;
                o.z = (cur[i].z - m) * r * g.z + b.z;
                o.w = (cur[i].w - m) * r * g.w + b.w;
                if (y32)
                    reinterpret_cast<float4*>(y32 + row * D)[c] =
                        (round_out & 1) ? make_float4(round_tf32_rn(o.x), round_tf32_rn(o.y), round_tf32_rn(o.z), round_tf32_rn(o.w)) : o;
                if (y16)
                    reinterpret_cast<uint2*>(y16 + row * D)[c] = (round_out & 2) ? make_uint2(pack_f16x2(o.x, o.y), pack_f16x2(o.z, o.w))
                                                                                : make_uint2(pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
            }
        }
#pragma unroll
        for (int i = 0; i < MAXC; ++i) cur[i] = nxt[i];
    }
}

// VAR 0: dx_add is fetched after the reductions (two dependent DRAM round trips per row); VAR 1: every load of the row is
// issued before the first reduction (one round trip, 24 more registers); VAR 2: VAR 1 + streaming (evict-first) hints on
// the operands nobody re-reads (saved x, the incoming fp32 gradient stream and its fp32 successor).
template <int MAXC, int VAR>
__global__ void __launch_bounds__(LN_WARPS * 32)
layernorm_bwd_kernel(const __nv_bfloat16* __restrict__ dy16, const float* __restrict__ dy32, const float* __restrict__ x,
                     const float* __restrict__ gamma, const float* __restrict__ mean, const float* __restrict__ rstd, const float* dx_add,
                     long long M, int D, float* dx32, __nv_bfloat16* __restrict__ dx16) {
    pdl_wait();
    pdl_trigger();
    const long long row = static_cast<long long>(blockIdx.x) * LN_WARPS + (threadIdx.x >> 5);
    if (row >= M) return;
    const int lane = threadIdx.x & 31;
    const int nch = D >> 2;
    const float4* xr = reinterpret_cast<const float4*>(x + row * D);
    const float4* ar = reinterpret_cast<const float4*>(dx_add ? dx_add + row * D : nullptr);
    const float4* g4 = reinterpret_cast<const float4*>(gamma);
    float4 gg[MAXC], xh[MAXC], ad[VAR >= 1 ? MAXC : 1];
#pragma unroll
    for (int i = 0; i < MAXC; ++i) {
        const int c = lane + 32 * i;
        if (c < nch) {
            if (dy16) {
                const uint2 raw = VAR == 2 ? __ldcs(reinterpret_cast<const uint2*>(dy16 + row * D) + c) : reinterpret_cast<const uint2*>(dy16 + row * D)[c];
                const float2 lo = unpack_bf16x2(raw.x), hi = unpack_bf16x2(raw.y);
                gg[i] = make_float4(lo.x, lo.y, hi.x, hi.y);
            } else {
                gg[i] = VAR == 2 ? __ldcs(reinterpret_cast<const float4*>(dy32 + row * D) + c) : reinterpret_cast<const float4*>(dy32 + row * D)[c];
            }
            xh[i] = VAR == 2 ? __ldcs(xr + c) : xr[c];
            if (VAR >= 1) ad[i] = ar ? (VAR == 2 ? __ldcs(ar + c) : ar[c]) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    const float mu = mean[row], rs = rstd[row];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < MAXC; ++i) {
        const int c = lane + 32 * i;
        if (c < nch) {
            const float4 g = __ldg(g4 + c);
            const float4 d = gg[i], xv = xh[i];
            gg[i] = make_float4(d.x * g.x, d.y * g.y, d.z * g.z, d.w * g.w);
            xh[i] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
            s1 += (gg[i].x + gg[i].y) + (gg[i].z + gg[i].w);
            s2 += (gg[i].x * xh[i].x + gg[i].y * xh[i].y) + (gg[i].z * xh[i].z + gg[i].w * xh[i].w);
        }
    }
    const float c1 = warp_sum(s1) / static_cast<float>(D);
    const float c2 = warp_sum(s2) / static_cast<float>(D);
#pragma unroll
    for (int i = 0; i < MAXC; ++i) {
        const int c = lane + 32 * i;
        if (c < nch) {
            float4 o;
            o.x = rs * (gg[i].x - c1 - xh[i].x * c2);
            o.y = rs * (gg[i].y - c1 - xh[i].y * c2);
            o.z = rs * (gg[i].z - c1 - xh[i].z * c2);
            o.w = rs * (gg[i].w - c1 - xh[i].w * c2);
            if (VAR >= 1) {
                o.x += ad[i].x; o.y += ad[i].y; o.z += ad[i].z; o.w += ad[i].w;
            } else if (dx_add) {
                const float4 a = ar[c];
                o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w;
            }
            if (dx32) {
                if (VAR == 2) __stcs(reinterpret_cast<float4*>(dx32 + row * D) + c, o);
                else reinterpret_cast<float4*>(dx32 + row * D)[c] = o;
            }
            if (dx16) reinterpret_cast<uint2*>(dx16 + row * D)[c] = make_uint2(pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
        }
    }
}

// VAR 3: two warps per row (64 lanes x <= 3 chunks): half the registers of VAR 1, twice the resident rows per SM; the two
// partial sums of a row meet in shared memory.  Every load of the row is issued before the reduction.
constexpr int LNW_ROWS = 4;     // rows per block
template <int MAXC, int WPR>
__global__ void __launch_bounds__(LNW_ROWS * WPR * 32)
layernorm_bwd_wide_kernel(const __nv_bfloat16* __restrict__ dy16, const float* __restrict__ dy32, const float* __restrict__ x,
                          const float* __restrict__ gamma, const float* __restrict__ mean, const float* __restrict__ rstd, const float* dx_add,
                          long long M, int D, float* dx32, __nv_bfloat16* __restrict__ dx16) {
    constexpr int TPR = WPR * 32;
    __shared__ float s_part[LNW_ROWS][WPR][2];
    pdl_wait();
    pdl_trigger();
    const int r = threadIdx.x / TPR, t = threadIdx.x - r * TPR, half = t >> 5;
    const long long row = static_cast<long long>(blockIdx.x) * LNW_ROWS + r;
    const bool ok = row < M;
    const int nch = D >> 2;
    const long long off = ok ? row * D : 0;
    const float4* xr = reinterpret_cast<const float4*>(x + off);
    const float4* ar = reinterpret_cast<const float4*>(dx_add ? dx_add + off : nullptr);
    const float4* g4 = reinterpret_cast<const float4*>(gamma);
    float4 gg[MAXC], xh[MAXC], ad[MAXC];
#pragma unroll
    for (int i = 0; i < MAXC; ++i) {
        const int c = t + TPR * i;
        if (ok && c < nch) {
            if (dy16) {
                const uint2 raw = __ldcs(reinterpret_cast<const uint2*>(dy16 + off) + c);
                const float2 lo = unpack_bf16x2(raw.x), hi = unpack_bf16x2(raw.y);
                gg[i] = make_float4(lo.x, lo.y, hi.x, hi.y);
            } else {
                gg[i] = __ldcs(reinterpret_cast<const float4*>(dy32 + off) + c);
            }
            xh[i] = __ldcs(xr + c);
            ad[i] = ar ? __ldcs(ar + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    const float mu = ok ? mean[row] : 0.f, rs = ok ? rstd[row] : 0.f;
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < MAXC; ++i) {
        const int c = t + TPR * i;
        if (ok && c < nch) {
            const float4 g = __ldg(g4 + c);
            const float4 d = gg[i], xv = xh[i];
            gg[i] = make_float4(d.x * g.x, d.y * g.y, d.z * g.z, d.w * g.w);
            xh[i] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
            s1 += (gg[i].x + gg[i].y) + (gg[i].z + gg[i].w);
            s2 += (gg[i].x * xh[i].x + gg[i].y * xh[i].y) + (gg[i].z * xh[i].z + gg[i].w * xh[i].w);
        }
    }
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    if ((threadIdx.x & 31) == 0) { s_part[r][half][0] = s1; s_part[r][half][1] = s2; }
    __syncthreads();
    float c1 = 0.f, c2 = 0.f;
#pragma unroll
    for (int w = 0; w < WPR; ++w) { c1 += s_part[r][w][0]; c2 += s_part[r][w][1]; }
    c1 /= static_cast<float>(D);
    c2 /= static_cast<float>(D);
#pragma unroll
    for (int i = 0; i < MAXC; ++i) {
        const int c = t + TPR * i;
        if (ok && c < nch) {
            float4 o;
            o.x = rs * (gg[i].x - c1 - xh[i].x * c2) + ad[i].x;
            o.y = rs * (gg[i].y - c1 - xh[i].y * c2) + ad[i].y;
            o.z = rs * (gg[i].z - c1 - xh[i].z * c2) + ad[i].z;
            o.w = rs * (gg[i].w - c1 - xh[i].w * c2) + ad[i].w;
            if (dx32) __stcs(reinterpret_cast<float4*>(dx32 + off) + c, o);
            if (dx16) reinterpret_cast<uint2*>(dx16 + off)[c] = make_uint2(pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
        }
    }
}

// forward, WPR warps per row (see layernorm_bwd_wide_kernel): more resident rows per SM than the one-warp form
template <int MAXC, int WPR>
__global__ void __launch_bounds__(LNW_ROWS * WPR * 32)
layernorm_fwd_wide_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta, float eps, long long M,
                          int D, float* __restrict__ y32, __nv_bfloat16* __restrict__ y16, float* __restrict__ mean_out,
                          float* __restrict__ rstd_out, int round_out) {
    constexpr int TPR = WPR * 32;
    __shared__ float s_part[LNW_ROWS][WPR][2];
    pdl_wait();
    pdl_trigger();
    const int r = threadIdx.x / TPR, t = threadIdx.x - r * TPR, w = t >> 5;
    const long long row = static_cast<long long>(blockIdx.x) * LNW_ROWS + r;
    const bool ok = row < M;
    const int nch = D >> 2;
    const long long off = ok ? row * D : 0;
    const float4* xr = reinterpret_cast<const float4*>(x + off);
    float4 v[MAXC];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < MAXC; ++i) {
        const int c = t + TPR * i;
        if (ok && c < nch) {
            v[i] = xr[c];
            sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        }
    }
    sum = warp_sum(sum);
    if ((threadIdx.x & 31) == 0) s_part[r][w][0] = sum;
    __syncthreads();
    float mean = 0.f;
#pragma unroll
    for (int k = 0; k < WPR; ++k) mean += s_part[r][k][0];
    mean /= static_cast<float>(D);
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < MAXC; ++i) {
        const int c = t + TPR * i;
        if (ok && c < nch) {
            const float a = v[i].x - mean, b = v[i].y - mean, cc = v[i].z - mean, d = v[i].w - mean;
            sq += (a * a + b * b) + (cc * cc + d * d);
        }
    }
    sq = warp_sum(sq);
    if ((threadIdx.x & 31) == 0) s_part[r][w][1] = sq;
    __syncthreads();
    float var = 0.f;
#pragma unroll
    for (int k = 0; k < WPR; ++k) var += s_part[r][k][1];
    const float rstd = 1.0f / sqrtf(var / static_cast<float>(D) + eps);
    if (ok && t == 0) {
        if (mean_out) mean_out[row] = mean;
        if (rstd_out) rstd_out[row] = rstd;
    }
    const float4* g4 = reinterpret_cast<const float4*>(gamma);
    const float4* b4 = reinterpret_cast<const float4*>(beta);
#pragma unroll
    for (int i = 0; i < MAXC; ++i) {
        const int c = t + TPR * i;
        if (ok && c < nch) {
            const float4 g = __ldg(g4 + c), b = __ldg(b4 + c);
            float4 o;
            o.x = (v[i].x - mean) * rstd * g.x + b.x;
            o.y = (v[i].y - mean) * rstd * g.y + b.y;
            o.z = (v[i].z - mean) * rstd * g.z + b.z;
            o.w = (v[i].w - mean) * rstd * g.w + b.w;
            if (y32)
                reinterpret_cast<float4*>(y32 + off)[c] =
                    (round_out & 1) ? make_float4(round_tf32_rn(o.x), round_tf32_rn(o.y), round_tf32_rn(o.z), round_tf32_rn(o.w)) : o;
            if (y16) reinterpret_cast<uint2*>(y16 + off)[c] = (round_out & 2) ? make_uint2(pack_f16x2(o.x, o.y), pack_f16x2(o.z, o.w)) : make_uint2(pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
        }
    }
}

// D = 64 (the CLIPSeg decoder, reduce_dim = 64): a row is ONE float4 per lane of a half-warp, so a warp-per-row kernel idles
// half its lanes and is pure launch / latency (13-17 us for 4 MB; twelve such launches per step).  Here a half-warp owns a
// row, a warp walks 2 x SM_ROWS rows with every load issued up front, reductions stay inside the 16 lanes.
constexpr int SM_ROWS = 4;       // row pairs per warp
__device__ __forceinline__ float half_warp_sum(float v) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void __launch_bounds__(256)
layernorm64_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta, float eps, long long M,
                       float* __restrict__ y32, __nv_bfloat16* __restrict__ y16, float* __restrict__ mean_out, float* __restrict__ rstd_out, int round_out) {
    pdl_wait();
    pdl_trigger();
    const int lane = threadIdx.x & 31, sub = lane >> 4, c = lane & 15;
    const long long base = (static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5)) * (2 * SM_ROWS) + sub;
    float4 v[SM_ROWS];
#pragma unroll
    for (int i = 0; i < SM_ROWS; ++i) {
        const long long row = base + 2 * i;
        v[i] = row < M ? __ldcs(reinterpret_cast<const float4*>(x + row * 64) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c), b = __ldg(reinterpret_cast<const float4*>(beta) + c);
#pragma unroll
    for (int i = 0; i < SM_ROWS; ++i) {
        const long long row = base + 2 * i;
        const float mean = half_warp_sum((v[i].x + v[i].y) + (v[i].z + v[i].w)) * (1.0f / 64.0f);
        const float a0 = v[i].x - mean, a1 = v[i].y - mean, a2 = v[i].z - mean, a3 = v[i].w - mean;
        const float rstd = 1.0f / sqrtf(half_warp_sum((a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3)) * (1.0f / 64.0f) + eps);
        if (row < M) {
            if (c == 0) {
                if (mean_out) mean_out[row] = mean;
                if (rstd_out) rstd_out[row] = rstd;
            }
            float4 o = make_float4(a0 * rstd * g.x + b.x, a1 * rstd * g.y + b.y, a2 * rstd * g.z + b.z, a3 * rstd * g.w + b.w);
            if (y32)
                reinterpret_cast<float4*>(y32 + row * 64)[c] =
                    (round_out & 1) ? make_float4(round_tf32_rn(o.x), round_tf32_rn(o.y), round_tf32_rn(o.z), round_tf32_rn(o.w)) : o;
            if (y16)
                reinterpret_cast<uint2*>(y16 + row * 64)[c] = (round_out & 2) ? make_uint2(pack_f16x2(o.x, o.y), pack_f16x2(o.z, o.w))
                                                                         : make_uint2(pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
        }
    }
}

__global__ void __launch_bounds__(256)
layernorm64_bwd_kernel(const __nv_bfloat16* __restrict__ dy16, const float* __restrict__ dy32, const float* __restrict__ x,
                       const float* __restrict__ gamma, const float* __restrict__ mean, const float* __restrict__ rstd, const float* dx_add,
                       long long M, float* dx32, __nv_bfloat16* __restrict__ dx16) {
    pdl_wait();
    pdl_trigger();
    const int lane = threadIdx.x & 31, sub = lane >> 4, c = lane & 15;
    const long long base = (static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5)) * (2 * SM_ROWS) + sub;
    float4 gg[SM_ROWS], xv[SM_ROWS], ad[SM_ROWS];
    float mu[SM_ROWS], rs[SM_ROWS];
#pragma unroll
    for (int i = 0; i < SM_ROWS; ++i) {
        const long long row = base + 2 * i;
        const bool ok = row < M;
        if (dy16) {
            const uint2 raw = ok ? reinterpret_cast<const uint2*>(dy16 + row * 64)[c] : make_uint2(0u, 0u);
            const float2 lo = unpack_bf16x2(raw.x), hi = unpack_bf16x2(raw.y);
            gg[i] = make_float4(lo.x, lo.y, hi.x, hi.y);
        } else {
            gg[i] = ok ? reinterpret_cast<const float4*>(dy32 + row * 64)[c] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        xv[i] = ok ? reinterpret_cast<const float4*>(x + row * 64)[c] : make_float4(0.f, 0.f, 0.f, 0.f);
        ad[i] = (ok && dx_add) ? reinterpret_cast<const float4*>(dx_add + row * 64)[c] : make_float4(0.f, 0.f, 0.f, 0.f);
        mu[i] = ok ? mean[row] : 0.f;
        rs[i] = ok ? rstd[row] : 0.f;
    }
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c);
#pragma unroll
    for (int i = 0; i < SM_ROWS; ++i) {
        const long long row = base + 2 * i;
        const float4 d = make_float4(gg[i].x * g.x, gg[i].y * g.y, gg[i].z * g.z, gg[i].w * g.w);
        const float4 xh = make_float4((xv[i].x - mu[i]) * rs[i], (xv[i].y - mu[i]) * rs[i], (xv[i].z - mu[i]) * rs[i], (xv[i].w - mu[i]) * rs[i]);
        const float c1 = half_warp_sum((d.x + d.y) + (d.z + d.w)) * (1.0f / 64.0f);
        const float c2 = half_warp_sum((d.x * xh.x + d.y * xh.y) + (d.z * xh.z + d.w * xh.w)) * (1.0f / 64.0f);
        if (row < M) {
            float4 o = make_float4(rs[i] * (d.x - c1 - xh.x * c2) + ad[i].x, rs[i] * (d.y - c1 - xh.y * c2) + ad[i].y,
                                   rs[i] * (d.z - c1 - xh.z * c2) + ad[i].z, rs[i] * (d.w - c1 - xh.w * c2) + ad[i].w);
            if (dx32) reinterpret_cast<float4*>(dx32 + row * 64)[c] = o;
            if (dx16) reinterpret_cast<uint2*>(dx16 + row * 64)[c] = make_uint2(pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
        }
    }
}

}  // namespace tvs

extern "C" __attribute__((visibility("default"))) int tvs_layernorm_fwd(const float* x, const float* gamma, const float* beta, float eps, int64_t M, int32_t D, float* y_f32,
                                 void* y_bf16, float* mean, float* rstd, int32_t round_tf32, void* stream) {
    using namespace tvs;
    TVS_REQUIRE(x && gamma && beta && (y_f32 || y_bf16), "tvs_layernorm_fwd: null pointer");
    TVS_REQUIRE(M > 0 && D > 0 && D % 4 == 0 && D <= 128 * LN_MAXC, "tvs_layernorm_fwd: D=%d must be a multiple of 4 and <= %d", D, 128 * LN_MAXC);
    const unsigned grid = static_cast<unsigned>((M + LN_WARPS - 1) / LN_WARPS);
    static const bool small_off = [] { const char* e = getenv("TVS_LN_SMALL"); return e && e[0] == '0'; }();     // TVS_LN_SMALL=0: generic kernels for D = 64 too
    if (D == 64 && M >= 512 && !small_off) {
        const unsigned g64 = static_cast<unsigned>((M + 8 * 2 * SM_ROWS - 1) / (8 * 2 * SM_ROWS));
        TVS_CUDA(launch_pdl(layernorm64_fwd_kernel, dim3(g64), dim3(256), 0, static_cast<cudaStream_t>(stream), 1, x, gamma, beta, eps,
                            static_cast<long long>(M), y_f32, static_cast<__nv_bfloat16*>(y_bf16), mean, rstd, round_tf32));
        return check_launch("layernorm64_fwd_kernel");
    }
    static const int variant = [] { const char* e = getenv("TVS_LN_FWD"); return e ? atoi(e) : 1; }();   // 1: warp-per-row kernels (two rows per warp for large M); 0: one row per warp always; 2 / 3: warps per row
    if (variant >= 2 && D <= 768 && D > 128) {
        const unsigned gridw = static_cast<unsigned>((M + LNW_ROWS - 1) / LNW_ROWS);
        if (variant == 2)
            TVS_CUDA(launch_pdl(layernorm_fwd_wide_kernel<3, 2>, dim3(gridw), dim3(LNW_ROWS * 64), 0, static_cast<cudaStream_t>(stream), 1, x, gamma, beta,
                                eps, static_cast<long long>(M), D, y_f32, static_cast<__nv_bfloat16*>(y_bf16), mean, rstd, round_tf32));
        else
            TVS_CUDA(launch_pdl(layernorm_fwd_wide_kernel<2, 3>, dim3(gridw), dim3(LNW_ROWS * 96), 0, static_cast<cudaStream_t>(stream), 1, x, gamma, beta,
                                eps, static_cast<long long>(M), D, y_f32, static_cast<__nv_bfloat16*>(y_bf16), mean, rstd, round_tf32));
    } else if (D <= 768 && M >= 8192 && variant == 4) {      // persistent, software-pipelined (one row in flight per warp)
        static const int ctas_per_sm = [] { const char* e = getenv("TVS_LN_PIPE_CTAS"); return e ? atoi(e) : 4; }();
        const long long want = (static_cast<long long>(M) + LN2_WARPS - 1) / LN2_WARPS;
        const long long cap = static_cast<long long>(sm_count()) * ctas_per_sm;
        TVS_CUDA(launch_pdl(layernorm_fwd_pipe_kernel<6>, dim3(static_cast<unsigned>(want < cap ? want : cap)), dim3(LN2_WARPS * 32), 0,
                            static_cast<cudaStream_t>(stream), 1, x, gamma, beta, eps, static_cast<long long>(M), D, y_f32,
                            static_cast<__nv_bfloat16*>(y_bf16), mean, rstd, round_tf32));
    } else if (D <= 768 && M >= 2048 && (variant == 1 || variant == 4)) {      // large row counts: two rows per warp
        const unsigned grid2 = static_cast<unsigned>((M + 2 * LN2_WARPS - 1) / (2 * LN2_WARPS));
        TVS_CUDA(launch_pdl(layernorm_fwd2_kernel<6>, dim3(grid2), dim3(LN2_WARPS * 32), 0, static_cast<cudaStream_t>(stream), 1, x, gamma, beta, eps,
                            static_cast<long long>(M), D, y_f32, static_cast<__nv_bfloat16*>(y_bf16), mean, rstd, round_tf32));
    } else if (D <= 1024)
        TVS_CUDA(launch_pdl(layernorm_fwd_kernel<8>, dim3(grid), dim3(LN_WARPS * 32), 0, static_cast<cudaStream_t>(stream), 1, x, gamma, beta, eps,
                            static_cast<long long>(M), D, y_f32, static_cast<__nv_bfloat16*>(y_bf16), mean, rstd, round_tf32));
    else
        TVS_CUDA(launch_pdl(layernorm_fwd_kernel<16>, dim3(grid), dim3(LN_WARPS * 32), 0, static_cast<cudaStream_t>(stream), 1, x, gamma, beta, eps,
                            static_cast<long long>(M), D, y_f32, static_cast<__nv_bfloat16*>(y_bf16), mean, rstd, round_tf32));
    return check_launch("layernorm_fwd_kernel");
}

extern "C" __attribute__((visibility("default"))) int tvs_layernorm_bwd(const void* dy_bf16, const float* dy_f32, const float* x, const float* gamma, const float* mean,
                                 const float* rstd, const float* dx_add_f32, int64_t M, int32_t D, float* dx_out_f32, void* dx_out_bf16,
                                 void* stream) {
    using namespace tvs;
    TVS_REQUIRE((dy_bf16 != nullptr) != (dy_f32 != nullptr), "tvs_layernorm_bwd: exactly one of dy_bf16 / dy_f32");
    TVS_REQUIRE(x && gamma && mean && rstd && (dx_out_f32 || dx_out_bf16), "tvs_layernorm_bwd: null pointer");
    TVS_REQUIRE(M > 0 && D > 0 && D % 4 == 0 && D <= 128 * LN_MAXC, "tvs_layernorm_bwd: D=%d must be a multiple of 4 and <= %d", D, 128 * LN_MAXC);
    const unsigned grid = static_cast<unsigned>((M + LN_WARPS - 1) / LN_WARPS);
    static const bool small_off = [] { const char* e = getenv("TVS_LN_SMALL"); return e && e[0] == '0'; }();
    if (D == 64 && M >= 512 && !small_off && dx_add_f32 != dx_out_f32) {
        const unsigned g64 = static_cast<unsigned>((M + 8 * 2 * SM_ROWS - 1) / (8 * 2 * SM_ROWS));
        TVS_CUDA(launch_pdl(layernorm64_bwd_kernel, dim3(g64), dim3(256), 0, static_cast<cudaStream_t>(stream), 1,
                            static_cast<const __nv_bfloat16*>(dy_bf16), dy_f32, x, gamma, mean, rstd, dx_add_f32, static_cast<long long>(M),
                            dx_out_f32, static_cast<__nv_bfloat16*>(dx_out_bf16)));
        return check_launch("layernorm64_bwd_kernel");
    }
    static const int variant = [] { const char* e = getenv("TVS_LN_BWD"); return e ? atoi(e) : 4; }();   // see the kernel comments
#define TVS_LN_BWD_LAUNCH(MAXC, VAR)                                                                                                      \
    TVS_CUDA(launch_pdl(layernorm_bwd_kernel<MAXC, VAR>, dim3(grid), dim3(LN_WARPS * 32), 0, static_cast<cudaStream_t>(stream), 1,        \
                        static_cast<const __nv_bfloat16*>(dy_bf16), dy_f32, x, gamma, mean, rstd, dx_add_f32, static_cast<long long>(M), D, \
                        dx_out_f32, static_cast<__nv_bfloat16*>(dx_out_bf16)))
    if (variant >= 3 && D <= 768 && D > 128) {
        const unsigned gridw = static_cast<unsigned>((M + LNW_ROWS - 1) / LNW_ROWS);
        if (variant == 3)
            TVS_CUDA(launch_pdl(layernorm_bwd_wide_kernel<3, 2>, dim3(gridw), dim3(LNW_ROWS * 64), 0, static_cast<cudaStream_t>(stream), 1,
                                static_cast<const __nv_bfloat16*>(dy_bf16), dy_f32, x, gamma, mean, rstd, dx_add_f32, static_cast<long long>(M), D,
                                dx_out_f32, static_cast<__nv_bfloat16*>(dx_out_bf16)));
        else
            TVS_CUDA(launch_pdl(layernorm_bwd_wide_kernel<2, 3>, dim3(gridw), dim3(LNW_ROWS * 96), 0, static_cast<cudaStream_t>(stream), 1,
                                static_cast<const __nv_bfloat16*>(dy_bf16), dy_f32, x, gamma, mean, rstd, dx_add_f32, static_cast<long long>(M), D,
                                dx_out_f32, static_cast<__nv_bfloat16*>(dx_out_bf16)));
    } else if (D <= 768) {
        if (variant == 0) TVS_LN_BWD_LAUNCH(6, 0); else if (variant == 1) TVS_LN_BWD_LAUNCH(6, 1); else TVS_LN_BWD_LAUNCH(6, 2);
    } else if (D <= 1024) {
        if (variant == 0) TVS_LN_BWD_LAUNCH(8, 0); else if (variant == 1) TVS_LN_BWD_LAUNCH(8, 1); else TVS_LN_BWD_LAUNCH(8, 2);
    } else {
        if (variant == 0) TVS_LN_BWD_LAUNCH(16, 0); else TVS_LN_BWD_LAUNCH(16, 1);
    }
#undef TVS_LN_BWD_LAUNCH
    return check_launch("layernorm_bwd_kernel");
}
