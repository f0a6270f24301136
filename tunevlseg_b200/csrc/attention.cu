// Fused (flash-style) multi-head self-attention, forward and backward, for the CLIPSeg towers and decoder.
//
// Replaces eager_attention_forward (transformers modeling_clipseg.py:256-276): softmax(q k^T * d^-0.5 + mask) v
// with the softmax in fp32.  The d^-0.5 scale is folded into Wq/bq by the host (exact: d is 64 or 16).
//
// Round-1 implementation: warp-level mma.sync m16n8k16 (bf16 in, fp32 accumulate), 64x64 tiles, K/V streamed
// through a double-buffered cp.async ring, online softmax in registers; scores never touch HBM.  The backward
// is two kernels (dK/dV per key tile, dQ per query tile) so that no atomics are needed; both recompute P from
// the saved log-sum-exp.  [A tcgen05/TMEM version of these kernels is the planned replacement.]
//
// layout: qkv bf16 [B*S, 3*E], E = H*HD, columns Q | K | V, head h at columns h*HD .. h*HD+HD-1 of each part.
#include <stdlib.h>

#include "common.cuh"
#include "tvs_b200.h"

namespace tvs {

constexpr int ATT_BM = 64;
constexpr int ATT_BN = 64;
constexpr int ATT_THREADS = 128;
constexpr float LOG2E = 1.4426950408889634f;

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
    int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ float ex2_fast(float x) {   // one MUFU.EX2; ex2(-inf) = 0
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// smem tile [64 rows][HD] bf16, 16-byte chunks XOR-swizzled by row so ldmatrix is conflict free for HD = 64
template <int HD>
__device__ __forceinline__ uint32_t tile_addr(uint32_t base, int row, int chunk) {
    return base + row * (HD * 2) + ((chunk ^ (row & (HD / 8 - 1))) << 4);
}

// global [rows x HD] slab (row stride ld elements) -> smem tile; rows >= S are zero filled
template <int HD>
__device__ __forceinline__ void load_tile(uint32_t smem_base, const __nv_bfloat16* gbase, long long ld, int row0, int S) {
    constexpr int CH = HD / 8;
    for (int idx = threadIdx.x; idx < 64 * CH; idx += ATT_THREADS) {
        int r = idx / CH, c = idx % CH;
        bool valid = (row0 + r) < S;
        const __nv_bfloat16* src = gbase + static_cast<long long>(valid ? row0 + r : 0) * ld + c * 8;
        cp_async16(tile_addr<HD>(smem_base, r, c), src, valid);
    }
}

// A fragments (16 rows x HD) of this warp's 16-row slice of a tile
template <int HD>
__device__ __forceinline__ void load_a_frags(uint32_t smem_base, int warp_row0, uint32_t (&a)[HD / 16][4]) {
    const int lane = threadIdx.x & 31;
    const int r = warp_row0 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
    for (int ks = 0; ks < HD / 16; ++ks) ldsm_x4(tile_addr<HD>(smem_base, r, ks * 2 + (lane >> 4)), a[ks][0], a[ks][1], a[ks][2], a[ks][3]);
}

// acc[nb] (16 x 64, 8 n-blocks) += A(16 x HD) * T^T where T is a [64][HD] tile (B operand read non-transposed)
template <int HD>
__device__ __forceinline__ void mma_a_tileT(float (&acc)[8][4], const uint32_t (&a)[HD / 16][4], uint32_t tile_base) {
    const int lane = threadIdx.x & 31;
    const int mi = lane >> 3;
#pragma unroll
    for (int nbp = 0; nbp < 4; ++nbp) {
        const int row = nbp * 16 + (mi >> 1) * 8 + (lane & 7);
#pragma unroll
        for (int ks = 0; ks < HD / 16; ++ks) {
            uint32_t b0, b1, b2, b3;
            ldsm_x4(tile_addr<HD>(tile_base, row, ks * 2 + (mi & 1)), b0, b1, b2, b3);
            mma_bf16(acc[2 * nbp], a[ks], b0, b1);
            mma_bf16(acc[2 * nbp + 1], a[ks], b2, b3);
        }
    }
}

// out[nb] (16 x HD, HD/8 n-blocks) += P(16 x 64, as A fragments p[4][4]) * T where T is a [64][HD] tile (B read transposed)
template <int HD>
__device__ __forceinline__ void mma_p_tile(float (&out)[HD / 8][4], const uint32_t (&p)[4][4], uint32_t tile_base) {
    const int lane = threadIdx.x & 31;
    const int mi = lane >> 3;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {  // 16 rows of T per step
        const int row = ks * 16 + (mi & 1) * 8 + (lane & 7);
#pragma unroll
        for (int nbp = 0; nbp < HD / 16; ++nbp) {
            uint32_t b0, b1, b2, b3;
            ldsm_x4_t(tile_addr<HD>(tile_base, row, nbp * 2 + (mi >> 1)), b0, b1, b2, b3);
            mma_bf16(out[2 * nbp], p[ks], b0, b1);
            mma_bf16(out[2 * nbp + 1], p[ks], b2, b3);
        }
    }
}

__device__ __forceinline__ void acc_to_afrag(const float (&s)[8][4], uint32_t (&p)[4][4]) {
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        p[t][0] = pack_bf16x2(s[2 * t][0], s[2 * t][1]);
        p[t][1] = pack_bf16x2(s[2 * t][2], s[2 * t][3]);
        p[t][2] = pack_bf16x2(s[2 * t + 1][0], s[2 * t + 1][1]);
        p[t][3] = pack_bf16x2(s[2 * t + 1][2], s[2 * t + 1][3]);
    }
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
// PLAIN: no causal mask and no key mask (the decoder's self-attention) - the per-element mask logic (a dozen integer
// instructions per score) disappears from every full key tile; with head dim 16 those instructions, not the MMAs, were
// the kernel (issue bound: 48 us for 2 GFLOP).
template <int HD, bool PLAIN>
__global__ void __launch_bounds__(ATT_THREADS)
attn_fwd_kernel(const __nv_bfloat16* __restrict__ qkv, int S, int H, int causal, const uint8_t* __restrict__ key_mask,
                __nv_bfloat16* __restrict__ out, float* __restrict__ out32, float* __restrict__ lse) {
    __shared__ __align__(128) __nv_bfloat16 sQ[64 * HD];
    __shared__ __align__(128) __nv_bfloat16 sK[2][64 * HD];
    __shared__ __align__(128) __nv_bfloat16 sV[2][64 * HD];
    const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int E = H * HD;
    const long long ld = 3LL * E;
    const __nv_bfloat16* qb = qkv + static_cast<long long>(b) * S * ld + h * HD;
    const __nv_bfloat16* kb = qb + E;
    const __nv_bfloat16* vb = qb + 2 * E;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, tq = lane & 3;
    const int q0 = qt * ATT_BM;
    const uint32_t sq = static_cast<uint32_t>(__cvta_generic_to_shared(sQ));
    const uint32_t sk[2] = {static_cast<uint32_t>(__cvta_generic_to_shared(sK[0])), static_cast<uint32_t>(__cvta_generic_to_shared(sK[1]))};
    const uint32_t sv[2] = {static_cast<uint32_t>(__cvta_generic_to_shared(sV[0])), static_cast<uint32_t>(__cvta_generic_to_shared(sV[1]))};

    int n_tiles = (S + ATT_BN - 1) / ATT_BN;
    if (causal) n_tiles = min(n_tiles, qt + 1);

    load_tile<HD>(sq, qb, ld, q0, S);
    load_tile<HD>(sk[0], kb, ld, 0, S);
    load_tile<HD>(sv[0], vb, ld, 0, S);
    cp_async_commit();

    uint32_t qa[HD / 16][4];
    float o[HD / 8][4];
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
    float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
    const uint8_t* km = key_mask ? key_mask + static_cast<long long>(b) * S : nullptr;

    for (int t = 0; t < n_tiles; ++t) {
        const int buf = t & 1;
        if (t + 1 < n_tiles) {
            load_tile<HD>(sk[buf ^ 1], kb, ld, (t + 1) * ATT_BN, S);
            load_tile<HD>(sv[buf ^ 1], vb, ld, (t + 1) * ATT_BN, S);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        if (t == 0) load_a_frags<HD>(sq, warp * 16, qa);

        float s[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
        mma_a_tileT<HD>(s, qa, sk[buf]);

        // mask + online softmax (rows g and g+8 of this warp's 16)
        const int k0 = t * ATT_BN;
        float mx[2] = {-INFINITY, -INFINITY};
        if (PLAIN && k0 + ATT_BN <= S) {
#pragma unroll
            for (int nb = 0; nb < 8; ++nb) {
                mx[0] = fmaxf(mx[0], fmaxf(s[nb][0], s[nb][1]));
                mx[1] = fmaxf(mx[1], fmaxf(s[nb][2], s[nb][3]));
            }
        } else {
#pragma unroll
            for (int nb = 0; nb < 8; ++nb) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int key = k0 + nb * 8 + 2 * tq + (e & 1);
                    const int q = q0 + warp * 16 + g + (e >> 1) * 8;
                    bool ok = key < S && (PLAIN || !causal || key <= q);
                    if (!PLAIN && ok && km) ok = km[key] != 0;
                    s[nb][e] = ok ? s[nb][e] : -INFINITY;
                    mx[e >> 1] = fmaxf(mx[e >> 1], s[nb][e]);
                }
            }
        }
        float m_use[2], corr[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
            const float m_new = fmaxf(m_run[r], mx[r]);
            m_use[r] = (m_new == -INFINITY) ? 0.f : m_new;
            corr[r] = exp2f((m_run[r] - m_use[r]) * LOG2E);
            m_run[r] = m_new;
            l_run[r] *= corr[r];
        }
        const float mL[2] = {m_use[0] * LOG2E, m_use[1] * LOG2E};
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float p = PLAIN ? ex2_fast(fmaf(s[nb][e], LOG2E, -mL[e >> 1])) : exp2f((s[nb][e] - m_use[e >> 1]) * LOG2E);
                s[nb][e] = p;
                l_run[e >> 1] += p;
            }
        }
#pragma unroll
        for (int i = 0; i < HD / 8; ++i) {
            o[i][0] *= corr[0]; o[i][1] *= corr[0]; o[i][2] *= corr[1]; o[i][3] *= corr[1];
        }
        uint32_t p[4][4];
        acc_to_afrag(s, p);
        mma_p_tile<HD>(o, p, sv[buf]);
        __syncthreads();
    }

#pragma unroll
    for (int r = 0; r < 2; ++r) {
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int q = q0 + warp * 16 + g + r * 8;
        if (q < S) {
            const float inv = l_run[r] > 0.f ? 1.0f / l_run[r] : 0.f;
            __nv_bfloat16* orow = out + (static_cast<long long>(b) * S + q) * E + h * HD;
#pragma unroll
            for (int nb = 0; nb < HD / 8; ++nb)
                *reinterpret_cast<uint32_t*>(orow + nb * 8 + 2 * tq) = pack_bf16x2(o[nb][2 * r] * inv, o[nb][2 * r + 1] * inv);
            if (out32) {
                float* frow = out32 + (static_cast<long long>(b) * S + q) * E + h * HD;
#pragma unroll
                for (int nb = 0; nb < HD / 8; ++nb)
                    *reinterpret_cast<float2*>(frow + nb * 8 + 2 * tq) =
                        make_float2(round_tf32_rn(o[nb][2 * r] * inv), round_tf32_rn(o[nb][2 * r + 1] * inv));      // feeds a tf32 GEMM
            }
            if (tq == 0) lse[(static_cast<long long>(b) * H + h) * S + q] = (m_run[r] == -INFINITY ? 0.f : m_run[r]) + logf(fmaxf(l_run[r], 1e-30f));
        }
    }
}

// ---------------------------------------------------------------------------------------------
// backward: delta = rowsum(dO * O)
// ---------------------------------------------------------------------------------------------
template <int HD>
__global__ void attn_delta_kernel(const __nv_bfloat16* __restrict__ out, const __nv_bfloat16* __restrict__ dout, int S, int H,
                                  long long rows, float* __restrict__ delta, int o_f16) {
    constexpr int CH = HD / 8;  // 16-byte chunks (lanes) per head
    const long long row = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    const int E = H * HD;
    const long long b = row / S, q = row % S;
    // E/8 is a multiple of CH and CH divides 32, so a head never straddles a pass; uniform trip count for the shuffles
    for (int c0 = 0; c0 < E / 8; c0 += 32) {
        const int c = c0 + lane;
        const bool active = c < E / 8;
        float acc = 0.f;
        if (active) {
            const uint4 a = *reinterpret_cast<const uint4*>(out + row * E + c * 8);
            const uint4 d = *reinterpret_cast<const uint4*>(dout + row * E + c * 8);
            const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, dw[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float2 x = o_f16 ? unpack_f16x2(aw[i]) : unpack_bf16x2(aw[i]), y = unpack_bf16x2(dw[i]);
                acc += x.x * y.x + x.y * y.y;
            }
        }
#pragma unroll
        for (int o = CH / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (active && (c % CH) == 0) delta[(b * H + c / CH) * S + q] = acc;
    }
}

// ---------------------------------------------------------------------------------------------
// backward: dQ (one CTA per query tile, loops over key tiles)
// ---------------------------------------------------------------------------------------------
template <int HD, bool PLAIN>
__global__ void __launch_bounds__(ATT_THREADS)
attn_bwd_dq_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ dout, const float* __restrict__ lse,
                   const float* __restrict__ delta, int S, int H, int causal, const uint8_t* __restrict__ key_mask,
                   __nv_bfloat16* __restrict__ dqkv) {
    __shared__ __align__(128) __nv_bfloat16 sQ[64 * HD];   // Q, then reused for dO
    __shared__ __align__(128) __nv_bfloat16 sK[2][64 * HD];
    __shared__ __align__(128) __nv_bfloat16 sV[2][64 * HD];
    const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int E = H * HD;
    const long long ld = 3LL * E;
    const __nv_bfloat16* qb = qkv + static_cast<long long>(b) * S * ld + h * HD;
    const __nv_bfloat16* kb = qb + E;
    const __nv_bfloat16* vb = qb + 2 * E;
    const __nv_bfloat16* dob = dout + static_cast<long long>(b) * S * E + h * HD;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, tq = lane & 3;
    const int q0 = qt * ATT_BM;
    const uint32_t sq = static_cast<uint32_t>(__cvta_generic_to_shared(sQ));
    const uint32_t sk[2] = {static_cast<uint32_t>(__cvta_generic_to_shared(sK[0])), static_cast<uint32_t>(__cvta_generic_to_shared(sK[1]))};
    const uint32_t sv[2] = {static_cast<uint32_t>(__cvta_generic_to_shared(sV[0])), static_cast<uint32_t>(__cvta_generic_to_shared(sV[1]))};

    int n_tiles = (S + ATT_BN - 1) / ATT_BN;
    if (causal) n_tiles = min(n_tiles, qt + 1);

    uint32_t qa[HD / 16][4], da[HD / 16][4];
    load_tile<HD>(sq, qb, ld, q0, S);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    load_a_frags<HD>(sq, warp * 16, qa);
    __syncthreads();
    load_tile<HD>(sq, dob, E, q0, S);
    load_tile<HD>(sk[0], kb, ld, 0, S);
    load_tile<HD>(sv[0], vb, ld, 0, S);
    cp_async_commit();

    float row_lse[2], row_delta[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int q = q0 + warp * 16 + g + r * 8;
        const long long idx = (static_cast<long long>(b) * H + h) * S + q;
        row_lse[r] = q < S ? lse[idx] : INFINITY;
        row_delta[r] = q < S ? delta[idx] : 0.f;
    }
    const float lseL[2] = {row_lse[0] * LOG2E, row_lse[1] * LOG2E};
    float dq[HD / 8][4];
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f;
    const uint8_t* km = key_mask ? key_mask + static_cast<long long>(b) * S : nullptr;

    for (int t = 0; t < n_tiles; ++t) {
        const int buf = t & 1;
        if (t + 1 < n_tiles) {
            load_tile<HD>(sk[buf ^ 1], kb, ld, (t + 1) * ATT_BN, S);
            load_tile<HD>(sv[buf ^ 1], vb, ld, (t + 1) * ATT_BN, S);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        if (t == 0) load_a_frags<HD>(sq, warp * 16, da);

        float s[8][4], dp[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
            dp[i][0] = dp[i][1] = dp[i][2] = dp[i][3] = 0.f;
        }
        mma_a_tileT<HD>(s, qa, sk[buf]);
        mma_a_tileT<HD>(dp, da, sv[buf]);
        const int k0 = t * ATT_BN;
        if (PLAIN && k0 + ATT_BN <= S) {
#pragma unroll
            for (int nb = 0; nb < 8; ++nb) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float p = ex2_fast(fmaf(s[nb][e], LOG2E, -lseL[e >> 1]));      // lse = +inf for q >= S -> 0
                    s[nb][e] = p * (dp[nb][e] - row_delta[e >> 1]);
                }
            }
        } else {
#pragma unroll
            for (int nb = 0; nb < 8; ++nb) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int key = k0 + nb * 8 + 2 * tq + (e & 1);
                    const int q = q0 + warp * 16 + g + (e >> 1) * 8;
                    bool ok = key < S && (PLAIN || !causal || key <= q);
                    if (!PLAIN && ok && km) ok = km[key] != 0;
                    const float p = ok ? exp2f((s[nb][e] - row_lse[e >> 1]) * LOG2E) : 0.f;
                    s[nb][e] = p * (dp[nb][e] - row_delta[e >> 1]);
                }
            }
        }
        uint32_t ds[4][4];
        acc_to_afrag(s, ds);
        mma_p_tile<HD>(dq, ds, sk[buf]);
        __syncthreads();
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int q = q0 + warp * 16 + g + r * 8;
        if (q < S) {
            __nv_bfloat16* drow = dqkv + (static_cast<long long>(b) * S + q) * ld + h * HD;
#pragma unroll
            for (int nb = 0; nb < HD / 8; ++nb)
                *reinterpret_cast<uint32_t*>(drow + nb * 8 + 2 * tq) = pack_bf16x2(dq[nb][2 * r], dq[nb][2 * r + 1]);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// backward: dK, dV (one CTA per key tile, loops over query tiles; works on the transposed problem)
// ---------------------------------------------------------------------------------------------
template <int HD, bool PLAIN>
__global__ void __launch_bounds__(ATT_THREADS)
attn_bwd_dkv_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ dout, const float* __restrict__ lse,
                    const float* __restrict__ delta, int S, int H, int causal, const uint8_t* __restrict__ key_mask,
                    __nv_bfloat16* __restrict__ dqkv) {
    __shared__ __align__(128) __nv_bfloat16 sA[2][64 * HD];   // Q tiles (K tile first)
    __shared__ __align__(128) __nv_bfloat16 sB[2][64 * HD];   // dO tiles (V tile first)
    __shared__ float sLse[2][64], sDelta[2][64];
    const int kt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int E = H * HD;
    const long long ld = 3LL * E;
    const __nv_bfloat16* qb = qkv + static_cast<long long>(b) * S * ld + h * HD;
    const __nv_bfloat16* kb = qb + E;
    const __nv_bfloat16* vb = qb + 2 * E;
    const __nv_bfloat16* dob = dout + static_cast<long long>(b) * S * E + h * HD;
    const float* lse_b = lse + (static_cast<long long>(b) * H + h) * S;
    const float* delta_b = delta + (static_cast<long long>(b) * H + h) * S;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, tq = lane & 3;
    const int k0 = kt * ATT_BN;
    const uint32_t sa[2] = {static_cast<uint32_t>(__cvta_generic_to_shared(sA[0])), static_cast<uint32_t>(__cvta_generic_to_shared(sA[1]))};
    const uint32_t sb[2] = {static_cast<uint32_t>(__cvta_generic_to_shared(sB[0])), static_cast<uint32_t>(__cvta_generic_to_shared(sB[1]))};

    const int n_qt = (S + ATT_BM - 1) / ATT_BM;
    const int t_begin = causal ? kt : 0;   // queries before this key tile never see it

    // stage K and V of this tile, pull them into A fragments
    uint32_t ka[HD / 16][4], va[HD / 16][4];
    load_tile<HD>(sa[0], kb, ld, k0, S);
    load_tile<HD>(sb[0], vb, ld, k0, S);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    load_a_frags<HD>(sa[0], warp * 16, ka);
    load_a_frags<HD>(sb[0], warp * 16, va);
    __syncthreads();

    auto stage = [&](int t, int buf) {
        load_tile<HD>(sa[buf], qb, ld, t * ATT_BM, S);
        load_tile<HD>(sb[buf], dob, E, t * ATT_BM, S);
        if (threadIdx.x < 64) {
            const int q = t * ATT_BM + threadIdx.x;
            sLse[buf][threadIdx.x] = q < S ? (PLAIN ? lse_b[q] * LOG2E : lse_b[q]) : INFINITY;
            sDelta[buf][threadIdx.x] = q < S ? delta_b[q] : 0.f;
        }
    };

    float dk[HD / 8][4], dv[HD / 8][4];
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) {
        dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.f;
        dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f;
    }
    bool key_ok[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int key = k0 + warp * 16 + g + r * 8;
        key_ok[r] = key < S && (!key_mask || key_mask[static_cast<long long>(b) * S + key] != 0);
    }

    if (t_begin < n_qt) {
        stage(t_begin, 0);
        cp_async_commit();
    }
    for (int t = t_begin; t < n_qt; ++t) {
        const int buf = (t - t_begin) & 1;
        if (t + 1 < n_qt) {
            stage(t + 1, buf ^ 1);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();

        float st[8][4], dpt[8][4];   // S^T and dP^T : rows = keys (this warp's 16), cols = 64 queries
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            st[i][0] = st[i][1] = st[i][2] = st[i][3] = 0.f;
            dpt[i][0] = dpt[i][1] = dpt[i][2] = dpt[i][3] = 0.f;
        }
        mma_a_tileT<HD>(st, ka, sa[buf]);
        mma_a_tileT<HD>(dpt, va, sb[buf]);
        const int q0 = t * ATT_BM;
        float pt[8][4];
        if (PLAIN) {
#pragma unroll
            for (int nb = 0; nb < 8; ++nb) {
                const float2 l2 = *reinterpret_cast<const float2*>(&sLse[buf][nb * 8 + 2 * tq]);      // pre-scaled by log2(e)
                const float2 d2 = *reinterpret_cast<const float2*>(&sDelta[buf][nb * 8 + 2 * tq]);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float p = key_ok[e >> 1] ? ex2_fast(fmaf(st[nb][e], LOG2E, -((e & 1) ? l2.y : l2.x))) : 0.f;   // lse = +inf for q >= S -> 0
                    pt[nb][e] = p;
                    st[nb][e] = p * (dpt[nb][e] - ((e & 1) ? d2.y : d2.x));
                }
            }
        } else {
#pragma unroll
            for (int nb = 0; nb < 8; ++nb) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int qi = nb * 8 + 2 * tq + (e & 1);
                    const int key = k0 + warp * 16 + g + (e >> 1) * 8;
                    const bool ok = key_ok[e >> 1] && (!causal || key <= q0 + qi);
                    const float p = ok ? exp2f((st[nb][e] - sLse[buf][qi]) * LOG2E) : 0.f;   // lse = +inf for q >= S -> 0
                    pt[nb][e] = p;
                    st[nb][e] = p * (dpt[nb][e] - sDelta[buf][qi]);
                }
            }
        }
        uint32_t pf[4][4], dsf[4][4];
        acc_to_afrag(pt, pf);
        acc_to_afrag(st, dsf);
        mma_p_tile<HD>(dv, pf, sb[buf]);    // dV += P^T dO
        mma_p_tile<HD>(dk, dsf, sa[buf]);   // dK += dS^T Q
        __syncthreads();
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int key = k0 + warp * 16 + g + r * 8;
        if (key < S) {
            __nv_bfloat16* drow = dqkv + (static_cast<long long>(b) * S + key) * ld + h * HD;
#pragma unroll
            for (int nb = 0; nb < HD / 8; ++nb) {
                *reinterpret_cast<uint32_t*>(drow + E + nb * 8 + 2 * tq) = pack_bf16x2(dk[nb][2 * r], dk[nb][2 * r + 1]);
                *reinterpret_cast<uint32_t*>(drow + 2 * E + nb * 8 + 2 * tq) = pack_bf16x2(dv[nb][2 * r], dv[nb][2 * r + 1]);
            }
        }
    }
}

template <int HD>
static int attn_fwd_launch(const void* qkv, int B, int S, int H, int causal, const uint8_t* km, void* out, float* out32, float* lse, cudaStream_t st) {
    dim3 grid((S + ATT_BM - 1) / ATT_BM, H, B);
    if (!causal && !km)
        attn_fwd_kernel<HD, true><<<grid, ATT_THREADS, 0, st>>>(static_cast<const __nv_bfloat16*>(qkv), S, H, 0, nullptr,
                                                                static_cast<__nv_bfloat16*>(out), out32, lse);
    else
        attn_fwd_kernel<HD, false><<<grid, ATT_THREADS, 0, st>>>(static_cast<const __nv_bfloat16*>(qkv), S, H, causal, km,
                                                                 static_cast<__nv_bfloat16*>(out), out32, lse);
    return check_launch("attn_fwd_kernel");
}

template <int HD>
static int attn_bwd_launch(const void* qkv, const void* out, const void* dout, const float* lse, int B, int S, int H, int causal,
                           const uint8_t* km, float* delta, void* dqkv, cudaStream_t st) {
    const long long rows = static_cast<long long>(B) * S;
    attn_delta_kernel<HD><<<static_cast<unsigned>((rows + 3) / 4), 128, 0, st>>>(static_cast<const __nv_bfloat16*>(out),
                                                                                  static_cast<const __nv_bfloat16*>(dout), S, H, rows, delta, 0);
    if (int rc = check_launch("attn_delta_kernel")) return rc;
    dim3 grid((S + ATT_BM - 1) / ATT_BM, H, B);
    const __nv_bfloat16* q_ = static_cast<const __nv_bfloat16*>(qkv);
    const __nv_bfloat16* do_ = static_cast<const __nv_bfloat16*>(dout);
    __nv_bfloat16* dq_ = static_cast<__nv_bfloat16*>(dqkv);
    const bool plain = !causal && !km;
    if (plain) attn_bwd_dkv_kernel<HD, true><<<grid, ATT_THREADS, 0, st>>>(q_, do_, lse, delta, S, H, 0, nullptr, dq_);
    else attn_bwd_dkv_kernel<HD, false><<<grid, ATT_THREADS, 0, st>>>(q_, do_, lse, delta, S, H, causal, km, dq_);
    if (int rc = check_launch("attn_bwd_dkv_kernel")) return rc;
    if (plain) attn_bwd_dq_kernel<HD, true><<<grid, ATT_THREADS, 0, st>>>(q_, do_, lse, delta, S, H, 0, nullptr, dq_);
    else attn_bwd_dq_kernel<HD, false><<<grid, ATT_THREADS, 0, st>>>(q_, do_, lse, delta, S, H, causal, km, dq_);
    return check_launch("attn_bwd_dq_kernel");
}

// tcgen05 / TMEM kernels for head dim 64 without masks (attention_sm100.cu)
bool attn_tc_enabled();
int attn_tc_fwd(const void* qkv, int B, int S, int H, void* out, float* out32, float* lse, cudaStream_t st, int o_f16);
int attn_tc_bwd(const void* qkv, const void* out_o, const void* dout, const float* lse, float* delta, int B, int S, int H, void* dqkv, cudaStream_t st,
                int row_begin, int o_f16);

}  // namespace tvs

extern "C" __attribute__((visibility("default"))) int tvs_attn_fwd(const void* qkv, int32_t B, int32_t S, int32_t H, int32_t hd, int32_t causal, const uint8_t* key_mask,
                            void* out, float* out_f32, float* lse, int32_t flags, void* stream) {
    using namespace tvs;
    TVS_REQUIRE(qkv && out && lse, "tvs_attn_fwd: null pointer");
    TVS_REQUIRE(B > 0 && S > 0 && H > 0, "tvs_attn_fwd: bad shape B=%d S=%d H=%d", B, S, H);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int o_f16 = (flags & TVS_ATTN_O_F16) ? 1 : 0;
    if (hd == 64 && !causal && !key_mask && attn_tc_enabled()) return attn_tc_fwd(qkv, B, S, H, out, out_f32, lse, st, o_f16);
    TVS_REQUIRE(!o_f16, "tvs_attn_fwd: TVS_ATTN_O_F16 is implemented by the tcgen05 kernels only (hd = 64, no masks)");
    if (hd == 64) return attn_fwd_launch<64>(qkv, B, S, H, causal, key_mask, out, out_f32, lse, st);
    if (hd == 16) return attn_fwd_launch<16>(qkv, B, S, H, causal, key_mask, out, out_f32, lse, st);
    set_error("tvs_attn_fwd: head dim %d not supported (64 or 16)", hd);
    return -1;
}

extern "C" __attribute__((visibility("default"))) int tvs_attn_bwd_tail(const void* qkv, const void* out, const void* dout, const float* lse, int32_t B, int32_t S, int32_t H,
                            int32_t hd, int32_t causal, const uint8_t* key_mask, float* delta, void* dqkv, int32_t row_begin, int32_t flags, void* stream) {
    using namespace tvs;
    const int o_f16 = (flags & TVS_ATTN_O_F16) ? 1 : 0;
    TVS_REQUIRE(row_begin >= 0 && row_begin < S, "tvs_attn_bwd_tail: row_begin %d outside [0, %d)", row_begin, S);
    TVS_REQUIRE(qkv && out && dout && lse && delta && dqkv, "tvs_attn_bwd_tail: null pointer");
    TVS_REQUIRE(B > 0 && S > 0 && H > 0, "tvs_attn_bwd_tail: bad shape B=%d S=%d H=%d", B, S, H);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (hd == 64 && !causal && !key_mask && attn_tc_enabled()) {
        static const bool fused_delta = [] { const char* e = getenv("TVS_ATTN_DELTA"); return !(e && e[0] == 'k'); }();   // TVS_ATTN_DELTA=kernel: separate pass
        if (row_begin < 128 && fused_delta) return attn_tc_bwd(qkv, out, dout, lse, delta, B, S, H, dqkv, st, row_begin, o_f16);   // delta comes out of the dQ kernel
        // tail-only backward: dK / dV of the last tile still need delta of EVERY query row
        const long long rows = static_cast<long long>(B) * S;
        attn_delta_kernel<64><<<static_cast<unsigned>((rows + 3) / 4), 128, 0, st>>>(static_cast<const __nv_bfloat16*>(out),
                                                                                      static_cast<const __nv_bfloat16*>(dout), S, H, rows, delta, o_f16);
        if (int rc = check_launch("attn_delta_kernel")) return rc;
        return attn_tc_bwd(qkv, nullptr, dout, lse, delta, B, S, H, dqkv, st, row_begin, o_f16);
    }
    TVS_REQUIRE(!o_f16, "tvs_attn_bwd: TVS_ATTN_O_F16 is implemented by the tcgen05 kernels only (hd = 64, no masks)");
    if (hd == 64) return attn_bwd_launch<64>(qkv, out, dout, lse, B, S, H, causal, key_mask, delta, dqkv, st);
    if (hd == 16) return attn_bwd_launch<16>(qkv, out, dout, lse, B, S, H, causal, key_mask, delta, dqkv, st);
    set_error("tvs_attn_bwd_tail: head dim %d not supported (64 or 16)", hd);
    return -1;
}

extern "C" __attribute__((visibility("default"))) int tvs_attn_bwd(const void* qkv, const void* out, const void* dout, const float* lse, int32_t B, int32_t S, int32_t H,
                            int32_t hd, int32_t causal, const uint8_t* key_mask, float* delta, void* dqkv, int32_t flags, void* stream) {
    return tvs_attn_bwd_tail(qkv, out, dout, lse, B, S, H, hd, causal, key_mask, delta, dqkv, 0, flags, stream);
}
