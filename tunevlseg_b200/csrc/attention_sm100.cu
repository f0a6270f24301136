// tcgen05 / TMEM / TMA flash attention for the CLIP vision tower: head dim 64, no mask, any sequence length.
//
// Replaces eager_attention_forward (transformers modeling_clipseg.py:256-276) and its backward for the shapes that
// carry 14 % of the step's FLOPs (S = 485 + n, 12 heads).  Scores live in TMEM and never touch shared or global memory.
//
// One CTA = one (batch, head, 128-row outer tile); 192 threads:
//   warps 0-3  element-wise stage: thread t owns TMEM lane t (one row): tcgen05.ld scores, exp2 / scaling in registers,
//              tcgen05.st the bf16 probabilities back into TMEM (two per 32-bit column) as the A operand of the next MMA
//   warp 4     TMA producer (3-D tensor maps [feature, token, batch]: rows past the sequence end are zero-filled)
//   warp 5     tcgen05.mma issuer (one lane) + TMEM allocation
// Every "score" MMA is 128x128x64 with both operands K-major in shared memory; every "accumulate" MMA is 128x64x128
// with A read from TMEM and B = the SAME shared-memory tile addressed MN-major (the 64 features are contiguous).
//
//   FWD  outer = Q tile, inner = K/V tiles.  Two passes over the keys instead of an online rescale: pass 1 recomputes
//        S = Q K^T per tile only for the row maximum, pass 2 recomputes S, writes P = exp2(S - m) and accumulates
//        O += P V in TMEM.  (The exp units, not the tensor core, bound this kernel; the second Q K^T is nearly free and
//        removes the accumulator read-modify-write.)  256 TMEM columns -> two CTAs per SM overlap MMA and exp.
//   DQ   outer = Q, dO tile; inner = K/V: S = Q K^T, dP = dO V^T, dS = P o (dP - delta) -> TMEM, dQ += dS K.
//   DKV  outer = K, V tile; inner = Q/dO (+ lse, delta rows): S^T = K Q^T, dP^T = V dO^T, P^T and dS^T -> TMEM,
//        dV += P^T dO, dK += dS^T Q.
// The 1/sqrt(d) scale is folded into Wq by the host.  lse is the natural-log-sum-exp per row, delta = rowsum(dO o O).
#include <stdlib.h>

#include "common.cuh"
#include "sm100_ptx.cuh"
#include "tvs_b200.h"

namespace tvs {
using namespace ptx;

constexpr int AT = 128;                    // tile rows
constexpr int AHD = 64;                    // head dim
constexpr int ATILE = AT * AHD * 2;        // 16 KB per [128 x 64] bf16 tile
constexpr int ATC_THREADS = 192;
constexpr float L2E = 1.4426950408889634f;
enum { MODE_FWD = 0, MODE_DQ = 1, MODE_DKV = 2 };

// ATC_EXPERIMENT == 8: per-CTA timeline of the backward kernels (tools/microbench/attn_timeline.cu reads it back): the CTAs
// (ot, h, b) = (0..1, 0, 0) log (clock64, event, warp, iteration) at every hand-off.  Not compiled into the product.
#ifndef ATC_EXPERIMENT
#define ATC_EXPERIMENT 0
#endif
#if ATC_EXPERIMENT == 8
// [mode 0..2][cta 0..1][warp 0..9][128 events][2]: every warp owns its slots (no atomics: a returning atomic would stall the warp
// for an L2 round trip at every event); stores are fire-and-forget
__device__ unsigned long long g_tl[3 * 2 * 10 * 512 * 2];
#define TL(ev, it)                                                                                                          \
    do {                                                                                                                    \
        if (blockIdx.y == 0 && blockIdx.z == 0 && blockIdx.x < 2 && (threadIdx.x & 31) == 0 && tl_n < 512) {                \
            unsigned long long* p_ = g_tl + ((((MODE * 2 + blockIdx.x) * 10 + (threadIdx.x >> 5)) * 512 + tl_n) << 1);      \
            p_[0] = clock64();                                                                                              \
            p_[1] = ((unsigned long long)(ev) << 16) | ((it) & 255) | (1ull << 40);                                         \
            ++tl_n;                                                                                                         \
        }                                                                                                                   \
    } while (0)
#else
#define TL(ev, it) do { } while (0)
#endif

// explicit shared-window load: the dynamic shared-memory base is re-aligned through integer arithmetic, after which the compiler
// only knows a GENERIC pointer and emits LD.E (address-space check per access) instead of LDS
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ float fast_exp2(float x) {   // one MUFU.EX2, flushes denormals; exp2(-inf) = 0
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// inner tile rows: 64 per step, so that the TMEM footprint is 128 columns in the forward (S + O: three CTAs per SM,
// bounded by registers) and 256 in the backward kernels (S + dP + accumulators: two CTAs per SM).  Co-resident CTAs
// overlap one CTA's exp stage, prologue and epilogue with another CTA's MMAs.
__device__ __forceinline__ void store_row_bf16_32(__nv_bfloat16* dst, const uint32_t (&a)[32]) {
    // 32 fp32 accumulators -> 32 bf16 = 64 bytes = 2 x 256-bit stores
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        uint32_t w[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) w[i] = pack_bf16x2(__uint_as_float(a[16 * q + 2 * i]), __uint_as_float(a[16 * q + 2 * i + 1]));
        st_global_256(dst + 16 * q, w);
    }
}

#ifndef ATC_BWD_STAGES
#define ATC_BWD_STAGES 2
#endif
template <int MODE>
struct AtcSmem {
    static constexpr int TI = 64;
    static constexpr int ITILE = TI * AHD * 2;
    static constexpr int N_OUTER = MODE == MODE_FWD ? 1 : 2;
    // inner-tile ring depth.  Measured (round 2, B = 32, S = 489): 2 / 3 / 4 stages = 171.5 / 175.0 / 173.0 us for the backward
    // pair - the kernels do not wait for their tiles, so the default stays at two stages (ATC_BWD_STAGES to A/B)
    static constexpr int NST = MODE == MODE_FWD ? 2 : ATC_BWD_STAGES;
    static constexpr int OUTER = 0;
    static constexpr int INNER = N_OUTER * ATILE;                 // NST stages x 2 tiles
    static constexpr int VEC = INNER + NST * 2 * ITILE;           // [NST][2][TI] floats (DKV: lse, delta; DQ: 2 x 128 delta halves)
    static constexpr int BAR = VEC + NST * 2 * TI * 4;
    static constexpr int TOTAL = BAR + 16 * 8 + 1024;             // barriers + tmem slot + alignment slack
};

__device__ __forceinline__ void store_row_bf16_64(__nv_bfloat16* dst, const uint32_t (&a)[32], const uint32_t (&b)[32], float scale, int f16) {
    // 64 fp32 accumulators (two 32-column TMEM loads) -> 64 bf16 (or IEEE fp16: TVS_ATTN_O_F16) = 128 bytes = 4 x 256-bit stores
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        uint32_t w[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int e = 16 * q + 2 * i;
            const float x0 = __uint_as_float(e < 32 ? a[e] : b[e - 32]) * scale;
            const float x1 = __uint_as_float(e + 1 < 32 ? a[e + 1] : b[e + 1 - 32]) * scale;
            w[i] = f16 ? pack_f16x2(x0, x1) : pack_bf16x2(x0, x1);
        }
        st_global_256(dst + 16 * q, w);
    }
}

// Backward modes run EIGHT element-wise warps: two per TMEM lane quarter, each owning 32 of the 64 score columns of a step
// (no row-wise reduction is needed in the backward: p and dS are element-wise given lse and delta).  With one warp per
// scheduler and CTA the TMEM load -> exp -> pack -> TMEM store chain of a step was exposed; four warps per scheduler
// (two CTAs per SM) hide each other's latencies.  Each warp writes its packed P / dS into the first 16 columns of the 32
// columns it has just read itself, so no store can land on a column another warp still has to load.
constexpr int ATC_THREADS_BWD = 320;
template <int MODE>
__global__ void __launch_bounds__(MODE == MODE_FWD ? ATC_THREADS : ATC_THREADS_BWD, MODE == MODE_FWD ? 3 : 2)
attn_tc_kernel(const __grid_constant__ CUtensorMap map_qkv, const __grid_constant__ CUtensorMap map_do,
               const __grid_constant__ CUtensorMap map_qkv_in, const __grid_constant__ CUtensorMap map_do_in, int S, int H,
               __nv_bfloat16* __restrict__ out, float* __restrict__ out32, float* __restrict__ lse_out,
               const float* __restrict__ lse_in, const float* __restrict__ delta_in, __nv_bfloat16* __restrict__ dqkv, int ot0, int o_f16) {
    using L = AtcSmem<MODE>;
#if ATC_EXPERIMENT == 8
    unsigned tl_n = 0;
#endif
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint8_t* s_outer0 = smem + L::OUTER;
    uint8_t* s_outer1 = s_outer0 + ATILE;                       // only when N_OUTER == 2
    uint8_t* s_inner = smem + L::INNER;                         // stage s: tile0 at s*2*ATILE, tile1 at +ATILE
    float* s_vec = reinterpret_cast<float*>(smem + L::VEC);     // [stage][which][128]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::BAR);
    uint64_t* outer_full = bars;
    constexpr int NST = L::NST;
    uint64_t* in_full = bars + 1;               // [NST]
    uint64_t* in_empty = bars + 1 + NST;        // [NST]
    uint64_t* s_full = bars + 1 + 2 * NST;
    uint64_t* ew_done = s_full + 1;
    uint64_t* acc_full = s_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_full + 3);
    static_assert((1 + 2 * NST + 4) * 8 <= 16 * 8, "barrier block too small");

    const int ot = blockIdx.x + ot0, h = blockIdx.y, b = blockIdx.z;     // ot0 > 0: only the outer tiles from ot0 on (tvs_attn_bwd_tail)
    const int E = H * AHD;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int EW_WARPS = MODE == MODE_FWD ? 4 : 8;       // element-wise warps; then the TMA producer, then the MMA issuer
    constexpr int W_TMA = EW_WARPS, W_MMA = EW_WARPS + 1;
    constexpr int TI = L::TI, ITILE = L::ITILE;
    const int n_in = (S + TI - 1) / TI;
    const int n_it = MODE == MODE_FWD ? 2 * n_in : n_in;
    constexpr uint32_t TMEM_COLS = MODE == MODE_FWD ? 128 : 256;
    constexpr uint32_t C_S = 0, C_DP = TI;                          // S [0,64); BWD: dP [64,128)
    constexpr uint32_t C_ACC0 = MODE == MODE_FWD ? 64 : 128;        // O / dQ / dV
    constexpr uint32_t C_ACC1 = 192;                                // dK
    const long long bh = static_cast<long long>(b) * H + h;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&map_qkv);
        if (MODE != MODE_FWD) tma_prefetch_desc(&map_do);
        mbar_init(outer_full, 1);
        for (int s = 0; s < NST; ++s) {
            mbar_init(&in_full[s], MODE == MODE_DKV ? 2 : 1);
            mbar_init(&in_empty[s], 1);
        }
        mbar_init(s_full, 1);
#if ATC_EXPERIMENT == 1 || ATC_EXPERIMENT == 4
        mbar_init(ew_done, MODE == MODE_FWD ? EW_WARPS * 32 : EW_WARPS);
#else
        mbar_init(ew_done, EW_WARPS * 32);
#endif
        mbar_init(acc_full, 1);
        fence_barrier_init();
    }
    if (warp == W_MMA) tmem_alloc(tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = *tmem_slot;
    pdl_wait();        // set-up above is private; q / k / v (and dO, lse, delta) come from the preceding kernels
    pdl_trigger();
    TL(30, 0);

    if (warp == W_TMA) {
        // ------------------------------------------------------------------ TMA producer
        const int cq = h * AHD, ck = E + h * AHD, cv = 2 * E + h * AHD;
        if (elect_one()) {
            mbar_expect_tx(outer_full, L::N_OUTER * ATILE);
            if (MODE == MODE_FWD) {
                tma_load_3d(s_outer0, &map_qkv, cq, ot * AT, b, outer_full);
            } else if (MODE == MODE_DQ) {
                tma_load_3d(s_outer0, &map_qkv, cq, ot * AT, b, outer_full);
                tma_load_3d(s_outer1, &map_do, h * AHD, ot * AT, b, outer_full);
            } else {
                tma_load_3d(s_outer0, &map_qkv, ck, ot * AT, b, outer_full);
                tma_load_3d(s_outer1, &map_qkv, cv, ot * AT, b, outer_full);
            }
        }
        for (int it = 0; it < n_it; ++it) {
            const int stage = it % NST, par = (it / NST) & 1;
            const int j = MODE == MODE_FWD ? it % n_in : it;
            uint8_t* t0 = s_inner + stage * 2 * ITILE;
            uint8_t* t1 = t0 + ITILE;
            mbar_wait(&in_empty[stage], par ^ 1);       // whole warp, converged: uniform loop state
            TL(20, it);
            if (elect_one()) {
                if (MODE == MODE_FWD) {
                    const bool second = it >= n_in;
                    mbar_expect_tx(&in_full[stage], second ? 2 * ITILE : ITILE);
                    tma_load_3d(t0, &map_qkv_in, ck, j * TI, b, &in_full[stage]);
                    if (second) tma_load_3d(t1, &map_qkv_in, cv, j * TI, b, &in_full[stage]);
                } else if (MODE == MODE_DQ) {
                    mbar_expect_tx(&in_full[stage], 2 * ITILE);
                    tma_load_3d(t0, &map_qkv_in, ck, j * TI, b, &in_full[stage]);
                    tma_load_3d(t1, &map_qkv_in, cv, j * TI, b, &in_full[stage]);
                } else {
                    mbar_expect_tx(&in_full[stage], 2 * ITILE);
                    tma_load_3d(t0, &map_qkv_in, cq, j * TI, b, &in_full[stage]);
                    tma_load_3d(t1, &map_do_in, h * AHD, j * TI, b, &in_full[stage]);
                }
            }
            if (MODE == MODE_DKV) {
                __syncwarp();
                float* v_lse = s_vec + stage * 2 * TI;
                float* v_del = v_lse + TI;
#pragma unroll
                for (int k = 0; k < TI / 32; ++k) {
                    const int qi = lane + 32 * k, q = j * TI + qi;
                    v_lse[qi] = q < S ? lse_in[bh * S + q] * L2E : INFINITY;
                    v_del[qi] = q < S ? delta_in[bh * S + q] : 0.f;
                }
                __syncwarp();
                if (elect_one()) mbar_arrive(&in_full[stage]);
            }
            __syncwarp();
        }
    } else if (warp == W_MMA) {
        // ------------------------------------------------------------------ MMA issuer
        {
            constexpr uint32_t idesc_s = umma_idesc_bf16(AT, TI, 0, 0);
            constexpr uint32_t idesc_acc = umma_idesc_bf16(AT, AHD, 0, 1);
            // backward: the packed A operand of K-step k (16 inner rows = 8 columns) - the two element-wise halves leave their
            // 16 packed columns at offsets 0 and 32 of the 64-column block they read
            auto PK = [](int k) -> uint32_t { return k < 2 ? 8u * k : 32u + 8u * (k - 2); };
            mbar_wait(outer_full, 0);
            tc_fence_after();
            TL(31, 0);
            const uint64_t a0 = umma_desc_sw128(smem_u32(s_outer0));
            const uint64_t a1 = umma_desc_sw128(smem_u32(s_outer1));
            for (int it = 0; it < n_it; ++it) {
                const int stage = it % NST, par = (it / NST) & 1;
                mbar_wait(&in_full[stage], par);
                tc_fence_after();
                TL(1, it);
                const uint64_t b0 = umma_desc_sw128(smem_u32(s_inner + stage * 2 * ITILE));
                const uint64_t b1 = umma_desc_sw128(smem_u32(s_inner + stage * 2 * ITILE + ITILE));
                const bool last = it == n_it - 1;
                if (MODE == MODE_FWD) {
                    if (it > 0 && (it - 1) < n_in) {   // the previous iteration was a max-only pass: S must have been read
                        mbar_wait(ew_done, (it - 1) & 1);
                        tc_fence_after();
                    }
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) umma_ss(tb + C_S, a0 + 2 * k, b0 + 2 * k, idesc_s, k > 0);
                        umma_commit(s_full);
                        if (it < n_in) umma_commit(&in_empty[stage]);
                    }
                    __syncwarp();
                    if (it >= n_in) {
                        mbar_wait(ew_done, it & 1);
                        tc_fence_after();
                        if (elect_one()) {
#pragma unroll
                            for (int k = 0; k < TI / 16; ++k)   // O += P V : 16 keys per MMA = 8 TMEM columns of P, 2048 bytes of V
                                umma_ts(tb + C_ACC0, tb + C_S + 8 * k, b1 + 128 * k, idesc_acc, (it > n_in || k > 0) ? 1u : 0u);
                            umma_commit(&in_empty[stage]);
                            if (last) umma_commit(acc_full);
                        }
                        __syncwarp();
                    }
                } else {
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) umma_ss(tb + C_S, a0 + 2 * k, b0 + 2 * k, idesc_s, k > 0);
#pragma unroll
                        for (int k = 0; k < 4; ++k) umma_ss(tb + C_DP, a1 + 2 * k, b1 + 2 * k, idesc_s, k > 0);
                        umma_commit(s_full);
                    }
                    __syncwarp();
                    TL(2, it);
#if ATC_EXPERIMENT != 7
                    mbar_wait(ew_done, it & 1);
#endif
                    tc_fence_after();
                    TL(3, it);
                    if (elect_one()) {
                        if (MODE == MODE_DQ) {
#pragma unroll
                            for (int k = 0; k < TI / 16; ++k)   // dQ += dS K_j   (dS: packed columns [0,16) and [32,48) of the S block)
                                umma_ts(tb + C_ACC0, tb + C_S + PK(k), b0 + 128 * k, idesc_acc, (it > 0 || k > 0) ? 1u : 0u);
                        } else {
#pragma unroll
                            for (int k = 0; k < TI / 16; ++k)   // dV += P^T dO_i
                                umma_ts(tb + C_ACC0, tb + C_S + PK(k), b1 + 128 * k, idesc_acc, (it > 0 || k > 0) ? 1u : 0u);
#pragma unroll
                            for (int k = 0; k < TI / 16; ++k)   // dK += dS^T Q_i
                                umma_ts(tb + C_ACC1, tb + C_DP + PK(k), b0 + 128 * k, idesc_acc, (it > 0 || k > 0) ? 1u : 0u);
                        }
                        umma_commit(&in_empty[stage]);
                        if (last) umma_commit(acc_full);
                    }
                    __syncwarp();
                    TL(4, it);
                }
            }
        }
    } else {
        // ------------------------------------------------------------------ element-wise stage (thread = TMEM lane = row;
        // backward: two warps per lane quarter, `half` selects the 32-column half of every 64-column block)
        const int quarter = warp & 3, half = warp >> 2;          // half is 0 in the forward (four warps)
        const int tid = quarter * 32 + lane;
        const uint32_t tl = tb + (static_cast<uint32_t>(quarter * 32) << 16);
        const int row = ot * AT + tid;
        const bool row_ok = row < S;
        if (MODE == MODE_FWD) {
            float m = -INFINITY;
            for (int it = 0; it < n_in; ++it) {            // pass 1: row maximum
                mbar_wait(s_full, it & 1);
                tc_fence_after();
                const int k0 = it * TI;
                uint32_t r0[32], r1[32];                    // both 32-column chunks in flight, one wait
                tmem_ld32(tl + C_S, r0);
                tmem_ld32(tl + C_S + 32, r1);
                tmem_ld_wait();
                if (k0 + TI <= S) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) m = fmaxf(m, fmaxf(__uint_as_float(r0[i]), __uint_as_float(r1[i])));
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        if (k0 + i < S) m = fmaxf(m, __uint_as_float(r0[i]));
                        if (k0 + 32 + i < S) m = fmaxf(m, __uint_as_float(r1[i]));
                    }
                }
                tc_fence_before();
                mbar_arrive(ew_done);
            }
            const float mL = (m == -INFINITY) ? 0.f : m * L2E;
            float l = 0.f;
            for (int it = n_in; it < n_it; ++it) {         // pass 2: P = exp2(S - m), row sum
                mbar_wait(s_full, it & 1);
                tc_fence_after();
                const int k0 = (it - n_in) * TI;
                uint32_t r0[32], r1[32];
                tmem_ld32(tl + C_S, r0);
                tmem_ld32(tl + C_S + 32, r1);
                tmem_ld_wait();
                const bool full = k0 + TI <= S;
#pragma unroll
                for (int c = 0; c < 2; ++c) {                 // 32 score columns -> 16 packed bf16x2 columns, in place
                    uint32_t pk[16];
#pragma unroll
                    for (int i = 0; i < 32; i += 2) {
                        float p0 = fast_exp2(fmaf(__uint_as_float(c == 0 ? r0[i] : r1[i]), L2E, -mL));
                        float p1 = fast_exp2(fmaf(__uint_as_float(c == 0 ? r0[i + 1] : r1[i + 1]), L2E, -mL));
                        if (!full) {
                            if (k0 + 32 * c + i >= S) p0 = 0.f;
                            if (k0 + 32 * c + i + 1 >= S) p1 = 0.f;
                        }
                        l += p0 + p1;
                        pk[i / 2] = pack_bf16x2(p0, p1);
                    }
                    tmem_st16(tl + C_S + 16 * c, pk);
                }
                tmem_st_wait();
                tc_fence_before();
                mbar_arrive(ew_done);
            }
            mbar_wait(acc_full, 0);
            tc_fence_after();
            uint32_t o0[32], o1[32];
            tmem_ld32(tl + C_ACC0, o0);
            tmem_ld32(tl + C_ACC0 + 32, o1);
            tmem_ld_wait();
            if (row_ok) {
                const float inv = l > 0.f ? 1.0f / l : 0.f;
                const long long tok = static_cast<long long>(b) * S + row;
                store_row_bf16_64(out + tok * E + h * AHD, o0, o1, inv, o_f16);
                if (out32) {
                    float* f = out32 + tok * E + h * AHD;
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        uint32_t w[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int e = 8 * q + i;
                            w[i] = __float_as_uint(round_tf32_rn(__uint_as_float(e < 32 ? o0[e] : o1[e - 32]) * inv));     // feeds a tf32 GEMM
                        }
                        st_global_256(f + 8 * q, w);
                    }
                }
                lse_out[bh * S + row] = (m == -INFINITY ? 0.f : m) + logf(fmaxf(l, 1e-30f));
            }
        } else {
            // backward: rows are queries (DQ) or keys (DKV); this thread owns columns [32 half, 32 half + 32) of S and dP
            const float my_lse = (MODE == MODE_DQ && row_ok) ? lse_in[bh * S + row] * L2E : INFINITY;
            float my_del = 0.f;
            if (MODE == MODE_DQ) {
                if (out != nullptr) {
                    // delta = rowsum(dO o O) computed here (the separate delta kernel and its launch gap disappear): this
                    // thread's half of the dO row comes from the swizzled outer tile the MMAs use, its half of the O row
                    // from global memory; the two halves meet through shared memory (s_vec is unused in this mode)
                    mbar_wait(outer_full, 0);
                    float part = 0.f;
                    const long long tok = static_cast<long long>(b) * S + (row_ok ? row : 0);
                    const uint4* orow = reinterpret_cast<const uint4*>(out + tok * E + h * AHD + 32 * half);
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int chunk = 4 * half + c;       // 16-byte chunk of the 128-byte row; stored at chunk ^ (row & 7)
                        const uint4 g = *reinterpret_cast<const uint4*>(s_outer1 + tid * 128 + ((chunk ^ (tid & 7)) << 4));
                        const uint4 o = row_ok ? orow[c] : make_uint4(0u, 0u, 0u, 0u);
                        const uint32_t gw[4] = {g.x, g.y, g.z, g.w}, ow[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float2 a = unpack_bf16x2(gw[i]), bb = o_f16 ? unpack_f16x2(ow[i]) : unpack_bf16x2(ow[i]);
                            part = fmaf(a.x, bb.x, part);
                            part = fmaf(a.y, bb.y, part);
                        }
                    }
                    s_vec[half * AT + tid] = part;
                    asm volatile("bar.sync 1, 256;" ::: "memory");
                    my_del = s_vec[tid] + s_vec[AT + tid];
                    if (half == 0 && row_ok) out32[bh * S + row] = my_del;      // the dK / dV kernel that follows reads it
                } else if (row_ok) {
                    my_del = delta_in[bh * S + row];
                }
            }
            const uint32_t tS = tl + C_S + 32 * half, tD = tl + C_DP + 32 * half;
            TL(9, 0);
            for (int it = 0; it < n_it; ++it) {
                mbar_wait(s_full, it & 1);
                tc_fence_after();
                TL(10, it);
                const int c0 = it * TI + 32 * half;          // first inner row of this half (key for DQ, query for DKV)
                const float* v_lse = s_vec + (it % NST) * 2 * TI + 32 * half;
                const float* v_del = v_lse + TI;
                uint32_t sA[32], dA[32];
#if ATC_EXPERIMENT == 6
#pragma unroll
                for (int i = 0; i < 32; ++i) { sA[i] = it + i; dA[i] = it - i; }
#else
                tmem_ld32(tS, sA);
                tmem_ld32(tD, dA);
                tmem_ld_wait();
#endif
                TL(11, it);
                uint32_t pkp[16], pks[16];
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                    const float s0 = __uint_as_float(sA[i]), s1 = __uint_as_float(sA[i + 1]);
                    const float g0 = __uint_as_float(dA[i]), g1 = __uint_as_float(dA[i + 1]);
                    float p0, p1, d0, d1;
                    // No per-element masks: a key past the sequence end has an all-zero K / V row (TMA zero fill), so its finite
                    // p and dS multiply zeros in dQ += dS K; a key ROW past the end (DKV) only feeds dK / dV rows that are never
                    // stored; a query past the end has lse = +inf, hence p = dS = 0.  (The masks cost ~3 of ~9 issue slots per
                    // score, and the element-wise warps are issue-bound: ncu round 2.)
#if ATC_EXPERIMENT >= 2      // timing experiments only (results are wrong)
                    p0 = fmaf(s0, L2E, -my_lse); p1 = fmaf(s1, L2E, -my_lse); d0 = p0 * (g0 - my_del); d1 = p1 * (g1 - my_del);
                    if (false) {
#else
                    if (MODE == MODE_DQ) {
#endif
                        p0 = fast_exp2(fmaf(s0, L2E, -my_lse));
                        p1 = fast_exp2(fmaf(s1, L2E, -my_lse));
                        d0 = p0 * (g0 - my_del);
                        d1 = p1 * (g1 - my_del);
                    } else {
                        const float2 l2 = *reinterpret_cast<const float2*>(v_lse + i);      // lse = +inf for queries past the end -> p = 0
                        const float2 e2 = *reinterpret_cast<const float2*>(v_del + i);
                        p0 = fast_exp2(fmaf(s0, L2E, -l2.x));
                        p1 = fast_exp2(fmaf(s1, L2E, -l2.y));
                        d0 = p0 * (g0 - e2.x);
                        d1 = p1 * (g1 - e2.y);
                    }
#if ATC_EXPERIMENT >= 3
                    pks[i / 2] = sA[i] ^ dA[i + 1]; pkp[i / 2] = sA[i + 1] ^ dA[i];
                    if (false)
#endif
                    {
                        pks[i / 2] = pack_bf16x2(d0, d1);
                        if (MODE == MODE_DKV) pkp[i / 2] = pack_bf16x2(p0, p1);
                    }
                }
                // the packed values overwrite the first 16 of the 32 columns this thread has just read (both loads completed)
#if ATC_EXPERIMENT == 5 || ATC_EXPERIMENT == 6
                if (pks[3] == 0x12345678u && pkp[5] == 0x9abcdef0u) tmem_st16(tS, pks);
                else if (false)
#endif
                if (MODE == MODE_DQ) {
                    tmem_st16(tS, pks);          // dS
                } else {
                    tmem_st16(tS, pkp);          // P^T
                    tmem_st16(tD, pks);          // dS^T
                }
                TL(12, it);
                tmem_st_wait();
                tc_fence_before();
                TL(13, it);
#if ATC_EXPERIMENT == 1 || ATC_EXPERIMENT == 4
                __syncwarp();
                if (lane == 0) mbar_arrive(ew_done);
#else
                mbar_arrive(ew_done);
#endif
            }
            TL(14, n_it);
            mbar_wait(acc_full, 0);
            tc_fence_after();
            TL(32, 0);
            uint32_t a0[32];
            tmem_ld32(tl + C_ACC0 + 32 * half, a0);
            tmem_ld_wait();
            const long long tok = static_cast<long long>(b) * S + row;
            __nv_bfloat16* drow = dqkv + tok * 3 * E + h * AHD + 32 * half;
            if (MODE == MODE_DQ) {
                if (row_ok) store_row_bf16_32(drow, a0);
            } else {
                if (row_ok) store_row_bf16_32(drow + 2 * E, a0);     // dV
                tmem_ld32(tl + C_ACC1 + 32 * half, a0);
                tmem_ld_wait();
                if (row_ok) store_row_bf16_32(drow + E, a0);         // dK
            }
        }
    }

    TL(33, 0);
    tc_fence_before();
    __syncthreads();
    if (warp == W_MMA) {
        tc_fence_after();
        tmem_dealloc(tb, TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------------
// backward, PERSISTENT version (round 2, the default).  A per-CTA timeline of attn_tc_kernel<DQ|DKV> (clock64 stamps at every
// hand-off, tools/microbench/attn_timeline.cu) showed where its 172 us per layer go: a CTA lives ~23.7 k clk, of which ~5 k
// pass before its first MMA (tensor-map fetch + the TMA round trip of the outer tiles, with all 296 resident CTAs asking at
// once) and ~1 k in the epilogue - a quarter of every CTA's life with the tensor pipe and the exponential units idle, six
// rounds of CTAs per SM.  (Also measured there: a tcgen05.mma 128 x 64 x 16 with both operands in shared memory takes 48 clk,
// not 32 - 6 KB of operand reads per instruction against 128 B/clk - so S / dP cost 384 clk per step and the pair of kernels
// cannot go below ~50 us per layer; the TMEM read-out (515-740 B/clk/SM measured) and the MUFU (16 / clk) are not the bound.)
// Here 2 x #SM CTAs stay resident and walk the (outer tile, head, sample) items round-robin: barriers, TMEM and tensor maps are
// set up once, the outer tiles are double-buffered and those of item k + 1 are requested while item k computes, the inner ring
// never drains between items, and the accumulator read-out of item k overlaps the first S / dP MMAs of item k + 1.
// Same arithmetic, same operand layouts and the same per-item instruction order as attn_tc_kernel, so results are bit-identical.
// ------------------------------------------------------------------------------------------------
#ifndef ATC_BWDP_STAGES
#define ATC_BWDP_STAGES 2
#endif
struct BwdPSmem {
    static constexpr int TI = 64;
    static constexpr int ITILE = TI * AHD * 2;                    // 8 KB
    static constexpr int NST = ATC_BWDP_STAGES;
    static constexpr int OUTER = 0;                               // [2 buffers][2 tiles] of ATILE
    static constexpr int INNER = 4 * ATILE;                       // [NST][2 tiles] of ITILE
    static constexpr int VEC = INNER + NST * 2 * ITILE;           // DKV: [NST][2][TI] floats (lse, delta); DQ: [2][128] delta halves
    static constexpr int VEC_BYTES = (NST * 2 * TI > 2 * AT ? NST * 2 * TI : 2 * AT) * 4;
    static constexpr int BAR = VEC + VEC_BYTES;
    static constexpr int NBAR = 4 + 2 * NST + 3;                  // outer_full[2], outer_empty[2], in_full / in_empty[NST], s_full, ew_done, acc_full
    static constexpr int TOTAL = BAR + (NBAR + 1) * 8 + 1024;     // barriers + tmem slot + alignment slack
};

template <int MODE>
__global__ void __launch_bounds__(ATC_THREADS_BWD, 2)
attn_tc_bwdp_kernel(const __grid_constant__ CUtensorMap map_qkv, const __grid_constant__ CUtensorMap map_do,
                    const __grid_constant__ CUtensorMap map_qkv_in, const __grid_constant__ CUtensorMap map_do_in, int S, int H, int n_ot,
                    int n_items, const __nv_bfloat16* __restrict__ out, float* __restrict__ delta_out, const float* __restrict__ lse_in,
                    const float* __restrict__ delta_in, __nv_bfloat16* __restrict__ dqkv, int ot0, int o_f16) {
    static_assert(MODE == MODE_DQ || MODE == MODE_DKV, "backward modes only");
    using L = BwdPSmem;
#if ATC_EXPERIMENT == 8
    unsigned tl_n = 0;
#endif
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint8_t* s_outer = smem + L::OUTER;                         // buffer u: tile0 at u*2*ATILE, tile1 at +ATILE
    uint8_t* s_inner = smem + L::INNER;                         // stage s: tile0 at s*2*ITILE, tile1 at +ITILE
    float* s_vec = reinterpret_cast<float*>(smem + L::VEC);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::BAR);
    constexpr int NST = L::NST, TI = L::TI, ITILE = L::ITILE;
    uint64_t* outer_full = bars;                // [2]
    uint64_t* outer_empty = bars + 2;           // [2]
    uint64_t* in_full = bars + 4;               // [NST]
    uint64_t* in_empty = bars + 4 + NST;        // [NST]
    uint64_t* s_full = bars + 4 + 2 * NST;
    uint64_t* ew_done = s_full + 1;
    uint64_t* acc_full = s_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_full + 3);

    const int E = H * AHD;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int EW_WARPS = 8, W_TMA = EW_WARPS, W_MMA = EW_WARPS + 1;
    const int n_in = (S + TI - 1) / TI;
    constexpr uint32_t TMEM_COLS = 256;
    constexpr uint32_t C_S = 0, C_DP = TI, C_ACC0 = 128, C_ACC1 = 192;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&map_qkv);
        tma_prefetch_desc(&map_do);
        tma_prefetch_desc(&map_qkv_in);
        tma_prefetch_desc(&map_do_in);
        for (int u = 0; u < 2; ++u) {
            mbar_init(&outer_full[u], 1);
            mbar_init(&outer_empty[u], 1);
        }
        for (int s = 0; s < NST; ++s) {
            mbar_init(&in_full[s], MODE == MODE_DKV ? 2 : 1);
            mbar_init(&in_empty[s], 1);
        }
        mbar_init(s_full, 1);
        mbar_init(ew_done, EW_WARPS * 32);
        mbar_init(acc_full, 1);
        fence_barrier_init();
    }
    if (warp == W_MMA) tmem_alloc(tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = *tmem_slot;
    pdl_wait();        // set-up above is private; q / k / v, dO, lse (and delta) come from the preceding kernels
    pdl_trigger();
    TL(30, 0);

    // item w -> (outer tile, head, sample); the outer tile runs fastest so that neighbouring CTAs share K / V (or Q / dO) in L2
    auto item = [&](int w, int& ot, int& h, int& b) {
        ot = ot0 + w % n_ot;
        const int hb = w / n_ot;
        h = hb % H;
        b = hb / H;
    };

    if (warp == W_TMA) {
        // ------------------------------------------------------------------ TMA producer
        auto load_outer = [&](int k, int w) {          // outer tiles of the k-th item of this CTA into buffer k & 1
            int ot, h, b;
            item(w, ot, h, b);
            const int u = k & 1;
            mbar_wait(&outer_empty[u], ((k >> 1) & 1) ^ 1);
            if (elect_one()) {
                uint8_t* o0 = s_outer + u * 2 * ATILE;
                mbar_expect_tx(&outer_full[u], 2 * ATILE);
                if (MODE == MODE_DQ) {
                    tma_load_3d(o0, &map_qkv, h * AHD, ot * AT, b, &outer_full[u]);
                    tma_load_3d(o0 + ATILE, &map_do, h * AHD, ot * AT, b, &outer_full[u]);
                } else {
                    tma_load_3d(o0, &map_qkv, E + h * AHD, ot * AT, b, &outer_full[u]);
                    tma_load_3d(o0 + ATILE, &map_qkv, 2 * E + h * AHD, ot * AT, b, &outer_full[u]);
                }
            }
            __syncwarp();
        };
        int g = 0, k = 0;
        if (static_cast<int>(blockIdx.x) < n_items) load_outer(0, blockIdx.x);
        for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++k) {
            int ot, h, b;
            item(w, ot, h, b);
            const long long bh = static_cast<long long>(b) * H + h;
            const int cq = h * AHD, ck = E + h * AHD, cv = 2 * E + h * AHD;
            const int w_next = w + gridDim.x;
            bool prefetched = w_next >= n_items;
            for (int it = 0; it < n_in; ++it, ++g) {
                const int stage = g % NST, par = (g / NST) & 1;
                uint8_t* t0 = s_inner + stage * 2 * ITILE;
                uint8_t* t1 = t0 + ITILE;
                mbar_wait(&in_empty[stage], par ^ 1);       // whole warp, converged: uniform loop state
                TL(20, it);
                if (elect_one()) {
                    mbar_expect_tx(&in_full[stage], 2 * ITILE);
                    if (MODE == MODE_DQ) {
                        tma_load_3d(t0, &map_qkv_in, ck, it * TI, b, &in_full[stage]);
                        tma_load_3d(t1, &map_qkv_in, cv, it * TI, b, &in_full[stage]);
                    } else {
                        tma_load_3d(t0, &map_qkv_in, cq, it * TI, b, &in_full[stage]);
                        tma_load_3d(t1, &map_do_in, h * AHD, it * TI, b, &in_full[stage]);
                    }
                }
                if (MODE == MODE_DKV) {
                    __syncwarp();
                    float* v_lse = s_vec + stage * 2 * TI;
                    float* v_del = v_lse + TI;
#pragma unroll
                    for (int c = 0; c < TI / 32; ++c) {
                        const int qi = lane + 32 * c, q = it * TI + qi;
                        v_lse[qi] = q < S ? lse_in[bh * S + q] * L2E : INFINITY;
                        v_del[qi] = q < S ? delta_in[bh * S + q] : 0.f;
                    }
                    __syncwarp();
                    if (elect_one()) mbar_arrive(&in_full[stage]);
                }
                __syncwarp();
                // once the ring has wrapped inside this item the previous item has retired (its last MMAs released the stage we
                // have just been given), so its outer buffer is free: ask for the next item's outer tiles now, a whole item ahead
                if (!prefetched && it == NST) {
                    load_outer(k + 1, w_next);
                    prefetched = true;
                }
            }
            if (!prefetched) load_outer(k + 1, w_next);      // short sequences: may wait for this item's MMAs
        }
    } else if (warp == W_MMA) {
        // ------------------------------------------------------------------ MMA issuer
        constexpr uint32_t idesc_s = umma_idesc_bf16(AT, TI, 0, 0);
        constexpr uint32_t idesc_acc = umma_idesc_bf16(AT, AHD, 0, 1);
        auto PK = [](int kk) -> uint32_t { return kk < 2 ? 8u * kk : 32u + 8u * (kk - 2); };
        int g = 0, k = 0;
        for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++k) {
            const int u = k & 1;
            mbar_wait(&outer_full[u], (k >> 1) & 1);
            tc_fence_after();
            TL(31, k);
            const uint64_t a0 = umma_desc_sw128(smem_u32(s_outer + u * 2 * ATILE));
            const uint64_t a1 = umma_desc_sw128(smem_u32(s_outer + u * 2 * ATILE + ATILE));
            for (int it = 0; it < n_in; ++it, ++g) {
                const int stage = g % NST, par = (g / NST) & 1;
                mbar_wait(&in_full[stage], par);
                tc_fence_after();
                TL(1, it);
                const uint64_t b0 = umma_desc_sw128(smem_u32(s_inner + stage * 2 * ITILE));
                const uint64_t b1 = umma_desc_sw128(smem_u32(s_inner + stage * 2 * ITILE + ITILE));
                const bool last = it == n_in - 1;
                if (elect_one()) {
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) umma_ss(tb + C_S, a0 + 2 * kk, b0 + 2 * kk, idesc_s, kk > 0);
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) umma_ss(tb + C_DP, a1 + 2 * kk, b1 + 2 * kk, idesc_s, kk > 0);
                    umma_commit(s_full);
                }
                __syncwarp();
                TL(2, it);
                // ew_done of step (k, 0) also tells that the element-wise warps have read the accumulators of item k - 1 out of
                // TMEM (their epilogue precedes this step in program order), so the first accumulate MMA may overwrite them
                mbar_wait(ew_done, g & 1);
                tc_fence_after();
                TL(3, it);
                if (elect_one()) {
                    if (MODE == MODE_DQ) {
#pragma unroll
                        for (int kk = 0; kk < TI / 16; ++kk)   // dQ += dS K_j
                            umma_ts(tb + C_ACC0, tb + C_S + PK(kk), b0 + 128 * kk, idesc_acc, (it > 0 || kk > 0) ? 1u : 0u);
                    } else {
#pragma unroll
                        for (int kk = 0; kk < TI / 16; ++kk)   // dV += P^T dO_i
                            umma_ts(tb + C_ACC0, tb + C_S + PK(kk), b1 + 128 * kk, idesc_acc, (it > 0 || kk > 0) ? 1u : 0u);
#pragma unroll
                        for (int kk = 0; kk < TI / 16; ++kk)   // dK += dS^T Q_i
                            umma_ts(tb + C_ACC1, tb + C_DP + PK(kk), b0 + 128 * kk, idesc_acc, (it > 0 || kk > 0) ? 1u : 0u);
                    }
                    umma_commit(&in_empty[stage]);
                    if (last) {
                        umma_commit(acc_full);
                        umma_commit(&outer_empty[u]);
                    }
                }
                __syncwarp();
                TL(4, it);
            }
        }
    } else {
        // ------------------------------------------------------------------ element-wise stage (see attn_tc_kernel)
        const int quarter = warp & 3, half = warp >> 2;
        const int tid = quarter * 32 + lane;
        const uint32_t tl = tb + (static_cast<uint32_t>(quarter * 32) << 16);
        const uint32_t tS = tl + C_S + 32 * half, tD = tl + C_DP + 32 * half;
        int g = 0, k = 0;
        for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++k) {
            int ot, h, b;
            item(w, ot, h, b);
            const long long bh = static_cast<long long>(b) * H + h;
            const int row = ot * AT + tid;
            const bool row_ok = row < S;
            const int u = k & 1;
            const float my_lse = (MODE == MODE_DQ && row_ok) ? lse_in[bh * S + row] * L2E : INFINITY;
            float my_del = 0.f;
            if (MODE == MODE_DQ) {
                if (out != nullptr) {
                    // delta = rowsum(dO o O) from the swizzled dO tile of this item and the O row in global memory (attn_tc_kernel)
                    mbar_wait(&outer_full[u], (k >> 1) & 1);
                    const uint8_t* s_do = s_outer + u * 2 * ATILE + ATILE;
                    float part = 0.f;
                    const long long tok = static_cast<long long>(b) * S + (row_ok ? row : 0);
                    const uint4* orow = reinterpret_cast<const uint4*>(out + tok * E + h * AHD + 32 * half);
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int chunk = 4 * half + c;
                        const uint4 gq = *reinterpret_cast<const uint4*>(s_do + tid * 128 + ((chunk ^ (tid & 7)) << 4));
                        const uint4 o = row_ok ? orow[c] : make_uint4(0u, 0u, 0u, 0u);
                        const uint32_t gw[4] = {gq.x, gq.y, gq.z, gq.w}, ow[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float2 a = unpack_bf16x2(gw[i]), bb = o_f16 ? unpack_f16x2(ow[i]) : unpack_bf16x2(ow[i]);
                            part = fmaf(a.x, bb.x, part);
                            part = fmaf(a.y, bb.y, part);
                        }
                    }
                    s_vec[half * AT + tid] = part;
                    asm volatile("bar.sync 1, 256;" ::: "memory");
                    my_del = s_vec[tid] + s_vec[AT + tid];
                    if (half == 0 && row_ok) delta_out[bh * S + row] = my_del;      // the dK / dV kernel that follows reads it
                } else if (row_ok) {
                    my_del = delta_in[bh * S + row];
                }
            }
            TL(9, k);
            for (int it = 0; it < n_in; ++it, ++g) {
                mbar_wait(s_full, g & 1);
                tc_fence_after();
                TL(10, it);
                const uint32_t v_lse = smem_u32(s_vec) + static_cast<uint32_t>(((g % NST) * 2 * TI + 32 * half) * 4);
                const uint32_t v_del = v_lse + TI * 4;
                uint32_t sA[32], dA[32];
                tmem_ld32(tS, sA);
                tmem_ld32(tD, dA);
                tmem_ld_wait();
                TL(11, it);
                uint32_t pkp[16], pks[16];
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                    const float s0 = __uint_as_float(sA[i]), s1 = __uint_as_float(sA[i + 1]);
                    const float g0 = __uint_as_float(dA[i]), g1 = __uint_as_float(dA[i + 1]);
                    float p0, p1, d0, d1;
                    if (MODE == MODE_DQ) {
                        p0 = fast_exp2(fmaf(s0, L2E, -my_lse));
                        p1 = fast_exp2(fmaf(s1, L2E, -my_lse));
                        d0 = p0 * (g0 - my_del);
                        d1 = p1 * (g1 - my_del);
                    } else {
                        // per-query lse / delta: the same address in every lane (broadcast).  128-bit loads, four queries each:
                        // with 64-bit loads the 32 LDS per thread and step were what the dK / dV element-wise stage waited for
#if ATC_EXPERIMENT == 9      // timing experiment: no shared-memory loads at all
                        const float4 l4 = make_float4(my_lse, my_lse, my_lse, my_lse), e4 = make_float4(my_del, my_del, my_del, my_del);
#else
                        const float4 l4 = lds128(v_lse + (i & ~3) * 4);
                        const float4 e4 = lds128(v_del + (i & ~3) * 4);
#endif
                        const bool hi = (i & 2) != 0;
                        p0 = fast_exp2(fmaf(s0, L2E, -(hi ? l4.z : l4.x)));
                        p1 = fast_exp2(fmaf(s1, L2E, -(hi ? l4.w : l4.y)));
                        d0 = p0 * (g0 - (hi ? e4.z : e4.x));
                        d1 = p1 * (g1 - (hi ? e4.w : e4.y));
                    }
                    pks[i / 2] = pack_bf16x2(d0, d1);
                    if (MODE == MODE_DKV) pkp[i / 2] = pack_bf16x2(p0, p1);
                }
                if (MODE == MODE_DQ) {
                    tmem_st16(tS, pks);          // dS
                } else {
                    tmem_st16(tS, pkp);          // P^T
                    tmem_st16(tD, pks);          // dS^T
                }
                TL(12, it);
                tmem_st_wait();
                tc_fence_before();
                TL(13, it);
                mbar_arrive(ew_done);
            }
            TL(14, k);
            mbar_wait(acc_full, k & 1);
            tc_fence_after();
            TL(32, k);
            uint32_t a0[32];
            tmem_ld32(tl + C_ACC0 + 32 * half, a0);
            tmem_ld_wait();
            const long long tok = static_cast<long long>(b) * S + row;
            __nv_bfloat16* drow = dqkv + tok * 3 * E + h * AHD + 32 * half;
            if (MODE == MODE_DQ) {
                if (row_ok) store_row_bf16_32(drow, a0);
            } else {
                if (row_ok) store_row_bf16_32(drow + 2 * E, a0);     // dV
                tmem_ld32(tl + C_ACC1 + 32 * half, a0);
                tmem_ld_wait();
                if (row_ok) store_row_bf16_32(drow + E, a0);         // dK
            }
            tc_fence_before();      // orders these TMEM reads before the ew_done arrive of the next item's first step
            TL(33, k);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == W_MMA) {
        tc_fence_after();
        tmem_dealloc(tb, TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------------
// backward, second version: 32-row inner tiles with TWO score buffers in TMEM and a software-pipelined MMA stream.
// In attn_tc_kernel<DQ|DKV> a CTA alternates MMA (S, dP) -> element-wise -> MMA (accumulate) on ONE 64-column score buffer;
// ncu (round 2): tensor pipe 23-25 %, warps stalled on TMEM loads, each kernel at ~55 % of its TMEM read-out floor
// (64 B/clk/SM), because the element-wise stage of a CTA idles while its own score MMAs run and the co-resident CTA only
// partly fills the hole.  Here S / dP of step it + 1 are queued before the accumulate MMAs of step it, so the tensor core
// works on the next score tile while the exponentials of the current one run; same 256 TMEM columns (two buffers of
// 32 + 32 columns, accumulators at 128 / 192), two CTAs per SM, 4-stage inner ring of 4 KB tiles.
// ------------------------------------------------------------------------------------------------
constexpr int BT = 32;                          // inner rows per step
constexpr int BITILE = BT * AHD * 2;            // 4 KB
constexpr int BST = 4;                          // inner stages
struct Bwd2Smem {
    static constexpr int OUTER = 0;
    static constexpr int INNER = 2 * ATILE;
    static constexpr int VEC = INNER + BST * 2 * BITILE;            // DKV: [BST][2][BT] floats (lse, delta); DQ: [2][128] floats (delta halves)
    static constexpr int BAR = VEC + 2 * AT * 4;
    static constexpr int TOTAL = BAR + 16 * 8 + 1024;
};

template <int MODE>
__global__ void __launch_bounds__(ATC_THREADS_BWD, 2)
attn_tc_bwd2_kernel(const __grid_constant__ CUtensorMap map_qkv, const __grid_constant__ CUtensorMap map_do,
                    const __grid_constant__ CUtensorMap map_qkv_in, const __grid_constant__ CUtensorMap map_do_in, int S, int H,
                    const __nv_bfloat16* __restrict__ out, float* __restrict__ delta_out, const float* __restrict__ lse_in,
                    const float* __restrict__ delta_in, __nv_bfloat16* __restrict__ dqkv, int ot0, int o_f16) {
    using L = Bwd2Smem;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint8_t* s_outer0 = smem + L::OUTER;
    uint8_t* s_outer1 = s_outer0 + ATILE;
    uint8_t* s_inner = smem + L::INNER;                         // stage s: tile0 at s * 2 * BITILE, tile1 at + BITILE
    float* s_vec = reinterpret_cast<float*>(smem + L::VEC);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::BAR);
    uint64_t* outer_full = bars;
    uint64_t* in_full = bars + 1;      // [BST]
    uint64_t* in_empty = bars + 5;     // [BST]
    uint64_t* s_full = bars + 9;       // [2]
    uint64_t* ew_done = bars + 11;     // [2]
    uint64_t* acc_full = bars + 13;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);

    const int ot = blockIdx.x + ot0, h = blockIdx.y, b = blockIdx.z;
    const int E = H * AHD;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int W_TMA = 8, W_MMA = 9;
    const int n_it = (S + BT - 1) / BT;
    constexpr uint32_t TMEM_COLS = 256, C_ACC0 = 128, C_ACC1 = 192;
    auto CS = [](int bf) -> uint32_t { return 64u * bf; };
    auto CDP = [](int bf) -> uint32_t { return 64u * bf + 32u; };
    const long long bh = static_cast<long long>(b) * H + h;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&map_qkv);
        tma_prefetch_desc(&map_do);
        mbar_init(outer_full, 1);
        for (int s = 0; s < BST; ++s) {
            mbar_init(&in_full[s], MODE == MODE_DKV ? 2 : 1);
            mbar_init(&in_empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&s_full[s], 1);
            mbar_init(&ew_done[s], 256);
        }
        mbar_init(acc_full, 1);
        fence_barrier_init();
    }
    if (warp == W_MMA) tmem_alloc(tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = *tmem_slot;
    pdl_wait();
    pdl_trigger();

    if (warp == W_TMA) {
        const int cq = h * AHD, ck = E + h * AHD, cv = 2 * E + h * AHD;
        if (elect_one()) {
            mbar_expect_tx(outer_full, 2 * ATILE);
            if (MODE == MODE_DQ) {
                tma_load_3d(s_outer0, &map_qkv, cq, ot * AT, b, outer_full);
                tma_load_3d(s_outer1, &map_do, h * AHD, ot * AT, b, outer_full);
            } else {
                tma_load_3d(s_outer0, &map_qkv, ck, ot * AT, b, outer_full);
                tma_load_3d(s_outer1, &map_qkv, cv, ot * AT, b, outer_full);
            }
        }
        for (int it = 0; it < n_it; ++it) {
            const int stage = it % BST, par = (it / BST) & 1;
            uint8_t* t0 = s_inner + stage * 2 * BITILE;
            uint8_t* t1 = t0 + BITILE;
            mbar_wait(&in_empty[stage], par ^ 1);
            if (elect_one()) {
                mbar_expect_tx(&in_full[stage], 2 * BITILE);
                if (MODE == MODE_DQ) {
                    tma_load_3d(t0, &map_qkv_in, ck, it * BT, b, &in_full[stage]);
                    tma_load_3d(t1, &map_qkv_in, cv, it * BT, b, &in_full[stage]);
                } else {
                    tma_load_3d(t0, &map_qkv_in, cq, it * BT, b, &in_full[stage]);
                    tma_load_3d(t1, &map_do_in, h * AHD, it * BT, b, &in_full[stage]);
                }
            }
            if (MODE == MODE_DKV) {
                __syncwarp();
                float* v_lse = s_vec + stage * 2 * BT;
                const int q = it * BT + lane;
                v_lse[lane] = q < S ? lse_in[bh * S + q] * L2E : INFINITY;
                v_lse[BT + lane] = q < S ? delta_in[bh * S + q] : 0.f;
                __syncwarp();
                if (elect_one()) mbar_arrive(&in_full[stage]);
            }
            __syncwarp();
        }
    } else if (warp == W_MMA) {
        constexpr uint32_t idesc_s = umma_idesc_bf16(AT, BT, 0, 0);
        constexpr uint32_t idesc_acc = umma_idesc_bf16(AT, AHD, 0, 1);
        mbar_wait(outer_full, 0);
        tc_fence_after();
        const uint64_t a0 = umma_desc_sw128(smem_u32(s_outer0));
        const uint64_t a1 = umma_desc_sw128(smem_u32(s_outer1));
        // accumulate MMAs of step j: the packed A operand of K-step k (16 inner rows = 8 columns) sits at the start of the 16
        // columns the element-wise warps of half k read themselves
        auto issue_acc = [&](int j) {
            const int bfj = j & 1, stg = j % BST;
            mbar_wait(&ew_done[bfj], (j >> 1) & 1);
            tc_fence_after();
            if (elect_one()) {
                const uint64_t b0 = umma_desc_sw128(smem_u32(s_inner + stg * 2 * BITILE));
                const uint64_t b1 = umma_desc_sw128(smem_u32(s_inner + stg * 2 * BITILE + BITILE));
                if (MODE == MODE_DQ) {
#pragma unroll
                    for (int k = 0; k < BT / 16; ++k)   // dQ += dS K_j
                        umma_ts(tb + C_ACC0, tb + 64 * bfj + 16 * k, b0 + 128 * k, idesc_acc, (j > 0 || k > 0) ? 1u : 0u);
                } else {
#pragma unroll
                    for (int k = 0; k < BT / 16; ++k)   // dV += P^T dO_i
                        umma_ts(tb + C_ACC0, tb + 64 * bfj + 16 * k, b1 + 128 * k, idesc_acc, (j > 0 || k > 0) ? 1u : 0u);
#pragma unroll
                    for (int k = 0; k < BT / 16; ++k)   // dK += dS^T Q_i
                        umma_ts(tb + C_ACC1, tb + 64 * bfj + 32 + 16 * k, b0 + 128 * k, idesc_acc, (j > 0 || k > 0) ? 1u : 0u);
                }
                umma_commit(&in_empty[stg]);
                if (j == n_it - 1) umma_commit(acc_full);
            }
            __syncwarp();
        };
        for (int it = 0; it < n_it; ++it) {
            const int stage = it % BST, par = (it / BST) & 1, bf = it & 1;
            mbar_wait(&in_full[stage], par);
            tc_fence_after();
            // score buffer bf was last read (as packed P / dS) by the accumulate MMAs of step it - 2: issued earlier in program order
            if (elect_one()) {
                const uint64_t b0 = umma_desc_sw128(smem_u32(s_inner + stage * 2 * BITILE));
                const uint64_t b1 = umma_desc_sw128(smem_u32(s_inner + stage * 2 * BITILE + BITILE));
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_ss(tb + 64 * bf, a0 + 2 * k, b0 + 2 * k, idesc_s, k > 0);
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_ss(tb + 64 * bf + 32, a1 + 2 * k, b1 + 2 * k, idesc_s, k > 0);
                umma_commit(&s_full[bf]);
            }
            __syncwarp();
            if (it >= 1) issue_acc(it - 1);
        }
        issue_acc(n_it - 1);
    } else {
        // element-wise warps: thread = TMEM lane = row (query for DQ, key for DKV); `half` = which 16 of the 32 score columns
        const int quarter = warp & 3, half = warp >> 2;
        const int tid = quarter * 32 + lane;
        const uint32_t tl = tb + (static_cast<uint32_t>(quarter * 32) << 16);
        const int row = ot * AT + tid;
        const bool row_ok = row < S;
        const float my_lse = (MODE == MODE_DQ && row_ok) ? lse_in[bh * S + row] * L2E : INFINITY;
        float my_del = 0.f;
        if (MODE == MODE_DQ) {
            if (out != nullptr) {
                // delta = rowsum(dO o O): this thread's half of the dO row from the swizzled outer tile, its half of the O row from
                // global memory; the halves meet through shared memory
                mbar_wait(outer_full, 0);
                float part = 0.f;
                const long long tok = static_cast<long long>(b) * S + (row_ok ? row : 0);
                const uint4* orow = reinterpret_cast<const uint4*>(out + tok * E + h * AHD + 32 * half);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int chunk = 4 * half + c;
                    const uint4 g = *reinterpret_cast<const uint4*>(s_outer1 + tid * 128 + ((chunk ^ (tid & 7)) << 4));
                    const uint4 o = row_ok ? orow[c] : make_uint4(0u, 0u, 0u, 0u);
                    const uint32_t gw[4] = {g.x, g.y, g.z, g.w}, ow[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float2 a = unpack_bf16x2(gw[i]), bb = o_f16 ? unpack_f16x2(ow[i]) : unpack_bf16x2(ow[i]);
                        part = fmaf(a.x, bb.x, part);
                        part = fmaf(a.y, bb.y, part);
                    }
                }
                s_vec[half * AT + tid] = part;
                asm volatile("bar.sync 1, 256;" ::: "memory");
                my_del = s_vec[tid] + s_vec[AT + tid];
                if (half == 0 && row_ok) delta_out[bh * S + row] = my_del;
            } else if (row_ok) {
                my_del = delta_in[bh * S + row];
            }
        }
        for (int it = 0; it < n_it; ++it) {
            const int bf = it & 1, stage = it % BST;
            mbar_wait(&s_full[bf], (it >> 1) & 1);
            tc_fence_after();
            const uint32_t tS = tl + 64 * bf + 16 * half, tD = tS + 32;
            const int c0 = it * BT + 16 * half;
            const float* v_lse = s_vec + stage * 2 * BT + 16 * half;
            const float* v_del = v_lse + BT;
            uint32_t sA[16], dA[16];
            tmem_ld16(tS, sA);
            tmem_ld16(tD, dA);
            tmem_ld_wait();
            uint32_t pkp[8], pks[8];
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
                const float s0 = __uint_as_float(sA[i]), s1 = __uint_as_float(sA[i + 1]);
                const float g0 = __uint_as_float(dA[i]), g1 = __uint_as_float(dA[i + 1]);
                float p0, p1, d0, d1;
                if (MODE == MODE_DQ) {
                    p0 = fast_exp2(fmaf(s0, L2E, -my_lse));
                    p1 = fast_exp2(fmaf(s1, L2E, -my_lse));
                    d0 = p0 * (g0 - my_del);
                    d1 = p1 * (g1 - my_del);
                } else {
                    const float2 l2 = *reinterpret_cast<const float2*>(v_lse + i);
                    const float2 e2 = *reinterpret_cast<const float2*>(v_del + i);
                    p0 = fast_exp2(fmaf(s0, L2E, -l2.x));
                    p1 = fast_exp2(fmaf(s1, L2E, -l2.y));
                    d0 = p0 * (g0 - e2.x);
                    d1 = p1 * (g1 - e2.y);
                }
                pks[i / 2] = pack_bf16x2(d0, d1);
                if (MODE == MODE_DKV) pkp[i / 2] = pack_bf16x2(p0, p1);
            }
            if (MODE == MODE_DQ) {
                tmem_st8(tS, pks);           // dS
            } else {
                tmem_st8(tS, pkp);           // P^T
                tmem_st8(tD, pks);           // dS^T
            }
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(&ew_done[bf]);
        }
        mbar_wait(acc_full, 0);
        tc_fence_after();
        uint32_t a0r[32];
        tmem_ld32(tl + C_ACC0 + 32 * half, a0r);
        tmem_ld_wait();
        const long long tok = static_cast<long long>(b) * S + row;
        __nv_bfloat16* drow = dqkv + tok * 3 * E + h * AHD + 32 * half;
        if (MODE == MODE_DQ) {
            if (row_ok) store_row_bf16_32(drow, a0r);
        } else {
            if (row_ok) store_row_bf16_32(drow + 2 * E, a0r);     // dV
            tmem_ld32(tl + C_ACC1 + 32 * half, a0r);
            tmem_ld_wait();
            if (row_ok) store_row_bf16_32(drow + E, a0r);         // dK
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == W_MMA) {
        tc_fence_after();
        tmem_dealloc(tb, TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------------
// dedicated forward kernel: as MODE_FWD above (two passes over the keys, P kept in TMEM) but with TWO score buffers in
// TMEM and a software-pipelined MMA stream - Q K_{it+1}^T is queued before P_{it} V_{it}, so the tensor core works on
// the next score tile while the exp stage of the current one runs; the per-iteration critical path is the exp stage
// alone instead of MMA -> barrier -> exp -> barrier -> MMA.  4-stage K/V ring.  256 TMEM columns, two CTAs per SM.
// ------------------------------------------------------------------------------------------------
constexpr int FST = 4;                          // K/V stages
constexpr int FTI = 64;                         // keys per step
constexpr int FITILE = FTI * AHD * 2;           // 8 KB
struct FwdSmem {
    static constexpr int Q = 0;
    static constexpr int INNER = ATILE;
    static constexpr int BAR = INNER + FST * 2 * FITILE;
    static constexpr int TOTAL = BAR + 16 * 8 + 1024;
};

__global__ void __launch_bounds__(ATC_THREADS, 2)
attn_tc_fwd_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_in, int S, int H,
                   __nv_bfloat16* __restrict__ out, float* __restrict__ out32, float* __restrict__ lse_out, int o_f16) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint8_t* s_q = smem + FwdSmem::Q;
    uint8_t* s_inner = smem + FwdSmem::INNER;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + FwdSmem::BAR);
    uint64_t* q_full = bars;
    uint64_t* in_full = bars + 1;       // [FST]
    uint64_t* in_empty = bars + 5;      // [FST]
    uint64_t* s_full = bars + 9;        // [2]
    uint64_t* ew_done = bars + 11;      // [2]
    uint64_t* acc_full = bars + 13;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);

    const int ot = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int E = H * AHD;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_in = (S + FTI - 1) / FTI;
    const int n_it = 2 * n_in;
    constexpr uint32_t TMEM_COLS = 256, C_O = 128;
    const long long bh = static_cast<long long>(b) * H + h;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&map_q);
        tma_prefetch_desc(&map_in);
        mbar_init(q_full, 1);
        for (int s = 0; s < FST; ++s) {
            mbar_init(&in_full[s], 1);
            mbar_init(&in_empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&s_full[s], 1);
            mbar_init(&ew_done[s], 128);
        }
        mbar_init(acc_full, 1);
        fence_barrier_init();
    }
    if (warp == 5) tmem_alloc(tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = *tmem_slot;
    pdl_wait();        // set-up above is private; q / k / v (and dO, lse, delta) come from the preceding kernels
    pdl_trigger();

    if (warp == 4) {
        const int cq = h * AHD, ck = E + h * AHD, cv = 2 * E + h * AHD;
        if (elect_one()) {
            mbar_expect_tx(q_full, ATILE);
            tma_load_3d(s_q, &map_q, cq, ot * AT, b, q_full);
        }
        for (int it = 0; it < n_it; ++it) {
            const int stage = it % FST, par = (it / FST) & 1;
            const int j = it % n_in;
            const bool second = it >= n_in;
            mbar_wait(&in_empty[stage], par ^ 1);
            if (elect_one()) {
                uint8_t* t0 = s_inner + stage * 2 * FITILE;
                mbar_expect_tx(&in_full[stage], second ? 2 * FITILE : FITILE);
                tma_load_3d(t0, &map_in, ck, j * FTI, b, &in_full[stage]);
                if (second) tma_load_3d(t0 + FITILE, &map_in, cv, j * FTI, b, &in_full[stage]);
            }
            __syncwarp();
        }
    } else if (warp == 5) {
        constexpr uint32_t idesc_s = umma_idesc_bf16(AT, FTI, 0, 0);
        constexpr uint32_t idesc_acc = umma_idesc_bf16(AT, AHD, 0, 1);
        mbar_wait(q_full, 0);
        tc_fence_after();
        const uint64_t a0 = umma_desc_sw128(smem_u32(s_q));
        auto issue_pv = [&](int j) {     // O += P_j V_j ; releases the K/V stage of iteration j
            const int bfj = j & 1, stg = j % FST;
            mbar_wait(&ew_done[bfj], (j >> 1) & 1);
            tc_fence_after();
            if (elect_one()) {
                const uint64_t bv = umma_desc_sw128(smem_u32(s_inner + stg * 2 * FITILE + FITILE));
#pragma unroll
                for (int k = 0; k < FTI / 16; ++k)
                    umma_ts(tb + C_O, tb + 64 * bfj + 8 * k, bv + 128 * k, idesc_acc, (j > n_in || k > 0) ? 1u : 0u);
                umma_commit(&in_empty[stg]);
                if (j == n_it - 1) umma_commit(acc_full);
            }
            __syncwarp();
        };
        for (int it = 0; it < n_it; ++it) {
            const int stage = it % FST, par = (it / FST) & 1, bf = it & 1;
            mbar_wait(&in_full[stage], par);
            tc_fence_after();
            if (it >= 2 && (it - 2) < n_in) {     // this score buffer was last read by a max-only iteration
                mbar_wait(&ew_done[bf], ((it - 2) >> 1) & 1);
                tc_fence_after();
            }
            if (elect_one()) {
                const uint64_t bk = umma_desc_sw128(smem_u32(s_inner + stage * 2 * FITILE));
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_ss(tb + 64 * bf, a0 + 2 * k, bk + 2 * k, idesc_s, k > 0);
                umma_commit(&s_full[bf]);
                if (it < n_in) umma_commit(&in_empty[stage]);
            }
            __syncwarp();
            if (it - 1 >= n_in) issue_pv(it - 1);
        }
        issue_pv(n_it - 1);
    } else {
        const int tid = threadIdx.x;
        const uint32_t tl = tb + (static_cast<uint32_t>(warp * 32) << 16);
        const int row = ot * AT + tid;
        const bool row_ok = row < S;
        float m = -INFINITY;
        for (int it = 0; it < n_in; ++it) {            // pass 1: row maximum
            const int bf = it & 1;
            mbar_wait(&s_full[bf], (it >> 1) & 1);
            tc_fence_after();
            const int k0 = it * FTI;
            uint32_t r0[32], r1[32];
            tmem_ld32(tl + 64 * bf, r0);
            tmem_ld32(tl + 64 * bf + 32, r1);
            tmem_ld_wait();
            if (k0 + FTI <= S) {
#pragma unroll
                for (int i = 0; i < 32; ++i) m = fmaxf(m, fmaxf(__uint_as_float(r0[i]), __uint_as_float(r1[i])));
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    if (k0 + i < S) m = fmaxf(m, __uint_as_float(r0[i]));
                    if (k0 + 32 + i < S) m = fmaxf(m, __uint_as_float(r1[i]));
                }
            }
            tc_fence_before();
            mbar_arrive(&ew_done[bf]);
        }
        const float mL = (m == -INFINITY) ? 0.f : m * L2E;
        float l = 0.f;
        for (int it = n_in; it < n_it; ++it) {         // pass 2: P = exp2(S - m) in place, row sum
            const int bf = it & 1;
            mbar_wait(&s_full[bf], (it >> 1) & 1);
            tc_fence_after();
            const int k0 = (it - n_in) * FTI;
            uint32_t r0[32], r1[32];
            tmem_ld32(tl + 64 * bf, r0);
            tmem_ld32(tl + 64 * bf + 32, r1);
            tmem_ld_wait();
            const bool full = k0 + FTI <= S;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                uint32_t pk[16];
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                    float p0 = fast_exp2(fmaf(__uint_as_float(c == 0 ? r0[i] : r1[i]), L2E, -mL));
                    float p1 = fast_exp2(fmaf(__uint_as_float(c == 0 ? r0[i + 1] : r1[i + 1]), L2E, -mL));
                    if (!full) {
                        if (k0 + 32 * c + i >= S) p0 = 0.f;
                        if (k0 + 32 * c + i + 1 >= S) p1 = 0.f;
                    }
                    l += p0 + p1;
                    pk[i / 2] = pack_bf16x2(p0, p1);
                }
                tmem_st16(tl + 64 * bf + 16 * c, pk);
            }
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(&ew_done[bf]);
        }
        mbar_wait(acc_full, 0);
        tc_fence_after();
        uint32_t o0[32], o1[32];
        tmem_ld32(tl + C_O, o0);
        tmem_ld32(tl + C_O + 32, o1);
        tmem_ld_wait();
        if (row_ok) {
            const float inv = l > 0.f ? 1.0f / l : 0.f;
            const long long tok = static_cast<long long>(b) * S + row;
            store_row_bf16_64(out + tok * E + h * AHD, o0, o1, inv, o_f16);
            if (out32) {
                float* f = out32 + tok * E + h * AHD;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    uint32_t w[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int e = 8 * q + i;
                        w[i] = __float_as_uint(round_tf32_rn(__uint_as_float(e < 32 ? o0[e] : o1[e - 32]) * inv));     // feeds a tf32 GEMM
                    }
                    st_global_256(f + 8 * q, w);
                }
            }
            lse_out[bh * S + row] = (m == -INFINITY ? 0.f : m) + logf(fmaxf(l, 1e-30f));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 5) {
        tc_fence_after();
        tmem_dealloc(tb, TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------------
// one-pass forward: the two-pass kernel above reads every fp32 score out of TMEM twice (734 MB per layer at ~64 B/clk/SM
// = ~45 us, its floor).  Here each score tile is read once: the running row maximum is kept STALE until a tile exceeds it
// by more than 2^8 (probabilities then stay below 256, exact in fp32 and safe in bf16), and only then is O rescaled in
// TMEM - by the element-wise warp itself, after the previous P V has completed (o_done).  With real attention logits the
// first tile fixes the maximum and rescales are rare.  Same pipeline otherwise: two score buffers, Q K_{j+1}^T queued
// before P_j V_j, 4-stage K/V ring, 256 TMEM columns, two CTAs per SM.
// ------------------------------------------------------------------------------------------------
constexpr float RESCALE_LOG2 = 8.0f;

__global__ void __launch_bounds__(ATC_THREADS, 2)
attn_tc_fwd1_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_in, int S, int H,
                    __nv_bfloat16* __restrict__ out, float* __restrict__ out32, float* __restrict__ lse_out, int o_f16) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint8_t* s_q = smem + FwdSmem::Q;
    uint8_t* s_inner = smem + FwdSmem::INNER;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + FwdSmem::BAR);
    uint64_t* q_full = bars;
    uint64_t* in_full = bars + 1;       // [FST]
    uint64_t* in_empty = bars + 5;      // [FST]
    uint64_t* s_full = bars + 9;        // [2]
    uint64_t* ew_done = bars + 11;      // [2]
    uint64_t* o_done = bars + 13;       // one phase per P V
    uint64_t* acc_full = bars + 14;     // completes once, with the LAST P V
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 15);

    const int ot = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int E = H * AHD;
    const int warp = threadIdx.x >> 5;
    const int n_it = (S + FTI - 1) / FTI;
    constexpr uint32_t TMEM_COLS = 256, C_O = 128;
    const long long bh = static_cast<long long>(b) * H + h;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&map_q);
        tma_prefetch_desc(&map_in);
        mbar_init(q_full, 1);
        for (int s = 0; s < FST; ++s) {
            mbar_init(&in_full[s], 1);
            mbar_init(&in_empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&s_full[s], 1);
            mbar_init(&ew_done[s], 128);
        }
        mbar_init(o_done, 1);
        mbar_init(acc_full, 1);
        fence_barrier_init();
    }
    if (warp == 5) tmem_alloc(tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = *tmem_slot;
    pdl_wait();
    pdl_trigger();

    if (warp == 4) {
        const int cq = h * AHD, ck = E + h * AHD, cv = 2 * E + h * AHD;
        if (elect_one()) {
            mbar_expect_tx(q_full, ATILE);
            tma_load_3d(s_q, &map_q, cq, ot * AT, b, q_full);
        }
        for (int it = 0; it < n_it; ++it) {
            const int stage = it % FST, par = (it / FST) & 1;
            mbar_wait(&in_empty[stage], par ^ 1);
            if (elect_one()) {
                uint8_t* t0 = s_inner + stage * 2 * FITILE;
                mbar_expect_tx(&in_full[stage], 2 * FITILE);
                tma_load_3d(t0, &map_in, ck, it * FTI, b, &in_full[stage]);
                tma_load_3d(t0 + FITILE, &map_in, cv, it * FTI, b, &in_full[stage]);
            }
            __syncwarp();
        }
    } else if (warp == 5) {
        constexpr uint32_t idesc_s = umma_idesc_bf16(AT, FTI, 0, 0);
        constexpr uint32_t idesc_acc = umma_idesc_bf16(AT, AHD, 0, 1);
        mbar_wait(q_full, 0);
        tc_fence_after();
        const uint64_t a0 = umma_desc_sw128(smem_u32(s_q));
        auto issue_pv = [&](int j) {     // O += P_j V_j ; releases the K/V stage of step j, completes phase j of o_done
            const int bfj = j & 1, stg = j % FST;
            mbar_wait(&ew_done[bfj], (j >> 1) & 1);
            tc_fence_after();
            if (elect_one()) {
                const uint64_t bv = umma_desc_sw128(smem_u32(s_inner + stg * 2 * FITILE + FITILE));
#pragma unroll
                for (int k = 0; k < FTI / 16; ++k)
                    umma_ts(tb + C_O, tb + 64 * bfj + 8 * k, bv + 128 * k, idesc_acc, (j > 0 || k > 0) ? 1u : 0u);
                umma_commit(&in_empty[stg]);
                umma_commit(o_done);
                if (j == n_it - 1) umma_commit(acc_full);
            }
            __syncwarp();
        };
        for (int it = 0; it < n_it; ++it) {
            const int stage = it % FST, par = (it / FST) & 1, bf = it & 1;
            mbar_wait(&in_full[stage], par);
            tc_fence_after();
            // score buffer bf was last read as P by P V of step it - 2: issued earlier in program order, after its ew_done
            if (elect_one()) {
                const uint64_t bk = umma_desc_sw128(smem_u32(s_inner + stage * 2 * FITILE));
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_ss(tb + 64 * bf, a0 + 2 * k, bk + 2 * k, idesc_s, k > 0);
                umma_commit(&s_full[bf]);
            }
            __syncwarp();
            if (it >= 1) issue_pv(it - 1);
        }
        issue_pv(n_it - 1);
    } else {
        const int tid = threadIdx.x;
        const uint32_t tl = tb + (static_cast<uint32_t>(warp * 32) << 16);
        const int row = ot * AT + tid;
        const bool row_ok = row < S;
        float mL = -INFINITY;          // (stale) running maximum, in log2 units
        float l = 0.f;
        for (int it = 0; it < n_it; ++it) {
            const int bf = it & 1;
            mbar_wait(&s_full[bf], (it >> 1) & 1);
            tc_fence_after();
            const int k0 = it * FTI;
            uint32_t r0[32], r1[32];
            tmem_ld32(tl + 64 * bf, r0);
            tmem_ld32(tl + 64 * bf + 32, r1);
            tmem_ld_wait();
            const bool full = k0 + FTI <= S;
            float mt = -INFINITY;
            if (full) {
#pragma unroll
                for (int i = 0; i < 32; ++i) mt = fmaxf(mt, fmaxf(__uint_as_float(r0[i]), __uint_as_float(r1[i])));
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    if (k0 + i < S) mt = fmaxf(mt, __uint_as_float(r0[i]));
                    if (k0 + 32 + i < S) mt = fmaxf(mt, __uint_as_float(r1[i]));
                }
            }
            const float mtL = mt * L2E;
            const bool need = mtL > mL + RESCALE_LOG2;      // always true on the first tile (mL = -inf, key 0 is valid)
            if (__any_sync(0xffffffffu, need)) {
                const float scale = need ? fast_exp2(mL - mtL) : 1.0f;      // exp2(-inf) = 0 on the first tile
                if (it > 0) {
                    // the previous P V must have landed in O before it is rescaled; the next one is not issued before ew_done
                    mbar_wait(o_done, (it - 1) & 1);
                    tc_fence_after();
                    uint32_t o0[32], o1[32];
                    tmem_ld32(tl + C_O, o0);
                    tmem_ld32(tl + C_O + 32, o1);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        o0[i] = __float_as_uint(__uint_as_float(o0[i]) * scale);
                        o1[i] = __float_as_uint(__uint_as_float(o1[i]) * scale);
                    }
                    tmem_st32(tl + C_O, o0);
                    tmem_st32(tl + C_O + 32, o1);
                }
                l *= scale;
                if (need) mL = mtL;
            }
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                uint32_t pk[16];
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                    float p0 = fast_exp2(fmaf(__uint_as_float(c == 0 ? r0[i] : r1[i]), L2E, -mL));
                    float p1 = fast_exp2(fmaf(__uint_as_float(c == 0 ? r0[i + 1] : r1[i + 1]), L2E, -mL));
                    if (!full) {
                        if (k0 + 32 * c + i >= S) p0 = 0.f;
                        if (k0 + 32 * c + i + 1 >= S) p1 = 0.f;
                    }
                    l += p0 + p1;
                    pk[i / 2] = pack_bf16x2(p0, p1);
                }
                tmem_st16(tl + 64 * bf + 16 * c, pk);
            }
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(&ew_done[bf]);
        }
        // NOT o_done: a parity wait cannot tell phase n_it - 1 from phase n_it - 3.  When this warp leaves the loop only P V of
        // step n_it - 3 is known to have completed (it precedes the last Q K^T in the MMA stream); if P V (n_it - 2) is still in
        // flight, o_done's CURRENT parity differs from (n_it - 1) & 1 and the wait falls through - O was then read two
        // accumulations early (round 2: non-reproducible CRIS decoder attention, S = 676, one run in ~5).  acc_full
        // completes exactly once, with the last P V.
        mbar_wait(acc_full, 0);
        tc_fence_after();
        uint32_t o0[32], o1[32];
        tmem_ld32(tl + C_O, o0);
        tmem_ld32(tl + C_O + 32, o1);
        tmem_ld_wait();
        if (row_ok) {
            const float inv = l > 0.f ? 1.0f / l : 0.f;
            const long long tok = static_cast<long long>(b) * S + row;
            store_row_bf16_64(out + tok * E + h * AHD, o0, o1, inv, o_f16);
            if (out32) {
                float* f = out32 + tok * E + h * AHD;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    uint32_t w[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int e = 8 * q + i;
                        w[i] = __float_as_uint(round_tf32_rn(__uint_as_float(e < 32 ? o0[e] : o1[e - 32]) * inv));     // feeds a tf32 GEMM
                    }
                    st_global_256(f + 8 * q, w);
                }
            }
            lse_out[bh * S + row] = (mL == -INFINITY ? 0.f : mL * 0.6931471805599453f) + logf(fmaxf(l, 1e-30f));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 5) {
        tc_fence_after();
        tmem_dealloc(tb, TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------------
// one-pass forward, PERSISTENT (round 2, the default): attn_tc_fwd1_kernel's arithmetic and pipeline with 2 x #SM resident CTAs
// walking the (query tile, head, sample) items - see attn_tc_bwdp_kernel for what a per-CTA timeline showed (a quarter of a
// one-item CTA's life is set-up and the first TMA round trip).  Q is double-buffered and requested one item ahead, the K / V
// ring and the two score buffers never drain between items, and the O read-out of item k overlaps Q K^T of item k + 1.
// Measured: 57.9 -> 49.8 us per layer (B = 32, S = 489, H = 12).  A variant with EIGHT softmax warps (two per TMEM lane quarter,
// 32 score columns each, tile maximum and row sum exchanged through shared memory behind a 64-thread named barrier per step)
// was slower, 53.6 us: the extra barrier and exchange per step cost more than the shorter per-thread chain saves.
// ------------------------------------------------------------------------------------------------
struct FwdPSmem {
    static constexpr int Q = 0;                                   // [2] x ATILE
    static constexpr int INNER = 2 * ATILE;                       // [FST][2] x FITILE
    static constexpr int BAR = INNER + FST * 2 * FITILE;
    static constexpr int NBAR = 4 + 2 * FST + 6;                  // q_full[2], q_empty[2], in_full / in_empty[FST], s_full[2], ew_done[2], o_done, acc_full
    static constexpr int TOTAL = BAR + (NBAR + 1) * 8 + 1024;
};

__global__ void __launch_bounds__(ATC_THREADS, 2)
attn_tc_fwd1p_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_in, int S, int H, int n_ot, int n_items,
                     __nv_bfloat16* __restrict__ out, float* __restrict__ out32, float* __restrict__ lse_out, int o_f16) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint8_t* s_q = smem + FwdPSmem::Q;
    uint8_t* s_inner = smem + FwdPSmem::INNER;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + FwdPSmem::BAR);
    uint64_t* q_full = bars;                    // [2]
    uint64_t* q_empty = bars + 2;               // [2]
    uint64_t* in_full = bars + 4;               // [FST]
    uint64_t* in_empty = bars + 4 + FST;        // [FST]
    uint64_t* s_full = bars + 4 + 2 * FST;      // [2]
    uint64_t* ew_done = s_full + 2;             // [2]
    uint64_t* o_done = s_full + 4;              // one phase per P V
    uint64_t* acc_full = s_full + 5;            // one phase per item, with its LAST P V
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_full + 6);

    const int E = H * AHD;
    const int warp = threadIdx.x >> 5;
    const int n_it = (S + FTI - 1) / FTI;
    constexpr uint32_t TMEM_COLS = 256, C_O = 128;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&map_q);
        tma_prefetch_desc(&map_in);
        for (int u = 0; u < 2; ++u) {
            mbar_init(&q_full[u], 1);
            mbar_init(&q_empty[u], 1);
            mbar_init(&s_full[u], 1);
            mbar_init(&ew_done[u], 128);
        }
        for (int s = 0; s < FST; ++s) {
            mbar_init(&in_full[s], 1);
            mbar_init(&in_empty[s], 1);
        }
        mbar_init(o_done, 1);
        mbar_init(acc_full, 1);
        fence_barrier_init();
    }
    if (warp == 5) tmem_alloc(tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = *tmem_slot;
    pdl_wait();
    pdl_trigger();

    auto item = [&](int w, int& ot, int& h, int& b) {
        ot = w % n_ot;
        const int hb = w / n_ot;
        h = hb % H;
        b = hb / H;
    };

    if (warp == 4) {
        // ------------------------------------------------------------------ TMA producer
        auto load_q = [&](int k, int w) {
            int ot, h, b;
            item(w, ot, h, b);
            const int u = k & 1;
            mbar_wait(&q_empty[u], ((k >> 1) & 1) ^ 1);
            if (elect_one()) {
                mbar_expect_tx(&q_full[u], ATILE);
                tma_load_3d(s_q + u * ATILE, &map_q, h * AHD, ot * AT, b, &q_full[u]);
            }
            __syncwarp();
        };
        int g = 0, k = 0;
        if (static_cast<int>(blockIdx.x) < n_items) load_q(0, blockIdx.x);
        for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++k) {
            int ot, h, b;
            item(w, ot, h, b);
            const int ck = E + h * AHD, cv = 2 * E + h * AHD;
            const int w_next = w + gridDim.x;
            bool prefetched = w_next >= n_items;
            for (int it = 0; it < n_it; ++it, ++g) {
                const int stage = g % FST, par = (g / FST) & 1;
                mbar_wait(&in_empty[stage], par ^ 1);
                if (elect_one()) {
                    uint8_t* t0 = s_inner + stage * 2 * FITILE;
                    mbar_expect_tx(&in_full[stage], 2 * FITILE);
                    tma_load_3d(t0, &map_in, ck, it * FTI, b, &in_full[stage]);
                    tma_load_3d(t0 + FITILE, &map_in, cv, it * FTI, b, &in_full[stage]);
                }
                __syncwarp();
                // the ring has wrapped inside this item: its step 0 has retired, hence the whole previous item - the Q buffer of
                // item k + 1 (last used by item k - 1) is free
                if (!prefetched && it == FST) {
                    load_q(k + 1, w_next);
                    prefetched = true;
                }
            }
            if (!prefetched) load_q(k + 1, w_next);
        }
    } else if (warp == 5) {
        // ------------------------------------------------------------------ MMA issuer
        constexpr uint32_t idesc_s = umma_idesc_bf16(AT, FTI, 0, 0);
        constexpr uint32_t idesc_acc = umma_idesc_bf16(AT, AHD, 0, 1);
        int g = 0, k = 0;
        for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++k) {
            const int u = k & 1;
            mbar_wait(&q_full[u], (k >> 1) & 1);
            tc_fence_after();
            const uint64_t a0 = umma_desc_sw128(smem_u32(s_q + u * ATILE));
            // O += P_j V_j for step j of this item (global step gj); releases its K / V stage, completes one phase of o_done.
            // ew_done of the item's step 0 also tells that the element-wise warps have read O of the previous item out of TMEM
            auto issue_pv = [&](int j, int gj) {
                const int bfj = gj & 1, stg = gj % FST;
                mbar_wait(&ew_done[bfj], (gj >> 1) & 1);
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t bv = umma_desc_sw128(smem_u32(s_inner + stg * 2 * FITILE + FITILE));
#pragma unroll
                    for (int kk = 0; kk < FTI / 16; ++kk)
                        umma_ts(tb + C_O, tb + 64 * bfj + 8 * kk, bv + 128 * kk, idesc_acc, (j > 0 || kk > 0) ? 1u : 0u);
                    umma_commit(&in_empty[stg]);
                    umma_commit(o_done);
                    if (j == n_it - 1) {
                        umma_commit(acc_full);
                        umma_commit(&q_empty[u]);
                    }
                }
                __syncwarp();
            };
            for (int it = 0; it < n_it; ++it, ++g) {
                const int stage = g % FST, par = (g / FST) & 1, bf = g & 1;
                mbar_wait(&in_full[stage], par);
                tc_fence_after();
                // score buffer bf was last read as P by P V of global step g - 2: issued earlier in program order, after its ew_done
                if (elect_one()) {
                    const uint64_t bk = umma_desc_sw128(smem_u32(s_inner + stage * 2 * FITILE));
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) umma_ss(tb + 64 * bf, a0 + 2 * kk, bk + 2 * kk, idesc_s, kk > 0);
                    umma_commit(&s_full[bf]);
                }
                __syncwarp();
                if (it >= 1) issue_pv(it - 1, g - 1);
            }
            issue_pv(n_it - 1, g - 1);
        }
    } else {
        // ------------------------------------------------------------------ softmax warps (thread = TMEM lane = query row)
        const int tid = threadIdx.x;
        const uint32_t tl = tb + (static_cast<uint32_t>(warp * 32) << 16);
        int g = 0, k = 0;
        for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++k) {
            int ot, h, b;
            item(w, ot, h, b);
            const long long bh = static_cast<long long>(b) * H + h;
            const int row = ot * AT + tid;
            const bool row_ok = row < S;
            float mL = -INFINITY;          // (stale) running maximum, in log2 units
            float l = 0.f;
            for (int it = 0; it < n_it; ++it, ++g) {
                const int bf = g & 1;
                mbar_wait(&s_full[bf], (g >> 1) & 1);
                tc_fence_after();
                const int k0 = it * FTI;
                uint32_t r0[32], r1[32];
                tmem_ld32(tl + 64 * bf, r0);
                tmem_ld32(tl + 64 * bf + 32, r1);
                tmem_ld_wait();
                const bool full = k0 + FTI <= S;
                float mt = -INFINITY;
                if (full) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) mt = fmaxf(mt, fmaxf(__uint_as_float(r0[i]), __uint_as_float(r1[i])));
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        if (k0 + i < S) mt = fmaxf(mt, __uint_as_float(r0[i]));
                        if (k0 + 32 + i < S) mt = fmaxf(mt, __uint_as_float(r1[i]));
                    }
                }
                const float mtL = mt * L2E;
                const bool need = mtL > mL + RESCALE_LOG2;      // always true on the first tile (mL = -inf, key 0 is valid)
                if (__any_sync(0xffffffffu, need)) {
                    const float scale = need ? fast_exp2(mL - mtL) : 1.0f;      // exp2(-inf) = 0 on the first tile
                    if (it > 0) {
                        // the previous P V (global step g - 1) must have landed in O before it is rescaled; when s_full of step g has
                        // completed, every P V up to step g - 2 has, so the parity wait cannot alias
                        mbar_wait(o_done, (g - 1) & 1);
                        tc_fence_after();
                        uint32_t o0[32], o1[32];
                        tmem_ld32(tl + C_O, o0);
                        tmem_ld32(tl + C_O + 32, o1);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            o0[i] = __float_as_uint(__uint_as_float(o0[i]) * scale);
                            o1[i] = __float_as_uint(__uint_as_float(o1[i]) * scale);
                        }
                        tmem_st32(tl + C_O, o0);
                        tmem_st32(tl + C_O + 32, o1);
                    }
                    l *= scale;
                    if (need) mL = mtL;
                }
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    uint32_t pk[16];
#pragma unroll
                    for (int i = 0; i < 32; i += 2) {
                        float p0 = fast_exp2(fmaf(__uint_as_float(c == 0 ? r0[i] : r1[i]), L2E, -mL));
                        float p1 = fast_exp2(fmaf(__uint_as_float(c == 0 ? r0[i + 1] : r1[i + 1]), L2E, -mL));
                        if (!full) {
                            if (k0 + 32 * c + i >= S) p0 = 0.f;
                            if (k0 + 32 * c + i + 1 >= S) p1 = 0.f;
                        }
                        l += p0 + p1;
                        pk[i / 2] = pack_bf16x2(p0, p1);
                    }
                    tmem_st16(tl + 64 * bf + 16 * c, pk);
                }
                tmem_st_wait();
                tc_fence_before();
                mbar_arrive(&ew_done[bf]);
            }
            mbar_wait(acc_full, k & 1);       // one phase per item (completes with the item's last P V)
            tc_fence_after();
            uint32_t o0[32], o1[32];
            tmem_ld32(tl + C_O, o0);
            tmem_ld32(tl + C_O + 32, o1);
            tmem_ld_wait();
            tc_fence_before();                // orders the O read before the ew_done arrive of the next item's first step
            if (row_ok) {
                const float inv = l > 0.f ? 1.0f / l : 0.f;
                const long long tok = static_cast<long long>(b) * S + row;
                store_row_bf16_64(out + tok * E + h * AHD, o0, o1, inv, o_f16);
                if (out32) {
                    float* f = out32 + tok * E + h * AHD;
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        uint32_t wv[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int e = 8 * q + i;
                            wv[i] = __float_as_uint(round_tf32_rn(__uint_as_float(e < 32 ? o0[e] : o1[e - 32]) * inv));     // feeds a tf32 GEMM
                        }
                        st_global_256(f + 8 * q, wv);
                    }
                }
                lse_out[bh * S + row] = (mL == -INFINITY ? 0.f : mL * 0.6931471805599453f) + logf(fmaxf(l, 1e-30f));
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 5) {
        tc_fence_after();
        tmem_dealloc(tb, TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn3)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn3 encode_fn3() {
    static EncodeTiledFn3 fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn3>(p);
    }
    return fn;
}

// bf16 tensor [B][S][cols] -> 3-D map {cols, S, B}, box {64, 128, 1}, 128-byte swizzle; rows >= S read as zeros
static int make_tmap3(CUtensorMap* map, const void* ptr, int cols, int S, int B, int box_rows = AT) {
    EncodeTiledFn3 fn = encode_fn3();
    TVS_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled not available from the driver");
    cuuint64_t dims[3] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(S), static_cast<cuuint64_t>(B)};
    cuuint64_t strides[2] = {static_cast<cuuint64_t>(cols) * 2, static_cast<cuuint64_t>(S) * cols * 2};
    cuuint32_t box[3] = {AHD, static_cast<cuuint32_t>(box_rows), 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    TVS_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (3-D) failed with %d (cols=%d S=%d B=%d)", (int)r, cols, S, B);
    return 0;
}

template <int MODE>
static int launch_atc(const CUtensorMap& mq, const CUtensorMap& md, const CUtensorMap& mqi, const CUtensorMap& mdi, int B, int S, int H,
                      __nv_bfloat16* out, float* out32, float* lse_out,
                      const float* lse_in, const float* delta_in, __nv_bfloat16* dqkv, cudaStream_t st, int ot0 = 0, int o_f16 = 0) {
    using L = AtcSmem<MODE>;
    auto kern = attn_tc_kernel<MODE>;
    static bool attr_set = false;
    if (!attr_set) {
        TVS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
        attr_set = true;
    }
    dim3 grid((S + AT - 1) / AT - ot0, H, B);
    TVS_CUDA(launch_pdl(kern, grid, dim3(MODE == MODE_FWD ? ATC_THREADS : ATC_THREADS_BWD), L::TOTAL, st, 1, mq, md, mqi, mdi, S, H, out, out32, lse_out, lse_in, delta_in, dqkv, ot0, o_f16));
    return check_launch(MODE == MODE_FWD ? "attn_tc_kernel<fwd>" : (MODE == MODE_DQ ? "attn_tc_kernel<dq>" : "attn_tc_kernel<dkv>"));
}

bool attn_tc_enabled() {
    static const bool off = [] { const char* e = getenv("TVS_ATTN"); return e && e[0] == 'm'; }();   // TVS_ATTN=mma -> legacy kernels
    return !off;
}

int attn_tc_fwd(const void* qkv, int B, int S, int H, void* out, float* out32, float* lse, cudaStream_t st, int o_f16) {
    CUtensorMap mq, mqi;
    if (int rc = make_tmap3(&mq, qkv, 3 * H * AHD, S, B)) return rc;
    if (int rc = make_tmap3(&mqi, qkv, 3 * H * AHD, S, B, 64)) return rc;
    static const bool legacy = [] { const char* e = getenv("TVS_ATTN_FWD"); return e && e[0] == '1'; }();   // TVS_ATTN_FWD=1: single-buffer variant
    if (legacy) return launch_atc<MODE_FWD>(mq, mq, mqi, mqi, B, S, H, static_cast<__nv_bfloat16*>(out), out32, lse, nullptr, nullptr, nullptr, st, 0, o_f16);
    static const bool two_pass = [] { const char* e = getenv("TVS_ATTN_FWD"); return e && e[0] == '2'; }();   // TVS_ATTN_FWD=2: two-pass pipelined variant
    static bool attr_set = false;
    if (!attr_set) {
        TVS_CUDA(cudaFuncSetAttribute(attn_tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FwdSmem::TOTAL));
        TVS_CUDA(cudaFuncSetAttribute(attn_tc_fwd1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FwdSmem::TOTAL));
        attr_set = true;
    }
    static const bool one_item = [] { const char* e = getenv("TVS_ATTN_FWD"); return e && e[0] == '3'; }();   // TVS_ATTN_FWD=3: one CTA per item (round 1 / early round 2)
    if (!two_pass && !one_item) {      // default: persistent one-pass kernel
        static bool attr_p = false;
        if (!attr_p) {
            TVS_CUDA(cudaFuncSetAttribute(attn_tc_fwd1p_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FwdPSmem::TOTAL));
            attr_p = true;
        }
        const int n_ot = (S + AT - 1) / AT;
        const long long n_items = static_cast<long long>(n_ot) * H * B;
        TVS_REQUIRE(n_items < (1LL << 31), "attention forward: too many (tile, head, sample) items");
        const int resident = 2 * sm_count();
        dim3 pgrid(static_cast<unsigned>(n_items < resident ? n_items : resident));
        TVS_CUDA(launch_pdl(attn_tc_fwd1p_kernel, pgrid, dim3(ATC_THREADS), FwdPSmem::TOTAL, st, 1, mq, mqi, S, H, n_ot, static_cast<int>(n_items),
                            static_cast<__nv_bfloat16*>(out), out32, lse, o_f16));
        return check_launch("attn_tc_fwd1p_kernel");
    }
    dim3 grid((S + AT - 1) / AT, H, B);
    if (!two_pass) {
        TVS_CUDA(launch_pdl(attn_tc_fwd1_kernel, grid, dim3(ATC_THREADS), FwdSmem::TOTAL, st, 1, mq, mqi, S, H, static_cast<__nv_bfloat16*>(out), out32, lse, o_f16));
        return check_launch("attn_tc_fwd1_kernel");
    }
    TVS_CUDA(launch_pdl(attn_tc_fwd_kernel, grid, dim3(ATC_THREADS), FwdSmem::TOTAL, st, 1, mq, mqi, S, H, static_cast<__nv_bfloat16*>(out), out32, lse, o_f16));
    return check_launch("attn_tc_fwd_kernel");
}

template <int MODE>
static int launch_bwd2(const CUtensorMap& mq, const CUtensorMap& md, const CUtensorMap& mqi, const CUtensorMap& mdi, int B, int S, int H,
                       const __nv_bfloat16* out, float* delta_out, const float* lse_in, const float* delta_in, __nv_bfloat16* dqkv, cudaStream_t st,
                       int ot0, int o_f16) {
    auto kern = attn_tc_bwd2_kernel<MODE>;
    static bool attr_set = false;
    if (!attr_set) {
        TVS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Bwd2Smem::TOTAL));
        attr_set = true;
    }
    dim3 grid((S + AT - 1) / AT - ot0, H, B);
    TVS_CUDA(launch_pdl(kern, grid, dim3(ATC_THREADS_BWD), Bwd2Smem::TOTAL, st, 1, mq, md, mqi, mdi, S, H, out, delta_out, lse_in, delta_in, dqkv, ot0, o_f16));
    return check_launch(MODE == MODE_DQ ? "attn_tc_bwd2_kernel<dq>" : "attn_tc_bwd2_kernel<dkv>");
}

template <int MODE>
static int launch_bwdp(const CUtensorMap& mq, const CUtensorMap& md, const CUtensorMap& mqi, const CUtensorMap& mdi, int B, int S, int H,
                       const __nv_bfloat16* out, float* delta_out, const float* lse_in, const float* delta_in, __nv_bfloat16* dqkv, cudaStream_t st,
                       int ot0, int o_f16) {
    auto kern = attn_tc_bwdp_kernel<MODE>;
    static bool attr_set = false;
    if (!attr_set) {
        TVS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, BwdPSmem::TOTAL));
        attr_set = true;
    }
    const int n_ot = (S + AT - 1) / AT - ot0;
    const long long n_items = static_cast<long long>(n_ot) * H * B;
    TVS_REQUIRE(n_items < (1LL << 31), "attention backward: too many (tile, head, sample) items");
    const int resident = 2 * sm_count();          // two CTAs per SM (TMEM: 2 x 256 columns; shared memory: 2 x ~99 KB)
    dim3 grid(static_cast<unsigned>(n_items < resident ? n_items : resident));
    TVS_CUDA(launch_pdl(kern, grid, dim3(ATC_THREADS_BWD), BwdPSmem::TOTAL, st, 1, mq, md, mqi, mdi, S, H, n_ot, static_cast<int>(n_items), out, delta_out,
                        lse_in, delta_in, dqkv, ot0, o_f16));
    return check_launch(MODE == MODE_DQ ? "attn_tc_bwdp_kernel<dq>" : "attn_tc_bwdp_kernel<dkv>");
}

// delta must already hold rowsum(dO o O)
// row_begin > 0: dqkv is produced only for the 128-row tiles that contain rows >= row_begin (queries for dQ, keys for dK / dV)
// out_o != nullptr (whole-sequence backward only): delta is computed by the dQ kernel from O and dO and written for the dK / dV kernel
int attn_tc_bwd(const void* qkv, const void* out_o, const void* dout, const float* lse, float* delta, int B, int S, int H, void* dqkv, cudaStream_t st,
                int row_begin, int o_f16) {
    const int ot0 = row_begin / AT;
    // TVS_ATTN_BWD=2: the double-buffered 32-row kernels (attn_tc_bwd2_kernel).  Measured equal to the single-buffer kernels
    // (170.1 vs 170.0 us per layer, B = 32, S = 489): removing the MMA -> element-wise -> MMA serialisation did not move the
    // time, so that chain is not what bounds the backward; the default stays the round-1 pair.
    static const bool v2 = [] { const char* e = getenv("TVS_ATTN_BWD"); return e && e[0] == '2'; }();
    if (v2) {
        CUtensorMap mq, md, mqi, mdi;      // 128-row boxes for the outer tiles, 32-row boxes for the inner ones
        if (int rc = make_tmap3(&mq, qkv, 3 * H * AHD, S, B)) return rc;
        if (int rc = make_tmap3(&md, dout, H * AHD, S, B)) return rc;
        if (int rc = make_tmap3(&mqi, qkv, 3 * H * AHD, S, B, BT)) return rc;
        if (int rc = make_tmap3(&mdi, dout, H * AHD, S, B, BT)) return rc;
        __nv_bfloat16* dq = static_cast<__nv_bfloat16*>(dqkv);
        if (out_o != nullptr && ot0 == 0) {      // dQ first: it produces delta on the way
            if (int rc = launch_bwd2<MODE_DQ>(mq, md, mqi, mdi, B, S, H, static_cast<const __nv_bfloat16*>(out_o), delta, lse, delta, dq, st, 0, o_f16)) return rc;
            return launch_bwd2<MODE_DKV>(mq, md, mqi, mdi, B, S, H, nullptr, nullptr, lse, delta, dq, st, 0, 0);
        }
        if (int rc = launch_bwd2<MODE_DKV>(mq, md, mqi, mdi, B, S, H, nullptr, nullptr, lse, delta, dq, st, ot0, 0)) return rc;
        return launch_bwd2<MODE_DQ>(mq, md, mqi, mdi, B, S, H, nullptr, nullptr, lse, delta, dq, st, ot0, 0);
    }
    CUtensorMap mq, md, mqi, mdi;      // 128-row boxes for the outer tiles, 64-row boxes for the inner ones
    if (int rc = make_tmap3(&mq, qkv, 3 * H * AHD, S, B)) return rc;
    if (int rc = make_tmap3(&md, dout, H * AHD, S, B)) return rc;
    if (int rc = make_tmap3(&mqi, qkv, 3 * H * AHD, S, B, 64)) return rc;
    if (int rc = make_tmap3(&mdi, dout, H * AHD, S, B, 64)) return rc;
    // default: the persistent kernels (attn_tc_bwdp_kernel); TVS_ATTN_BWD=1: one CTA per (tile, head, sample) as in round 1
    static const bool v1 = [] { const char* e = getenv("TVS_ATTN_BWD"); return e && e[0] == '1'; }();
    if (!v1) {
        __nv_bfloat16* dq = static_cast<__nv_bfloat16*>(dqkv);
        if (out_o != nullptr && ot0 == 0) {      // dQ first: it produces delta on the way
            if (int rc = launch_bwdp<MODE_DQ>(mq, md, mqi, mdi, B, S, H, static_cast<const __nv_bfloat16*>(out_o), delta, lse, delta, dq, st, 0, o_f16)) return rc;
            return launch_bwdp<MODE_DKV>(mq, md, mqi, mdi, B, S, H, nullptr, nullptr, lse, delta, dq, st, 0, 0);
        }
        if (int rc = launch_bwdp<MODE_DKV>(mq, md, mqi, mdi, B, S, H, nullptr, nullptr, lse, delta, dq, st, ot0, 0)) return rc;
        return launch_bwdp<MODE_DQ>(mq, md, mqi, mdi, B, S, H, nullptr, nullptr, lse, delta, dq, st, ot0, 0);
    }
    if (out_o != nullptr && ot0 == 0) {
        // dQ first: it produces delta on the way; dK / dV (disjoint columns of dqkv) follow in the stream
        if (int rc = launch_atc<MODE_DQ>(mq, md, mqi, mdi, B, S, H, static_cast<__nv_bfloat16*>(const_cast<void*>(out_o)), delta, nullptr, lse, delta,
                                         static_cast<__nv_bfloat16*>(dqkv), st, 0, o_f16))
            return rc;
        return launch_atc<MODE_DKV>(mq, md, mqi, mdi, B, S, H, nullptr, nullptr, nullptr, lse, delta, static_cast<__nv_bfloat16*>(dqkv), st, 0);
    }
    if (int rc = launch_atc<MODE_DKV>(mq, md, mqi, mdi, B, S, H, nullptr, nullptr, nullptr, lse, delta, static_cast<__nv_bfloat16*>(dqkv), st, ot0)) return rc;
    return launch_atc<MODE_DQ>(mq, md, mqi, mdi, B, S, H, nullptr, nullptr, nullptr, lse, delta, static_cast<__nv_bfloat16*>(dqkv), st, ot0);
}

}  // namespace tvs

#if ATC_EXPERIMENT == 8
extern "C" __attribute__((visibility("default"))) int tvs_debug_timeline(unsigned long long* host, int max_words, int reset) {
    const int words = 3 * 2 * 10 * 512 * 2;
    cudaDeviceSynchronize();
    if (host && max_words >= words) cudaMemcpyFromSymbol(host, tvs::g_tl, sizeof(unsigned long long) * words);
    if (reset) {
        void* p = nullptr;
        cudaGetSymbolAddress(&p, tvs::g_tl);
        cudaMemset(p, 0, sizeof(unsigned long long) * words);
    }
    return words;
}
#endif
