// Fused feed-forward block of the CLIPSeg decoder layer (reduce_dim D = 64, FFN width F = 2048) on tcgen05 / TMEM.
//
// Replaces CLIPSegMLP inside CLIPSegDecoderLayer (transformers modeling_clipseg.py:341-354 as used by :421-431):
//   forward   out = x + relu(x W1^T + b1) W2^T + b2          (the residual of the post-LN block, hf:427-429)
//   dgrad     dx  = g + ((g W2) o [x W1^T + b1 > 0]) W1      (no weight gradient: the decoder is frozen)
// As two GEMMs the [M, 2048] hidden activation costs 128 MB (fp32) per layer and direction at M = 15 648 and the K = 64 /
// N = 64 GEMMs run at 50-120 TFLOP/s; here the hidden units never leave the SM.  The loop has the shape of the attention
// kernels (attention_sm100.cu) with W1 in the role of K, W2^T in the role of V and relu in the role of the softmax:
//   per 64 hidden units j:  S = X W1_j^T (TMEM)  ->  P = relu(S + b1_j) written back to TMEM as the A operand
//                           ->  ACC += P W2T_j   (B = the W2T_j tile addressed MN-major)
//   dgrad:                  S as above, dH = G W2T_j^T, P = dH o [S + b1_j > 0], ACC += P W1_j (the W1_j tile MN-major).
// The decoder decides the logits directly, so bf16 operands are not enough (DESIGN.md "Precision policy"): every fp32
// operand is split into a bf16 head and a bf16 tail (x = hi + lo, |lo| <= 2^-9 |x|) and each product is the three MMAs
// hi*hi + lo*hi + hi*lo (error ~2^-17 relative, tighter than the kind::tf32 GEMMs this replaces); all MMAs are
// kind::f16 with fp32 accumulation.  SPLIT = false keeps the heads only.
//
// One CTA = 128 rows; 320 threads: warps 0-7 element-wise stage (two warps per TMEM lane quarter, each owning half of the
// 64 hidden units of a step: with one warp per scheduler every dependent-instruction latency was exposed, 36 / 46 us;
// two warps per scheduler: see DESIGN.md), warp 8 TMA producer of the weight tiles (3-stage ring), warp 9 MMA issuer.
// The activation rows are read as fp32 by their own threads, split and written into the 128-byte-swizzled operand tiles
// by hand (no activation tensor map, no separate split pass).  P has its own TMEM columns, so no store of the
// element-wise stage can land on a score column another warp has not read yet.
#include "common.cuh"
#include "sm100_ptx.cuh"
#include "tvs_b200.h"

namespace tvs {
using namespace ptx;

constexpr int FT = 128;                  // rows per CTA
constexpr int FD = 64;                   // model width (one 128-byte swizzle row of bf16)
constexpr int FC = 64;                   // hidden units per step
constexpr int F_ATILE = FT * FD * 2;     // 16 KB
constexpr int F_WTILE = FC * FD * 2;     // 8 KB
constexpr int F_NST = 3;                 // weight stages
constexpr int F_THREADS = 320;
constexpr int F_EW = 256;                // element-wise threads
constexpr int F_MAXF = 4096;             // b1 lives in shared memory

template <bool BWD, bool SPLIT>
struct FfnSmem {
    static constexpr int NP = SPLIT ? 2 : 1;                  // parts per operand (hi, lo)
    static constexpr int NA = (BWD ? 2 : 1) * NP;             // activation tiles: x (hi, lo) [, g (hi, lo)]
    static constexpr int NW = 2 * NP;                         // weight tiles per stage: W1 (hi, lo), W2T (hi, lo)
    static constexpr int A = 0;
    static constexpr int W = NA * F_ATILE;
    static constexpr int BIAS = W + F_NST * NW * F_WTILE;
    static constexpr int BAR = BIAS + F_MAXF * 4;
    static constexpr int TOTAL = BAR + 16 * 8 + 1024;
};

__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bar_sync_ew() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// one fp32 row of 64 -> bf16 head (and tail) rows of the swizzled [128][64] operand tiles: 16-byte chunk c of row r sits at
// chunk position c ^ (r & 7) (CU_TENSOR_MAP_SWIZZLE_128B / UMMA layout SWIZZLE_128B on a 1024-byte aligned tile)
template <bool SPLIT>
__device__ __forceinline__ void stage_row(const float* __restrict__ src, bool ok, int r, uint8_t* tile_hi, uint8_t* tile_lo, int c_begin, int c_end) {
#pragma unroll 4
    for (int c = c_begin; c < c_end; ++c) {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
        if (ok) {
            a = reinterpret_cast<const float4*>(src)[2 * c];
            b = reinterpret_cast<const float4*>(src)[2 * c + 1];
        }
        const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const __nv_bfloat16 h0 = __float2bfloat16_rn(v[2 * i]), h1 = __float2bfloat16_rn(v[2 * i + 1]);
            hi[i] = static_cast<uint32_t>(__bfloat16_as_ushort(h0)) | (static_cast<uint32_t>(__bfloat16_as_ushort(h1)) << 16);
            if (SPLIT) lo[i] = pack_bf16x2(v[2 * i] - __bfloat162float(h0), v[2 * i + 1] - __bfloat162float(h1));
        }
        const int pos = r * 128 + ((c ^ (r & 7)) << 4);
        *reinterpret_cast<uint4*>(tile_hi + pos) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        if (SPLIT) *reinterpret_cast<uint4*>(tile_lo + pos) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
}

template <bool BWD, bool SPLIT>
__global__ void __launch_bounds__(F_THREADS, 1)
ffn_tc_kernel(const __grid_constant__ CUtensorMap map_w1_hi, const __grid_constant__ CUtensorMap map_w1_lo,
              const __grid_constant__ CUtensorMap map_w2t_hi, const __grid_constant__ CUtensorMap map_w2t_lo, const float* __restrict__ x,
              const float* __restrict__ g, const float* __restrict__ b1, const float* __restrict__ b2, long long M, int F,
              float* __restrict__ out) {
    using L = FfnSmem<BWD, SPLIT>;
    constexpr int NP = L::NP, NW = L::NW;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint8_t* s_a = smem + L::A;              // x hi [, x lo] [, g hi [, g lo]]
    uint8_t* s_w = smem + L::W;              // stage s: W1 hi [, W1 lo], W2T hi [, W2T lo]
    float* s_b1 = reinterpret_cast<float*>(smem + L::BIAS);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::BAR);
    uint64_t* a_ready = bars;
    uint64_t* w_full = bars + 1;        // [F_NST]
    uint64_t* w_empty = bars + 4;       // [F_NST]
    uint64_t* s_full = bars + 7;        // [2]
    uint64_t* ew_done = bars + 9;       // [2]
    uint64_t* acc_full = bars + 11;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

    const int warp = threadIdx.x >> 5;
    const int n_it = F / FC;
    // TMEM: buffer bf at columns [128 bf, 128 bf + 128): S at +0, dH at +64; accumulator at 256 (64 columns);
    // P of buffer bf at 320 + 64 bf: packed head in columns [0, 32), packed tail in [32, 64)
    constexpr uint32_t TMEM_COLS = 512, C_DH = 64, C_ACC = 256, C_P = 320, C_LO = 32;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&map_w1_hi);
        tma_prefetch_desc(&map_w2t_hi);
        if (SPLIT) {
            tma_prefetch_desc(&map_w1_lo);
            tma_prefetch_desc(&map_w2t_lo);
        }
        mbar_init(a_ready, F_EW);
        for (int s = 0; s < F_NST; ++s) {
            mbar_init(&w_full[s], 1);
            mbar_init(&w_empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&s_full[s], 1);
            mbar_init(&ew_done[s], F_EW);
        }
        mbar_init(acc_full, 1);
        fence_barrier_init();
    }
    if (warp == 9) tmem_alloc(tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = *tmem_slot;
    pdl_wait();
    pdl_trigger();

    if (warp == 8) {
        // ------------------------------------------------------------------ TMA producer: weight tiles of 64 hidden units
        for (int it = 0; it < n_it; ++it) {
            const int stage = it % F_NST, par = (it / F_NST) & 1;
            mbar_wait(&w_empty[stage], par ^ 1);
            if (elect_one()) {
                uint8_t* t = s_w + stage * NW * F_WTILE;
                mbar_expect_tx(&w_full[stage], NW * F_WTILE);
                tma_load_3d(t, &map_w1_hi, 0, it * FC, 0, &w_full[stage]);
                if (SPLIT) tma_load_3d(t + F_WTILE, &map_w1_lo, 0, it * FC, 0, &w_full[stage]);
                tma_load_3d(t + NP * F_WTILE, &map_w2t_hi, 0, it * FC, 0, &w_full[stage]);
                if (SPLIT) tma_load_3d(t + (NP + 1) * F_WTILE, &map_w2t_lo, 0, it * FC, 0, &w_full[stage]);
            }
            __syncwarp();
        }
    } else if (warp == 9) {
        // ------------------------------------------------------------------ MMA issuer
        constexpr uint32_t idesc_s = umma_idesc_bf16(FT, FC, 0, 0);
        constexpr uint32_t idesc_acc = umma_idesc_bf16(FT, FD, 0, 1);
        mbar_wait(a_ready, 0);
        tc_fence_after();
        uint64_t ax[2], ag[2];
        for (int p = 0; p < NP; ++p) {
            ax[p] = umma_desc_sw128(smem_u32(s_a + p * F_ATILE));
            ag[p] = umma_desc_sw128(smem_u32(s_a + (NP + p) * F_ATILE));
        }
        // products of a split pair, most significant first: (hi, hi), (lo, hi), (hi, lo)
        constexpr int NPROD = SPLIT ? 3 : 1;
        constexpr int PA[3] = {0, 1, 0}, PB[3] = {0, 0, 1};
        auto issue_acc = [&](int j) {     // ACC += P_j * (W2T_j | W1_j) ; releases the weight stage of step j
            const int bfj = j & 1, stg = j % F_NST;
            mbar_wait(&ew_done[bfj], (j >> 1) & 1);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t wbase = smem_u32(s_w + stg * NW * F_WTILE + (BWD ? 0 : NP * F_WTILE));
                bool first = (j == 0);
#pragma unroll
                for (int pr = 0; pr < NPROD; ++pr) {
                    const uint32_t pa = tb + C_P + 64 * bfj + (PA[pr] ? C_LO : 0);
                    const uint64_t bw = umma_desc_sw128(wbase + PB[pr] * F_WTILE);
#pragma unroll
                    for (int k = 0; k < FC / 16; ++k) {   // 16 hidden units per MMA = 8 TMEM columns of P, 16 rows (2048 bytes) of the tile
                        umma_ts(tb + C_ACC, pa + 8 * k, bw + 128 * k, idesc_acc, first ? 0u : 1u);
                        first = false;
                    }
                }
                umma_commit(&w_empty[stg]);
                if (j == n_it - 1) umma_commit(acc_full);
            }
            __syncwarp();
        };
        for (int it = 0; it < n_it; ++it) {
            const int stage = it % F_NST, par = (it / F_NST) & 1, bf = it & 1;
            mbar_wait(&w_full[stage], par);
            tc_fence_after();
            // S / dH of buffer bf were last read by the element-wise stage of step it - 2, which finished before the accumulate
            // MMAs of that step were issued (ew_done); P of buffer bf is rewritten only after s_full[bf] of this step, a
            // commit that also covers those accumulate MMAs
            if (elect_one()) {
                const uint32_t w1 = smem_u32(s_w + stage * NW * F_WTILE);
                const uint32_t w2 = w1 + NP * F_WTILE;
#pragma unroll
                for (int pr = 0; pr < NPROD; ++pr) {
                    const uint64_t bw = umma_desc_sw128(w1 + PB[pr] * F_WTILE);
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_ss(tb + 128 * bf, ax[PA[pr]] + 2 * k, bw + 2 * k, idesc_s, (pr > 0 || k > 0) ? 1u : 0u);
                }
                if (BWD) {
#pragma unroll
                    for (int pr = 0; pr < NPROD; ++pr) {
                        const uint64_t bw = umma_desc_sw128(w2 + PB[pr] * F_WTILE);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_ss(tb + 128 * bf + C_DH, ag[PA[pr]] + 2 * k, bw + 2 * k, idesc_s, (pr > 0 || k > 0) ? 1u : 0u);
                    }
                }
                umma_commit(&s_full[bf]);
            }
            __syncwarp();
            if (it >= 1) issue_acc(it - 1);
        }
        issue_acc(n_it - 1);
    } else {
        // ------------------------------------------------------------------ element-wise stage: thread = (row = TMEM lane, column half)
        const int lane = threadIdx.x & 31;
        const int quarter = warp & 3, half = warp >> 2;
        const int r = quarter * 32 + lane;
        const uint32_t tl = tb + (static_cast<uint32_t>(quarter * 32) << 16);
        const long long row = static_cast<long long>(blockIdx.x) * FT + r;
        const bool row_ok = row < M;
        if (BWD) {
            if (half == 0) stage_row<SPLIT>(x + row * FD, row_ok, r, s_a, s_a + F_ATILE, 0, 8);
            else stage_row<SPLIT>(g + row * FD, row_ok, r, s_a + NP * F_ATILE, s_a + (NP + 1) * F_ATILE, 0, 8);
        } else {
            stage_row<SPLIT>(x + row * FD, row_ok, r, s_a, s_a + F_ATILE, 4 * half, 4 * half + 4);
        }
        for (int i = threadIdx.x; i < F; i += F_EW) s_b1[i] = b1[i];
        fence_proxy_async_smem();       // the operand tiles were written through the generic proxy; the MMA reads them through the async proxy
        mbar_arrive(a_ready);
        bar_sync_ew();                  // s_b1 complete
        for (int it = 0; it < n_it; ++it) {
            const int bf = it & 1;
            // this step's 32 fc1 biases: explicit ld.shared (through the re-aligned generic base the compiler emitted LD.E)
            float bj[32];
            {
                const uint32_t ba = smem_u32(s_b1 + it * FC + 32 * half);
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(bj[4 * q]), "=f"(bj[4 * q + 1]), "=f"(bj[4 * q + 2]), "=f"(bj[4 * q + 3]) : "r"(ba + 16 * q));
            }
            mbar_wait(&s_full[bf], (it >> 1) & 1);
            tc_fence_after();
            const uint32_t tS = tl + 128 * bf + 32 * half;
            const uint32_t tP = tl + C_P + 64 * bf + 16 * half;
            uint32_t sv[32], dv[32];
            tmem_ld32(tS, sv);
            if (BWD) tmem_ld32(tS + C_DH, dv);
            tmem_ld_wait();
            uint32_t ph[16], pl[16];
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
                float h0 = __uint_as_float(sv[i]) + bj[i], h1 = __uint_as_float(sv[i + 1]) + bj[i + 1];
                if (BWD) {
                    h0 = h0 > 0.f ? __uint_as_float(dv[i]) : 0.f;
                    h1 = h1 > 0.f ? __uint_as_float(dv[i + 1]) : 0.f;
                } else {
                    h0 = fmaxf(h0, 0.f);
                    h1 = fmaxf(h1, 0.f);
                }
                const __nv_bfloat16 a0 = __float2bfloat16_rn(h0), a1 = __float2bfloat16_rn(h1);
                ph[i / 2] = static_cast<uint32_t>(__bfloat16_as_ushort(a0)) | (static_cast<uint32_t>(__bfloat16_as_ushort(a1)) << 16);
                if (SPLIT) pl[i / 2] = pack_bf16x2(h0 - __bfloat162float(a0), h1 - __bfloat162float(a1));
            }
            tmem_st16(tP, ph);
            if (SPLIT) tmem_st16(tP + C_LO, pl);
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(&ew_done[bf]);
        }
        mbar_wait(acc_full, 0);
        tc_fence_after();
        uint32_t acc[32];
        tmem_ld32(tl + C_ACC + 32 * half, acc);
        tmem_ld_wait();
        if (row_ok) {
            const float* res = (BWD ? g : x) + row * FD + 32 * half;      // residual: x (forward), g (dgrad)
            float* o = out + row * FD + 32 * half;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 ra = reinterpret_cast<const float4*>(res)[2 * q], rb = reinterpret_cast<const float4*>(res)[2 * q + 1];
                const float rr[8] = {ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, rb.z, rb.w};
                uint32_t w[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float v = __uint_as_float(acc[8 * q + i]) + rr[i];
                    if (!BWD) v += __ldg(b2 + 32 * half + 8 * q + i);
                    w[i] = __float_as_uint(v);
                }
                st_global_256(o + 8 * q, w);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 9) {
        tc_fence_after();
        tmem_dealloc(tb, TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFnF)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFnF encode_fn_f() {
    static EncodeTiledFnF fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFnF>(p);
    }
    return fn;
}

// bf16 weight [F][64] -> 3-D map {64, F, 1}, box {64, 64, 1}, 128-byte swizzle
static int make_wmap(CUtensorMap* map, const void* ptr, int F) {
    EncodeTiledFnF fn = encode_fn_f();
    TVS_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled not available from the driver");
    cuuint64_t dims[3] = {FD, static_cast<cuuint64_t>(F), 1};
    cuuint64_t strides[2] = {FD * 2, static_cast<cuuint64_t>(F) * FD * 2};
    cuuint32_t box[3] = {FD, FC, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    TVS_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (ffn weight) failed with %d (F=%d)", (int)r, F);
    return 0;
}

template <bool BWD, bool SPLIT>
static int launch_ffn(const void* w1_hi, const void* w1_lo, const void* w2t_hi, const void* w2t_lo, const float* x, const float* g, const float* b1,
                      const float* b2, long long M, int F, float* out, cudaStream_t st) {
    using L = FfnSmem<BWD, SPLIT>;
    CUtensorMap m1h, m1l, m2h, m2l;
    if (int rc = make_wmap(&m1h, w1_hi, F)) return rc;
    if (int rc = make_wmap(&m2h, w2t_hi, F)) return rc;
    if (int rc = make_wmap(&m1l, SPLIT ? w1_lo : w1_hi, F)) return rc;
    if (int rc = make_wmap(&m2l, SPLIT ? w2t_lo : w2t_hi, F)) return rc;
    auto kern = ffn_tc_kernel<BWD, SPLIT>;
    static bool attr_set = false;
    if (!attr_set) {
        TVS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
        attr_set = true;
    }
    const unsigned grid = static_cast<unsigned>((M + FT - 1) / FT);
    TVS_CUDA(launch_pdl(kern, dim3(grid), dim3(F_THREADS), L::TOTAL, st, 1, m1h, m1l, m2h, m2l, x, g, b1, b2, M, F, out));
    return check_launch(BWD ? "ffn_tc_kernel<bwd>" : "ffn_tc_kernel<fwd>");
}

}  // namespace tvs

extern "C" __attribute__((visibility("default"))) int tvs_ffn64_fwd(const float* x, const void* w1_hi, const void* w1_lo, const void* w2t_hi, const void* w2t_lo,
                                                                    const float* b1, const float* b2, int64_t M, int32_t D, int32_t F, float* out, void* stream) {
    using namespace tvs;
    TVS_REQUIRE(x && w1_hi && w2t_hi && b1 && b2 && out, "tvs_ffn64_fwd: null pointer");
    TVS_REQUIRE((w1_lo == nullptr) == (w2t_lo == nullptr), "tvs_ffn64_fwd: give both weight tails or neither");
    TVS_REQUIRE(D == FD, "tvs_ffn64_fwd: model width must be %d (got %d)", FD, D);
    TVS_REQUIRE(M > 0 && F > 0 && F % FC == 0 && F <= F_MAXF, "tvs_ffn64_fwd: F=%d must be a multiple of %d and <= %d", F, FC, F_MAXF);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (w1_lo) return launch_ffn<false, true>(w1_hi, w1_lo, w2t_hi, w2t_lo, x, nullptr, b1, b2, M, F, out, st);
    return launch_ffn<false, false>(w1_hi, nullptr, w2t_hi, nullptr, x, nullptr, b1, b2, M, F, out, st);
}

extern "C" __attribute__((visibility("default"))) int tvs_ffn64_bwd(const float* x, const float* g, const void* w1_hi, const void* w1_lo, const void* w2t_hi,
                                                                    const void* w2t_lo, const float* b1, int64_t M, int32_t D, int32_t F, float* dx, void* stream) {
    using namespace tvs;
    TVS_REQUIRE(x && g && w1_hi && w2t_hi && b1 && dx, "tvs_ffn64_bwd: null pointer");
    TVS_REQUIRE((w1_lo == nullptr) == (w2t_lo == nullptr), "tvs_ffn64_bwd: give both weight tails or neither");
    TVS_REQUIRE(D == FD, "tvs_ffn64_bwd: model width must be %d (got %d)", FD, D);
    TVS_REQUIRE(M > 0 && F > 0 && F % FC == 0 && F <= F_MAXF, "tvs_ffn64_bwd: F=%d must be a multiple of %d and <= %d", F, FC, F_MAXF);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (w1_lo) return launch_ffn<true, true>(w1_hi, w1_lo, w2t_hi, w2t_lo, x, g, b1, nullptr, M, F, dx, st);
    return launch_ffn<true, false>(w1_hi, nullptr, w2t_hi, nullptr, x, g, b1, nullptr, M, F, dx, st);
}
