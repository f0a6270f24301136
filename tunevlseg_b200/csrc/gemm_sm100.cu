// tcgen05 / TMEM / TMA GEMM for sm_100a:  C[M,N] = epilogue(A[M,K] * W[N,K]^T), bf16 in, fp32 accumulate.
//
// One persistent CTA per SM, 256 threads, warp-specialised:
//   warp 0 (one lane)  TMA producer: A tile 128x64 and W tile BNx64 per k-block into a STAGES-deep smem ring
//                      (128-byte swizzle), completion on `full[stage]` mbarriers;
//   warp 1 (one lane)  MMA issuer: 4 x tcgen05.mma (128 x BN x 16, cta_group::1) per k-block, accumulator in
//                      TMEM (two buffers of BN columns so the epilogue of tile i overlaps the MMAs of tile i+1),
//                      tcgen05.commit releases the smem slot (`empty[stage]`) and publishes `tmem_full[buf]`;
//   warp 2             TMEM allocate / free;
//   warps 4-7          epilogue: tcgen05.ld 32 lanes x 32 columns per warp, fused bias / activation /
//                      activation-gradient / residual, stores fp32 and/or bf16 rows.
// Tail handling: TMA zero-fills rows >= M, >= N and columns >= K; the epilogue predicates its stores.
//
// Replaces the nn.Linear calls of HF CLIPSeg (modeling_clipseg.py:297-300, :345-353) and their dgrad.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "tvs_b200.h"

namespace tvs {

constexpr int BM = 128;
constexpr int BK_BYTES = 128;  // one k-block = one swizzle-128B row: 64 bf16 or 32 fp32 (tf32) elements
constexpr int MMAS_PER_KB = 4; // 4 x (16 bf16 | 8 tf32) = 32 bytes of K per tcgen05.mma
constexpr int GEMM_THREADS = 384;   // warps 0-3: TMA / MMA / TMEM alloc / spare; warps 4-11: two epilogue groups of four warps

struct GemmEpilogue {
    const float* bias;
    const float* residual;
    long long ldr;
    float* out_f32;
    long long ldo32;
    __nv_bfloat16* out_bf16;
    long long ldo16;
    __nv_bfloat16* pre_bf16;
    long long ldpre;
    const __nv_bfloat16* aux_bf16;
    long long ldaux;
    int act;
    int vec_ok;      // every leading dimension % 4 == 0 and every pointer 16-byte aligned -> 128-bit staged epilogue
    int vec256_ok;   // 32-byte aligned rows for every operand -> direct 256-bit epilogue (the default)
    int staged;      // force the shared-memory staged (coalesced) epilogue
    int round_out;   // out_f32 is rounded to nearest tf32 (cvt.rna): it only feeds further kind::tf32 MMAs, which truncate
    int fmt_a, fmt_b;   // kind::f16 operand formats of A and W in the instruction descriptor: 0 = IEEE fp16, 1 = bf16 (ignored by kind::tf32)
    int out16_f16;      // the 16-bit activation output (out_bf16) is written as IEEE fp16 (11-bit significand) instead of bf16
    // deep-prompt overwrite fused into the residual epilogue (EPI_RES_F32): output rows whose position inside their sample,
    // row % ovr_S, lies in [ovr_row0, ovr_row0 + ovr_n) receive ovr_ctx[(row / ovr_S) * ovr_bs + (row % ovr_S - ovr_row0) * N + col]
    // instead of the GEMM result (base_multimodal_clipseg.py:394-398: the prompt rows are re-written after every block < depth)
    const float* ovr_ctx;
    long long ovr_bs;
    int ovr_S, ovr_row0, ovr_n;
    // implicit-GEMM 3x3 convolution (conv_wp != 0): A is a zero-bordered channels-last image [B, conv_hp, conv_wp, C] seen as
    // a [B*hp*wp, C] matrix; k-block kb belongs to tap kb / conv_kpt and loads the A rows SHIFTED by that tap's offset
    // in the flattened image (the physical zero border makes every interior pixel's 9 taps correct); the epilogue maps
    // the padded row to the unpadded output row and skips border rows.
    int conv_wp, conv_hp, conv_kpt;
    // stream-K (specialised pair kernels, static schedule): the (tile, k-block) iteration space is cut into one contiguous
    // range per cluster; a tile cut in two is finished by the cluster that holds its FIRST k-blocks (it reaches the tile last),
    // the other cluster computes the later k-blocks first and parks its raw fp32 accumulator in sk_ws (launch_gemm_inst)
    int pre_dgelu;   // TVS_GEMM_PRE_DGELU: pre_bf16 receives QuickGELU'(v) instead of v
    float* sk_ws;
    int* sk_flags;
    int sk_q;        // iterations (k-blocks) per cluster; 0 = off (round-robin whole tiles)
};

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t addr = smem_u32(bar);
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(addr),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            smem_u32(smem_dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}
// same tile delivered to the same shared-memory offset of every CTA in `mask`; each destination's mbarrier (same offset) gets the bytes
__device__ __forceinline__ void tma_load_2d_mc(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%2, %3}], [%4], %5;" ::"r"(
            smem_u32(smem_dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)), "h"(mask)
        : "memory");
}
// One elected lane of a CONVERGED warp.  The producer / MMA warps keep all 32 lanes in the loop and predicate only the
// issuing instructions: the loop state then stays warp-uniform (uniform registers feed UTMALDG / UTCHMMA directly);
// running the loop in a single divergent lane costs an R2UR round trip per operand and made the MMA issue loop, not
// the tensor pipe, the bottleneck (~70 % utilisation).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// UMMA shared-memory descriptor, K-major operand, 128-byte swizzle (cute::UMMA::SmemDescriptor layout):
//   [0,14) start>>4, [16,30) LBO>>4 (unused for swizzled K-major), [32,46) SBO>>4 = 1024B (8 rows x 128B),
//   [46,48) version = 1 (sm_100), [61,64) layout = 2 (SWIZZLE_128B).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
    d |= static_cast<uint64_t>(1) << 16;
    d |= static_cast<uint64_t>(1024 >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}
// Instruction descriptor for kind::f16 (cute::UMMA::InstrDescriptor): D=f32, A=B=bf16, both K-major.
// fmt: 1 = BF16 (kind::f16), 2 = TF32 (kind::tf32, operands are fp32 in shared memory, top 19 bits used)
__host__ __device__ constexpr uint32_t umma_idesc(int m, int n, uint32_t fmt_a, uint32_t fmt_b) {
    return (1u << 4) | (fmt_a << 7) | (fmt_b << 10) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// arrive on the barrier at this offset in every CTA of `mask` once the MMAs issued so far have completed
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
                 "h"(mask)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// epilogue.  TMEM gives every thread one ROW of the accumulator (32 consecutive columns per tcgen05.ld); storing
// rows straight to global memory makes each warp store touch 32 different lines.  So each epilogue warp transposes
// its 32x32 chunk through a private shared-memory stage (row stride 36 floats: conflict-free for the 128-bit row
// writes and for the 128-bit reads below) and then works on COALESCED positions: a quarter-warp covers one 128-byte
// row segment, a warp instruction covers 4 rows, 8 passes cover the chunk.  All fused operand reads (bias, residual,
// saved pre-activation) and all output writes use the same coalesced mapping.
// ---------------------------------------------------------------------------------------------
constexpr int EPI_LD = 36;                       // floats per staged row
constexpr int EPI_WARP_FLOATS = 32 * EPI_LD;     // 4608 bytes per epilogue warp

__device__ __forceinline__ float rn_tf32f(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
__device__ __forceinline__ bool is_dact(int act) { return act == TVS_ACT_DQGELU || act == TVS_ACT_DRELU || act == TVS_ACT_MULAUX; }

__device__ __forceinline__ float epi_act(const GemmEpilogue& ep, float v, float aux) {
    if (ep.act == TVS_ACT_QGELU) return quick_gelu(v);
    if (ep.act == TVS_ACT_RELU) return fmaxf(v, 0.0f);
    if (ep.act == TVS_ACT_DQGELU) return v * quick_gelu_grad(aux);
    if (ep.act == TVS_ACT_DRELU) return aux > 0.0f ? v : 0.0f;
    if (ep.act == TVS_ACT_MULAUX) return v * aux;
    return v;
}

// 4 consecutive columns [col, col+4) of one row, all vector accesses 16-byte (f32) / 8-byte (bf16) aligned
__device__ __forceinline__ void epilogue_vec4(const GemmEpilogue& ep, float4 v, float4 bias, long long row, int col) {
    v.x += bias.x; v.y += bias.y; v.z += bias.z; v.w += bias.w;
    if (ep.pre_bf16) {
        const float4 s4 = ep.pre_dgelu ? make_float4(quick_gelu_grad(v.x), quick_gelu_grad(v.y), quick_gelu_grad(v.z), quick_gelu_grad(v.w)) : v;
        *reinterpret_cast<uint2*>(ep.pre_bf16 + row * ep.ldpre + col) = make_uint2(pack_bf16x2(s4.x, s4.y), pack_bf16x2(s4.z, s4.w));
    }
    if (ep.act != TVS_ACT_NONE) {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        if (is_dact(ep.act)) {
            const uint2 q = __ldg(reinterpret_cast<const uint2*>(ep.aux_bf16 + row * ep.ldaux + col));
            const float2 lo = unpack_bf16x2(q.x), hi = unpack_bf16x2(q.y);
            a = make_float4(lo.x, lo.y, hi.x, hi.y);
        }
        v.x = epi_act(ep, v.x, a.x); v.y = epi_act(ep, v.y, a.y); v.z = epi_act(ep, v.z, a.z); v.w = epi_act(ep, v.w, a.w);
    }
    if (ep.residual) {
        const float4 r = *reinterpret_cast<const float4*>(ep.residual + row * ep.ldr + col);
        v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
    }
    if (ep.act == TVS_ACT_RES_RELU) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
    if (ep.out_f32)
        *reinterpret_cast<float4*>(ep.out_f32 + row * ep.ldo32 + col) =
            ep.round_out ? make_float4(rn_tf32f(v.x), rn_tf32f(v.y), rn_tf32f(v.z), rn_tf32f(v.w)) : v;
    if (ep.out_bf16)
        *reinterpret_cast<uint2*>(ep.out_bf16 + row * ep.ldo16 + col) =
            ep.out16_f16 ? make_uint2(pack_f16x2(v.x, v.y), pack_f16x2(v.z, v.w)) : make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
}

// scalar, fully predicated (ragged N such as the 25-tap additive map, or unaligned leading dimensions)
__device__ __forceinline__ void epilogue_scalar(const GemmEpilogue& ep, float x, long long row, int c) {
    if (ep.bias) x += __ldg(ep.bias + c);
    if (ep.pre_bf16) ep.pre_bf16[row * ep.ldpre + c] = __float2bfloat16(ep.pre_dgelu ? quick_gelu_grad(x) : x);
    if (ep.act != TVS_ACT_NONE) x = epi_act(ep, x, is_dact(ep.act) ? __bfloat162float(ep.aux_bf16[row * ep.ldaux + c]) : 0.f);
    if (ep.residual) x += ep.residual[row * ep.ldr + c];
    if (ep.act == TVS_ACT_RES_RELU) x = fmaxf(x, 0.f);
    if (ep.out_f32) ep.out_f32[row * ep.ldo32 + c] = ep.round_out ? rn_tf32f(x) : x;
    if (ep.out_bf16) {
        if (ep.out16_f16) reinterpret_cast<__half*>(ep.out_bf16)[row * ep.ldo16 + c] = __float2half_rn(x);
        else ep.out_bf16[row * ep.ldo16 + c] = __float2bfloat16(x);
    }
}

// ---- direct epilogue: the thread keeps its TMEM row and moves 32 bytes per instruction (LDG.256 / STG.256, sm_100) ----
// With cta_group::1 and 128-row tiles the tensor core already consumes most of the 128 B/clk shared-memory
// bandwidth for its operands, so staging the accumulator through shared memory (below) slows the MMA down more than
// coalescing helps (measured: 2x slower).  Full 32-byte sectors per thread keep the L2 write path efficient enough.
__device__ __forceinline__ void ld256(const void* p, uint32_t (&r)[8]) {
    asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "l"(p));
}
__device__ __forceinline__ void ld256_cg(const void* p, uint32_t (&r)[8]) {      // L2 only: data another SM has just written
    asm volatile("ld.global.cg.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "l"(p)
                 : "memory");
}
__device__ __forceinline__ void st256(void* p, const uint32_t (&r)[8]) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]),
                 "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void st_bf16_row32(__nv_bfloat16* p, const float (&v)[32]) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        uint32_t w[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) w[i] = pack_bf16x2(v[16 * h + 2 * i], v[16 * h + 2 * i + 1]);
        st256(p + 16 * h, w);
    }
}
__device__ __forceinline__ void st_f16_row32(__nv_bfloat16* p, const float (&v)[32]) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        uint32_t w[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) w[i] = pack_f16x2(v[16 * h + 2 * i], v[16 * h + 2 * i + 1]);
        st256(p + 16 * h, w);
    }
}

// The global operands a chunk's epilogue needs (saved pre-activation, residual).  They are fetched ONE CHUNK AHEAD so
// that their ~1 us latency overlaps the previous chunk's math and stores instead of stalling every chunk.
struct EpiExtras {
    uint32_t aux[16];
    uint32_t res[32];
};
__device__ __forceinline__ void load_extras(const GemmEpilogue& ep, long long row, int col0, EpiExtras& e) {
    if (is_dact(ep.act)) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            uint32_t a[8];
            ld256(ep.aux_bf16 + row * ep.ldaux + col0 + 16 * h, a);
#pragma unroll
            for (int i = 0; i < 8; ++i) e.aux[8 * h + i] = a[i];
        }
    }
    if (ep.residual) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            uint32_t r[8];
            ld256(ep.residual + row * ep.ldr + col0 + 8 * q, r);
#pragma unroll
            for (int i = 0; i < 8; ++i) e.res[8 * q + i] = r[i];
        }
    }
}

// one row (this thread's TMEM lane), 32 consecutive columns starting at col0; requires ep.vec256_ok and col0 + 32 <= N
__device__ __forceinline__ void epilogue_direct(const GemmEpilogue& ep, const uint32_t (&acc)[32], long long row, int col0, const EpiExtras& ex) {
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(acc[i]);
    if (ep.bias) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            uint32_t b[8];
            ld256(ep.bias + col0 + 8 * q, b);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[8 * q + i] += __uint_as_float(b[i]);
        }
    }
    if (ep.pre_bf16) {
        if (ep.pre_dgelu) {
            float gd[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) gd[i] = quick_gelu_grad(v[i]);
            st_bf16_row32(ep.pre_bf16 + row * ep.ldpre + col0, gd);
        } else {
            st_bf16_row32(ep.pre_bf16 + row * ep.ldpre + col0, v);
        }
    }
    if (ep.act == TVS_ACT_QGELU) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = quick_gelu(v[i]);
    } else if (ep.act == TVS_ACT_RELU) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.0f);
    } else if (is_dact(ep.act)) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const float2 u = unpack_bf16x2(ex.aux[i]);
            v[2 * i] = epi_act(ep, v[2 * i], u.x);
            v[2 * i + 1] = epi_act(ep, v[2 * i + 1], u.y);
        }
    }
    if (ep.residual) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] += __uint_as_float(ex.res[i]);
    }
    if (ep.act == TVS_ACT_RES_RELU) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.0f);
    }
    if (ep.out_f32) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            uint32_t o[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = __float_as_uint(ep.round_out ? rn_tf32f(v[8 * q + i]) : v[8 * q + i]);
            st256(ep.out_f32 + row * ep.ldo32 + col0 + 8 * q, o);
        }
    }
    if (ep.out_bf16) {
        if (ep.out16_f16) st_f16_row32(ep.out_bf16 + row * ep.ldo16 + col0, v);
        else st_bf16_row32(ep.out_bf16 + row * ep.ldo16 + col0, v);
    }
}

// ---- specialised epilogues (template parameter EPI of the kernel) -------------------------------------------------------
// The generic epilogue above decides everything from runtime flags; on a kernel of ~8 000 SASS instructions the compiler's
// unswitching of those flags is fragile (measured in round 2: adding two cold-path branches re-rolled the dGELU path into
// four branches PER ELEMENT, 93 -> 167 us, and cost the plain bf16 epilogue 8 %).  The shapes that carry the step get
// straight-line epilogues with no flag tests at all; everything else (ragged N, convolutions, tf32, rare flag sets) keeps the
// generic code.  All of them require: 32-byte aligned rows (vec256_ok), N % 32 == 0.
enum { EPI_GENERIC = 0, EPI_OUT_BF16 = 1 /* (+bias) -> bf16 */, EPI_RES_F32 = 2 /* (+bias) + residual -> f32 */,
       EPI_FC1 = 3 /* + bias -> bf16 pre-activation; QuickGELU -> fp16 */, EPI_DQGELU = 4 /* * QuickGELU'(aux bf16) -> bf16 */ };

template <int EPI>
__device__ __forceinline__ void epilogue_spec(const GemmEpilogue& ep, const uint32_t (&acc)[32], long long row, int col0, const uint32_t (&ext)[32],
                                              uint32_t s_bias /* shared-window address of this chunk's 32 bias values, or 0 */,
                                              bool ovr = false /* EPI_RES_F32: ext holds the prompt row that replaces the result */) {
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(acc[i]);
    if (EPI != EPI_DQGELU && s_bias != 0) {              // kernel-uniform; 8 broadcast LDS.128 (explicit ld.shared: through the
#pragma unroll                                           // re-aligned generic base the compiler emitted LD.E, round 2)
        for (int q = 0; q < 8; ++q) {
            float4 b;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "r"(s_bias + 16 * q));
            v[4 * q] += b.x; v[4 * q + 1] += b.y; v[4 * q + 2] += b.z; v[4 * q + 3] += b.w;
        }
    }
    if (EPI == EPI_OUT_BF16) {
        st_bf16_row32(ep.out_bf16 + row * ep.ldo16 + col0, v);
    } else if (EPI == EPI_RES_F32) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            uint32_t o[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = ovr ? ext[8 * q + i] : __float_as_uint(v[8 * q + i] + __uint_as_float(ext[8 * q + i]));
            st256(ep.out_f32 + row * ep.ldo32 + col0 + 8 * q, o);
        }
    } else if (EPI == EPI_FC1) {
        if (ep.pre_dgelu) {      // kernel-uniform: save QuickGELU'(u) instead of u - the sigmoid is in registers here, and the dgrad
            float gd[32];        // through the activation becomes one multiply per element (TVS_ACT_MULAUX) instead of a MUFU + 8 ops
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const float t = 1.702f * v[i];
                const float sg = sigmoid_tanh(t);
                gd[i] = sg * fmaf(t, 1.0f - sg, 1.0f);
                v[i] *= sg;
            }
            st_bf16_row32(ep.pre_bf16 + row * ep.ldpre + col0, gd);
        } else {
            st_bf16_row32(ep.pre_bf16 + row * ep.ldpre + col0, v);
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = quick_gelu(v[i]);
        }
        st_f16_row32(ep.out_bf16 + row * ep.ldo16 + col0, v);
    } else if (EPI == EPI_DQGELU) {
        if (ep.act == TVS_ACT_MULAUX) {      // kernel-uniform: aux already holds the derivative
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const float2 u = unpack_bf16x2(ext[i]);
                v[2 * i] *= u.x;
                v[2 * i + 1] *= u.y;
            }
        } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const float2 u = unpack_bf16x2(ext[i]);
                v[2 * i] *= quick_gelu_grad(u.x);
                v[2 * i + 1] *= quick_gelu_grad(u.y);
            }
        }
        st_bf16_row32(ep.out_bf16 + row * ep.ldo16 + col0, v);
    }
}

// the global operands of one chunk of a specialised epilogue (residual row segment / saved pre-activation), issued early
template <int EPI>
__device__ __forceinline__ void spec_load_ext(const GemmEpilogue& ep, long long row, int col0, uint32_t (&ext)[32], const float* ovr_row = nullptr) {
    if (EPI == EPI_DQGELU) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            uint32_t a8[8];
            ld256(ep.aux_bf16 + row * ep.ldaux + col0 + 16 * h, a8);
#pragma unroll
            for (int i = 0; i < 8; ++i) ext[8 * h + i] = a8[i];
        }
    } else if (EPI == EPI_RES_F32) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            uint32_t r8[8];
            ld256(ovr_row ? ovr_row + col0 + 8 * q : ep.residual + row * ep.ldr + col0 + 8 * q, r8);
#pragma unroll
            for (int i = 0; i < 8; ++i) ext[8 * q + i] = r8[i];
        }
    }
}

// one 32x32 chunk: acc = this thread's row (lane) of the chunk; stage = this warp's private staging area
__device__ __forceinline__ void epilogue_chunk(const GemmEpilogue& ep, const uint32_t (&acc)[32], float* stage, long long row0, int col0,
                                               int M, int N, int lane) {
    float4* srow = reinterpret_cast<float4*>(stage + lane * EPI_LD);
#pragma unroll
    for (int i = 0; i < 8; ++i)
        srow[i] = make_float4(__uint_as_float(acc[4 * i]), __uint_as_float(acc[4 * i + 1]), __uint_as_float(acc[4 * i + 2]), __uint_as_float(acc[4 * i + 3]));
    __syncwarp();
    const int c4 = (lane & 7) * 4, rsub = lane >> 3;
    const int col = col0 + c4;
    if (ep.vec_ok && col + 4 <= N) {
        const float4 bias = ep.bias ? __ldg(reinterpret_cast<const float4*>(ep.bias + col)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int p = 0; p < 8; ++p) {
            const int rr = p * 4 + rsub;
            const long long row = row0 + rr;
            if (row < M) epilogue_vec4(ep, *reinterpret_cast<const float4*>(stage + rr * EPI_LD + c4), bias, row, col);
        }
    } else if (col < N) {
#pragma unroll 1
        for (int p = 0; p < 8; ++p) {
            const int rr = p * 4 + rsub;
            const long long row = row0 + rr;
            if (row < M)
                for (int j = 0; j < 4; ++j)
                    if (col + j < N) epilogue_scalar(ep, stage[rr * EPI_LD + c4 + j], row, col + j);
        }
    }
    __syncwarp();
}

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
constexpr int SCHED_D = 4;      // depth of the dynamic tile ring

template <int BN, int STAGES, int CL>
struct GemmSmem {
    static constexpr int A_BYTES = BM * BK_BYTES;
    static constexpr int B_BYTES = (BN / CL) * BK_BYTES;        // cta_group::2: each CTA of the pair holds half of the W tile
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
    static constexpr int SCHED_OFFSET = BAR_OFFSET + (2 * STAGES + 4) * 8 + 16;   // tile ring: full[4], empty[4] barriers, 4 tile indices
    static constexpr int EPI_OFFSET = SCHED_OFFSET + 2 * SCHED_D * 8 + SCHED_D * 4;   // 16-byte aligned
    static constexpr int TOTAL = EPI_OFFSET + 4 * EPI_WARP_FLOATS * 4 + 1024;     // + alignment slack
};

// CL = 2: the kernel runs as CTA PAIRS (cta_group::2).  The pair computes a 256 x BN tile: each CTA loads its own 128
// rows of A and HALF of the W tile, the leader CTA issues one tcgen05.mma.cta_group::2 (M = 256) that reads both CTAs'
// shared memory, and each CTA's TMEM receives its own 128 x BN accumulator.  An SM can take in roughly 64-70 B/clk from
// L2 (measured: the 1-CTA 128x256 tile needs 96 B/clk at full MMA rate and saturates at ~70 % tensor utilisation;
// multicasting W between two independent CTAs did not help because every SM still RECEIVES the whole tile); the pair
// needs 64 B/clk per SM and also halves the shared-memory operand reads per SM.
//   full[s]     lives in the leader: its own arrive.expect_tx(2 x stage bytes) + the peer's remote arrive; both CTAs' TMA
//               loads (cta_group::2) complete_tx on it
//   empty[s], tmem_full[a]  local to each CTA, signalled by the leader's multicast tcgen05.commit
//   tmem_empty[a]  lives in the leader, 4 epilogue warps x 2 CTAs arrive on it
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    // default semantics (release at CTA scope), as cutlass::arch::ClusterBarrier::arrive(cta_id): a cluster-scope release costs a full
    // memory barrier (~1 us) on the producer's critical path (measured: pair mode ran at 33 % tensor utilisation with it)
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_2d_cg2(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint32_t bar_cluster_addr) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            smem_u32(smem_dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(bar_cluster_addr)
        : "memory");
}
__device__ __forceinline__ void umma_bf16_cg2(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_tf32_cg2(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit_cg2(uint64_t* bar) {   // arrives on this offset in BOTH CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
                 "h"(static_cast<uint16_t>(3))
                 : "memory");
}


// ---------------------------------------------------------------------------------------------
// dynamic tile scheduler
// ---------------------------------------------------------------------------------------------
// The kernel is persistent with one CTA (pair) per SM.  With a static round-robin tile list a CTA that STARTS late
// finishes late by the same amount, and the whole kernel with it: inside the training step the text tower runs on a
// second stream, and whenever one of its ~250 small kernels holds a few SMs at the moment a vision-tower GEMM is launched
// the GEMM's tail grows by that kernel's duration (measured: the 15648x3072x768 GEMM 89 -> 103 us, QKV 52 -> 71 us).
// With a shared counter a late CTA simply takes fewer tiles.  [Result: no gain, see launch_gemm_cl - the static list stays
// the default and this path is an opt-in experiment.]  Warp 3 of the (leader) CTA draws tile indices from a
// global counter and publishes them through a 4-deep shared-memory ring (in pair mode also into the peer's ring, remote
// store + cluster-scope release arrive); producer, MMA issuer and the eight epilogue warps each read every entry and
// release it on the leader's `empty` barrier.  -1 ends the loops.  The cluster that draws the last sentinel resets the
// counter pair, so a slot is clean again when its kernel has finished (graph replays reuse the slot of their node).
__device__ __forceinline__ void mbar_wait_acq_cluster(uint64_t* bar, uint32_t parity) {
    uint32_t addr = smem_u32(bar);
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP_C:\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE_C;\n\t"
        "bra WAIT_LOOP_C;\n\t"
        "WAIT_DONE_C:\n\t"
        "}" ::"r"(addr),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void st_cluster_u32(uint32_t cluster_addr, uint32_t v) {
    asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(cluster_addr), "r"(v) : "memory");
}
__device__ __forceinline__ void mbar_arrive_release_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}

template <int CL>
struct TileIter {
    // static part
    int cur, step, num;
    // dynamic part
    bool dyn;
    volatile int* s_tile;
    uint64_t* sf;
    uint64_t* se;
    uint32_t slot, par, cta_rank;

    // k-block range and role of the tile next() returned: role 0 = whole tile, 1 = owner of a cut tile (k-blocks [0, kb1); the
    // cluster after this one parked [kb1, num_kb)), 2 = helper (k-blocks [kb0, num_kb): park the accumulator in slot `cluster`)
    int kb0, kb1, role;
    int sk_pos, sk_end, num_kb;      // stream-K: this cluster's range of the linearised (tile, k-block) space; sk_end < 0 = off

    __device__ __forceinline__ int next() {      // called by every lane of a converged warp
        if (!dyn && sk_end >= 0) {
            if (sk_pos >= sk_end) return -1;
            const int t = sk_pos / num_kb;
            kb0 = sk_pos - t * num_kb;
            kb1 = min(num_kb, kb0 + (sk_end - sk_pos));
            sk_pos += kb1 - kb0;
            role = (kb0 == 0 && kb1 == num_kb) ? 0 : (kb0 == 0 ? 1 : 2);
            return t;
        }
        kb0 = 0; kb1 = num_kb; role = 0;
        if (!dyn) {
            const int t = cur;
            cur += step;
            return t < num ? t : -1;
        }
        if (CL == 2) {
            mbar_wait_acq_cluster(&sf[slot], par);
            asm volatile("fence.acq_rel.cluster;" ::: "memory");
        } else {
            mbar_wait(&sf[slot], par);
        }
        const int t = s_tile[slot];
        __syncwarp();                                                  // every lane has read the entry
        if ((threadIdx.x & 31) == 0) {
            if (CL == 2 && cta_rank != 0) mbar_arrive_cluster(mapa_u32(smem_u32(&se[slot]), 0));
            else mbar_arrive(&se[slot]);
        }
        if (++slot == SCHED_D) { slot = 0; par ^= 1; }
        return t;
    }
};

template <int BN, int STAGES, bool TF32, int CL, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w, int M,
                         int N, int K, GemmEpilogue ep, int* __restrict__ sched) {
    using L = GemmSmem<BN, STAGES, CL>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + STAGES * L::A_BYTES;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full = empty_bar + STAGES;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    uint64_t* sched_full = reinterpret_cast<uint64_t*>(smem + L::SCHED_OFFSET);
    uint64_t* sched_empty = sched_full + SCHED_D;
    volatile int* sched_tile = reinterpret_cast<volatile int*>(sched_empty + SCHED_D);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int num_m = (M + BM - 1) / BM;
    const int num_n = (N + BN - 1) / BN;
    const uint32_t cta_rank = CL == 2 ? cluster_ctarank() : 0;
    const int first_tile = blockIdx.x / CL, tile_step = gridDim.x / CL;   // cluster id / number of clusters
    const int num_tiles = ((num_m + CL - 1) / CL) * num_n;                 // tiles of CL stacked 128-row blocks
    constexpr int BK = TF32 ? BK_BYTES / 4 : BK_BYTES / 2;   // elements of K per k-block
    const int num_kb = (K + BK - 1) / BK;
    constexpr uint32_t TMEM_COLS = 2 * BN;  // power of two >= 32 for BN in {64,128,256}

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_w);
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], CL);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(&tmem_full[0], 1);
        mbar_init(&tmem_full[1], 1);
        mbar_init(&tmem_empty[0], 8 * CL);
        mbar_init(&tmem_empty[1], 8 * CL);
#pragma unroll
        for (int d = 0; d < SCHED_D; ++d) {
            mbar_init(&sched_full[d], 1);
            mbar_init(&sched_empty[d], CL == 2 ? 19 : 10);     // producer + MMA issuer + 8 epilogue warps (+ the peer's producer and 8 epilogue warps)
        }
        fence_barrier_init();
    }
    if (warp == 2) {
        if (CL == 2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    if (CL == 2) cluster_sync_all();   // the peer's barriers must be initialised before anything is multicast to them
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();        // everything above is private set-up; from here on the kernel reads what its predecessor wrote
    pdl_trigger();

    TileIter<CL> tiles;
    tiles.cur = first_tile; tiles.step = tile_step; tiles.num = num_tiles;
    tiles.dyn = sched != nullptr; tiles.s_tile = sched_tile; tiles.sf = sched_full; tiles.se = sched_empty;
    tiles.slot = 0; tiles.par = 0; tiles.cta_rank = cta_rank;
    tiles.num_kb = num_kb; tiles.kb0 = 0; tiles.kb1 = num_kb; tiles.role = 0; tiles.sk_pos = 0; tiles.sk_end = -1;
    if (EPI != EPI_GENERIC && sched == nullptr && ep.sk_q > 0) {
        const long long total = static_cast<long long>(num_tiles) * num_kb, lo = static_cast<long long>(first_tile) * ep.sk_q;
        tiles.sk_pos = static_cast<int>(lo < total ? lo : total);
        tiles.sk_end = static_cast<int>(lo + ep.sk_q < total ? lo + ep.sk_q : total);
    }

    if (warp == 3) {
        if (sched != nullptr && cta_rank == 0) {
            // ------------------------------------------------------------------ tile scheduler
            uint32_t slot = 0, par = 0;
            for (;;) {
                mbar_wait(&sched_empty[slot], par ^ 1);
                int t = 0;
                if (lane == 0) {
                    t = atomicAdd(&sched[0], 1);
                    if (t >= num_tiles) t = -1;
                    sched_tile[slot] = t;
                    if (CL == 2) {
                        st_cluster_u32(mapa_u32(smem_u32(const_cast<int*>(&sched_tile[slot])), 1), static_cast<uint32_t>(t));
                        asm volatile("fence.acq_rel.cluster;" ::: "memory");
                        mbar_arrive_release_cluster(mapa_u32(smem_u32(&sched_full[slot]), 1));
                    }
                    mbar_arrive(&sched_full[slot]);
                }
                t = __shfl_sync(0xffffffffu, t, 0);
                if (t < 0) break;
                if (++slot == SCHED_D) { slot = 0; par ^= 1; }
            }
            // every cluster draws exactly one sentinel; the last one to do so puts the counter pair back to zero
            if (lane == 0) {
                const int clusters = gridDim.x / CL;
                if (atomicAdd(&sched[1], 1) == clusters - 1) {
                    sched[1] = 0;
                    __threadfence();
                    sched[0] = 0;
                }
            }
        }
    } else if (warp == 0) {
        {
            uint32_t stage = 0, phase = 0;
            for (int tile = tiles.next(); tile >= 0; tile = tiles.next()) {
                const int m_blk = (tile / num_n) * CL + cta_rank, n_blk = tile % num_n;
                for (int kb = tiles.kb0; kb < tiles.kb1; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    int a_col = kb * BK, a_row = m_blk * BM;
                    if (ep.conv_wp) {       // tap (ky, kx) = rows shifted by (ky - 1) * wp + (kx - 1); rows outside the matrix are zero-filled by TMA
                        const int tap = kb / ep.conv_kpt;
                        a_col = (kb - tap * ep.conv_kpt) * BK;
                        a_row += (tap / 3 - 1) * ep.conv_wp + (tap % 3 - 1);
                    }
                    if (!elect_one()) {
                    } else if (CL == 2) {
                        const uint32_t leader_full = mapa_u32(smem_u32(&full_bar[stage]), 0);
                        if (cta_rank == 0) mbar_expect_tx(&full_bar[stage], 2 * L::STAGE_BYTES);
                        else mbar_arrive_cluster(leader_full);
                        tma_load_2d_cg2(smem_a + stage * L::A_BYTES, &tmap_a, a_col, a_row, leader_full);
                        tma_load_2d_cg2(smem_b + stage * L::B_BYTES, &tmap_w, kb * BK, n_blk * BN + cta_rank * (BN / 2), leader_full);   // (BN/2)-row box
                    } else {
                        mbar_expect_tx(&full_bar[stage], L::STAGE_BYTES);
                        tma_load_2d(smem_a + stage * L::A_BYTES, &tmap_a, a_col, a_row, &full_bar[stage]);
                        tma_load_2d(smem_b + stage * L::B_BYTES, &tmap_w, kb * BK, n_blk * BN, &full_bar[stage]);
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (cta_rank == 0) {
            // kind::f16 takes fp16 and bf16 operands, chosen per operand by the descriptor (runtime: no extra template instances)
            const uint32_t idesc = TF32 ? umma_idesc(BM * CL, BN, 2u, 2u) : umma_idesc(BM * CL, BN, static_cast<uint32_t>(ep.fmt_a), static_cast<uint32_t>(ep.fmt_b));
            uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
            for (int tile = tiles.next(); tile >= 0; tile = tiles.next()) {
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                const int kb_first = tiles.kb0, kb_last = tiles.kb1 - 1;
                for (int kb = kb_first; kb <= kb_last; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint64_t da = umma_desc_sw128(smem_u32(smem_a + stage * L::A_BYTES));
                    const uint64_t db = umma_desc_sw128(smem_u32(smem_b + stage * L::B_BYTES));
                    if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < MMAS_PER_KB; ++k) {
                        // advance 32 bytes (16 bf16 / 8 tf32) along K inside the 128-byte swizzle row: +2 in 16-byte units
                        const uint32_t accum = (kb != kb_first || k != 0) ? 1u : 0u;
                        if (CL == 2) {
                            if (TF32) umma_tf32_cg2(d_tmem, da + 2 * k, db + 2 * k, idesc, accum);
                            else umma_bf16_cg2(d_tmem, da + 2 * k, db + 2 * k, idesc, accum);
                        } else {
                            if (TF32) umma_tf32(d_tmem, da + 2 * k, db + 2 * k, idesc, accum);
                            else umma_bf16(d_tmem, da + 2 * k, db + 2 * k, idesc, accum);
                        }
                    }
                    if (CL == 2) {
                        umma_commit_cg2(&empty_bar[stage]);
                        if (kb == kb_last) umma_commit_cg2(&tmem_full[acc]);
                    } else {
                        umma_commit(&empty_bar[stage]);
                        if (kb == kb_last) umma_commit(&tmem_full[acc]);
                    }
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            }
        }
    } else if (warp >= 4) {
        // Two epilogue groups (warps 4-7 and 8-11).  A warp may only touch the TMEM lane quarter warp % 4, so both groups
        // cover all 128 rows and split the 32-column chunks between them (even / odd).  With a single group the
        // epilogue's serial chain (TMEM load -> operand loads -> math -> stores, ~1 us per chunk) was longer than the
        // MMA time of a tile and capped the tensor pipe at 55 % (ncu); two warps per scheduler also hide each other's
        // latencies.
        const int ew = (warp - 4) & 3;     // TMEM lane quarter = warp % 4
        const int grp = (warp - 4) >> 2;   // 0 / 1
        uint32_t acc = 0, acc_phase = 0;
        if (EPI != EPI_GENERIC) {
            // ---- straight-line epilogue of a hot shape (host guarantees vec256_ok, N % 32 == 0, no convolution) ----
            // Everything the math needs besides the accumulator is requested BEFORE the wait it would otherwise follow:
            //   * the bias slice of this warp's chunks goes global -> registers before the wait for the MMAs of the tile, then
            //     into a private shared-memory strip (broadcast LDS in the math instead of four exposed 256-bit global loads per
            //     chunk: the first version lost ~0.3 us per chunk there);
            //   * the residual / saved pre-activation segment of chunk k + 1 is requested while chunk k's TMEM load is in flight
            //     (double-buffered registers), the first chunk's before the wait for the tile.
            constexpr int NCH = BN / 64;                 // chunks of 32 columns per group and tile (pairs {2g, 2g+1}, {2g+4, 2g+5})
            float* s_bias = reinterpret_cast<float*>(smem + L::EPI_OFFSET) + (warp - 4) * (NCH * 32);     // the staged path's area is unused here
            for (int tile = tiles.next(); tile >= 0; tile = tiles.next()) {
                const int m_blk = (tile / num_n) * CL + cta_rank, n_blk = tile % num_n;
                const long long row0 = static_cast<long long>(m_blk) * BM + ew * 32;
                const long long orow = row0 + lane;
                const bool row_ok = orow < M;
                const int role = tiles.role;          // stream-K: 0 whole tile, 1 owner of a cut tile, 2 helper (raw accumulator -> sk_ws)
                const bool use_bias = EPI != EPI_DQGELU && ep.bias != nullptr && role != 2;
                // the parked accumulator of a cut tile: slot = the HELPER's cluster id, [CL * 128 rows][BN] fp32; one flag per
                // (slot, CTA of the pair, epilogue warp): the same warp of the same CTA rank reads exactly what it wrote
                const int sk_slot = first_tile + (role == 1 ? 1 : 0);
                float* sk_row = ep.sk_ws + (static_cast<long long>(sk_slot) * (CL * BM) + cta_rank * BM + ew * 32 + lane) * BN;
                volatile int* sk_flag = ep.sk_flags + (sk_slot * 2 + static_cast<int>(cta_rank)) * 8 + (warp - 4);
                float bias_r[NCH];
#pragma unroll
                for (int k = 0; k < NCH; ++k) {
                    const int col = n_blk * BN + (grp * 2 + (k & 1) + (k >> 1) * 4) * 32 + lane;
                    bias_r[k] = (use_bias && col < N) ? __ldg(ep.bias + col) : 0.f;
                }
                const float* ovr_row = nullptr;       // this thread's output row is a prompt row: its replacement values
                if (EPI == EPI_RES_F32 && ep.ovr_ctx != nullptr && row_ok) {
                    const long long smp = orow / ep.ovr_S;
                    const int pr = static_cast<int>(orow - smp * ep.ovr_S) - ep.ovr_row0;
                    if (pr >= 0 && pr < ep.ovr_n) ovr_row = ep.ovr_ctx + smp * ep.ovr_bs + static_cast<long long>(pr) * N;
                }
                uint32_t ext[2][32];
                if (role != 2 && row_ok && n_blk * BN + grp * 64 < N) spec_load_ext<EPI>(ep, orow, n_blk * BN + grp * 64, ext[0], ovr_row);
                mbar_wait(&tmem_full[acc], acc_phase);
                tc_fence_after();
                if (role == 1) {        // the helper parked its part at the very start of its range, this tile is our last: no real wait
                    while (*sk_flag == 0) { }
                    __threadfence();
                }
                if (use_bias) {
                    __syncwarp();                        // the previous tile's reads of the strip are done
#pragma unroll
                    for (int k = 0; k < NCH; ++k) s_bias[k * 32 + lane] = bias_r[k];
                    __syncwarp();
                }
                const uint32_t t_row = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + acc * BN;
#pragma unroll
                for (int k = 0; k < NCH; ++k) {
                    const int cc = grp * 2 + (k & 1) + (k >> 1) * 4;
                    const int col0 = n_blk * BN + cc * 32;
                    if (row0 < M && col0 < N) {          // warp-uniform
                        uint32_t r[32];
                        tmem_ld_32x32(t_row + cc * 32, r);
                        if (role != 2 && k + 1 < NCH) {
                            const int coln = n_blk * BN + (grp * 2 + ((k + 1) & 1) + ((k + 1) >> 1) * 4) * 32;
                            if (row_ok && coln < N) spec_load_ext<EPI>(ep, orow, coln, ext[(k + 1) & 1], ovr_row);
                        }
                        uint32_t pr[4][8];
                        if (role == 1 && row_ok) {
#pragma unroll
                            for (int q = 0; q < 4; ++q) ld256_cg(sk_row + cc * 32 + 8 * q, pr[q]);
                        }
                        tmem_ld_wait();
                        if (role == 1 && row_ok) {      // + the k-blocks the helper computed (fixed order: deterministic)
#pragma unroll
                            for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) + __uint_as_float(pr[i >> 3][i & 7]));
                        }
                        if (role == 2) {
                            if (row_ok) {
#pragma unroll
                                for (int q = 0; q < 4; ++q) {
                                    uint32_t w8[8];
#pragma unroll
                                    for (int i = 0; i < 8; ++i) w8[i] = r[8 * q + i];
                                    st256(sk_row + cc * 32 + 8 * q, w8);
                                }
                            }
                        } else if (row_ok) epilogue_spec<EPI>(ep, r, orow, col0, ext[k & 1], use_bias ? smem_u32(s_bias + k * 32) : 0u, ovr_row != nullptr);
                    }
                }
                tc_fence_before();
                if (role == 2) __threadfence();       // every lane's parked values are visible before the flag is
                __syncwarp();
                if (lane == 0) {
                    if (role == 2) *sk_flag = 1;
                    if (role == 1) *sk_flag = 0;      // consumed: the slot is clean for the next launch (stream order)
                    if (CL == 2 && cta_rank != 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tmem_empty[acc]), 0));
                    else mbar_arrive(&tmem_empty[acc]);
                }
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            }
        } else
        for (int tile = tiles.next(); tile >= 0; tile = tiles.next()) {
            const int m_blk = (tile / num_n) * CL + cta_rank, n_blk = tile % num_n;
            mbar_wait(&tmem_full[acc], acc_phase);
            tc_fence_after();
            const long long row0 = static_cast<long long>(m_blk) * BM + ew * 32;
            const uint32_t t_row = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + acc * BN;
            float* stage = reinterpret_cast<float*>(smem + L::EPI_OFFSET) + ew * EPI_WARP_FLOATS;
            const bool direct = ep.vec256_ok && !ep.staged;
            bool row_ok = row0 + lane < M;
            long long orow = row0 + lane;          // the row of the output (and residual / aux) matrices this accumulator row belongs to
            if (ep.conv_wp) {
                const int hpwp = ep.conv_hp * ep.conv_wp;
                const long long b = orow / hpwp;
                const int rem = static_cast<int>(orow - b * hpwp);
                const int yp = rem / ep.conv_wp, xp = rem - yp * ep.conv_wp;
                row_ok = row_ok && yp >= 1 && yp <= ep.conv_hp - 2 && xp >= 1 && xp <= ep.conv_wp - 2;     // border rows are padding
                orow = (b * (ep.conv_hp - 2) + (yp - 1)) * (ep.conv_wp - 2) + (xp - 1);
            }
            const bool extras = direct && row_ok && (ep.residual != nullptr || is_dact(ep.act));
            // a group takes PAIRS of adjacent chunks (64 columns): for the bf16 arrays a thread then reads / writes whole
            // 128-byte lines within a few hundred cycles instead of a quarter of a line per visit
#pragma unroll 1
            for (int cc = grp * 2; cc < BN / 32; cc += (cc & 1) ? 3 : 1) {
                const int c = cc;
                const int col0 = n_blk * BN + c * 32;
                if (row0 >= M || col0 >= N) break;      // warp-uniform: nothing of this chunk is inside the matrix
                uint32_t r[32];
                tmem_ld_32x32(t_row + c * 32, r);
                if (direct && col0 + 32 <= N) {
                    EpiExtras ex;
                    if (extras) load_extras(ep, orow, col0, ex);     // in flight together with the TMEM load
                    tmem_ld_wait();
                    if (row_ok) epilogue_direct(ep, r, orow, col0, ex);
                } else {
                    tmem_ld_wait();
                    // ragged / unaligned chunks go through the shared-memory stage, which exists once per lane quarter:
                    // group 1 hands them to group 0 (tiny GEMMs only)
                    if (grp == 0) epilogue_chunk(ep, r, stage, row0, col0, M, N, lane);
                }
            }
            if (grp == 0 && !(direct && N % 32 == 0)) {
                // pick up the chunks that group 1 skipped because they need the staged path
#pragma unroll 1
                for (int c = 2; c < BN / 32; c += (c & 1) ? 3 : 1) {
                    const int col0 = n_blk * BN + c * 32;
                    if (row0 >= M || col0 >= N) break;
                    if (direct && col0 + 32 <= N) continue;
                    uint32_t r[32];
                    tmem_ld_32x32(t_row + c * 32, r);
                    tmem_ld_wait();
                    epilogue_chunk(ep, r, stage, row0, col0, M, N, lane);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (CL == 2 && cta_rank != 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tmem_empty[acc]), 0));
                else mbar_arrive(&tmem_empty[acc]);
            }
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
        }
    }

    tc_fence_before();
    if (CL == 2) cluster_sync_all();   // do not exit while the peer can still signal this CTA's barriers
    else __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        if (CL == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// 2-D row-major [rows, cols] (bf16 or fp32) with leading dimension ld (elements); box = box_rows x 128 bytes, 128B swizzle.
static int make_tmap(CUtensorMap* map, const void* ptr, long long rows, long long cols, long long ld, int box_rows, bool f32) {
    EncodeTiledFn fn = encode_fn();
    TVS_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled not available from the driver");
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
    const int esz = f32 ? 4 : 2;
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * esz};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(BK_BYTES / esz), static_cast<cuuint32_t>(box_rows)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    TVS_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with %d (rows=%lld cols=%lld ld=%lld)", (int)r, rows, cols, ld);
    return 0;
}

// counter pairs {next tile, clusters done} of the dynamic tile scheduler: zero at module load, self-resetting (see the
// kernel), handed out round-robin - a pair is reused 4096 scheduled launches later, long after its kernel has finished
constexpr int SCHED_SLOTS = 4096;
__device__ int g_sched[2 * SCHED_SLOTS];

static int* sched_slot() {
    static int* base[64] = {};
    static unsigned seq = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    if (!base[dev]) {
        void* p = nullptr;
        if (cudaGetSymbolAddress(&p, g_sched) != cudaSuccess) return nullptr;
        base[dev] = static_cast<int*>(p);
    }
    return base[dev] + 2 * (seq++ % SCHED_SLOTS);
}

// stream-K work space: one parked accumulator tile (CL * 128 x 256 fp32) and 2 x 8 flags per cluster, per (device, stream) so that
// GEMMs on different streams never share a slot.  Allocated on first use - never inside a stream capture (the call then runs
// without stream-K; the warm-up steps that precede a capture have allocated it by then).
struct SkWs { int dev; cudaStream_t stream; float* ws; int* flags; };
static bool sk_workspace(cudaStream_t stream, int clusters, float** ws, int** flags) {
    static SkWs table[16];
    static int n = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return false;
    for (int i = 0; i < n; ++i)
        if (table[i].dev == dev && table[i].stream == stream) { *ws = table[i].ws; *flags = table[i].flags; return true; }
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(stream, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone || n == 16) return false;
    const int slots = sm_count() + 1;          // clusters + 1: the owner of the last cluster never reads, but keep the index in range
    (void)clusters;
    float* w = nullptr;
    int* f = nullptr;
    if (cudaMalloc(&w, static_cast<size_t>(slots) * 256 * 256 * sizeof(float)) != cudaSuccess) { cudaGetLastError(); return false; }
    if (cudaMalloc(&f, static_cast<size_t>(slots) * 16 * sizeof(int)) != cudaSuccess) { cudaGetLastError(); cudaFree(w); return false; }
    cudaMemset(f, 0, static_cast<size_t>(slots) * 16 * sizeof(int));
    table[n++] = SkWs{dev, stream, w, f};
    *ws = w; *flags = f;
    return true;
}

static thread_local int g_last_variant = 0;
template <int BN, int STAGES, bool TF32, int CL, int EPI>
static int launch_gemm_inst(const tvs_gemm_args& a, const GemmEpilogue& ep_in, cudaStream_t stream) {
    GemmEpilogue ep = ep_in;
    using L = GemmSmem<BN, STAGES, CL>;
    g_last_variant = (BN << 16) | (STAGES << 8) | (EPI << 5) | ((TF32 ? 1 : 0) << 4) | CL;
    CUtensorMap ta, tw;
    if (int rc = make_tmap(&ta, a.A, a.M, a.conv_h ? a.K / 9 : a.K, a.lda, BM, TF32)) return rc;
    if (int rc = make_tmap(&tw, a.W, a.N, a.K, a.ldw, BN / CL, TF32)) return rc;
    auto kern = gemm_bf16_tcgen05_kernel<BN, STAGES, TF32, CL, EPI>;
    static bool attr_set = false;
    if (!attr_set) {
        TVS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
        attr_set = true;
    }
    const int tiles = (((a.M + BM - 1) / BM + CL - 1) / CL) * ((a.N + BN - 1) / BN);
    const int max_clusters = sm_count() / CL;
    const int clusters = tiles < max_clusters ? tiles : max_clusters;
    // TVS_GEMM_SCHED: "static" round-robin tiles; "cl1" / "cl2" dynamic only for single-CTA / pair kernels; "memset" also zeroes the
    // counter pair in the stream before every launch (debug switches)
    // Default: static.  Measured on the B = 32 step (same-box pairs): dynamic for single-CTA kernels only 10.03 ms vs 9.95 ms static, dynamic
    // for the pair kernels 10.50 ms (nine remote arrivals per tile and peer) - the interference from the text-tower stream is not a
    // late-start tail, so the scheduler buys nothing; and a pair cluster that draws NO tile (forced with "force" + TVS_GEMM_CLUSTER=2)
    // hangs at exit.  Kept as an experiment switch ("dyn", "cl1") - never the default.
    static const char mode = [] { const char* e = getenv("TVS_GEMM_SCHED"); return e ? (e[0] == 'c' ? e[2] : e[0]) : 's'; }();
    const bool dyn_ok = mode == 'd' || mode == 'm' || mode == 'f' || (mode == '1' && CL == 1) || (mode == '2' && CL == 2);
    int* sched = (dyn_ok && (tiles > clusters || mode == 'f')) ? sched_slot() : nullptr;      // one tile per cluster needs no scheduler
    const int launch_clusters = mode == 'f' ? max_clusters : clusters;      // "force": every SM gets a CTA, most of them draw no tile at all (test switch)
    if (sched && mode == 'm') TVS_CUDA(cudaMemsetAsync(sched, 0, 2 * sizeof(int), stream));
    // stream-K for the specialised pair kernels (the tower shapes at M = B * S): whole-tile round robin costs ceil(tiles / clusters)
    // rounds - 3 instead of 2.51 for the N = 768 shapes, 11 instead of 10.05 for N = 3072 at M = 15 648.  OPT-IN (TVS_GEMM_STREAM_K /
    // TVS_GEMM_STREAMK=1), because it measured SLOWER: N = 3072 plain 54.6 -> 62.1 us, fc2 65.7 -> 80.1, QKV 43.2 -> 51.4.  The kernels
    // are bound by L2 -> SM ingress and epilogue traffic, not by the MMA count: a last round with fewer active pairs runs faster than
    // a full one (so the quantisation costs less than the tile count suggests), while every pair pays two extra fp32 accumulator
    // transfers (park + fetch, 128 KB per CTA each) for its two cut tiles.
    ep.sk_q = 0; ep.sk_ws = nullptr; ep.sk_flags = nullptr;
    static const bool sk_env = [] { const char* e = getenv("TVS_GEMM_STREAMK"); return e && e[0] == '1'; }();
    if (EPI != EPI_GENERIC && CL == 2 && (sk_env || (a.reserved & TVS_GEMM_STREAM_K)) && sched == nullptr && launch_clusters == max_clusters && tiles > max_clusters && tiles % max_clusters != 0) {
        constexpr int BKE = TF32 ? BK_BYTES / 4 : BK_BYTES / 2;
        const long long num_kb = (a.K + BKE - 1) / BKE;
        const long long total = static_cast<long long>(tiles) * num_kb;
        const long long q = (total + max_clusters - 1) / max_clusters;
        if (q >= num_kb && total < (1LL << 31) && sk_workspace(stream, max_clusters, &ep.sk_ws, &ep.sk_flags)) ep.sk_q = static_cast<int>(q);
    }
    if (ep.sk_q > 0) g_last_variant |= 1 << 28;      // reported as " stream-k" (tvs_gemm_last_variant)
    TVS_CUDA(launch_pdl(kern, dim3(launch_clusters * CL), dim3(GEMM_THREADS), L::TOTAL, stream, CL, ta, tw, a.M, a.N, a.K, ep, sched));
    return check_launch("gemm_bf16_tcgen05_kernel");
}

// Which straight-line epilogue (if any) covers this call.  TVS_GEMM_EPI=generic forces the runtime-flag epilogue (A/B switch).
static int pick_epi(const tvs_gemm_args& a, const GemmEpilogue& ep, bool tf32) {
    static const bool generic_only = [] { const char* e = getenv("TVS_GEMM_EPI"); return e && e[0] == 'g'; }();
    if (generic_only || tf32 || a.conv_h || !ep.vec256_ok || ep.staged || a.N % 32 != 0 || ep.round_out) return EPI_GENERIC;
    const bool o16 = ep.out_bf16 != nullptr, o32 = ep.out_f32 != nullptr, res = ep.residual != nullptr, pre = ep.pre_bf16 != nullptr;
    if (ep.act == TVS_ACT_NONE && o16 && !ep.out16_f16 && !o32 && !res && !pre) return EPI_OUT_BF16;
    if (ep.act == TVS_ACT_NONE && o32 && res && !o16 && !pre) return EPI_RES_F32;
    if (ep.act == TVS_ACT_QGELU && o16 && ep.out16_f16 && pre && !o32 && !res) return EPI_FC1;
    if ((ep.act == TVS_ACT_DQGELU || ep.act == TVS_ACT_MULAUX) && o16 && !ep.out16_f16 && !o32 && !res && !pre && !ep.bias) return EPI_DQGELU;
    return EPI_GENERIC;
}

template <int BN, int STAGES, bool TF32, int CL>
static int launch_gemm_cl(const tvs_gemm_args& a, const GemmEpilogue& ep, cudaStream_t stream) {
    if constexpr (!TF32 && BN >= 128) {       // the instances the vision tower runs at M = B * S: specialised epilogues
        switch (pick_epi(a, ep, TF32)) {
            case EPI_OUT_BF16: return launch_gemm_inst<BN, STAGES, TF32, CL, EPI_OUT_BF16>(a, ep, stream);
            case EPI_RES_F32: return launch_gemm_inst<BN, STAGES, TF32, CL, EPI_RES_F32>(a, ep, stream);
            case EPI_FC1: return launch_gemm_inst<BN, STAGES, TF32, CL, EPI_FC1>(a, ep, stream);
            case EPI_DQGELU: return launch_gemm_inst<BN, STAGES, TF32, CL, EPI_DQGELU>(a, ep, stream);
            default: break;
        }
    }
    return launch_gemm_inst<BN, STAGES, TF32, CL, EPI_GENERIC>(a, ep, stream);
}

// CTA pairs pay off when the GEMM fills the machine; tiny problems keep independent CTAs
template <int BN, int STAGES, bool TF32>
static int launch_gemm(const tvs_gemm_args& a, const GemmEpilogue& ep, cudaStream_t stream) {
    static const int mode = [] { const char* e = getenv("TVS_GEMM_CLUSTER"); return e ? atoi(e) : -1; }();   // -1 auto, 1 off, 2 on
    const long long tiles = static_cast<long long>((a.M + BM - 1) / BM) * ((a.N + BN - 1) / BN);
    constexpr int STAGES2 = BN == 256 ? 6 : 8;     // 32 KB / 24 KB per stage per CTA in pair mode
    // pairs win on the compute-heavy shapes; the memory-bound N = K = 768 residual GEMM is slightly better with single CTAs
    // (measured at M = 15 648, N = K = 768: with the fp32 residual in / out 38.7 us single vs 39.6 us pairs; with a plain 16-bit output
    // - the out-proj dgrad - 23.0 us single vs 21.0 us pairs, so that case takes the pairs too)
    const bool heavy = static_cast<long long>(a.N) * a.K >= 768LL * 1024 ||
                       (static_cast<long long>(a.N) * a.K >= 512LL * 1024 && a.out_f32 == nullptr && a.residual == nullptr && a.pre_bf16 == nullptr);
    const bool use2 = BN >= 128 && (mode == 2 || (mode == -1 && tiles >= 2LL * sm_count() && heavy));
    if (use2) return launch_gemm_cl<(BN >= 128 ? BN : 128), STAGES2, TF32, 2>(a, ep, stream);
    return launch_gemm_cl<BN, STAGES, TF32, 1>(a, ep, stream);
}

static int pick_tile_n(int M, int N, int K) {
    if (N <= 64) return 64;
    const int sms = sm_count();
    const int num_m = (M + BM - 1) / BM;
    // small problems (text tower, B*S ~ 400 rows): the serial K loop of one CTA is the critical path, so prefer the
    // narrowest tile - twice the CTAs and half the MMA time per k-block
    if (static_cast<long long>(num_m) * ((N + 127) / 128) < sms) return 64;
    if (N <= 128) return 128;
    // Compute-heavy shapes that will run as CTA pairs (launch_gemm): 256-wide tiles.  The pair kernels are bound by L2 -> SM
    // ingress (~64 B/clk/SM, ncu round 2); a 256 x 256 pair tile needs 64 B/clk per SM, a 256 x 128 one 96 B/clk, which
    // outweighs the wave quantisation the estimate below optimises (measured at M = 15 648: N = 768, K = 3072 70.7 -> 66.6 us
    // with the fp32 residual epilogue, 61.7 -> 57.7 plain; N = 2304, K = 768 47.7 -> 44.5; N = 768, K = 2304 47.6 -> 44.7;
    // the memory-bound N = K = 768 stays on single-CTA 128-wide tiles: 38.0 vs 43.1).
    if (N % 256 == 0 && static_cast<long long>(N) * K >= 768LL * 1024 && static_cast<long long>(num_m) * (N / 256) >= 2LL * sms) return 256;
    auto eff = [&](int bn) {
        long long tiles = static_cast<long long>(num_m) * ((N + bn - 1) / bn);
        long long waves = (tiles + sms - 1) / sms;
        double wave_eff = static_cast<double>(tiles) / static_cast<double>(waves * sms);
        double pad_eff = static_cast<double>(N) / static_cast<double>(((N + bn - 1) / bn) * bn);
        // 128x128 tiles read 128 B/clk of smem per MMA (the limit); 128x256 needs 96 B/clk
        return wave_eff * pad_eff * (bn == 256 ? 1.0 : 0.97);
    };
    return eff(256) >= eff(128) ? 256 : 128;
}

}  // namespace tvs

extern "C" __attribute__((visibility("default"))) int32_t tvs_gemm_last_variant(void) { return tvs::g_last_variant; }

extern "C" __attribute__((visibility("default"))) int tvs_gemm_bf16(const tvs_gemm_args* args, void* stream) {
    using namespace tvs;
    TVS_REQUIRE(args != nullptr, "tvs_gemm_bf16: null args");
    const tvs_gemm_args& a = *args;
    TVS_REQUIRE(a.M > 0 && a.N > 0 && a.K > 0, "tvs_gemm_bf16: bad shape M=%d N=%d K=%d", a.M, a.N, a.K);
    TVS_REQUIRE(a.ab_dtype == TVS_AB_BF16 || a.ab_dtype == TVS_AB_TF32 || a.ab_dtype == TVS_AB_F16,
                "tvs_gemm_bf16: ab_dtype must be TVS_AB_BF16, TVS_AB_TF32 or TVS_AB_F16 (got %d; mixed fp16 x bf16 operands are illegal on sm_100)", a.ab_dtype);
    const int kal = a.ab_dtype == TVS_AB_TF32 ? 4 : 8;   // 16-byte rows for TMA
    TVS_REQUIRE(a.K % kal == 0 && a.lda % kal == 0 && a.ldw % kal == 0, "tvs_gemm_bf16: K, lda, ldw must be multiples of %d (K=%d lda=%lld ldw=%lld)",
                kal, a.K, (long long)a.lda, (long long)a.ldw);
    TVS_REQUIRE((reinterpret_cast<uintptr_t>(a.A) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.W) & 15) == 0,
                "tvs_gemm_bf16: A and W must be 16-byte aligned");
    TVS_REQUIRE(a.out_f32 || a.out_bf16 || a.pre_bf16, "tvs_gemm_bf16: no output");
    TVS_REQUIRE(!(a.act == TVS_ACT_DQGELU || a.act == TVS_ACT_DRELU || a.act == TVS_ACT_MULAUX) || a.aux_bf16, "tvs_gemm_bf16: aux_bf16 required for derivative epilogues");
    TVS_REQUIRE(!(a.reserved & TVS_GEMM_PRE_DGELU) || (a.act == TVS_ACT_QGELU && a.pre_bf16), "tvs_gemm_bf16: TVS_GEMM_PRE_DGELU needs TVS_ACT_QGELU and pre_bf16");
    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    const bool vec_ok = (!a.out_f32 || (a.ldo32 % 4 == 0 && al16(a.out_f32))) && (!a.out_bf16 || (a.ldo16 % 4 == 0 && al16(a.out_bf16))) &&
                        (!a.pre_bf16 || (a.ldpre % 4 == 0 && al16(a.pre_bf16))) && (!a.aux_bf16 || (a.ldaux % 4 == 0 && al16(a.aux_bf16))) &&
                        (!a.residual || (a.ldr % 4 == 0 && al16(a.residual))) && (!a.bias || al16(a.bias));
    GemmEpilogue ep;
    ep.bias = a.bias;
    ep.residual = a.residual;
    ep.ldr = a.ldr;
    ep.out_f32 = a.out_f32;
    ep.ldo32 = a.ldo32;
    ep.out_bf16 = static_cast<__nv_bfloat16*>(a.out_bf16);
    ep.ldo16 = a.ldo16;
    ep.pre_bf16 = static_cast<__nv_bfloat16*>(a.pre_bf16);
    ep.ldpre = a.ldpre;
    ep.aux_bf16 = static_cast<const __nv_bfloat16*>(a.aux_bf16);
    ep.ldaux = a.ldaux;
    ep.act = a.act;
    ep.round_out = a.reserved & TVS_GEMM_ROUND_OUT_TF32;
    ep.out16_f16 = (a.reserved & TVS_GEMM_OUT16_F16) ? 1 : 0;
    ep.pre_dgelu = (a.reserved & TVS_GEMM_PRE_DGELU) ? 1 : 0;
    ep.fmt_a = ep.fmt_b = a.ab_dtype == TVS_AB_F16 ? 0 : 1;
    ep.ovr_ctx = a.ovr_ctx;
    ep.ovr_bs = a.ovr_batch_stride;
    ep.ovr_S = a.ovr_S;
    ep.ovr_row0 = a.ovr_row0;
    ep.ovr_n = a.ovr_n;
    if (a.ovr_ctx) {
        TVS_REQUIRE(a.ovr_S > 0 && a.ovr_row0 >= 0 && a.ovr_n > 0 && a.ovr_row0 + a.ovr_n <= a.ovr_S && a.M % a.ovr_S == 0,
                    "tvs_gemm_bf16: prompt overwrite needs M = B * S and 0 <= row0, row0 + n <= S (M=%d S=%d row0=%d n=%d)", a.M, a.ovr_S, a.ovr_row0, a.ovr_n);
        TVS_REQUIRE((reinterpret_cast<uintptr_t>(a.ovr_ctx) & 31) == 0 && a.N % 8 == 0 && a.ovr_batch_stride % 8 == 0,
                    "tvs_gemm_bf16: prompt overwrite needs 32-byte aligned context rows");
    }
    ep.conv_wp = ep.conv_hp = ep.conv_kpt = 0;
    ep.vec_ok = vec_ok ? 1 : 0;
    auto al32 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 31) == 0; };
    ep.vec256_ok = ((!a.out_f32 || (a.ldo32 % 8 == 0 && al32(a.out_f32))) && (!a.out_bf16 || (a.ldo16 % 16 == 0 && al32(a.out_bf16))) &&
                    (!a.pre_bf16 || (a.ldpre % 16 == 0 && al32(a.pre_bf16))) && (!a.aux_bf16 || (a.ldaux % 16 == 0 && al32(a.aux_bf16))) &&
                    (!a.residual || (a.ldr % 8 == 0 && al32(a.residual))) && (!a.bias || al32(a.bias)))
                       ? 1 : 0;
    static const bool force_staged = [] { const char* e = getenv("TVS_GEMM_EPILOGUE"); return e && e[0] == 's'; }();
    ep.staged = force_staged ? 1 : 0;
    if (a.conv_h) {
        const int bk = a.ab_dtype == TVS_AB_TF32 ? BK_BYTES / 4 : BK_BYTES / 2;
        const long long hpwp = static_cast<long long>(a.conv_h + 2) * (a.conv_w + 2);
        TVS_REQUIRE(a.conv_h > 0 && a.conv_w > 0 && a.K % 9 == 0 && (a.K / 9) % bk == 0 && a.lda == a.K / 9,
                    "tvs_gemm_bf16 (conv): K must be 9 * C with C a multiple of %d and lda == C (K=%d lda=%lld)", bk, a.K, (long long)a.lda);
        TVS_REQUIRE(a.M % hpwp == 0 && hpwp < (1LL << 31), "tvs_gemm_bf16 (conv): M=%d must be B * (H+2) * (W+2)", a.M);
        TVS_REQUIRE(ep.vec256_ok && a.N % 32 == 0 && !ep.staged, "tvs_gemm_bf16 (conv): needs N %% 32 == 0 and 32-byte aligned output rows");
        ep.conv_wp = a.conv_w + 2;
        ep.conv_hp = a.conv_h + 2;
        ep.conv_kpt = (a.K / 9) / bk;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    int bn = a.tile_n ? a.tile_n : pick_tile_n(a.M, a.N, a.K);
    if (a.ovr_ctx) {        // only the straight-line residual epilogue implements the fused overwrite
        if (!a.tile_n && bn < 128) bn = 128;
        TVS_REQUIRE(a.ab_dtype != TVS_AB_TF32 && bn >= 128 && pick_epi(a, ep, false) == EPI_RES_F32,
                    "tvs_gemm_bf16: the fused prompt overwrite is implemented for the 16-bit residual GEMM (bias + f32 residual -> f32, "
                    "N %% 32 == 0, aligned rows, tile_n >= 128); use tvs_prompt_overwrite otherwise");
    }
    const bool tf32 = a.ab_dtype == TVS_AB_TF32;
    switch (bn) {
        case 64: return tf32 ? launch_gemm<64, 8, true>(a, ep, s) : launch_gemm<64, 8, false>(a, ep, s);
        case 128: return tf32 ? launch_gemm<128, 6, true>(a, ep, s) : launch_gemm<128, 6, false>(a, ep, s);
        case 256: return tf32 ? launch_gemm<256, 4, true>(a, ep, s) : launch_gemm<256, 4, false>(a, ep, s);
        default: set_error("tvs_gemm_bf16: tile_n must be 0, 64, 128 or 256 (got %d)", bn); return -1;
    }
}
