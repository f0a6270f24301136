// Per-CTA timeline of the attention backward kernels (needs a library built with -DATC_EXPERIMENT=8):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I include tools/microbench/attn_timeline.cu -L tools/ab/e8 -ltvs_b200 -o tools/microbench/attn_timeline
//   LD_LIBRARY_PATH=tools/ab/e8 tools/microbench/attn_timeline
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <random>
#include <vector>

#include "tvs_b200.h"

extern "C" int tvs_debug_timeline(unsigned long long* host, int max_events, int reset);

int main() {
    const int B = 32, S = 489, H = 12, hd = 64, E = H * hd;
    const size_t M = size_t(B) * S;
    std::mt19937 rng(1);
    std::normal_distribution<float> nd(0.f, 1.f);
    std::vector<__nv_bfloat16> hq(M * 3 * E), hdo(M * E);
    for (auto& v : hq) v = __float2bfloat16(nd(rng) * 0.5f);
    for (auto& v : hdo) v = __float2bfloat16(nd(rng) * 0.1f);
    __nv_bfloat16 *qkv, *out, *dout, *dqkv;
    float *lse, *delta;
    cudaMalloc(&qkv, hq.size() * 2);
    cudaMalloc(&out, M * E * 2);
    cudaMalloc(&dout, M * E * 2);
    cudaMalloc(&dqkv, hq.size() * 2);
    cudaMalloc(&lse, size_t(B) * H * S * 4);
    cudaMalloc(&delta, size_t(B) * H * S * 4);
    cudaMemcpy(qkv, hq.data(), hq.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dout, hdo.data(), hdo.size() * 2, cudaMemcpyHostToDevice);
    if (tvs_attn_fwd(qkv, B, S, H, hd, 0, nullptr, out, nullptr, lse, 0, nullptr)) { printf("fwd: %s\n", tvs_last_error()); return 1; }
    for (int i = 0; i < 3; ++i)
        if (tvs_attn_bwd(qkv, out, dout, lse, B, S, H, hd, 0, nullptr, delta, dqkv, 0, nullptr)) { printf("bwd: %s\n", tvs_last_error()); return 1; }
    tvs_debug_timeline(nullptr, 0, 1);
    tvs_attn_bwd(qkv, out, dout, lse, B, S, H, hd, 0, nullptr, delta, dqkv, 0, nullptr);
    const int words = 3 * 2 * 10 * 512 * 2;
    std::vector<unsigned long long> ev(words);
    tvs_debug_timeline(ev.data(), words, 1);
    struct E_ { unsigned long long t; int evt, warp, it; };
    for (int mode = 1; mode <= 2; ++mode)
        for (int cta = 0; cta < 2; ++cta) {
            std::vector<E_> c;
            for (int w = 0; w < 10; ++w)
                for (int i = 0; i < 512; ++i) {
                    const unsigned long long* p = &ev[size_t((((mode * 2 + cta) * 10 + w) * 512 + i)) * 2];
                    if (!(p[1] >> 40)) break;
                    c.push_back({p[0], int((p[1] >> 16) & 0xffff), w, int(p[1] & 255)});
                }
            std::stable_sort(c.begin(), c.end(), [](const E_& a, const E_& b) { return a.t < b.t; });
            printf("---- %s kernel, CTA %d: %zu events (evt: 30 start, 31 outer tiles in, 20 TMA got empty stage, 1 MMA got inner tiles, 2 S/dP issued, "
                   "3 MMA got ew_done, 4 acc issued, 9 EW ready, 10 EW got s_full, 11 tmem ld done, 12 math + st issued, 13 st done, 14 EW loop end, "
                   "32 acc_full, 33 end)\n", mode == 1 ? "dQ" : "dK/dV", cta, c.size());
            if (c.empty()) continue;
            const unsigned long long t0 = c[0].t;
            unsigned long long prev = t0;
            for (auto& e : c) {
                if (e.warp != 0 && e.warp != 4 && e.warp != 8 && e.warp != 9) continue;      // EW warps 0 and 4, TMA (8), MMA (9)
                printf("  t=%7llu (+%5llu) warp %d evt %2d it %d\n", e.t - t0, e.t - prev, e.warp, e.evt, e.it);
                prev = e.t;
            }
        }
    return 0;
}
