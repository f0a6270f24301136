"""GPU: the HBM-bound kernels of the step, one by one, at the bench shapes: CUDA-event timing over rotating buffers larger
than L2 (or an explicit L2 flush), algorithmic bytes / duration against the measured copy bandwidth.
    python tools/kernel_bench.py [name ...]        (also the workload for the ncu --set full captures under profiles/)
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tunevlseg_b200 import abi  # noqa: E402

PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6452.2
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)


def timeit(fn, nbytes, name, reps=20):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()                              # evict L2 (256 MB written)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    us = ts[len(ts) // 2]
    print(f"{name:46s} {us:8.1f} us  {nbytes / 1e6:8.1f} MB  {nbytes / us / 1e3:8.1f} GB/s  {100 * nbytes / us / 1e3 / PEAK:5.1f}% of {PEAK:.0f}", flush=True)


def rnd(*shape, dtype=torch.float32, scale=1.0):
    return (torch.randn(*shape, device=dev, generator=g) * scale).to(dtype)


def bench_dicebce(B, HW):
    N = HW * HW
    lg, mk = rnd(B, 1, HW, HW, scale=3.0), (torch.rand(B, 1, HW, HW, device=dev, generator=g) < 0.3).float()
    parts, counts = torch.empty(B, 4, dtype=torch.float64, device=dev), torch.empty(B, 3, dtype=torch.int64, device=dev)
    conf, loss = torch.zeros(4, dtype=torch.int64, device=dev), torch.empty(1, device=dev)
    scratch = torch.empty(abi.dicebce_scratch_bytes(B, N), dtype=torch.uint8, device=dev)
    timeit(lambda: abi.dicebce_metrics_fwd(lg, mk, 0.5, 1.0, 0.2, parts, counts, conf, loss, scratch), 8 * B * N, f"dicebce_metrics_fwd B={B} {HW}^2")
    dl, gs = torch.empty_like(lg), torch.ones(1, device=dev)
    timeit(lambda: abi.dicebce_bwd(lg, mk, parts, gs, 1.0, 0.2, dl), 12 * B * N, f"dicebce_bwd B={B} {HW}^2")


def bench_ln(M, D):
    x, gm, bt = rnd(M, D), rnd(D), rnd(D)
    y16, mean, rstd = torch.empty(M, D, dtype=torch.float16, device=dev), torch.empty(M, device=dev), torch.empty(M, device=dev)
    timeit(lambda: abi.layernorm_fwd(x, gm, bt, 1e-5, y_bf16=y16, mean=mean, rstd=rstd), M * D * 6, f"layernorm_fwd [{M},{D}] f32 -> f16")
    dy, add = rnd(M, D, dtype=torch.bfloat16), rnd(M, D)
    dx, dx16 = torch.empty(M, D, device=dev), torch.empty(M, D, dtype=torch.bfloat16, device=dev)
    timeit(lambda: abi.layernorm_bwd(dy, x, gm, mean, rstd, dx_add=add, dx_f32=dx, dx_bf16=dx16), M * D * (2 + 4 + 4 + 4 + 2),
           f"layernorm_bwd [{M},{D}] (+add, f32+bf16 out)")


def bench_head(B, G=22, P=16, Dr=64, ks=5):
    G2, H, KK = G * G, G * P, ks * ks
    tconv, addmap = rnd(B * G2, P * P), rnd(B * G2, 32)
    bt, ba, ratio = rnd(1), rnd(1), torch.full((1,), 0.5, device=dev)
    logits, add_out = torch.empty(B, 1, H, H, device=dev), torch.empty(B, H, H, device=dev)
    nb = (tconv.numel() + B * G2 * KK + logits.numel() + add_out.numel()) * 4
    timeit(lambda: abi.head_fwd(tconv, addmap[:, :KK], bt, ba, ratio, abi.BLEND_RATIO, B, G, P, ks, logits, add_out), nb, f"head_fwd B={B} (ratio blend)")
    dl = rnd(B, 1, H, H, scale=1e-3)
    dt16 = torch.empty(B * G2, P * P, dtype=torch.bfloat16, device=dev)
    dam, dba, dr_ = torch.zeros(B * G2, 32, device=dev), torch.zeros(1, device=dev), torch.zeros(1, device=dev)
    nb = (dl.numel() + tconv.numel() + add_out.numel() + B * G2 * KK) * 4 + dt16.numel() * 2
    timeit(lambda: abi.head_bwd(dl, tconv, add_out, bt, ratio, abi.BLEND_RATIO, B, G, P, ks, dt16, dam[:, :KK], dba, dr_), nb, f"head_bwd B={B} (ratio blend)")


def bench_misc(B=32, S=489, D=768):
    img = rnd(B, 3, 352, 352)
    cols = torch.empty(B * 484, 768, dtype=torch.float16, device=dev)
    timeit(lambda: abi.im2col_patches(img, 16, cols), img.numel() * 4 + cols.numel() * 2, "im2col_patches B=32 352^2 -> f16")
    pe, cls, pos, ctx, h = rnd(B * 484, D), rnd(D), rnd(485, D), rnd(4, D), torch.empty(B * S, D, device=dev)
    timeit(lambda: abi.vision_assemble(pe, cls, pos, ctx, B, 484, 4, D, h), (pe.numel() + h.numel()) * 4, "vision_assemble B=32")
    x, mul = rnd(B, S, 64), rnd(B, 64)
    dy, dx, dm, da = rnd(B, S, 64), torch.empty(B, S, 64, device=dev), torch.empty(B, 64, device=dev), torch.empty(B, 64, device=dev)
    timeit(lambda: abi.film_bwd(dy, x, mul, dx, dm, da), 3 * x.numel() * 4, "film_bwd [32,489,64]")
    n = 772_000 // 4 * 4
    p, gr, m, v = rnd(n), rnd(n), torch.zeros(n, device=dev), torch.zeros(n, device=dev)
    timeit(lambda: abi.adamw_flat(p, gr, m, v, 1e-3, 0.9, 0.999, 1e-8, 0.0, 1), 7 * n * 4, "adamw_flat 0.77 M params")
    a, b16 = rnd(B * S, D), torch.empty(B * S, D, dtype=torch.bfloat16, device=dev)
    timeit(lambda: abi.cast_bf16(a, b16), a.numel() * 6, "cast_bf16 [15648,768]")
    dyw, xw, dw = rnd(B * 484, 25), rnd(B * 484, 64), torch.zeros(25, 64, device=dev)
    timeit(lambda: abi.wgrad_small(dyw, xw, dw), (dyw.numel() + xw.numel()) * 4, "wgrad_small [15488,25]x[15488,64]")


ALL = {"dicebce32": lambda: bench_dicebce(32, 352), "dicebce256": lambda: bench_dicebce(256, 416), "ln": lambda: bench_ln(15648, 768),
       "ln64": lambda: bench_ln(15648, 64), "head": lambda: bench_head(32), "misc": bench_misc}
if __name__ == "__main__":
    torch.cuda.set_device(0)
    abi.require_device()
    for k in (sys.argv[1:] or list(ALL)):
        ALL[k]()
