"""Time one tower GEMM shape through the C ABI (CUDA events, 50 launches): python tools/gemm_shape_bench.py N K epilogue [tile_n]
epilogue: plain | res | fc1 | fc1grad | dqgelu | mulaux.  Environment switches of the library (TVS_GEMM_EPILOGUE, TVS_GEMM_EPI, ...) apply."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tunevlseg_b200 import abi  # noqa: E402

N, K, epi = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
tile_n = int(sys.argv[4]) if len(sys.argv) > 4 else 0
M = 32 * 489
dt = torch.bfloat16 if epi in ("plain", "dqgelu", "mulaux") else torch.float16
A = (torch.randn(M, K, device="cuda") * 0.5).to(dt)
W = (torch.randn(N, K, device="cuda") * K ** -0.5).to(dt)
bias = torch.randn(N, device="cuda") * 0.1
kw = {}
if epi == "plain":
    kw = dict(out_bf16=torch.empty(M, N, dtype=torch.bfloat16, device="cuda"))
elif epi == "res":
    kw = dict(bias=bias, residual=torch.randn(M, N, device="cuda"), out_f32=torch.empty(M, N, device="cuda"))
elif epi == "fc1":
    kw = dict(bias=bias, pre_bf16=torch.empty(M, N, dtype=torch.bfloat16, device="cuda"), out_bf16=torch.empty(M, N, dtype=dt, device="cuda"), act=abi.ACT_QGELU)
elif epi == "fc1grad":
    kw = dict(bias=bias, pre_bf16=torch.empty(M, N, dtype=torch.bfloat16, device="cuda"), out_bf16=torch.empty(M, N, dtype=dt, device="cuda"), act=abi.ACT_QGELU,
              pre_is_grad=True)
elif epi == "mulaux":
    kw = dict(aux_bf16=torch.rand(M, N, device="cuda").to(torch.bfloat16), out_bf16=torch.empty(M, N, dtype=torch.bfloat16, device="cuda"), act=abi.ACT_MULAUX)
elif epi == "dqgelu":
    kw = dict(aux_bf16=torch.randn(M, N, device="cuda").to(torch.bfloat16), out_bf16=torch.empty(M, N, dtype=torch.bfloat16, device="cuda"), act=abi.ACT_DQGELU)
for _ in range(5):
    abi.gemm(A, W, tile_n=tile_n, **kw)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
for _ in range(50):
    abi.gemm(A, W, tile_n=tile_n, **kw)
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1e3 / 50
print(f"M={M} N={N} K={K} {epi} tile_n={tile_n}: {us:.1f} us, {2.0 * M * N * K / us / 1e6:.0f} TFLOP/s [{abi.gemm_last_variant()}]")
