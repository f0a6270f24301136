"""CPU experiment (no GPU): which rounding sites make up the VPT logit error at full geometry?

Runs the product engine over tests/fake_abi.py (the CPU emulation of the C ABI that reproduces the device's operand
rounding) and ablates one rounding site at a time.  Prints max-abs and RMS logit error against the fp32 oracle.
    python tools/precision_attribution.py [case] [B]
"""
import contextlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["TVS_TEXT_STREAM"] = "0"

from oracle import clipseg as OC  # noqa: E402
from tests import fake_abi  # noqa: E402
from tests.helpers import FULL, build_net, make_batch, oracle_head, oracle_state  # noqa: E402
from tunevlseg_b200 import abi, engine  # noqa: E402


class MP:
    def setattr(self, obj, name, val, raising=True):
        setattr(obj, name, val)


fake_abi.install(MP())


class _NoStream:
    def wait_stream(self, other): ...


torch.cuda.current_stream = lambda *a, **k: _NoStream()
torch.cuda.stream = lambda s: contextlib.nullcontext()
torch.Tensor.record_stream = lambda self, s: None

case = sys.argv[1] if len(sys.argv) > 1 else "vpt"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
spec = FULL
weights = OC.init_weights(spec, seed=7)
img, ids, am, mask = make_batch(spec, B, 8, 4)


def run(tag, setup=None):
    net = build_net(case, spec, weights, seed=3)
    st, head = oracle_state(case, net, spec), oracle_head(net)
    undo = setup() if setup else None
    with torch.no_grad():
        logits = net(text_input={"input_ids": ids, "attention_mask": am}, image_input=img)
        ref = OC.net_forward(weights, spec, st, head, ids, am, img)
    if undo:
        undo()
    e = (logits - ref)
    print(f"{tag:50s} max {e.abs().max().item():.5f}  rms {e.pow(2).mean().sqrt().item():.5f}  |logit|max {ref.abs().max().item():.2f}", flush=True)


orig_gemm = fake_abi.gemm
orig_attn = fake_abi.attn_fwd
orig_ln = fake_abi.layernorm_fwd
orig_e = engine._e
orig_bf = engine._bf


def exact_vision_gemm_inputs():
    """bf16 GEMMs take fp32 operands (weights + activations exact): all vision-tower operand rounding removed."""
    def _e(shape, dtype, like):
        return orig_e(shape, torch.float32 if dtype == torch.bfloat16 else dtype, like)

    def gemm(A, W, **kw):
        if A.dtype == torch.float32 and getattr(W, "_was_bf16", False):
            # emulate exact: bypass tf32 truncation
            v = A @ W.t()
            return _finish(v, kw)
        return orig_gemm(A, W, **kw)
    engine._e = _e
    return lambda: setattr(engine, "_e", orig_e)


def patch_alloc(pred):
    """allocate fp32 instead of bf16 for buffers selected by their shape predicate"""
    def _e(shape, dtype, like):
        if dtype == torch.bfloat16 and pred(shape):
            return orig_e(shape, torch.float32, like)
        return orig_e(shape, dtype, like)
    engine._e = _e
    return lambda: setattr(engine, "_e", orig_e)


# fake gemm variant that accepts mixed dtypes (fp32 activation with bf16 weights, or fp32 weights kept exact)
def gemm_mixed(A, W, **kw):
    if A.dtype != W.dtype:
        Wf = W.float()
        Af = A.float()
        A2 = Af if A.dtype == torch.float32 else Af
        return orig_gemm(A2.to(torch.bfloat16) if False else _Exact(A2), _Exact(Wf), **kw)
    return orig_gemm(A, W, **kw)


class _Exact(torch.Tensor):
    pass


def install_exact_gemm(exact_weights=False, exact_acts=False, only=None):
    """Vision-tower bf16 GEMMs: optionally keep weights / activations exact (fp32, no truncation)."""
    def gemm(A, W, *, bias=None, residual=None, out_f32=None, out_bf16=None, pre_bf16=None, aux_bf16=None, act=abi.ACT_NONE, tile_n=0, round_out=False, conv_hw=None):
        wt = getattr(W, "_exact", None)
        is_bf16_path = W.dtype == torch.bfloat16 or wt is not None
        if not is_bf16_path or (only and W.shape not in only):
            if wt is not None:
                W = W
            if A.dtype != W.dtype:
                A = A.to(W.dtype)
            return orig_gemm(A, W, bias=bias, residual=residual, out_f32=out_f32, out_bf16=out_bf16, pre_bf16=pre_bf16, aux_bf16=aux_bf16, act=act, round_out=round_out)
        Wr = wt if (exact_weights and wt is not None) else W.float()
        Ar = A.float() if (exact_acts or A.dtype == torch.bfloat16) else A.to(torch.bfloat16).float()
        v = Ar @ Wr.t()
        if bias is not None:
            v = v + bias
        if pre_bf16 is not None:
            pre_bf16.copy_(v)
        if act == abi.ACT_QGELU:
            v = v * torch.sigmoid(1.702 * v)
        if residual is not None:
            v = v + residual
        for o in (out_f32, out_bf16):
            if o is not None:
                o.copy_(v)
    abi.gemm = gemm

    def _bf(w):
        t = w.detach().to(torch.bfloat16).contiguous()
        t._exact = w.detach().float().contiguous()
        return t
    engine._bf = _bf

    def undo():
        abi.gemm = orig_gemm
        engine._bf = orig_bf
    return undo


M_rows = None

run("baseline (device emulation)")
run("exact vision weights", lambda: install_exact_gemm(exact_weights=True))


def s_acts():
    u1 = install_exact_gemm(exact_acts=True)
    u2 = patch_alloc(lambda shape: True)

    def attn(qkv, B_, S, H, hd, causal, key_mask, out, lse, out_f32=None):
        o, l = fake_abi._attn_ref(qkv.float(), B_, S, H, hd, causal, key_mask)
        out.copy_(o); lse.copy_(l)
        if out_f32 is not None:
            out_f32.copy_(fake_abi._tf32_rn(o))
    abi.attn_fwd = attn
    return lambda: (u1(), u2(), setattr(abi, "attn_fwd", orig_attn))


run("exact activations everywhere (bf16 buffers -> fp32)", s_acts)


def s_both():
    u1 = install_exact_gemm(exact_acts=True, exact_weights=True)
    u2 = patch_alloc(lambda shape: True)

    def attn(qkv, B_, S, H, hd, causal, key_mask, out, lse, out_f32=None):
        o, l = fake_abi._attn_ref(qkv.float(), B_, S, H, hd, causal, key_mask)
        out.copy_(o); lse.copy_(l)
        if out_f32 is not None:
            out_f32.copy_(fake_abi._tf32_rn(o))
    abi.attn_fwd = attn
    return lambda: (u1(), u2(), setattr(abi, "attn_fwd", orig_attn))


run("exact weights + activations (bf16 path exact)", s_both)

# single activation sites (vision tower shapes: qkv [M, 2304], fc1 act [M, 3072], ln / att [M, 768])
for tag, pred in (("qkv buffer fp32", lambda s: s[-1] == 2304), ("fc1 activation fp32", lambda s: s[-1] == 3072),
                  ("LN out + attention out fp32 (D=768 buffers)", lambda s: s[-1] == 768)):
    def s(pred=pred):
        u1 = install_exact_gemm(exact_acts=True)
        u2 = patch_alloc(pred)

        def attn(qkv, B_, S, H, hd, causal, key_mask, out, lse, out_f32=None):
            o, l = fake_abi._attn_ref(qkv.float(), B_, S, H, hd, causal, key_mask)
            out.copy_(o); lse.copy_(l)
            if out_f32 is not None:
                out_f32.copy_(fake_abi._tf32_rn(o))
        abi.attn_fwd = attn
        return lambda: (u1(), u2(), setattr(abi, "attn_fwd", orig_attn))
    run(tag, s)


# ---- proposal: fp16 (11-bit mantissa) instead of bf16 for the forward operands, qkv stays bf16 ----------------------------
def s_f16(all_acts=False):
    F16 = torch.float16

    def gemm(A, W, *, bias=None, residual=None, out_f32=None, out_bf16=None, pre_bf16=None, aux_bf16=None, act=abi.ACT_NONE, tile_n=0, round_out=False, conv_hw=None):
        w16 = getattr(W, "_f16", None)
        if w16 is None:
            if A.dtype != W.dtype:
                A = A.to(W.dtype)
            return orig_gemm(A, W, bias=bias, residual=residual, out_f32=out_f32, out_bf16=out_bf16, pre_bf16=pre_bf16, aux_bf16=aux_bf16, act=act, round_out=round_out)
        v = A.float() @ w16.float().t()
        if bias is not None:
            v = v + bias
        if pre_bf16 is not None:
            pre_bf16.copy_(v)
        if act == abi.ACT_QGELU:
            v = v * torch.sigmoid(1.702 * v)
        if residual is not None:
            v = v + residual
        for o in (out_f32, out_bf16):
            if o is not None:
                o.copy_(v)
    abi.gemm = gemm

    def _bf(w):
        t = w.detach().to(torch.bfloat16).contiguous()
        t._f16 = w.detach().to(F16).contiguous()
        return t
    engine._bf = _bf

    def _e(shape, dtype, like):
        if dtype == torch.bfloat16 and (shape[-1] in (768, 3072, 512) or all_acts):   # LN out / att out / fc1 act (and patch cols = 768)
            return orig_e(shape, F16, like)
        return orig_e(shape, dtype, like)
    engine._e = _e

    def attn(qkv, B_, S, H, hd, causal, key_mask, out, lse, out_f32=None):
        o, l = fake_abi._attn_ref(qkv.float(), B_, S, H, hd, causal, key_mask)
        out.copy_(o); lse.copy_(l)
        if out_f32 is not None:
            out_f32.copy_(fake_abi._tf32_rn(o))
    abi.attn_fwd = attn

    def undo():
        abi.gemm, engine._bf, engine._e, abi.attn_fwd = orig_gemm, orig_bf, orig_e, orig_attn
    return undo


run("fp16 weights + LN/att/fc1-act fp16, qkv bf16", s_f16)
run("fp16 everything incl. qkv", lambda: s_f16(True))
