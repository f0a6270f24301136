"""CRIS (CLIP-RN50 + FPN neck + vision-language decoder + dynamic-conv projector) on the sm_100a kernels.

Reference: src/models/components/cris_model/{__init__,clip,layers}.py and src/models/core_models/coop/coop_cris.py.

Layout: every activation is a channels-last matrix ``[B*H*W, C]`` in HBM, so 1x1 convolutions are plain tcgen05 GEMMs
and k x k convolutions are ``im2col`` + GEMM with the (eval-mode) BatchNorm folded into the weight rows and ReLU /
residual in the GEMM epilogue.  The image encoder is frozen and nothing upstream of it needs a gradient, so it runs
forward-only in bf16 with an fp32 residual stream; neck, decoder and projector carry the gradient that flows back to
the text prompts (through ``state`` and the word features) and run kind::tf32 MMAs on fp32 activations, dgrad only
(there is no weight gradient anywhere except the tiny additive layer).

Backward orchestration is left to autograd over a handful of primitives (each a ``torch.autograd.Function`` whose
forward and backward are C-ABI kernel calls); glue that carries no arithmetic of note (channel concatenation, adding
the positional tables, the 13x13 ``f5 * state`` product) stays in torch.
"""
from __future__ import annotations

from types import SimpleNamespace

import torch

from . import abi
from .engine import BF16, F32, PackedLayer, _e, _phase, encoder_layer_bwd, encoder_layer_fwd, rn_act, tf32_rn

BN_EPS = 1e-5
LN_EPS = 1e-5
# Operand type of the frozen CLIP-RN50 GEMMs.  fp32 activations + kind::tf32 MMAs: the 2e-2 logit bar is not reachable
# with bf16 operands here (measured on the emulated path: 0.035-0.046 vs 0.007-0.010), because the pooled image
# feature feeds CoCoOp's meta-net and the text state steers a dynamic convolution.
RN_DTYPE = F32


def _up8(n: int) -> int:
    return (n + 7) // 8 * 8


# ------------------------------------------------------------------------------------------------------------------
# packed frozen operands
# ------------------------------------------------------------------------------------------------------------------
class ConvOp:
    """Conv2d(bias=False) [+ BatchNorm2d (eval)] [+ ReLU] as a GEMM operand: w [Cout, Kp] with K ordered (ky, kx, cin)
    and padded to a multiple of 8, bias f32 [Cout]; ``w_t`` [Kp, Cout] for the dgrad when ``dgrad``."""

    def __init__(self, sd, conv_key, bn_prefix, dtype, dgrad, stride=1, conv_bias=None, relu=True):
        w = sd[conv_key].detach().to(F32)
        cout, cin, k, _ = w.shape
        if bn_prefix is not None:
            a = sd[f"{bn_prefix}.weight"].to(F32) / torch.sqrt(sd[f"{bn_prefix}.running_var"].to(F32) + BN_EPS)
            b = sd[f"{bn_prefix}.bias"].to(F32) - sd[f"{bn_prefix}.running_mean"].to(F32) * a
        else:
            a, b = torch.ones(cout, device=w.device), torch.zeros(cout, device=w.device)
        if conv_bias is not None:
            b = b + a * sd[conv_bias].to(F32)
        w2 = (w * a[:, None, None, None]).permute(0, 2, 3, 1).reshape(cout, k * k * cin)
        self.k, self.stride, self.pad, self.cin, self.cout, self.relu = k, stride, k // 2, cin, cout, relu
        self.K, self.Kp = k * k * cin, _up8(k * k * cin)
        wp = torch.zeros((cout, self.Kp), dtype=F32, device=w.device)
        wp[:, : self.K] = w2
        self.w = tf32_rn(wp) if dtype == F32 else wp.to(dtype).contiguous()     # forward operand: rounded to nearest tf32
        self.bias = b.contiguous()
        self.w_t = wp.t().contiguous() if dgrad else None      # dgrad: tf32 truncation is far inside the gradient tolerance
        # implicit-GEMM path (tvs_gemm_bf16 conv mode): 3x3 / stride 1 with the channel count a multiple of one k-block
        bk = 32 if dtype == F32 else 64
        self.implicit = k == 3 and stride == 1 and cin % bk == 0 and cout % 32 == 0
        self.implicit_dgrad = dgrad and self.implicit and cout % 32 == 0 and cin % 32 == 0
        if self.implicit_dgrad:
            # dx[q] = sum_t dy[q - off(t)] W_t^T: a 3x3 convolution of dy with the taps flipped (off(8 - t) = -off(t))
            w4 = (w * a[:, None, None, None]).permute(0, 2, 3, 1)                              # [Cout, ky, kx, Cin]
            self.w_dg = w4.flip(1, 2).permute(3, 1, 2, 0).reshape(cin, 9 * cout).contiguous()  # [Cin, (ky', kx', Cout)]
            self.w_t = None


class LinearOp:
    """Frozen nn.Linear as GEMM operands.  The output dimension is zero-padded to a multiple of 8 (``n_pad``) so that
    the dgrad GEMM, whose K is that dimension, meets the tensor-map alignment (proj.txt has 256*9 + 1 = 2305 outputs)."""

    def __init__(self, w, b=None, dtype=F32, dgrad=True):
        w = w.detach().to(F32)
        self.n = w.shape[0]
        self.n_pad = _up8(self.n)
        if self.n_pad != self.n:
            w = torch.cat((w, w.new_zeros((self.n_pad - self.n, w.shape[1]))))
            if b is not None:
                b = torch.cat((b.detach().to(F32), w.new_zeros(self.n_pad - self.n)))
        self.w = tf32_rn(w) if dtype == F32 else w.to(dtype).contiguous()
        self.bias = None if b is None else b.detach().to(F32).contiguous()
        self.w_t = w.t().contiguous() if dgrad else None


class LnOp:
    def __init__(self, sd, prefix):
        self.g, self.b = sd[f"{prefix}.weight"].detach().to(F32).contiguous(), sd[f"{prefix}.bias"].detach().to(F32).contiguous()


def _mha_ops(sd, prefix, heads):
    """nn.MultiheadAttention in_proj / out_proj as q (scaled), k, v, o LinearOps (+ fused qk rows for self-attention)."""
    wi, bi = sd[f"{prefix}.in_proj_weight"].detach().to(F32), sd[f"{prefix}.in_proj_bias"].detach().to(F32)
    D = wi.shape[1]
    sc = (D // heads) ** -0.5
    q = LinearOp(wi[:D] * sc, bi[:D] * sc)
    k = LinearOp(wi[D:2 * D], bi[D:2 * D])
    v = LinearOp(wi[2 * D:], bi[2 * D:])
    qk = LinearOp(torch.cat((wi[:D] * sc, wi[D:2 * D])), torch.cat((bi[:D] * sc, bi[D:2 * D])))
    o = LinearOp(sd[f"{prefix}.out_proj.weight"], sd[f"{prefix}.out_proj.bias"])
    return SimpleNamespace(q=q, k=k, v=v, qk=qk, o=o, heads=heads, hd=D // heads, D=D)


def _fma32(a, b, c):
    """fl32(a * b + c) with ONE rounding (x86 / CUDA fma), evaluated exactly with rationals."""
    import numpy as np
    from fractions import Fraction

    exact = Fraction(float(a)) * Fraction(float(b)) + Fraction(float(c))
    guess = np.float32(float(exact))
    cands = [np.nextafter(guess, np.float32(-np.inf)), guess, np.nextafter(guess, np.float32(np.inf))]
    best = min(cands, key=lambda v: (abs(Fraction(float(v)) - exact), int(np.float32(v).view(np.uint32)) & 1))
    return np.float32(best)


def _bicubic_tables(n_in: int, n_out: int, device, align_corners: bool = True, aten_cpu: bool = False):
    """Per-output taps of torch's bicubic (A = -0.75) with border clamping - and their transpose.  align_corners=True:
    src = o (n_in - 1) / (n_out - 1); False: src = (o + 0.5) n_in / n_out - 0.5 (aten upsample_bicubic2d).  The source
    index and the cubic coefficients are evaluated in fp32, operation for operation as aten does.

    ``aten_cpu``: the predict tail's reference is ``TF.resize`` on CPU tensors (Lightning moves predictions to the host,
    src/utils/save_utils.py:74-104).  ATen's x86 CPU kernel (UpSampleKernel.cpp, built with FMA contraction) evaluates
    ``src = fma(scale, o + 0.5, -0.5)``, ``conv1(x) = fl(fl(fma(1.25, x, -2.25) * x) * x) + 1`` and
    ``conv2(x) = fl(fma(fma(-0.75, x, 3.75), x, -6) * x) + 3``; pinned against ``F.interpolate`` itself in
    tests/test_oracle_resize_u8.py, these tables reproduce its weights to the last bit."""
    import numpy as np

    f = np.float32
    A = f(-0.75)
    if align_corners:
        scale = f(n_in - 1) / f(n_out - 1) if n_out > 1 else f(0)
    else:
        scale = f(n_in) / f(n_out)
    idx = torch.zeros((n_out, 4), dtype=torch.int32)
    wt = torch.zeros((n_out, 4), dtype=torch.float32)

    if aten_cpu:
        def c1(x):
            return f(f(f(_fma32(f(1.25), x, f(-2.25)) * x) * x) + f(1))

        def c2(x):
            return f(f(_fma32(_fma32(A, x, f(3.75)), x, f(-6)) * x) + f(3))
    else:
        def c1(x):          # cubic_convolution1
            return ((A + f(2)) * x - (A + f(3))) * x * x + f(1)

        def c2(x):          # cubic_convolution2
            return ((A * x - f(5) * A) * x + f(8) * A) * x - f(4) * A

    for o in range(n_out):
        if align_corners:
            src = scale * f(o)
        elif aten_cpu:
            src = _fma32(scale, f(o) + f(0.5), f(-0.5))
        else:
            src = scale * (f(o) + f(0.5)) - f(0.5)
        fl = int(np.floor(src))
        t = f(src - f(fl))
        t = min(max(t, f(0)), f(1))
        ws = (c2(f(t + f(1))), c1(t), c1(f(f(1) - t)), c2(f(f(f(1) - t) + f(1))))
        for a in range(4):
            idx[o, a] = min(max(fl - 1 + a, 0), n_in - 1)
            wt[o, a] = float(ws[a])
    rev = [[] for _ in range(n_in)]
    for o in range(n_out):
        for a in range(4):
            rev[int(idx[o, a])].append((o, float(wt[o, a])))
    mt = max(len(r) for r in rev)
    t_idx = torch.zeros((n_in, mt), dtype=torch.int32)
    t_w = torch.zeros((n_in, mt), dtype=torch.float32)
    cnt = torch.zeros((n_in,), dtype=torch.int32)
    for i, r in enumerate(rev):
        cnt[i] = len(r)
        for j, (o, wv) in enumerate(r):
            t_idx[i, j], t_w[i, j] = o, wv
    return idx.to(device), wt.to(device), t_idx.to(device), t_w.to(device), cnt.to(device), mt


def resample_tables(hi, wi, ho, wo, device, align_corners: bool = True, aten_cpu: bool = False):
    iy, wy, ty, twy, cy, mty = _bicubic_tables(hi, ho, device, align_corners, aten_cpu)
    ix, wx, tx, twx, cx, mtx = _bicubic_tables(wi, wo, device, align_corners, aten_cpu)
    mt = max(mty, mtx)

    def padto(t):
        out = torch.zeros((t.shape[0], mt), dtype=t.dtype, device=device)
        out[:, : t.shape[1]] = t
        return out.contiguous()

    return dict(iy=iy, wy=wy, ix=ix, wx=wx, ntaps=4, ty=padto(ty), twy=padto(twy), cy=cy, tx=padto(tx), twx=padto(twx), cx=cx, max_taps=mt)


class PackedCris:
    """Kernel-ready frozen operands of a CRIS model, built once per device from its ``state_dict``."""

    def __init__(self, sd, *, image_size, input_resolution, rn_layers, dec_layers, dec_heads):
        dev = sd["backbone.token_embedding.weight"].device
        self.image_size = image_size
        v = "backbone.visual"
        self.width = sd[f"{v}.conv3.weight"].shape[0]
        RN = RN_DTYPE
        self.stem = [ConvOp(sd, f"{v}.conv1.weight", f"{v}.bn1", RN, False, stride=2),
                     ConvOp(sd, f"{v}.conv2.weight", f"{v}.bn2", RN, False),
                     ConvOp(sd, f"{v}.conv3.weight", f"{v}.bn3", RN, False)]
        self.blocks = []
        for li, n in enumerate(rn_layers, start=1):
            for bi in range(n):
                p = f"{v}.layer{li}.{bi}"
                ds = ConvOp(sd, f"{p}.downsample.0.weight", f"{p}.downsample.1", RN, False, relu=False) \
                    if f"{p}.downsample.0.weight" in sd else None
                self.blocks.append(SimpleNamespace(
                    c1=ConvOp(sd, f"{p}.conv1.weight", f"{p}.bn1", RN, False), c2=ConvOp(sd, f"{p}.conv2.weight", f"{p}.bn2", RN, False),
                    c3=ConvOp(sd, f"{p}.conv3.weight", f"{p}.bn3", RN, False, relu=False), ds=ds,
                    stride=2 if (bi == 0 and li > 1) else 1, stage=li, last=(bi == n - 1)))
        a = f"{v}.attnpool"
        ed = sd[f"{a}.q_proj.weight"].shape[0]
        self.rn_heads = self.width * 32 // 64
        sc = (ed // self.rn_heads) ** -0.5
        self.ap_qkv = LinearOp(torch.cat((sd[f"{a}.q_proj.weight"] * sc, sd[f"{a}.k_proj.weight"], sd[f"{a}.v_proj.weight"])),
                               torch.cat((sd[f"{a}.q_proj.bias"] * sc, sd[f"{a}.k_proj.bias"], sd[f"{a}.v_proj.bias"])), dtype=RN, dgrad=False)
        self.ap_c = LinearOp(sd[f"{a}.c_proj.weight"], sd[f"{a}.c_proj.bias"], dtype=RN, dgrad=False)
        self.ap_connect = ConvOp(sd, f"{a}.connect.0.weight", f"{a}.connect.1", RN, False, relu=False)
        self.ap_pos_raw = sd[f"{a}.positional_embedding"].detach().to(F32)
        self.ap_sd = input_resolution // 32
        self._ap_pos = {}
        self.embed_dim = sd[f"{a}.c_proj.weight"].shape[0]

        # text encoder: same pre-LN / QuickGELU block as the CLIPSeg text tower -> reuse its packed layer (tf32)
        self.t_width = sd["backbone.ln_final.weight"].shape[0]
        self.t_heads = self.t_width // 64
        n_t = len({k.split(".")[3] for k in sd if k.startswith("backbone.transformer.resblocks.")})
        self.t_layers = []
        D = self.t_width
        for i in range(n_t):
            b = f"backbone.transformer.resblocks.{i}"
            wi, bi_ = sd[f"{b}.attn.in_proj_weight"], sd[f"{b}.attn.in_proj_bias"]
            lin = lambda w_, b_: SimpleNamespace(weight=w_, bias=b_)   # noqa: E731
            layer = SimpleNamespace(
                self_attn=SimpleNamespace(q_proj=lin(wi[:D], bi_[:D]), k_proj=lin(wi[D:2 * D], bi_[D:2 * D]), v_proj=lin(wi[2 * D:], bi_[2 * D:]),
                                          out_proj=lin(sd[f"{b}.attn.out_proj.weight"], sd[f"{b}.attn.out_proj.bias"])),
                mlp=SimpleNamespace(fc1=lin(sd[f"{b}.mlp.c_fc.weight"], sd[f"{b}.mlp.c_fc.bias"]),
                                    fc2=lin(sd[f"{b}.mlp.c_proj.weight"], sd[f"{b}.mlp.c_proj.bias"])),
                layer_norm1=lin(sd[f"{b}.ln_1.weight"], sd[f"{b}.ln_1.bias"]), layer_norm2=lin(sd[f"{b}.ln_2.weight"], sd[f"{b}.ln_2.bias"]))
            self.t_layers.append(PackedLayer(layer, self.t_heads, tf32=True, attn32=True))
        self.ln_final = LnOp(sd, "backbone.ln_final")
        self.t_proj = LinearOp(sd["backbone.text_projection"].t())           # state = pooled @ text_projection
        if self.t_proj.n != self.t_proj.n_pad:
            raise abi.TvsError(f"CLIP embed_dim {self.t_proj.n} must be a multiple of 8")
        self.pos_t = sd["backbone.positional_embedding"].detach().to(F32)

        # neck (layers.py:359-445); txt_proj = Linear(no bias) + BatchNorm1d + ReLU folded like a conv
        n = "neck"
        a_ = sd[f"{n}.txt_proj.1.weight"].to(F32) / torch.sqrt(sd[f"{n}.txt_proj.1.running_var"].to(F32) + BN_EPS)
        self.txt_proj = LinearOp(sd[f"{n}.txt_proj.0.weight"].to(F32) * a_[:, None],
                                 sd[f"{n}.txt_proj.1.bias"].to(F32) - sd[f"{n}.txt_proj.1.running_mean"].to(F32) * a_)
        cl = lambda name, dgrad=True: ConvOp(sd, f"{n}.{name}.0.weight", f"{n}.{name}.1", F32, dgrad)   # noqa: E731
        self.f1_v_proj, self.f2_v_proj, self.f3_v_proj = cl("f1_v_proj", False), cl("f2_v_proj", False), cl("f3_v_proj", False)
        self.f2_cat, self.f3_cat = cl("f2_cat"), cl("f3_cat")
        self.f4_proj5, self.f4_proj4, self.f4_proj3 = cl("f4_proj5"), cl("f4_proj4"), cl("f4_proj3")
        self.aggr = cl("aggr")
        self.coord0 = ConvOp(sd, f"{n}.coordconv.0.conv1.0.weight", f"{n}.coordconv.0.conv1.1", F32, True)
        self.coord1 = cl("coordconv.1")
        nl_a = sd[f"{n}.norm_layer.0.weight"].to(F32) / torch.sqrt(sd[f"{n}.norm_layer.0.running_var"].to(F32) + BN_EPS)
        self.nl_a = nl_a.contiguous()
        self.nl_b = (sd[f"{n}.norm_layer.0.bias"].to(F32) - sd[f"{n}.norm_layer.0.running_mean"].to(F32) * nl_a).contiguous()

        # decoder (layers.py:124-356)
        self.dec_heads = dec_heads
        self.dec = []
        for i in range(dec_layers):
            p = f"decoder.layers.{i}"
            self.dec.append(SimpleNamespace(
                sa=_mha_ops(sd, f"{p}.self_attn", dec_heads), ca=_mha_ops(sd, f"{p}.multihead_attn", dec_heads),
                norm1=LnOp(sd, f"{p}.norm1"), norm2=LnOp(sd, f"{p}.norm2"), norm3=LnOp(sd, f"{p}.norm3"),
                sa_norm=LnOp(sd, f"{p}.self_attn_norm"), ca_norm=LnOp(sd, f"{p}.cross_attn_norm"),
                ffn0=LinearOp(sd[f"{p}.ffn.0.weight"], sd[f"{p}.ffn.0.bias"]), ffn_ln=LnOp(sd, f"{p}.ffn.3"),
                ffn4=LinearOp(sd[f"{p}.ffn.4.weight"], sd[f"{p}.ffn.4.bias"])))
        self.dec_norm = LnOp(sd, "decoder.norm")
        self._pos = {}

        # projector (layers.py:69-119)
        self.pv1 = ConvOp(sd, "proj.vis.1.0.weight", "proj.vis.1.1", F32, True)
        self.pv3 = ConvOp(sd, "proj.vis.3.0.weight", "proj.vis.3.1", F32, True)
        self.pv4 = ConvOp(sd, "proj.vis.4.weight", None, F32, True, conv_bias="proj.vis.4.bias", relu=False)
        self.p_txt = LinearOp(sd["proj.txt.weight"], sd["proj.txt.bias"])
        self._tables = {}
        self.device = dev

    # ---- cached constant tables -------------------------------------------------------------------------------
    def attnpool_pos(self, H, W):
        key = (H, W)
        if key not in self._ap_pos:
            sd_, C = self.ap_sd, self.ap_pos_raw.shape[1]
            pos = self.ap_pos_raw[-sd_ * sd_:].reshape(1, sd_, sd_, C).permute(0, 3, 1, 2)
            pos = torch.nn.functional.interpolate(pos, size=(H, W), mode="bicubic", align_corners=False)
            self._ap_pos[key] = pos.flatten(2)[0].t().contiguous()          # (HW, C), built once per geometry
        return self._ap_pos[key]

    def positions(self, C, H, W, L):
        """Sine position tables of layers.py:149-236 as (HW, C) and (L, C) f32 (constants of the geometry)."""
        key = (C, H, W, L)
        if key not in self._pos:
            dev = self.device
            pe = torch.zeros(C, H, W, device=dev)
            half = C // 2
            mul = 1e-4 ** (torch.arange(0, half, 2, device=dev, dtype=F32) / half)
            aw = torch.arange(W, device=dev, dtype=F32)[:, None] * mul
            ah = torch.arange(H, device=dev, dtype=F32)[:, None] * mul
            pe[0:half:2] = torch.sin(aw).t()[:, None, :].expand(-1, H, -1)
            pe[1:half:2] = torch.cos(aw).t()[:, None, :].expand(-1, H, -1)
            pe[half::2] = torch.sin(ah).t()[:, :, None].expand(-1, -1, W)
            pe[half + 1::2] = torch.cos(ah).t()[:, :, None].expand(-1, -1, W)
            vpos = pe.reshape(C, H * W).t().contiguous()
            tp = torch.zeros(L, C, device=dev)
            ang = torch.arange(L, device=dev, dtype=F32)[:, None] * 1e-4 ** (torch.arange(0, C, 2, device=dev, dtype=F32) / C)
            tp[:, 0::2], tp[:, 1::2] = torch.sin(ang), torch.cos(ang)
            self._pos[key] = (vpos, tp.contiguous())
        return self._pos[key]

    def coord(self, B, H, W):
        key = ("coord", B, H, W)
        if key not in self._pos:
            yy, xx = torch.meshgrid(torch.linspace(-1, 1, H, device=self.device), torch.linspace(-1, 1, W, device=self.device), indexing="ij")
            self._pos[key] = torch.stack([xx, yy], dim=-1).reshape(1, H * W, 2).expand(B, -1, -1).reshape(B * H * W, 2).contiguous()
        return self._pos[key]

    def tables(self, hi, wi, ho, wo):
        key = (hi, wi, ho, wo)
        if key not in self._tables:
            self._tables[key] = resample_tables(hi, wi, ho, wo, self.device)
        return self._tables[key]


# ------------------------------------------------------------------------------------------------------------------
# forward-only image encoder (clip.py:18-274), bf16 GEMM operands + fp32 residual stream
# ------------------------------------------------------------------------------------------------------------------
def _conv_nograd(op: ConvOp, x, B, H, W, *, residual=None, act=None, x_rounded=False, round_out=False):
    """x: [B*H*W, Cin] in op.w.dtype -> ([B*Ho*Wo, Cout] same dtype, Ho, Wo).  tf32 operands are rounded to nearest:
    by the producer (``x_rounded``), inside im2col, or by an explicit pass; ``round_out`` rounds this conv's output in the
    GEMM epilogue for its consumers."""
    f32 = x.dtype == F32
    if op.implicit:
        xp = _e((B * (H + 2) * (W + 2), op.cin), x.dtype, x)
        abi.pad_nhwc(x, B, H, W, op.cin, xp, round_tf32=f32 and not x_rounded)
        y = _e((B * H * W, op.cout), x.dtype, x)
        abi.gemm(xp, op.w, bias=op.bias, residual=None if residual is None else residual.to(F32), out_f32=y if f32 else None,
                 out_bf16=None if f32 else y, act=(abi.ACT_RELU if op.relu else abi.ACT_NONE) if act is None else act,
                 round_out=round_out and f32, conv_hw=(H, W))
        return y, H, W
    if op.k == 1 and op.stride == 1:
        A, Ho, Wo = (rn_act(x) if (f32 and not x_rounded) else x), H, W
    else:
        Ho, Wo = (H + 2 * op.pad - op.k) // op.stride + 1, (W + 2 * op.pad - op.k) // op.stride + 1
        A = _e((B * Ho * Wo, op.Kp), x.dtype, x)
        abi.im2col_nhwc(x, B, H, W, op.cin, op.k, op.stride, op.pad, A, round_tf32=f32 and not x_rounded)
    y = _e((B * Ho * Wo, op.cout), x.dtype, x)
    if act is None:
        act = abi.ACT_RELU if op.relu else abi.ACT_NONE
    res32 = None if residual is None else residual.to(F32)
    abi.gemm(A, op.w, bias=op.bias, residual=res32, out_f32=y if f32 else None, out_bf16=None if f32 else y, act=act,
             round_out=round_out and f32)
    return y, Ho, Wo


@torch.no_grad()
@_phase("cris.encode_image")
def encode_image(pk: PackedCris, image):
    """image (B,3,H,W) f32 -> (v3, v4, v5) as f32 [B*h*w, C] matrices with their (h, w).

    Every activation of the frozen encoder only ever feeds tf32 GEMMs (and the residual adds), so each producer rounds
    its output to nearest tf32 once (GEMM epilogue / pooling kernel) instead of a separate rounding pass per consumer."""
    B, _, H, W = image.shape
    dt = pk.stem[0].w.dtype
    f32 = dt == F32
    x = image.permute(0, 2, 3, 1).contiguous().to(dt).view(B * H * W, 3)
    for i, op in enumerate(pk.stem):
        x, H, W = _conv_nograd(op, x, B, H, W, x_rounded=i > 0, round_out=True)
    C = pk.stem[-1].cout
    y = _e((B * (H // 2) * (W // 2), C), dt, x)
    abi.avgpool2_nhwc(x, B, H, W, C, y, round_tf32=f32)
    x, H, W = y, H // 2, W // 2
    outs = {}
    for blk in pk.blocks:
        o1, _, _ = _conv_nograd(blk.c1, x, B, H, W, x_rounded=True, round_out=True)
        o2, _, _ = _conv_nograd(blk.c2, o1, B, H, W, x_rounded=True, round_out=True)
        Ho, Wo, xin = H, W, x
        if blk.stride > 1:                      # anti-aliased stride: AvgPool after conv2 and in front of the shortcut conv
            Ho, Wo = H // 2, W // 2
            p2 = _e((B * Ho * Wo, blk.c2.cout), dt, x)
            abi.avgpool2_nhwc(o2, B, H, W, blk.c2.cout, p2, round_tf32=f32)
            o2 = p2
            xin = _e((B * Ho * Wo, blk.c1.cin), dt, x)
            abi.avgpool2_nhwc(x, B, H, W, blk.c1.cin, xin, round_tf32=f32)
        ident = xin if blk.ds is None else _conv_nograd(blk.ds, xin, B, Ho, Wo, x_rounded=True)[0]
        x, _, _ = _conv_nograd(blk.c3, o2, B, Ho, Wo, residual=ident, act=abi.ACT_RES_RELU, x_rounded=True, round_out=True)
        H, W = Ho, Wo
        if blk.last:
            outs[blk.stage] = (x.to(F32), H, W)
    # attention pool (clip.py:78-182): tokens + resized positions -> MHA -> c_proj, + connect(x) residual, ReLU
    x4, H4, W4 = outs[4]
    S, Ce = H4 * W4, x4.shape[1]
    res, _, _ = _conv_nograd(pk.ap_connect, x, B, H4, W4, x_rounded=True)
    t = (x4.view(B, S, Ce) + pk.attnpool_pos(H4, W4)).to(dt).view(B * S, Ce)
    qkv = _e((B * S, 3 * Ce), BF16, t)
    abi.gemm(rn_act(t) if f32 else t, pk.ap_qkv.w, bias=pk.ap_qkv.bias, out_bf16=qkv)
    att, att32 = _e((B * S, Ce), BF16, t), _e((B * S, Ce), F32, t)
    lse = _e((B, pk.rn_heads, S), F32, t)
    abi.attn_fwd(qkv, B, S, pk.rn_heads, Ce // pk.rn_heads, False, None, att, lse, out_f32=att32)
    v5 = _e((B * S, pk.embed_dim), F32, t)
    abi.gemm(att32 if f32 else att, pk.ap_c.w, bias=pk.ap_c.bias, residual=res.to(F32), out_f32=v5, act=abi.ACT_RES_RELU)
    return outs[2], outs[3], (v5, H4, W4)


# ------------------------------------------------------------------------------------------------------------------
# autograd primitives (fp32 activations, tf32 MMAs, dgrad only)
# ------------------------------------------------------------------------------------------------------------------
class ConvFn(torch.autograd.Function):
    """conv_layer of layers.py:14-26 on a channels-last matrix.  Saves only its output (ReLU mask): no wgrad."""

    @staticmethod
    def forward(ctx, x, op: ConvOp, B, H, W):
        x = x.contiguous()
        conv_hw = None
        if op.k == 1:
            A = rn_act(x)
        elif op.implicit:                 # no im2col matrix: zero-bordered copy + shifted TMA loads inside the GEMM
            A = _e((B * (H + 2) * (W + 2), op.cin), F32, x)
            abi.pad_nhwc(x, B, H, W, op.cin, A, round_tf32=True)
            conv_hw = (H, W)
        else:
            A = _e((B * H * W, op.Kp), F32, x)
            abi.im2col_nhwc(x, B, H, W, op.cin, op.k, 1, op.pad, A, round_tf32=True)
        y = _e((B * H * W, op.cout), F32, x)
        abi.gemm(A, op.w, bias=op.bias, out_f32=y, act=abi.ACT_RELU if op.relu else abi.ACT_NONE, conv_hw=conv_hw)
        ctx.op, ctx.geom = op, (B, H, W)
        ctx.save_for_backward(y if op.relu else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        op, (B, H, W) = ctx.op, ctx.geom
        (y,) = ctx.saved_tensors
        dy = dy.contiguous()
        if op.relu:
            dz = _e(tuple(dy.shape), F32, dy)
            abi.relu_mask(dy, y, dz)
        else:
            dz = dy
        if op.k == 1:
            dx = _e((B * H * W, op.cin), F32, dy)
            abi.gemm(dz, op.w_t[: op.cin], out_f32=dx)
        elif op.implicit_dgrad:
            dzp = _e((B * (H + 2) * (W + 2), op.cout), F32, dy)
            abi.pad_nhwc(dz, B, H, W, op.cout, dzp)
            dx = _e((B * H * W, op.cin), F32, dy)
            abi.gemm(dzp, op.w_dg, out_f32=dx, conv_hw=(H, W))
        else:
            dcol = _e((B * H * W, op.Kp), F32, dy)
            abi.gemm(dz, op.w_t, out_f32=dcol)
            cx = op.cin // 4 * 4                      # coordconv: the 2 coordinate channels carry no gradient
            dx = _e((B * H * W, cx), F32, dy)
            abi.col2im_nhwc(dcol, B, H, W, op.cin, cx, op.k, dx)
            if cx != op.cin:
                dx = torch.nn.functional.pad(dx, (0, op.cin - cx))
        return dx, None, None, None, None


def conv(x, op, B, H, W):
    if not x.requires_grad:
        with torch.no_grad():
            return ConvFn.apply(x, op, B, H, W)
    return ConvFn.apply(x, op, B, H, W)


class LinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, op: LinearOp, relu: bool):
        x2 = x.reshape(-1, x.shape[-1]).contiguous()
        y = _e((x2.shape[0], op.n_pad), F32, x)
        abi.gemm(rn_act(x2), op.w, bias=op.bias, out_f32=y, act=abi.ACT_RELU if relu else abi.ACT_NONE)
        ctx.op, ctx.relu, ctx.shape = op, relu, x.shape
        ctx.save_for_backward(y if relu else None)
        return y[:, : op.n].view(*x.shape[:-1], op.n) if op.n == op.n_pad else y[:, : op.n].reshape(*x.shape[:-1], op.n)

    @staticmethod
    def backward(ctx, dy):
        op = ctx.op
        (y,) = ctx.saved_tensors
        dy2 = dy.reshape(-1, dy.shape[-1])
        if op.n != op.n_pad:
            dy2 = torch.nn.functional.pad(dy2, (0, op.n_pad - op.n))
        dy2 = dy2.contiguous()
        if ctx.relu:
            dz = _e(tuple(dy2.shape), F32, dy2)
            abi.relu_mask(dy2, y, dz)
        else:
            dz = dy2
        dx = _e((dy2.shape[0], op.w.shape[1]), F32, dy2)
        abi.gemm(dz, op.w_t, out_f32=dx)
        return dx.view(ctx.shape), None, None


def linear(x, op, relu=False):
    return LinearFn.apply(x, op, relu)


class LayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, op: LnOp):
        x2 = x.reshape(-1, x.shape[-1]).contiguous()
        M = x2.shape[0]
        y, mean, rstd = _e(tuple(x2.shape), F32, x), _e((M,), F32, x), _e((M,), F32, x)
        abi.layernorm_fwd(x2, op.g, op.b, LN_EPS, y_f32=y, mean=mean, rstd=rstd)
        ctx.op = op
        ctx.save_for_backward(x2, mean, rstd)
        return y.view(x.shape)

    @staticmethod
    def backward(ctx, dy):
        x2, mean, rstd = ctx.saved_tensors
        dy2 = dy.reshape(-1, dy.shape[-1]).contiguous()
        dx = _e(tuple(x2.shape), F32, dy2)
        abi.layernorm_bwd(dy2, x2, ctx.op.g, mean, rstd, dx_f32=dx)
        return dx.view(dy.shape), None


def layer_norm(x, op):
    return LayerNormFn.apply(x, op)


class Upsample2xFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, B, H, W):
        x = x.contiguous()
        C = x.shape[1]
        y = _e((B * 4 * H * W, C), F32, x)
        abi.upsample2x_fwd(x, B, H, W, C, y)
        ctx.geom = (B, H, W, C)
        return y

    @staticmethod
    def backward(ctx, dy):
        B, H, W, C = ctx.geom
        dy = dy if dy.stride(1) == 1 else dy.contiguous()
        dx = _e((B * H * W, C), F32, dy)
        abi.upsample2x_bwd(dy, B, H, W, C, dx)
        return dx, None, None, None


def upsample2x(x, B, H, W):
    return Upsample2xFn.apply(x, B, H, W)


class SelfAttnFn(torch.autograd.Function):
    """nn.MultiheadAttention(q = k = x + pos, value = x) of layers.py:333-336, flash attention on the tcgen05 path."""

    @staticmethod
    def forward(ctx, xqk, xv, m, B, S):
        M, D = B * S, m.D
        ctx.shapes = (xqk.shape, xv.shape)
        xqk, xv = rn_act(xqk.reshape(M, D).contiguous()), rn_act(xv.reshape(M, D).contiguous())
        qkv = _e((M, 3 * D), BF16, xqk)
        abi.gemm(xqk, m.qk.w, bias=m.qk.bias, out_bf16=qkv[:, : 2 * D])
        abi.gemm(xv, m.v.w, bias=m.v.bias, out_bf16=qkv[:, 2 * D:])
        att, att32 = _e((M, D), BF16, xqk), _e((M, D), F32, xqk)
        lse = _e((B, m.heads, S), F32, xqk)
        abi.attn_fwd(qkv, B, S, m.heads, m.hd, False, None, att, lse, out_f32=att32)
        out = _e((M, D), F32, xqk)
        abi.gemm(att32, m.o.w, bias=m.o.bias, out_f32=out)              # att32 leaves the attention kernel rounded to tf32
        ctx.m, ctx.geom = m, (B, S)
        ctx.save_for_backward(qkv, att, lse)
        return out

    @staticmethod
    def backward(ctx, dout):
        m, (B, S) = ctx.m, ctx.geom
        qkv, att, lse = ctx.saved_tensors
        M, D = B * S, m.D
        datt = _e((M, D), BF16, dout)
        abi.gemm(dout.contiguous(), m.o.w_t, out_bf16=datt)
        dqkv = _e((M, 3 * D), BF16, dout)
        delta = _e((B, m.heads, S), F32, dout)
        abi.attn_bwd(qkv, att, datt, lse, B, S, m.heads, m.hd, False, None, delta, dqkv)
        dq32 = dqkv.to(F32)
        dxqk, dxv = _e((M, D), F32, dout), _e((M, D), F32, dout)
        abi.gemm(dq32[:, : 2 * D], m.qk.w_t, out_f32=dxqk)
        abi.gemm(dq32[:, 2 * D:], m.v.w_t, out_f32=dxv)
        return dxqk.view(ctx.shapes[0]), dxv.view(ctx.shapes[1]), None, None, None


class CrossAttnFn(torch.autograd.Function):
    """multihead_attn(query = vis + pos, key = txt + pos, value = txt, key_padding_mask) of layers.py:341-349."""

    @staticmethod
    def forward(ctx, xq, xk, xv, key_mask, m, B, Sq, Sk):
        D = m.D
        ctx.shapes = (xq.shape, xk.shape, xv.shape)
        xq, xk, xv = (rn_act(t.reshape(-1, D).contiguous()) for t in (xq, xk, xv))
        q, kv = _e((B * Sq, D), F32, xq), _e((B * Sk, 2 * D), F32, xq)
        abi.gemm(xq, m.q.w, bias=m.q.bias, out_f32=q)
        abi.gemm(xk, m.k.w, bias=m.k.bias, out_f32=kv[:, :D])
        abi.gemm(xv, m.v.w, bias=m.v.bias, out_f32=kv[:, D:])
        att = _e((B * Sq, D), F32, xq)
        lse = _e((B, m.heads, Sq), F32, xq)
        abi.cross_attn_fwd(q, kv[:, :D], kv[:, D:], key_mask, B, Sq, Sk, m.heads, m.hd, att, lse)
        out = _e((B * Sq, D), F32, xq)
        abi.gemm(att, m.o.w, bias=m.o.bias, out_f32=out)                # rounded to tf32 by the attention kernel
        ctx.m, ctx.geom, ctx.key_mask = m, (B, Sq, Sk), key_mask
        ctx.save_for_backward(q, kv, att, lse)
        return out

    @staticmethod
    def backward(ctx, dout):
        m, (B, Sq, Sk) = ctx.m, ctx.geom
        q, kv, att, lse = ctx.saved_tensors
        D = m.D
        datt = _e((B * Sq, D), F32, dout)
        abi.gemm(dout.contiguous(), m.o.w_t, out_f32=datt)
        dq, dkv = _e((B * Sq, D), F32, dout), _e((B * Sk, 2 * D), F32, dout)
        delta = _e((B, m.heads, Sq), F32, dout)
        abi.cross_attn_bwd(q, kv[:, :D], kv[:, D:], ctx.key_mask, att, datt, lse, B, Sq, Sk, m.heads, m.hd, dq, dkv[:, :D], dkv[:, D:], delta)
        dxq, dxk, dxv = _e((B * Sq, D), F32, dout), _e((B * Sk, D), F32, dout), _e((B * Sk, D), F32, dout)
        abi.gemm(dq, m.q.w_t, out_f32=dxq)
        abi.gemm(dkv[:, :D], m.k.w_t, out_f32=dxk)
        abi.gemm(dkv[:, D:], m.v.w_t, out_f32=dxv)
        return dxq.view(ctx.shapes[0]), dxk.view(ctx.shapes[1]), dxv.view(ctx.shapes[2]), None, None, None, None, None


class DynConvFn(torch.autograd.Function):
    """Projector tail (layers.py:104-119): per-sample 3x3 conv whose weights / bias are ``word``."""

    @staticmethod
    def forward(ctx, x, word, B, H, W):
        x, word = x.contiguous(), word.contiguous()
        C = x.shape[1]
        taps = _e((B * H * W, 9), F32, x)
        out = _e((B, 1, H, W), F32, x)
        abi.dynconv_fwd(x, word, word[:, C * 9:], B, H, W, C, taps, out)
        ctx.geom = (B, H, W, C)
        ctx.save_for_backward(x, word)
        return out

    @staticmethod
    def backward(ctx, dout):
        B, H, W, C = ctx.geom
        x, word = ctx.saved_tensors
        dout = dout.contiguous()
        dx = _e((B * H * W, C), F32, dout)
        chunks = max(1, min(64, (H * W + 255) // 256))
        part = _e((chunks, B, C * 9), F32, dout)
        abi.dynconv_bwd(dout, x, word, B, H, W, C, dx, part)
        dword = torch.cat((part.sum(0), dout.reshape(B, -1).sum(1, keepdim=True)), dim=1)
        return dx, dword, None, None, None


class TailFn(torch.autograd.Function):
    """coop_cris.py:235-242: bicubic (align_corners) upsampling of the prediction, the additive layer
    (Conv2d 1x1 no bias -> bilinear Upsample(size=img) -> Conv2d k x k replicate) on ``fq`` and the blend.
    The k x k conv is contracted with the channels at LOW resolution (see tvs_head_fwd); ``w0`` (mid, C) and ``w2``
    (1, mid, k, k) are trainable, so their (tiny) weight gradients are produced here."""

    @staticmethod
    @_phase("cris.tail.fwd")
    def forward(ctx, pred, fq, w0, w2, b2, ratio, pk: PackedCris, B, h, w, G):
        img = pk.image_size
        P = img // G
        if G * P != img:
            raise abi.TvsError(f"CRIS tail: image size {img} is not a multiple of the feature grid {G}")
        tab = pk.tables(h, w, img, img)
        big = _e((B * G * G, P * P), F32, pred)                       # bicubic map in the head's tiled layout
        abi.resample2d_fwd(pred.contiguous(), B, h, w, img, img, tab, P, big)
        logits = _e((B, 1, img, img), F32, pred)
        zero_b = torch.zeros(1, dtype=F32, device=pred.device)
        if w0 is None:
            abi.head_fwd(big, None, zero_b, None, None, abi.BLEND_NONE, B, G, P, 1, logits, None)
            ctx.blend = False
            ctx.geom = (B, h, w, G, P, 0, 0)
            ctx.pk = pk
            return logits
        mid, C = w0.shape[0], w0.shape[1]
        ks = w2.shape[-1]
        KK = ks * ks
        fq = fq.contiguous()
        w0m = w0.detach().to(F32).reshape(mid, C).contiguous()
        midf = _e((B * G * G, mid), F32, pred)
        abi.gemm(rn_act(fq), tf32_rn(w0m), out_f32=midf)
        wa = w2.detach().to(F32).reshape(mid, KK).t().contiguous()        # [KK, mid]
        addmap = torch.zeros((B * G * G, _up8(KK)), dtype=F32, device=pred.device)
        abi.gemm(rn_act(midf), tf32_rn(wa), out_f32=addmap[:, :KK])
        add_out = _e((B, img, img), F32, pred)
        r = ratio.detach().to(F32).reshape(1).contiguous()
        abi.head_fwd(big, addmap[:, :KK], zero_b, b2.detach().to(F32).contiguous(), r, abi.BLEND_RATIO, B, G, P, ks, logits, add_out)
        ctx.blend = True
        ctx.geom = (B, h, w, G, P, ks, mid)
        ctx.pk = pk
        ctx.shapes = (w0.shape, w2.shape, b2.shape, ratio.shape)
        ctx.save_for_backward(big, add_out, r, fq, w0m, midf, wa, zero_b)
        return logits

    @staticmethod
    @_phase("cris.tail.bwd")
    def backward(ctx, dlogits):
        pk = ctx.pk
        B, h, w, G, P, ks, mid = ctx.geom
        img = pk.image_size
        tab = pk.tables(h, w, img, img)
        dl = dlogits.contiguous().to(F32)
        dev = dl.device
        dbig = _e((B * G * G, P * P), BF16, dl)
        dpred = _e((B, 1, h, w), F32, dl)
        if not ctx.blend:
            zero_b = torch.zeros(1, dtype=F32, device=dev)
            abi.head_bwd(dl, None, None, zero_b, None, abi.BLEND_NONE, B, G, P, 1, dbig, None, None, None)
            abi.resample2d_bwd(dbig, B, h, w, img, img, tab, P, dpred)
            return dpred, None, None, None, None, None, None, None, None, None, None
        big, add_out, r, fq, w0m, midf, wa, zero_b = ctx.saved_tensors
        KK = ks * ks
        daddmap = torch.zeros((B * G * G, _up8(KK)), dtype=F32, device=dev)
        dba, dr_ = torch.zeros(1, dtype=F32, device=dev), torch.zeros(1, dtype=F32, device=dev)
        abi.head_bwd(dl, big, add_out, zero_b, r, abi.BLEND_RATIO, B, G, P, ks, dbig, daddmap[:, :KK], dba, dr_)
        abi.resample2d_bwd(dbig, B, h, w, img, img, tab, P, dpred)
        dwa = torch.zeros((KK, mid), dtype=F32, device=dev)
        abi.wgrad_small(daddmap[:, :KK], midf, dwa)
        dmid = _e((B * G * G, mid), F32, dl)
        wa_t = torch.zeros((mid, _up8(KK)), dtype=F32, device=dev)
        wa_t[:, :KK] = wa.t()
        abi.gemm(daddmap, wa_t, out_f32=dmid)
        dfq = _e(tuple(fq.shape), F32, dl)
        abi.gemm(dmid, w0m.t().contiguous(), out_f32=dfq)
        dw0 = (dmid.t() @ fq).reshape(ctx.shapes[0])          # (mid, C) weight gradient of the trainable 1x1 conv: 2 MFLOP
        return (dpred, dfq, dw0, dwa.t().reshape(ctx.shapes[1]), dba.reshape(ctx.shapes[2]), dr_.reshape(ctx.shapes[3]),
                None, None, None, None, None)


# ------------------------------------------------------------------------------------------------------------------
# text encoder with deep prompts (coop_cris.py:115-183)
# ------------------------------------------------------------------------------------------------------------------
class CrisTextFn(torch.autograd.Function):
    """emb (B,S,D): token + ctx embeddings + positions.  ctx_over (depth, n, D) or (depth, B, n, D): rows 1..n are
    re-written with ctx_over[idx] AFTER block idx < depth (0-based - block 0 included).  key_mask u8 (B,S) 1 = attend.
    Returns (ln_final(x) (B,S,D), text_projection(pooled) (B,E))."""

    @staticmethod
    @_phase("cris.text_encoder.fwd")
    def forward(ctx, emb, ctx_over, key_mask, pool_pos, pk: PackedCris, n_ctx: int):
        B, S, D = emb.shape
        x = emb.detach().to(F32).contiguous().view(B * S, D).clone()
        co = ctx_over.detach().to(F32).contiguous()
        depth = co.shape[0]
        saved = []
        for idx, layer in enumerate(pk.t_layers):
            x, sv = encoder_layer_fwd(layer, x, B, S, True, key_mask, LN_EPS, True)
            if idx < depth:
                abi.prompt_overwrite(x.view(B, S, D), 1, n_ctx, co[idx])
            saved.append(sv)
        words = _e((B * S, D), F32, x)
        mean_f, rstd_f = _e((B * S,), F32, x), _e((B * S,), F32, x)
        abi.layernorm_fwd(x, pk.ln_final.g, pk.ln_final.b, LN_EPS, y_f32=words, mean=mean_f, rstd=rstd_f)
        rows = torch.arange(B, device=x.device) * S + pool_pos.to(x.device)
        pooled = words.index_select(0, rows)
        state = _e((B, pk.t_proj.w.shape[0]), F32, x)
        abi.gemm(rn_act(pooled), pk.t_proj.w, out_f32=state)
        ctx.pk, ctx.saved, ctx.km = pk, saved, key_mask
        ctx.fin = (x, mean_f, rstd_f, rows)
        ctx.dims = (B, S, D, depth, n_ctx, tuple(co.shape))
        return words.view(B, S, D), state

    @staticmethod
    @_phase("cris.text_encoder.bwd")
    def backward(ctx, dwords, dstate):
        pk = ctx.pk
        B, S, D, depth, n, co_shape = ctx.dims
        x_last, mean_f, rstd_f, rows = ctx.fin
        dev = x_last.device
        dxf = torch.zeros((B * S, D), dtype=F32, device=dev) if dwords is None else dwords.contiguous().to(F32).view(B * S, D).clone()
        if dstate is not None:
            dpooled = _e((B, D), F32, x_last)
            abi.gemm(dstate.contiguous().to(F32), pk.t_proj.w_t, out_f32=dpooled)
            dxf.index_add_(0, rows, dpooled)
        g = _e((B * S, D), F32, x_last)
        abi.layernorm_bwd(dxf, x_last, pk.ln_final.g, mean_f, rstd_f, dx_f32=g)
        dco = torch.zeros(co_shape, dtype=F32, device=dev)
        for idx in range(len(pk.t_layers) - 1, -1, -1):
            if idx < depth:
                abi.prompt_grad(g.view(B, S, D), 1, n, dco[idx], zero_rows=True)
            g, _ = encoder_layer_bwd(pk.t_layers[idx], ctx.saved[idx], g, None, B, S, True, ctx.km)
        ctx.saved = None
        return g.view(B, S, D), dco, None, None, None, None


# ------------------------------------------------------------------------------------------------------------------
# neck / decoder / projector composition (torch autograd over the primitives above)
# ------------------------------------------------------------------------------------------------------------------
@_phase("cris.fpn")
def fpn(pk: PackedCris, vis, state, B):
    (v3, H3, W3), (v4, H4, W4), (v5, H5, W5) = vis
    s = linear(state, pk.txt_proj, relu=True)                                         # (B, C5)
    f5 = conv(v5, pk.f1_v_proj, B, H5, W5)                                              # no grad
    C5 = f5.shape[1]
    f5 = torch.relu((f5.view(B, H5 * W5, C5) * s[:, None, :]) * pk.nl_a + pk.nl_b).view(B * H5 * W5, C5)
    f4 = conv(v4, pk.f2_v_proj, B, H4, W4)                                              # no grad
    f4 = conv(torch.cat((f4, upsample2x(f5, B, H5, W5)), dim=1), pk.f2_cat, B, H4, W4)
    f3 = conv(v3, pk.f3_v_proj, B, H3, W3)                                              # no grad
    f3p = _e((B * H4 * W4, f3.shape[1]), F32, f3)
    abi.avgpool2_nhwc(f3, B, H3, W3, f3.shape[1], f3p)
    f3 = conv(torch.cat((f3p, f4), dim=1), pk.f3_cat, B, H4, W4)
    fq5 = upsample2x(conv(f5, pk.f4_proj5, B, H5, W5), B, H5, W5)
    fq4 = conv(f4, pk.f4_proj4, B, H4, W4)
    fq3 = conv(f3, pk.f4_proj3, B, H4, W4)
    fq = conv(torch.cat((fq3, fq4, fq5), dim=1), pk.aggr, B, H4, W4)
    fq = conv(torch.cat((fq, pk.coord(B, H4, W4)), dim=1), pk.coord0, B, H4, W4)
    return conv(fq, pk.coord1, B, H4, W4), H4, W4


@_phase("cris.transformer_decoder")
def transformer_decoder(pk: PackedCris, fq, words, key_mask, B, H, W):
    S, C = H * W, fq.shape[1]
    L = words.shape[1]
    vpos, tpos = pk.positions(C, H, W, L)
    vis = fq.view(B, S, C)
    wk = words + tpos
    for lyr in pk.dec:
        v2 = layer_norm(vis, lyr.norm1)
        sa = SelfAttnFn.apply(v2 + vpos, v2, lyr.sa, B, S).view(B, S, C)
        vis = vis + layer_norm(sa, lyr.sa_norm)
        v2 = layer_norm(vis, lyr.norm2)
        ca = CrossAttnFn.apply(v2 + vpos, wk, words, key_mask, lyr.ca, B, S, L).view(B, S, C)
        vis = vis + layer_norm(ca, lyr.ca_norm)
        v2 = layer_norm(vis, lyr.norm3)
        v2 = linear(layer_norm(linear(v2, lyr.ffn0, relu=True), lyr.ffn_ln), lyr.ffn4)
        vis = vis + v2
    return layer_norm(vis, pk.dec_norm).reshape(B * S, C)


@_phase("cris.projector")
def projector(pk: PackedCris, fq, state, B, H, W):
    x = conv(upsample2x(fq, B, H, W), pk.pv1, B, 2 * H, 2 * W)
    x = conv(upsample2x(x, B, 2 * H, 2 * W), pk.pv3, B, 4 * H, 4 * W)
    x = conv(x, pk.pv4, B, 4 * H, 4 * W)
    word = linear(state, pk.p_txt)
    return DynConvFn.apply(x, word, B, 4 * H, 4 * W)                                    # (B, 1, 4H, 4W)


def head_forward(pk: PackedCris, vis, words, state, key_mask, add_w0, add_w2, add_b2, ratio):
    """Everything after the two encoders: neck -> decoder -> projector -> bicubic + additive layer + blend."""
    B = words.shape[0]
    fq, H, W = fpn(pk, vis, state, B)
    fq = transformer_decoder(pk, fq, words, key_mask, B, H, W)
    pred = projector(pk, fq, state, B, H, W)
    return TailFn.apply(pred, fq, add_w0, add_w2, add_b2, ratio, pk, B, 4 * H, 4 * W, H)
