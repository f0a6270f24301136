"""The B200 execution engine of the CLIPSeg prompt-tuning step.

Everything here is host-side sequencing of the hand-written kernels in ``csrc/`` (through ``abi``): weight packing
(bf16 copies, fused QKV, transposed copies for dgrad), the per-layer forward / dgrad-only backward schedules, and
four ``torch.autograd.Function`` nodes that are the only places where autograd meets the kernels:

    VisionTowerFn   image (+ per-layer visual prompts)        -> the three decoder taps
    TextTowerFn     prompted token embeddings (+ deep prompts) -> conditional embedding (B, 512)
    DecoderFn       taps, conditional embedding, head params   -> logits (B, 1, H, W)
    DiceBceFn       logits, mask                               -> loss (+ integer metric counters, one pass)

The backbone is frozen, so every backward is dgrad only: there are no weight-gradient GEMMs; gradients reach the
prompt tables, the conditional embedding and the small trainable head.  Residual streams and LayerNorm statistics
are fp32, GEMM operands bf16 with fp32 accumulation in TMEM.

Reference arithmetic being replaced (see oracle/clipseg.py for the restatement the tests compare against):
  transformers modeling_clipseg.py:131-212 (vision embeddings), :256-338 (attention), :341-387 (encoder layer),
  :390-437 (decoder layer), :546-626 (decoder); /root/reference/src/models/core_models/coop/
  base_multimodal_clipseg.py:310-484 (prompted vision tower), :24-308 (prompted text tower),
  base_clipseg.py:82-172 and vpt_clipseg.py:237-319 (decoder + additive head).
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import torch

from . import abi

BF16 = torch.bfloat16
F16 = torch.float16
F32 = torch.float32

# Forward operands of the vision tower (LayerNorm / attention / GELU outputs, patch columns and the frozen weights) are IEEE
# fp16, not bf16: same tcgen05 kind::f16 rate, three more significand bits.  Measured on the device-exact CPU emulation
# (tools/precision_attribution.py, VPT full geometry): bf16 weights and the D=768 activations were 90 % of the logit
# error; fp16 takes the rms error from 3.2e-3 to 1.3e-3 (max over B=2: 0.0187 -> 0.0066).  Every such value is far inside
# fp16's range (|x| < 6.5e4); gradients, whose range is the issue, stay bf16.  TVS_F16=0 restores bf16 (A/B switch).
FWD16 = F16 if os.environ.get("TVS_F16", "1") != "0" else BF16


def _phase(name: str):
    """NVTX range around one phase of the step (TVS_NVTX=1; a no-op otherwise)."""
    def deco(fn):
        if not abi.NVTX:
            return fn

        def wrapped(*a, **k):
            with abi.nvtx_range(name):
                return fn(*a, **k)

        wrapped.__name__, wrapped.__doc__ = fn.__name__, fn.__doc__
        return wrapped
    return deco


def _e(shape, dtype, like):
    return torch.empty(shape, dtype=dtype, device=like.device)


def _bf(w: torch.Tensor) -> torch.Tensor:
    return w.detach().to(BF16).contiguous()


def _f(w: torch.Tensor) -> torch.Tensor:
    return w.detach().to(F32).contiguous()


def _h(w: torch.Tensor, dtype) -> torch.Tensor:
    return w.detach().to(dtype).contiguous()


def tf32_rn(w: torch.Tensor) -> torch.Tensor:
    """fp32 rounded to NEAREST tf32 (10-bit mantissa, ties away).  The kind::tf32 MMA truncates its operands, which is
    a systematic relative bias of about -3.4e-4 per GEMM; frozen weights are rounded once here, activations by
    ``rn_act`` / the im2col kernel."""
    w = w.detach().to(F32).contiguous()
    return ((w.view(torch.int32) + 0x1000) & ~0x1FFF).view(F32)


# TVS_FFN: "split" (default) fused decoder FFN with bf16 head + tail operands, "bf16" fused with heads only,
# "gemm" the two kind::tf32 GEMMs that materialise the hidden activation (kept for A/B measurements)
FFN_MODE = os.environ.get("TVS_FFN", "split")


# TVS_HEAD_BWD=kernel: the additive branch's gradient w.r.t. the low-resolution tap map by the stencil kernel
# (head_bwd_addmap_kernel, 128 us: barrier-bound); default: two tf32 GEMMs against the separable tap matrix (~30 us)
HEAD_BWD_GEMM = os.environ.get("TVS_HEAD_BWD", "gemm") != "kernel"


# TVS_FUSE_OVERWRITE=0: deep-prompt overwrite as a separate kernel after every block (A/B switch); default: in the fc2 epilogue
FUSE_OVERWRITE = os.environ.get("TVS_FUSE_OVERWRITE", "1") != "0"


# TVS_TAIL=0: run the bottom block's backward on every row (A/B measurements); default: prompt rows only
TAIL_PRUNE = os.environ.get("TVS_TAIL", "1") != "0"
# TVS_PRE_DGELU=1 (experiment, off): the fc1 epilogue of a 16-bit (vision) block saves QuickGELU'(u) instead of u and the fc2-dgrad
# epilogue only multiplies by it (TVS_GEMM_PRE_DGELU / TVS_ACT_MULAUX).  Measured at M = 15 648: the dgrad GEMM 79.6 -> 77.2 us,
# the fc1 GEMM 80.4 -> 84.9 us, step 9.37 vs 9.37 ms: what separates the dGELU epilogue from the plain one (56.3 us) is the 96 MB
# read of the saved tensor through the LSU (64-byte segments per thread), not the MUFU + 8 operations per element.
PRE_DGELU = os.environ.get("TVS_PRE_DGELU", "0") == "1"


def split_bf16(w: torch.Tensor, head_only: bool = False):
    """w = hi + lo with both parts bf16 (|lo| <= 2^-9 |w|): operands of the three-product MMAs of the fused FFN."""
    w = w.detach().to(F32).contiguous()
    hi = w.to(BF16)
    if head_only:
        return hi, None
    return hi, (w - hi.to(F32)).to(BF16)


def rn_act(a: torch.Tensor) -> torch.Tensor:
    """Round-to-nearest-tf32 copy of a 2-D fp32 GEMM operand (tvs_round_tf32)."""
    out = torch.empty(tuple(a.shape), dtype=F32, device=a.device)
    abi.round_tf32(a, out)
    return out


# ------------------------------------------------------------------------------------------------------------------
# weight packing
# ------------------------------------------------------------------------------------------------------------------
class PackedLayer:
    """bf16 operands of one transformer block (pre- or post-LN), forward and transposed (dgrad) copies."""

    def __init__(self, layer, heads: int, tf32: bool = False, attn32: bool = False, fwd16=BF16):
        """``fwd16``: 16-bit format of the FORWARD operands (weights here; the engine allocates LayerNorm / attention / GELU
        outputs to match): torch.float16 for the vision tower (see FWD16), bf16 otherwise.  dgrad copies (``*_t``) are bf16.
        ``tf32``: also keep fp32 forward operands; the block's forward GEMMs then run kind::tf32 on fp32
        activations (text tower and decoder: <3 % of the FLOPs but most of the logit rounding error).
        ``attn32`` (needs tf32, head dim 64, S <= 80): q / k / v stay fp32 and the attention runs in the short-key fp32
        kernel (tvs_cross_attn_*) - the CRIS text encoder, whose output steers a dynamic convolution."""
        sa, mlp = layer.self_attn, layer.mlp
        D = sa.q_proj.weight.shape[0]
        hd = D // heads
        scale = hd ** -0.5          # exact in bf16 for hd = 64 / 16 (powers of two)
        wq, bq = sa.q_proj.weight.detach() * scale, sa.q_proj.bias.detach() * scale
        wqkv = torch.cat((wq, sa.k_proj.weight.detach(), sa.v_proj.weight.detach()), dim=0)
        self.D, self.F, self.heads, self.hd = D, mlp.fc1.weight.shape[0], heads, hd
        self.act16 = fwd16
        self.wqkv = _h(wqkv, fwd16)                             # [3D, D]
        self.bqkv = _f(torch.cat((bq, sa.k_proj.bias.detach(), sa.v_proj.bias.detach())))
        self.wo, self.bo = _h(sa.out_proj.weight, fwd16), _f(sa.out_proj.bias)
        self.w1, self.b1 = _h(mlp.fc1.weight, fwd16), _f(mlp.fc1.bias)
        self.w2, self.b2 = _h(mlp.fc2.weight, fwd16), _f(mlp.fc2.bias)
        self.wqkv_t = _bf(wqkv.t())                             # [D, 3D]
        self.wo_t = _bf(sa.out_proj.weight.t())
        self.w1_t = _bf(mlp.fc1.weight.t())                     # [D, F]
        self.w2_t = _bf(mlp.fc2.weight.t())                     # [F, D]
        self.g1, self.be1 = _f(layer.layer_norm1.weight), _f(layer.layer_norm1.bias)
        self.g2, self.be2 = _f(layer.layer_norm2.weight), _f(layer.layer_norm2.bias)
        self.tf32, self.attn32 = tf32, attn32 and tf32
        if attn32 and tf32:
            self.wqkv_t32 = _f(wqkv.t())
        if tf32:
            pack = tf32_rn                                # forward tf32 operands are rounded to nearest (the MMA truncates)
            self.wqkv32, self.wo32 = pack(wqkv), pack(sa.out_proj.weight)
            self.w1_32, self.w2_32 = pack(mlp.fc1.weight), pack(mlp.fc2.weight)
            self.wo_t32 = _f(sa.out_proj.weight.t())
            self.w1_t32, self.w2_t32 = _f(mlp.fc1.weight.t()), _f(mlp.fc2.weight.t())
            if D == 64 and self.F % 64 == 0 and self.F <= 4096 and FFN_MODE != "gemm":
                # fused FFN (csrc/ffn_sm100.cu): bf16 head + tail of fc1.weight [F, D] and fc2.weight^T [F, D]
                self.ffn_w1 = split_bf16(mlp.fc1.weight, FFN_MODE == "bf16")
                self.ffn_w2t = split_bf16(mlp.fc2.weight.t(), FFN_MODE == "bf16")


class PackedClipSeg:
    """All frozen operands of a ``CLIPSegForImageSegmentation`` in kernel-ready form (built once per device)."""

    def __init__(self, model):
        cfg = model.config
        vc, tc = cfg.vision_config, cfg.text_config
        vm, tm, dec = model.clip.vision_model, model.clip.text_model, model.decoder
        self.eps = float(vc.layer_norm_eps)
        self.image_size, self.patch = vc.image_size, vc.patch_size
        self.grid = vc.image_size // vc.patch_size
        self.Dv, self.Dt, self.Dr = vc.hidden_size, tc.hidden_size, cfg.reduce_dim
        self.v_heads, self.t_heads, self.d_heads = vc.num_attention_heads, tc.num_attention_heads, cfg.decoder_num_attention_heads
        self.extract_layers = tuple(cfg.extract_layers)
        self.conditional_layer = cfg.conditional_layer
        self.max_pos = tc.max_position_embeddings
        self.eos_token_id = tc.eos_token_id
        for name, hd in (("vision", self.Dv // self.v_heads), ("text", self.Dt // self.t_heads), ("decoder", self.Dr // self.d_heads)):
            if hd not in (16, 64):
                raise abi.TvsError(f"{name} head dim {hd} unsupported by the attention kernels (64 or 16)")
        if cfg.use_complex_transposed_convolution:
            raise abi.TvsError("use_complex_transposed_convolution=True (refined decoder) is outside the supported path")
        # vision embeddings
        emb = vm.embeddings
        self.w_patch = _h(emb.patch_embedding.weight.reshape(self.Dv, -1), FWD16)    # [Dv, 3*P*P]
        self.cls = _f(emb.class_embedding)
        self.pos_v = _f(emb.position_embedding.weight)                               # [G*G+1, Dv]
        self.pre_g, self.pre_b = _f(vm.pre_layrnorm.weight), _f(vm.pre_layrnorm.bias)
        self.post_g, self.post_b = _f(vm.post_layernorm.weight), _f(vm.post_layernorm.bias)
        self.v_layers = [PackedLayer(l, self.v_heads, fwd16=FWD16) for l in vm.encoder.layers]
        self.w_vproj = _bf(model.clip.visual_projection.weight)                      # [proj, Dv]
        # text
        self.t_layers = [PackedLayer(l, self.t_heads, tf32=True) for l in tm.encoder.layers]
        self.fin_g, self.fin_b = _f(tm.final_layer_norm.weight), _f(tm.final_layer_norm.bias)
        self.w_tproj = tf32_rn(model.clip.text_projection.weight)                    # [proj, Dt] f32 (tf32 MMA)
        self.w_tproj_t = _f(model.clip.text_projection.weight.t())
        # decoder
        self.d_layers = [PackedLayer(l, self.d_heads, tf32=True) for l in dec.layers]
        self.w_red = [tf32_rn(r.weight) for r in dec.reduces]                        # [Dr, Dv] f32 (tf32 MMA, rounded to nearest)
        self.b_red = [_f(r.bias) for r in dec.reduces]
        self.w_red_t = [_bf(r.weight.t()) for r in dec.reduces]                      # [Dv, Dr]
        self.w_film = tf32_rn(torch.cat((dec.film_mul.weight, dec.film_add.weight), dim=0))  # [2Dr, proj] f32
        self.b_film = _f(torch.cat((dec.film_mul.bias, dec.film_add.bias)))
        self.w_film_t = _bf(torch.cat((dec.film_mul.weight, dec.film_add.weight), dim=0).t())  # [proj, 2Dr]
        tw = dec.transposed_convolution.weight.detach()                              # [Dr, 1, P, P]
        self.w_tconv = tf32_rn(tw.reshape(self.Dr, -1).t())                          # [P*P, Dr] f32 (tf32 MMA)
        self.w_tconv_t = _bf(tw.reshape(self.Dr, -1))                                # [Dr, P*P]
        self.b_tconv = _f(dec.transposed_convolution.bias)
        self._tap_mats: dict = {}

    def tap_matrix(self, ks: int) -> torch.Tensor:
        if ks not in self._tap_mats:
            self._tap_mats[ks] = head_tap_matrix(self.grid, self.patch, ks, self.cls.device)
        return self._tap_mats[ks]


def head_tap_matrix(G: int, P: int, ks: int, device) -> torch.Tensor:
    """Wm[(t, k), c] = weight with which pixel coordinate c, shifted by tap k and clamped (replicate padding), reads the low-res
    index t through ``Upsample(scale_factor=P, bilinear, align_corners=False)`` (base_clipseg.py:57-71).  The additive branch is
    separable in these weights, logits_add[Y, X] = sum wy(Y; ky, yi) wx(X; kx, xi) addmap[yi, xi, ky, kx], so its backward is two
    small GEMMs with this matrix (DecoderFn.backward).  fp32 arithmetic as in csrc/elementwise.cu ``bilin``; rows padded to a
    multiple of 8.  The weights are multiples of 1 / (2P): exact in tf32."""
    import numpy as np

    f = np.float32
    W, half = G * P, (ks - 1) // 2
    rows = (G * ks + 7) // 8 * 8
    m = np.zeros((rows, W), np.float32)
    for c in range(W):
        for k in range(ks):
            cc = min(max(c + k - half, 0), W - 1)
            src = max(f((f(cc) + f(0.5)) / f(P)) - f(0.5), f(0))
            i0 = min(int(src), G - 1)
            i1 = min(i0 + 1, G - 1)
            w1 = f(src - f(i0))
            m[i0 * ks + k, c] += f(1) - w1
            m[i1 * ks + k, c] += w1
    return torch.from_numpy(m).to(device)


def packed_for(model) -> PackedClipSeg:
    """Cache of PackedClipSeg on the HF module (frozen weights: packed once; call ``repack`` after loading new ones)."""
    dev = next(model.parameters()).device
    pk = getattr(model, "_tvs_packed", None)
    if pk is None or pk[0] != dev:
        abi.require_device()
        with torch.no_grad():
            pk = (dev, PackedClipSeg(model))
        object.__setattr__(model, "_tvs_packed", pk)
    return pk[1]


def repack(model) -> None:
    if hasattr(model, "_tvs_packed"):
        object.__delattr__(model, "_tvs_packed")


# ------------------------------------------------------------------------------------------------------------------
# transformer blocks
# ------------------------------------------------------------------------------------------------------------------
@dataclass
class Saved:
    """Activations one pre-LN block keeps for its dgrad."""
    x: torch.Tensor
    mean1: torch.Tensor
    rstd1: torch.Tensor
    qkv: torch.Tensor
    att: torch.Tensor
    lse: torch.Tensor
    x1: torch.Tensor
    mean2: torch.Tensor
    rstd2: torch.Tensor
    u: torch.Tensor


def encoder_layer_fwd(pk: PackedLayer, x, B, S, causal, key_mask, eps, save: bool, overwrite=None):
    """Pre-LN block (modeling_clipseg.py:357-387).  x: f32 [B*S, D] -> f32 [B*S, D].
    ``overwrite=(ctx, row0, n)``: the deep-prompt re-write of the block output (rows row0 .. row0+n-1 of every sample become
    ``ctx``, base_multimodal_clipseg.py:394-398) - fused into the fc2 epilogue on the 16-bit path, a separate kernel otherwise."""
    M, D, F = B * S, pk.D, pk.F
    hi = pk.tf32                       # fp32 activations + kind::tf32 MMAs for the small, precision-critical towers
    adt = F32 if hi else pk.act16      # 16-bit forward operands: fp16 in the vision tower (FWD16), q / k / v stay bf16
    ln = _e((M, D), adt, x)
    mean1, rstd1 = _e((M,), F32, x), _e((M,), F32, x)
    abi.layernorm_fwd(x, pk.g1, pk.be1, eps, y_f32=ln if hi else None, y_bf16=None if hi else ln, mean=mean1, rstd=rstd1, round_tf32=hi)
    lse = _e((B, pk.heads, S), F32, x)
    # tf32 operands are rounded to nearest by their producers: LayerNorm (flag), the attention kernels (always), the
    # fc1 epilogue (round_out)
    if pk.attn32:
        qkv = _e((M, 3 * D), F32, x)
        abi.gemm(ln, pk.wqkv32, bias=pk.bqkv, out_f32=qkv)
        att = att32 = _e((M, D), F32, x)
        abi.cross_attn_fwd(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], key_mask, B, S, S, pk.heads, pk.hd, att32, lse, causal=causal)
    else:
        qkv = _e((M, 3 * D), BF16, x)
        abi.gemm(ln, pk.wqkv32 if hi else pk.wqkv, bias=pk.bqkv, out_bf16=qkv)
        att = _e((M, D), BF16 if (hi or causal or key_mask is not None or pk.hd != 64) else pk.act16, x)   # fp16 out: tcgen05 kernels only
        att32 = _e((M, D), F32, x) if hi else None
        abi.attn_fwd(qkv, B, S, pk.heads, pk.hd, causal, key_mask, att, lse, out_f32=att32)
    x1 = _e((M, D), F32, x)
    abi.gemm(att32 if hi else att, pk.wo32 if hi else pk.wo, bias=pk.bo, residual=x, out_f32=x1)
    mean2, rstd2 = _e((M,), F32, x), _e((M,), F32, x)
    abi.layernorm_fwd(x1, pk.g2, pk.be2, eps, y_f32=ln if hi else None, y_bf16=None if hi else ln, mean=mean2, rstd=rstd2, round_tf32=hi)
    u = _e((M, F), BF16, x)
    a = _e((M, F), adt, x)
    # (TVS_PRE_DGELU=1: `u` receives QuickGELU'(pre-activation) and the fc2-dgrad epilogue multiplies by it - see PRE_DGELU)
    abi.gemm(ln, pk.w1_32 if hi else pk.w1, bias=pk.b1, pre_bf16=u if save else None, out_f32=a if hi else None,
             out_bf16=None if hi else a, act=abi.ACT_QGELU, round_out=hi, pre_is_grad=(save and not hi and PRE_DGELU))
    x2 = _e((M, D), F32, x)
    fuse = overwrite is not None and not hi and D % 32 == 0 and FUSE_OVERWRITE
    abi.gemm(a, pk.w2_32 if hi else pk.w2, bias=pk.b2, residual=x1, out_f32=x2,
             overwrite=(overwrite[0], S, overwrite[1], overwrite[2]) if fuse else None)
    if overwrite is not None and not fuse:
        abi.prompt_overwrite(x2.view(B, S, D), overwrite[1], overwrite[2], overwrite[0])
    sv = Saved(x, mean1, rstd1, qkv, att, lse, x1, mean2, rstd2, u) if save else None
    return x2, sv


def _tail(t: torch.Tensor, B: int, S: int, n: int) -> torch.Tensor:
    """The last n rows of every sample of a [B*S, ...] tensor as a contiguous [B*n, ...] copy."""
    return t.view(B, S, -1)[:, S - n:].reshape(B * n, -1) if t.dim() > 1 else t.view(B, S)[:, S - n:].reshape(B * n)


def encoder_layer_bwd(pk: PackedLayer, sv: Saved, g, g16, B, S, causal, key_mask, tail_rows: int = 0):
    """dgrad of the pre-LN block.  g / g16: f32 and bf16 copies of d(out) [B*S, D] -> (dx f32, dx bf16).
    tf32 layers (text tower) keep the gradient stream in fp32 and ignore / do not produce the bf16 copy.
    ``tail_rows`` = n > 0 (bottom block of a prompted tower): only the last n rows of every sample still need a gradient
    below this block, so the attention backward runs on the tiles holding them and the QKV dgrad / LayerNorm backward on
    the gathered [B*n, D] rows; returns (dx rows f32 [B*n, D], None)."""
    M, D, F = B * S, pk.D, pk.F
    hi = pk.tf32
    if hi:
        du = _e((M, F), F32, g)
        abi.gemm(g, pk.w2_t32, aux_bf16=sv.u, out_f32=du, act=abi.ACT_DQGELU)
        dln = _e((M, D), F32, g)
        abi.gemm(du, pk.w1_t32, out_f32=dln)
        g1 = _e((M, D), F32, g)
        abi.layernorm_bwd(dln, sv.x1, pk.g2, sv.mean2, sv.rstd2, dx_add=g, dx_f32=g1)
        datt = _e((M, D), F32 if pk.attn32 else BF16, g)
        abi.gemm(g1, pk.wo_t32, out_f32=datt if pk.attn32 else None, out_bf16=None if pk.attn32 else datt)
    else:
        du = _e((M, F), BF16, g)
        abi.gemm(g16, pk.w2_t, aux_bf16=sv.u, out_bf16=du, act=abi.ACT_MULAUX if PRE_DGELU else abi.ACT_DQGELU)
        dln = _e((M, D), BF16, g)
        abi.gemm(du, pk.w1_t, out_bf16=dln)
        g1, g1_16 = _e((M, D), F32, g), _e((M, D), BF16, g)
        abi.layernorm_bwd(dln, sv.x1, pk.g2, sv.mean2, sv.rstd2, dx_add=g, dx_f32=g1, dx_bf16=g1_16)
        datt = dln
        abi.gemm(g1_16, pk.wo_t, out_bf16=datt)
    delta = _e((B, pk.heads, S), F32, g)
    if pk.attn32:
        dqkv = _e((M, 3 * D), F32, g)
        q, k, v = sv.qkv[:, :D], sv.qkv[:, D:2 * D], sv.qkv[:, 2 * D:]
        abi.cross_attn_bwd(q, k, v, key_mask, sv.att, datt, sv.lse, B, S, S, pk.heads, pk.hd, dqkv[:, :D], dqkv[:, D:2 * D],
                           dqkv[:, 2 * D:], delta, causal=causal)
    elif tail_rows and not hi:
        n = tail_rows
        dqkv = _e((M, 3 * D), BF16, g)      # rows below the tile of row S - n stay unwritten and unread
        abi.attn_bwd(sv.qkv, sv.att, datt, sv.lse, B, S, pk.heads, pk.hd, causal, key_mask, delta, dqkv, row_begin=S - n)
        dln1 = _e((B * n, D), BF16, g)
        abi.gemm(_tail(dqkv, B, S, n), pk.wqkv_t, out_bf16=dln1)
        g0 = _e((B * n, D), F32, g)
        abi.layernorm_bwd(dln1, _tail(sv.x, B, S, n), pk.g1, _tail(sv.mean1, B, S, n), _tail(sv.rstd1, B, S, n),
                          dx_add=_tail(g1, B, S, n), dx_f32=g0)
        return g0, None
    else:
        dqkv = _e((M, 3 * D), BF16, g)
        abi.attn_bwd(sv.qkv, sv.att, datt, sv.lse, B, S, pk.heads, pk.hd, causal, key_mask, delta, dqkv)
    g0 = _e((M, D), F32, g)
    if hi:
        dln1 = _e((M, D), F32, g)
        abi.gemm(dqkv, pk.wqkv_t32 if pk.attn32 else pk.wqkv_t, out_f32=dln1)
        abi.layernorm_bwd(dln1, sv.x, pk.g1, sv.mean1, sv.rstd1, dx_add=g1, dx_f32=g0)
        return g0, None
    dln1 = _e((M, D), BF16, g)
    abi.gemm(dqkv, pk.wqkv_t, out_bf16=dln1)
    g0_16 = g1_16
    abi.layernorm_bwd(dln1, sv.x, pk.g1, sv.mean1, sv.rstd1, dx_add=g1, dx_f32=g0, dx_bf16=g0_16)
    return g0, g0_16


@dataclass
class SavedDec:
    qkv: torch.Tensor
    att: torch.Tensor
    lse: torch.Tensor
    s1: torch.Tensor
    mean1: torch.Tensor
    rstd1: torch.Tensor
    a: torch.Tensor
    s2: torch.Tensor
    mean2: torch.Tensor
    rstd2: torch.Tensor


def decoder_layer_fwd(pk: PackedLayer, x, B, S, eps):
    """Post-LN block with ReLU MLP (modeling_clipseg.py:390-437), fp32 activations, kind::tf32 GEMMs.
    Returns (y f32, saved)."""
    M, D, F = B * S, pk.D, pk.F
    qkv = _e((M, 3 * D), BF16, x)
    abi.gemm(rn_act(x), pk.wqkv32, bias=pk.bqkv, out_bf16=qkv)
    att, att32 = _e((M, D), BF16, x), _e((M, D), F32, x)
    lse = _e((B, pk.heads, S), F32, x)
    abi.attn_fwd(qkv, B, S, pk.heads, pk.hd, False, None, att, lse, out_f32=att32)
    s1 = _e((M, D), F32, x)
    abi.gemm(att32, pk.wo32, bias=pk.bo, residual=x, out_f32=s1)       # att32 leaves the attention kernel rounded to tf32
    y1 = _e((M, D), F32, x)
    mean1, rstd1 = _e((M,), F32, x), _e((M,), F32, x)
    abi.layernorm_fwd(s1, pk.g1, pk.be1, eps, y_f32=y1, mean=mean1, rstd=rstd1)
    s2 = _e((M, D), F32, x)
    if hasattr(pk, "ffn_w1"):
        a16 = y1                                                          # the dgrad recomputes the ReLU mask from y1
        abi.ffn64_fwd(y1, pk.ffn_w1, pk.ffn_w2t, pk.b1, pk.b2, s2)
    else:
        a32, a16 = _e((M, F), F32, x), _e((M, F), BF16, x)
        abi.gemm(rn_act(y1), pk.w1_32, bias=pk.b1, out_f32=a32, out_bf16=a16, act=abi.ACT_RELU, round_out=True)
        abi.gemm(a32, pk.w2_32, bias=pk.b2, residual=y1, out_f32=s2)
    y2 = _e((M, D), F32, x)
    mean2, rstd2 = _e((M,), F32, x), _e((M,), F32, x)
    abi.layernorm_fwd(s2, pk.g2, pk.be2, eps, y_f32=y2, mean=mean2, rstd=rstd2)
    return y2, SavedDec(qkv, att, lse, s1, mean1, rstd1, a16, s2, mean2, rstd2)


def decoder_layer_bwd(pk: PackedLayer, sv: SavedDec, g, B, S):
    """dgrad of the post-LN block (fp32 gradient stream, kind::tf32 GEMMs).  g: f32 d(out) -> f32 d(in)."""
    M, D, F = B * S, pk.D, pk.F
    ds2 = _e((M, D), F32, g)
    abi.layernorm_bwd(g, sv.s2, pk.g2, sv.mean2, sv.rstd2, dx_f32=ds2)
    dy1 = _e((M, D), F32, g)
    if hasattr(pk, "ffn_w1"):
        abi.ffn64_bwd(sv.a, ds2, pk.ffn_w1, pk.ffn_w2t, pk.b1, dy1)       # sv.a is y1 here
    else:
        da = _e((M, F), F32, g)
        abi.gemm(ds2, pk.w2_t32, aux_bf16=sv.a, out_f32=da, act=abi.ACT_DRELU)
        abi.gemm(da, pk.w1_t32, residual=ds2, out_f32=dy1)
    ds1 = ds2
    abi.layernorm_bwd(dy1, sv.s1, pk.g1, sv.mean1, sv.rstd1, dx_f32=ds1)
    datt = _e((M, D), BF16, g)
    abi.gemm(ds1, pk.wo_t32, out_bf16=datt)
    dqkv = _e((M, 3 * D), BF16, g)
    delta = _e((B, pk.heads, S), F32, g)
    abi.attn_bwd(sv.qkv, sv.att, datt, sv.lse, B, S, pk.heads, pk.hd, False, None, delta, dqkv)
    dx = dy1
    abi.gemm(dqkv, pk.wqkv_t, residual=ds1, out_f32=dx)
    return dx


# ------------------------------------------------------------------------------------------------------------------
# vision tower
# ------------------------------------------------------------------------------------------------------------------
def _vision_embed(pk: PackedClipSeg, image, ctx0):
    """patch-embed GEMM + CLS/position + prompt rows + pre-LN.  Returns (x0 f32 [B*S,Dv], h_pre, mean, rstd, S)."""
    B = image.shape[0]
    if tuple(image.shape[1:]) != (3, pk.image_size, pk.image_size):
        raise ValueError(f"expected images of shape (B, 3, {pk.image_size}, {pk.image_size}), got {tuple(image.shape)}")
    G2, D = pk.grid * pk.grid, pk.Dv
    n = 0 if ctx0 is None else ctx0.shape[-2]
    S = 1 + G2 + n
    cols = _e((B * G2, 3 * pk.patch * pk.patch), pk.w_patch.dtype, image)
    abi.im2col_patches(image.contiguous(), pk.patch, cols)
    pe = _e((B * G2, D), F32, image)
    abi.gemm(cols, pk.w_patch, out_f32=pe)
    h = _e((B * S, D), F32, image)
    abi.vision_assemble(pe, pk.cls, pk.pos_v, ctx0, B, G2, n, D, h)
    x0 = _e((B * S, D), F32, image)
    mean, rstd = _e((B * S,), F32, image), _e((B * S,), F32, image)
    abi.layernorm_fwd(h, pk.pre_g, pk.pre_b, pk.eps, y_f32=x0, mean=mean, rstd=rstd)
    return x0, h, mean, rstd, S


class VisionTowerFn(torch.autograd.Function):
    """Prompted CLIP vision tower: ctx (depth_used, n, Dv) appended last and re-written after every block
    idx < prompt_depth; runs max(extract_layers)+1 blocks; returns the decoder taps (each (B, S, Dv) f32).

    base_multimodal_clipseg.py:425-484 / :310-423 and vpt_clipseg.py:151-200 / :36-149.
    """

    @staticmethod
    @_phase("vision_tower.fwd")
    def forward(ctx, ctx_vis, image, pk: PackedClipSeg, prompt_depth: int):
        B = image.shape[0]
        n = ctx_vis.shape[1]
        need_grad = ctx_vis.requires_grad
        cv = ctx_vis.detach().to(F32).contiguous()
        x, h_pre, mean_pre, rstd_pre, S = _vision_embed(pk, image, cv[0])
        n_run = max(pk.extract_layers) + 1
        saved, taps = [], {}
        for idx in range(1, n_run + 1):
            x, sv = encoder_layer_fwd(pk.v_layers[idx - 1], x, B, S, False, None, pk.eps, need_grad,
                                      overwrite=(cv[idx], S - n, n) if idx < prompt_depth else None)
            saved.append(sv)
            if (idx - 1) in pk.extract_layers:
                taps[idx - 1] = x
        ctx.pk, ctx.saved, ctx.dims = pk, saved, (B, S, n, prompt_depth, n_run, ctx_vis.shape[0])
        ctx.pre = (h_pre, mean_pre, rstd_pre)
        outs = tuple(taps[i].view(B, S, -1) for i in pk.extract_layers)
        ctx.mark_non_differentiable()
        return outs

    @staticmethod
    @_phase("vision_tower.bwd")
    def backward(ctx, *dtaps):
        pk = ctx.pk
        B, S, n, depth, n_run, depth_alloc = ctx.dims
        D = pk.Dv
        dctx = torch.zeros((depth_alloc, n, D), dtype=F32, device=dtaps[0].device)
        tap_of = {layer + 1: k for k, layer in enumerate(pk.extract_layers)}   # block idx (1-based) -> tap slot
        g = g16 = None
        for idx in range(n_run, 0, -1):
            if idx in tap_of and dtaps[tap_of[idx]] is not None:
                dt = dtaps[tap_of[idx]].contiguous().view(B * S, D)
                if g is None:
                    g = dt.clone()
                    g16 = _e((B * S, D), BF16, g)
                    abi.cast_bf16(g, g16)
                    fresh = False
                else:
                    abi.add_f32(g, dt)
                    fresh = True
            else:
                fresh = False
            if g is None:
                continue
            if idx < depth:
                abi.prompt_grad(g.view(B, S, D), S - n, n, dctx[idx], zero_rows=True, dx_bf16=None if fresh else g16.view(B, S, D))
            if fresh:
                abi.cast_bf16(g, g16)
            tail = n if (idx == 1 and TAIL_PRUNE and n > 0) else 0     # below block 1 only the prompt rows carry a gradient
            g, g16 = encoder_layer_bwd(pk.v_layers[idx - 1], ctx.saved[idx - 1], g, g16, B, S, False, None, tail_rows=tail)
        h_pre, mean_pre, rstd_pre = ctx.pre
        if TAIL_PRUNE and n > 0:                # g holds the [B*n, D] prompt rows only
            dh = _e((B * n, D), F32, g)
            abi.layernorm_bwd(g, _tail(h_pre, B, S, n), pk.pre_g, _tail(mean_pre, B, S, n), _tail(rstd_pre, B, S, n), dx_f32=dh)
            abi.prompt_grad(dh.view(B, n, D), 0, n, dctx[0], zero_rows=False)
        else:
            dh = _e((B * S, D), F32, g)
            abi.layernorm_bwd(g, h_pre, pk.pre_g, mean_pre, rstd_pre, dx_f32=dh)
            abi.prompt_grad(dh.view(B, S, D), S - n, n, dctx[0], zero_rows=False)
        ctx.saved = None
        return dctx, None, None, None


@torch.no_grad()
def vision_tower_stock(pk: PackedClipSeg, image):
    """The stock (un-prompted, no-grad) CLIP vision tower of CoOp / CoCoOp: all blocks, taps + projected CLS feature
    (coop_clipseg.py:341-371)."""
    B = image.shape[0]
    x, _, _, _, S = _vision_embed(pk, image, None)
    taps = {}
    for i, layer in enumerate(pk.v_layers):
        x, _ = encoder_layer_fwd(layer, x, B, S, False, None, pk.eps, False)
        if i in pk.extract_layers:
            taps[i] = x
    cls_rows = _e((B, 1, pk.Dv), F32, x)
    abi.slice_rows(x.view(B, S, -1), 0, 1, y_f32=cls_rows)
    pooled16 = _e((B, pk.Dv), BF16, x)
    abi.layernorm_fwd(cls_rows.view(B, pk.Dv), pk.post_g, pk.post_b, pk.eps, y_bf16=pooled16)
    feats = _e((B, pk.w_vproj.shape[0]), F32, x)
    abi.gemm(pooled16, pk.w_vproj, out_f32=feats)
    return tuple(taps[i].view(B, S, -1) for i in pk.extract_layers), feats


# ------------------------------------------------------------------------------------------------------------------
# text tower
# ------------------------------------------------------------------------------------------------------------------
class TextTowerFn(torch.autograd.Function):
    """CLIP text tower on already-embedded (and prompt-inserted) tokens.

    emb: (B, S, Dt) f32 = token/ctx embeddings + positions; ctx_deep: (depth-1, n, Dt) or (depth-1, B, n, Dt)
    (rows 1..n are re-written after block idx < prompt_depth) or None; key_mask: u8 (B, S) or None;
    pool_pos: int64 (B,) row to pool after the final LayerNorm.  Returns text_projection(pooled) (B, proj).
    base_multimodal_clipseg.py:57-308 / coop_clipseg.py:75-339.
    """

    @staticmethod
    @_phase("text_tower.fwd")
    def forward(ctx, emb, ctx_deep, key_mask, pool_pos, pk: PackedClipSeg, n_ctx: int, defer=None):
        """``defer`` (a list): the tower's kernels are NOT enqueued now - a closure that enqueues them (writing into the returned
        tensor) is appended instead, to be called later on the same stream, before anything reads the result.  This separates
        WHEN the autograd node is created (its sequence number decides which tower's backward autograd enqueues first) from
        WHEN the forward kernels enter the stream / the captured graph (clipseg.BaseMultimodalCLIPSeg.model_forward)."""
        B, S, D = emb.shape
        need_grad = emb.requires_grad or (ctx_deep is not None and ctx_deep.requires_grad)
        emb_d = emb.detach()
        cd = None if ctx_deep is None else ctx_deep.detach().to(F32).contiguous()
        depth = 1 if cd is None else cd.shape[0] + 1
        km = None if key_mask is None else key_mask.to(torch.uint8).contiguous()
        cond = _e((B, pk.w_tproj.shape[0]), F32, emb_d)
        ctx.pk, ctx.km = pk, km
        ctx.dims = (B, S, D, depth, n_ctx, None if cd is None else tuple(cd.shape))

        def run():
            with torch.no_grad():
                x = emb_d.to(F32).contiguous().view(B * S, D).clone()
                saved = []
                for idx in range(1, len(pk.t_layers) + 1):
                    x, sv = encoder_layer_fwd(pk.t_layers[idx - 1], x, B, S, True, km, pk.eps, need_grad,
                                              overwrite=(cd[idx - 1], 1, n_ctx) if idx < depth else None)
                    saved.append(sv)
                xf = _e((B * S, D), F32, x)
                mean_f, rstd_f = _e((B * S,), F32, x), _e((B * S,), F32, x)
                abi.layernorm_fwd(x, pk.fin_g, pk.fin_b, pk.eps, y_f32=xf, mean=mean_f, rstd=rstd_f)
                rows = torch.arange(B, device=x.device) * S + pool_pos.to(x.device)
                pooled = xf.index_select(0, rows)
                abi.gemm(rn_act(pooled), pk.w_tproj, out_f32=cond)
                ctx.saved = saved
                ctx.fin = (x, mean_f, rstd_f, rows)

        if defer is None:
            run()
        else:
            defer.append(run)
        return cond

    @staticmethod
    @_phase("text_tower.bwd")
    def backward(ctx, dcond):
        pk = ctx.pk
        B, S, D, depth, n, cd_shape = ctx.dims
        x_last, mean_f, rstd_f, rows = ctx.fin
        dpooled = _e((B, D), F32, dcond)
        abi.gemm(dcond.contiguous().to(F32), pk.w_tproj_t, out_f32=dpooled)
        dxf = torch.zeros((B * S, D), dtype=F32, device=dcond.device)
        dxf.index_copy_(0, rows, dpooled)
        g = _e((B * S, D), F32, dxf)
        abi.layernorm_bwd(dxf, x_last, pk.fin_g, mean_f, rstd_f, dx_f32=g)
        dctx = None if cd_shape is None else torch.zeros(cd_shape, dtype=F32, device=dcond.device)
        for idx in range(len(pk.t_layers), 0, -1):
            if idx < depth:
                abi.prompt_grad(g.view(B, S, D), 1, n, dctx[idx - 1], zero_rows=True)
            g, _ = encoder_layer_bwd(pk.t_layers[idx - 1], ctx.saved[idx - 1], g, None, B, S, True, ctx.km)
        ctx.saved = None
        return g.view(B, S, D), dctx, None, None, None, None, None


@torch.no_grad()
def text_tower_stock(pk: PackedClipSeg, emb, key_mask, pool_pos):
    return TextTowerFn.apply(emb, None, key_mask, pool_pos, pk, 0)


# ------------------------------------------------------------------------------------------------------------------
# decoder
# ------------------------------------------------------------------------------------------------------------------
class DecoderFn(torch.autograd.Function):
    """FiLM-conditioned CLIPSeg decoder + transposed-conv head + (optional) additive upsample/conv branch.

    inputs : tap0, tap1, tap2 (B,S,Dv) f32 (in extract_layers order), cond (B,proj) f32,
             add_w (1,Dr,k,k) / add_b (1,) / ratio () trainable head parameters (or None), blend mode, n_strip.
    output : logits (B,1,H,W) f32.
    base_clipseg.py:82-172, vpt_clipseg.py:237-319, modeling_clipseg.py:599-626.
    """

    @staticmethod
    @_phase("decoder.fwd")
    def forward(ctx, tap0, tap1, tap2, cond, add_w, add_b, ratio, pk: PackedClipSeg, blend: int, n_strip: int, tconv_w=None, tconv_b=None):
        taps = (tap0, tap1, tap2)
        B, S, Dv = tap0.shape
        Dr, G, P = pk.Dr, pk.grid, pk.patch
        M = B * S
        acts = taps[::-1]
        cond_d = cond.detach().to(F32).contiguous()
        out = None
        saved_layers, film_saved = [], None
        for i, act in enumerate(acts):
            r = _e((M, Dr), F32, act)
            abi.gemm(rn_act(act.detach().contiguous().view(M, Dv)), pk.w_red[i], bias=pk.b_red[i], residual=out, out_f32=r)
            if i == pk.conditional_layer:
                film = _e((B, 2 * Dr), F32, act)
                abi.gemm(rn_act(cond_d), pk.w_film, bias=pk.b_film, out_f32=film)
                mul, add = film[:, :Dr].contiguous(), film[:, Dr:].contiguous()
                y = _e((M, Dr), F32, act)
                abi.film_fwd(r.view(B, S, Dr), mul, add, y.view(B, S, Dr), None)
                film_saved = (r, mul)
                r = y
            out, sv = decoder_layer_fwd(pk.d_layers[i], r, B, S, pk.eps)
            saved_layers.append(sv)
        G2 = G * G
        if 1 + G2 + n_strip != S:
            raise ValueError(f"decoder: sequence {S} != 1 + {G2} patches + {n_strip} prompt rows")
        feat32 = _e((B, G2, Dr), F32, out)
        abi.slice_rows(out.view(B, S, Dr), 1, G2, y_f32=feat32)
        tconv = _e((B * G2, P * P), F32, out)
        feat_rn = rn_act(feat32.view(B * G2, Dr))
        # no_freeze_last_layer (base_clipseg.py:73-80): the transposed convolution trains - read the live weight / bias
        w_tconv, b_tconv, w_tconv_t = pk.w_tconv, pk.b_tconv, pk.w_tconv_t
        if tconv_w is not None:
            tw = tconv_w.detach().to(F32).reshape(Dr, -1)                              # [Dr, P*P]
            w_tconv, w_tconv_t = tf32_rn(tw.t()), _bf(tw)
            b_tconv = tconv_b.detach().to(F32).contiguous()
        abi.gemm(feat_rn, w_tconv, out_f32=tconv)
        H = G * P
        logits = _e((B, 1, H, H), F32, out)
        addmap = add_out = wa16 = ratio_d = add_b_d = None
        ks = 0
        if blend != abi.BLEND_NONE:
            ks = add_w.shape[-1]
            KK = ks * ks
            # contract the channels of the k x k conv at LOW resolution: addmap[., ky*ks+kx] = feat . w[:, ky, kx]
            wa = add_w.detach().to(F32).reshape(Dr, KK).t().contiguous()           # [KK, Dr]
            wa16 = torch.zeros((32 * ((KK + 31) // 32), Dr), dtype=BF16, device=out.device)   # padded bf16 copy for the dgrad
            wa16[:KK] = wa.to(BF16)
            addmap = torch.zeros((B * G2, wa16.shape[0]), dtype=F32, device=out.device)
            abi.gemm(feat_rn, tf32_rn(wa), out_f32=addmap[:, :KK])
            add_out = _e((B, H, H), F32, out) if blend == abi.BLEND_RATIO else None
            ratio_d = None if ratio is None else ratio.detach().to(F32).reshape(1).contiguous()
            add_b_d = add_b.detach().to(F32).contiguous()
            abi.head_fwd(tconv, addmap[:, :KK], b_tconv, add_b_d, ratio_d, blend, B, G, P, ks, logits, add_out)
        else:
            abi.head_fwd(tconv, None, b_tconv, None, None, blend, B, G, P, 1, logits, None)
        ctx.pk, ctx.saved_layers, ctx.film_saved = pk, saved_layers, film_saved
        ctx.head = (tconv, add_out, ratio_d, wa16, feat32, ks)
        ctx.tconv = (b_tconv, w_tconv_t, None if tconv_w is None else tuple(tconv_w.shape), None if tconv_b is None else tuple(tconv_b.shape))
        ctx.dims = (B, S, Dv, blend, n_strip)
        ctx.has = (add_w is not None, add_b is not None, ratio is not None)
        ctx.shapes = (None if add_w is None else add_w.shape, None if add_b is None else add_b.shape,
                      None if ratio is None else ratio.shape)
        return logits

    @staticmethod
    @_phase("decoder.bwd")
    def backward(ctx, dlogits):
        pk = ctx.pk
        B, S, Dv, blend, n_strip = ctx.dims
        Dr, G, P = pk.Dr, pk.grid, pk.patch
        G2, M = G * G, B * S
        tconv, add_out, ratio_d, wa16, feat32, ks = ctx.head
        b_tconv, w_tconv_t, tw_shape, tb_shape = ctx.tconv
        dl = dlogits.contiguous().to(F32)
        dev = dl.device
        dtconv16 = _e((B * G2, P * P), BF16, dl)
        d_add_w = d_add_b = d_ratio = None
        dfeat = _e((B * G2, Dr), F32, dl)
        if blend != abi.BLEND_NONE:
            KK = ks * ks
            daddmap = torch.zeros((B * G2, wa16.shape[0]), dtype=F32, device=dev)
            dba = torch.zeros(1, dtype=F32, device=dev)
            dr_ = torch.zeros(1, dtype=F32, device=dev)
            Wimg = G * P
            if HEAD_BWD_GEMM and Wimg % 4 == 0:
                # daddmap[b, yi, xi, ky, kx] = wb * sum_{Y, X} dl[b, Y, X] wy(Y; ky, yi) wx(X; kx, xi): contract X, then Y, on the
                # tensor cores (kind::tf32; the operands are gradients: the MMA's truncation is far inside their tolerance)
                abi.head_bwd(dl, tconv, add_out, b_tconv, ratio_d, blend, B, G, P, ks, dtconv16, None, dba, dr_)
                wm = pk.tap_matrix(ks)                                   # [R, W], R = G * ks padded to 8
                R = wm.shape[0]
                t1 = _e((R, B * Wimg), F32, dl)
                abi.gemm(wm, dl.view(B * Wimg, Wimg), out_f32=t1)        # t1[(xi, kx), (b, Y)] = sum_X wx dl
                t1p = t1.view(R, B, Wimg).permute(1, 0, 2).contiguous()  # [(b, xi, kx), Y]
                c2 = _e((B * R, R), F32, dl)
                abi.gemm(t1p.view(B * R, Wimg), wm, out_f32=c2)          # c2[(b, xi, kx), (yi, ky)] = sum_Y wy t1
                dam = c2.view(B, R, R)[:, :G * ks, :G * ks].reshape(B, G, ks, G, ks).permute(0, 3, 1, 4, 2).reshape(B * G2, KK)
                daddmap[:, :KK] = dam * ratio_d if blend == abi.BLEND_RATIO else dam
            else:
                abi.head_bwd(dl, tconv, add_out, b_tconv, ratio_d, blend, B, G, P, ks, dtconv16, daddmap[:, :KK], dba, dr_)
            abi.gemm(dtconv16, w_tconv_t, out_f32=dfeat)
            da16 = _e(tuple(daddmap.shape), BF16, dl)
            abi.cast_bf16(daddmap, da16)
            abi.gemm(da16, wa16.t().contiguous(), residual=dfeat, out_f32=dfeat)
            dwa = torch.zeros((KK, Dr), dtype=F32, device=dev)
            abi.wgrad_small(daddmap[:, :KK], feat32.view(B * G2, Dr), dwa)
            d_add_w = dwa.t().reshape(ctx.shapes[0]) if ctx.has[0] else None
            d_add_b = dba.reshape(ctx.shapes[1]) if ctx.has[1] else None
            d_ratio = dr_.reshape(ctx.shapes[2]) if (ctx.has[2] and blend == abi.BLEND_RATIO) else None
        else:
            abi.head_bwd(dl, None, None, b_tconv, None, blend, B, G, P, 1, dtconv16, None, None, None)
            abi.gemm(dtconv16, w_tconv_t, out_f32=dfeat)
        d_tconv_w = d_tconv_b = None
        if tw_shape is not None:
            # weight gradient of the transposed convolution: dW[c, p] = sum_m feat[m, c] * dtconv[m, p] - one TN GEMM on the
            # transposed copies (contraction over the B*G*G patches), bf16 operands like every other gradient GEMM; the bias is
            # added to every pixel, so its gradient is the sum of dlogits (the blend scales both inside head_bwd's dtconv, and
            # d(logits)/d(bias) = (1 - ratio) under the ratio blend, 1 otherwise)
            Mp = B * G2
            Kp = (Mp + 7) // 8 * 8                                   # K of the GEMM: rows padded with zeros to a multiple of 8
            ft = torch.zeros((Dr, Kp), dtype=BF16, device=dev)
            ft[:, :Mp] = feat32.view(Mp, Dr).t()
            dt_t = torch.zeros((P * P, Kp), dtype=BF16, device=dev)
            dt_t[:, :Mp] = dtconv16.t()
            dwt = _e((Dr, P * P), F32, dl)
            abi.gemm(ft, dt_t, out_f32=dwt)
            d_tconv_w = dwt.reshape(tw_shape)
            scale = (1.0 - ratio_d) if (blend == abi.BLEND_RATIO and ratio_d is not None) else 1.0
            d_tconv_b = (dl.sum() * scale).reshape(tb_shape)
        g = _e((M, Dr), F32, dl)
        abi.unslice_rows(dfeat.view(B, G2, Dr), S, 1, g.view(B, S, Dr))
        need_taps = ctx.needs_input_grad[0] or ctx.needs_input_grad[1] or ctx.needs_input_grad[2]
        dtaps = [None, None, None]
        dcond = None
        for i in range(len(ctx.saved_layers) - 1, -1, -1):
            g = decoder_layer_bwd(pk.d_layers[i], ctx.saved_layers[i], g, B, S)
            if i == pk.conditional_layer:
                r, mul = ctx.film_saved
                dr = _e((M, Dr), F32, dl)
                dfilm = _e((B, 2 * Dr), F32, dl)
                dmul, dadd = _e((B, Dr), F32, dl), _e((B, Dr), F32, dl)
                abi.film_bwd(g.view(B, S, Dr), r.view(B, S, Dr), mul, dr.view(B, S, Dr), dmul, dadd)
                dfilm[:, :Dr] = dmul
                dfilm[:, Dr:] = dadd
                dcond = _e((B, pk.w_film_t.shape[0]), F32, dl)
                abi.gemm(dfilm.to(BF16), pk.w_film_t, out_f32=dcond)   # gradients stay on the bf16 path
                g = dr
            if need_taps:
                g16 = _e((M, Dr), BF16, dl)
                abi.cast_bf16(g, g16)
                dt = _e((M, Dv), F32, dl)
                abi.gemm(g16, pk.w_red_t[i], out_f32=dt)
                dtaps[len(dtaps) - 1 - i] = dt.view(B, S, Dv)
        ctx.saved_layers = None
        return dtaps[0], dtaps[1], dtaps[2], dcond, d_add_w, d_add_b, d_ratio, None, None, None, d_tconv_w, d_tconv_b


# ------------------------------------------------------------------------------------------------------------------
# loss + metric counters
# ------------------------------------------------------------------------------------------------------------------
class DiceBceFn(torch.autograd.Function):
    """MONAI DiceCELoss(sigmoid=True) fused with the Dice / IoU integer counters (one pass over logits and mask).

    forward(logits, mask, threshold, lambda_dice, lambda_ce, confmat) -> (loss (), counts int64 (B,3))
    ``confmat`` (int64 [4] = tn, fp, fn, tp with p > thr) is ACCUMULATED in place - it is the IoU metric state.
    """

    @staticmethod
    @_phase("dicebce.fwd")
    def forward(ctx, logits, mask, threshold, lambda_dice, lambda_ce, confmat):
        B = logits.shape[0]
        lg = logits.detach().to(F32).contiguous()
        mk = mask.detach().to(F32).contiguous()
        N = lg.numel() // B
        parts = torch.empty((B, 4), dtype=torch.float64, device=lg.device)
        counts = torch.empty((B, 3), dtype=torch.int64, device=lg.device)
        loss = torch.empty(1, dtype=F32, device=lg.device)
        scratch = torch.empty(abi.dicebce_scratch_bytes(B, N), dtype=torch.uint8, device=lg.device)
        if confmat is None:
            confmat = torch.zeros(4, dtype=torch.int64, device=lg.device)
        abi.dicebce_metrics_fwd(lg, mk, float(threshold), float(lambda_dice), float(lambda_ce), parts, counts, confmat, loss, scratch)
        ctx.save_for_backward(lg, mk, parts)
        ctx.lams = (float(lambda_dice), float(lambda_ce))
        ctx.shape = logits.shape
        ctx.mark_non_differentiable(counts)
        return loss.reshape(()), counts

    @staticmethod
    @_phase("dicebce.bwd")
    def backward(ctx, dloss, _dcounts):
        lg, mk, parts = ctx.saved_tensors
        dl = torch.empty_like(lg)
        gs = dloss.detach().to(F32).reshape(1).contiguous()
        abi.dicebce_bwd(lg, mk, parts, gs, ctx.lams[0], ctx.lams[1], dl)
        return dl.view(ctx.shape), None, None, None, None, None
