"""Oracle restatement of the CLIPSeg prompt-tuning forward (fp32, CPU, plain torch).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

``weights`` is a plain ``dict[str, Tensor]`` with the key names of
``transformers.CLIPSegForImageSegmentation.state_dict()`` (``clip.vision_model...``,
``clip.text_model...``, ``decoder...``), so a real checkpoint, a random-init HF
model or a fixture can be used unchanged.  ``head`` holds the reference wrapper's
own trainable tail: ``additive_decoder_layer.1.weight``/``.bias`` and
``residual_ratio`` (base_clipseg.py:58-72), or is ``None``.

Files followed
  third-party arithmetic (site-packages/transformers/models/clipseg/modeling_clipseg.py, v5.5.0):
    :131-212 vision embeddings   :215-253 text embeddings   :256-276 attention (fp32 softmax)
    :341-354 MLP (quick_gelu)    :357-387 pre-LN encoder layer   :390-437 post-LN decoder layer
    :546-626 decoder (reduces, FiLM, transposed conv)
  reference wrappers (/root/reference/src/models/core_models/coop):
    base_clipseg.py:82-199              decoder_forward, forward
    base_multimodal_clipseg.py:24-629   MaPLe / shared-* : vision first (10 layers, early exit), then text
    vpt_clipseg.py:36-395               VPT: stock text, prompted vision, ``logits += additive``
    coop_clipseg.py:40-484              CoOp / CoCoOp: stock 12-layer vision, pooled image feature, stock decoder
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch
import torch.nn.functional as F

from . import learners as L


@dataclass(frozen=True)
class ClipSegSpec:
    """The numbers of ``CLIPSegConfig`` that change arithmetic (configuration_clipseg.py:47-152)."""
    image_size: int = 352
    patch_size: int = 16
    v_hidden: int = 768
    v_heads: int = 12
    v_layers: int = 12
    v_mlp: int = 3072
    t_hidden: int = 512
    t_heads: int = 8
    t_layers: int = 12
    t_mlp: int = 2048
    vocab_size: int = 49408
    max_position_embeddings: int = 77
    projection_dim: int = 512
    reduce_dim: int = 64
    dec_heads: int = 4
    dec_mlp: int = 2048
    extract_layers: tuple = (3, 6, 9)
    conditional_layer: int = 0
    eos_token_id: int = 49407
    ln_eps: float = 1e-5

    @property
    def grid(self) -> int:
        return self.image_size // self.patch_size

    @classmethod
    def from_hf_config(cls, cfg) -> "ClipSegSpec":
        v, t = cfg.vision_config, cfg.text_config
        return cls(image_size=v.image_size, patch_size=v.patch_size, v_hidden=v.hidden_size,
                   v_heads=v.num_attention_heads, v_layers=v.num_hidden_layers, v_mlp=v.intermediate_size,
                   t_hidden=t.hidden_size, t_heads=t.num_attention_heads, t_layers=t.num_hidden_layers,
                   t_mlp=t.intermediate_size, vocab_size=t.vocab_size,
                   max_position_embeddings=t.max_position_embeddings, projection_dim=cfg.projection_dim,
                   reduce_dim=cfg.reduce_dim, dec_heads=cfg.decoder_num_attention_heads,
                   dec_mlp=cfg.decoder_intermediate_size, extract_layers=tuple(cfg.extract_layers),
                   conditional_layer=cfg.conditional_layer, eos_token_id=t.eos_token_id,
                   ln_eps=v.layer_norm_eps)


# ------------------------------------------------------------------------------------------------
# building blocks
# ------------------------------------------------------------------------------------------------
def quick_gelu(x: torch.Tensor) -> torch.Tensor:
    return x * torch.sigmoid(1.702 * x)


def _ln(w, prefix, x, eps):
    return F.layer_norm(x, (x.shape[-1],), w[f"{prefix}.weight"], w[f"{prefix}.bias"], eps)


def _lin(w, prefix, x):
    return F.linear(x, w[f"{prefix}.weight"], w.get(f"{prefix}.bias"))


def attention(w, prefix, x, heads, mask):
    """modeling_clipseg.py:256-338: softmax(q k^T * d^-0.5 + mask) v, softmax taken in fp32."""
    B, S, D = x.shape
    hd = D // heads

    def split(t):
        return t.view(B, S, heads, hd).transpose(1, 2)

    q, k, v = (split(_lin(w, f"{prefix}.{n}_proj", x)) for n in "qkv")
    s = (q @ k.transpose(-1, -2)) * hd ** -0.5
    if mask is not None:
        s = s + mask
    p = torch.softmax(s, dim=-1, dtype=torch.float32).to(q.dtype)
    o = (p @ v).transpose(1, 2).reshape(B, S, D)
    return _lin(w, f"{prefix}.out_proj", o)


def encoder_layer(w, prefix, x, heads, mask, eps, act=quick_gelu):
    """Pre-LN block (modeling_clipseg.py:357-387)."""
    x = x + attention(w, f"{prefix}.self_attn", _ln(w, f"{prefix}.layer_norm1", x, eps), heads, mask)
    h = _ln(w, f"{prefix}.layer_norm2", x, eps)
    return x + _lin(w, f"{prefix}.mlp.fc2", act(_lin(w, f"{prefix}.mlp.fc1", h)))


def decoder_layer(w, prefix, x, heads, eps):
    """Post-LN block with ReLU MLP (modeling_clipseg.py:390-437, :581-587)."""
    x = _ln(w, f"{prefix}.layer_norm1", x + attention(w, f"{prefix}.self_attn", x, heads, None), eps)
    h = _lin(w, f"{prefix}.mlp.fc2", F.relu(_lin(w, f"{prefix}.mlp.fc1", x)))
    return _ln(w, f"{prefix}.layer_norm2", x + h, eps)


def vision_embeddings(w, spec: ClipSegSpec, pixel_values):
    """modeling_clipseg.py:196-212 with grid == config (no position interpolation)."""
    pre = "clip.vision_model.embeddings"
    B = pixel_values.shape[0]
    if pixel_values.shape[-2:] != (spec.image_size, spec.image_size):
        raise ValueError("oracle covers the native resolution only (grid == position table)")
    pe = F.conv2d(pixel_values, w[f"{pre}.patch_embedding.weight"], stride=spec.patch_size)
    pe = pe.flatten(2).transpose(1, 2)
    cls = w[f"{pre}.class_embedding"].expand(B, 1, -1)
    return torch.cat((cls, pe), dim=1) + w[f"{pre}.position_embedding.weight"].unsqueeze(0)


def text_masks(attention_mask_2d, S, dtype=torch.float32):
    """Causal + padding additive masks (transformers.modeling_attn_mask_utils, ``finfo.min`` fill).

    The 4.x attention the reference was written for adds the two masks one after the other
    (base_multimodal_clipseg.py:205-222); their sum is what reaches the softmax.
    """
    neg = torch.finfo(dtype).min
    causal = torch.full((S, S), neg, dtype=dtype).triu(1)[None, None]
    if attention_mask_2d is None:
        return causal
    pad = torch.zeros(attention_mask_2d.shape[0], 1, 1, S, dtype=dtype)
    pad = pad.masked_fill(attention_mask_2d[:, None, None, :] == 0, neg)
    return causal + pad


# ------------------------------------------------------------------------------------------------
# towers
# ------------------------------------------------------------------------------------------------
def vision_tower_prompted(w, spec: ClipSegSpec, st: L.LearnerState, pixel_values):
    """VPT / MaPLe / shared-*: ctx appended LAST, pre-LN, 10 layers, deep overwrite of the last n rows.

    base_multimodal_clipseg.py:425-484 + :310-423 and vpt_clipseg.py:151-200 + :36-149.
    Returns the three decoder taps = hidden_states[i + 1] for i in extract_layers; every tap includes
    the already-overwritten prompt rows.
    """
    vm = "clip.vision_model"
    n = st.num_context
    h = vision_embeddings(w, spec, pixel_values)
    h = torch.cat((h, L.visual_context(st, 0).expand(h.size(0), -1, -1)), dim=1)
    h = _ln(w, f"{vm}.pre_layrnorm", h, spec.ln_eps)
    states = [h]
    last = max(spec.extract_layers)
    for idx in range(1, spec.v_layers + 1):
        h = encoder_layer(w, f"{vm}.encoder.layers.{idx - 1}", h, spec.v_heads, None, spec.ln_eps)
        if idx < st.prompt_depth:
            h = torch.cat((h[:, :-n], L.visual_context(st, idx).expand(h.size(0), -1, -1)), dim=1)
        states.append(h)
        if idx > last:
            break
    return tuple(states[i + 1] for i in spec.extract_layers)


def vision_tower_stock(w, spec: ClipSegSpec, pixel_values):
    """CoOp / CoCoOp: the stock HF vision model, all layers (coop_clipseg.py:341-371).

    Returns (taps, image_features) with image_features = visual_projection(post_layernorm(CLS)).
    """
    vm = "clip.vision_model"
    h = _ln(w, f"{vm}.pre_layrnorm", vision_embeddings(w, spec, pixel_values), spec.ln_eps)
    states = [h]
    for i in range(spec.v_layers):
        h = encoder_layer(w, f"{vm}.encoder.layers.{i}", h, spec.v_heads, None, spec.ln_eps)
        states.append(h)
    pooled = _ln(w, f"{vm}.post_layernorm", h[:, 0], spec.ln_eps)
    feats = F.linear(pooled, w["clip.visual_projection.weight"])
    return tuple(states[i + 1] for i in spec.extract_layers), feats


def text_tower(w, spec: ClipSegSpec, st: L.LearnerState | None, input_ids, attention_mask, image_features=None):
    """Text branch -> conditional embedding (B, projection_dim).

    ``st is None`` or a purely visual learner = stock HF path (vpt_clipseg.py:348).  Otherwise
    base_multimodal_clipseg.py:24-308 / coop_clipseg.py:40-339: ctx inserted after BOS, position
    embeddings for L+n positions, n ones prepended to the padding mask, rows 1..n overwritten after
    layer idx < prompt_depth (1-based), EOS pooled at ``argmax + n`` clamped to max_position-1.
    """
    tm = "clip.text_model"
    prompted = st is not None and st.is_textual
    n = st.num_context if prompted else 0
    emb = F.embedding(input_ids, w[f"{tm}.embeddings.token_embedding.weight"])
    if prompted:
        emb = L.insert_textual_context(st, emb, spec.max_position_embeddings, image_features)
        if attention_mask is not None:
            attention_mask = L.attention_mask_for_context(st, attention_mask, spec.max_position_embeddings)
    S = emb.shape[1]
    h = emb + w[f"{tm}.embeddings.position_embedding.weight"][:S]
    mask = text_masks(attention_mask, S)
    for idx in range(1, spec.t_layers + 1):
        h = encoder_layer(w, f"{tm}.encoder.layers.{idx - 1}", h, spec.t_heads, mask, spec.ln_eps)
        if prompted and idx < st.prompt_depth:
            ctx = L.textual_context(st, idx, image_features)
            if ctx.ndim == 2:
                ctx = ctx.expand(h.size(0), -1, -1)
            h = torch.cat((h[:, :1], ctx, h[:, n + 1:]), dim=1)
    h = _ln(w, f"{tm}.final_layer_norm", h, spec.ln_eps)
    ids = input_ids.to(torch.int)
    pre = ids if spec.eos_token_id == 2 else (ids == spec.eos_token_id).int()
    pos = pre.argmax(dim=-1) + n
    if prompted:
        pos = torch.clamp(pos, max=spec.max_position_embeddings - 1)
    pooled = h[torch.arange(h.shape[0]), pos]
    return F.linear(pooled, w["clip.text_projection.weight"])


# ------------------------------------------------------------------------------------------------
# decoder
# ------------------------------------------------------------------------------------------------
def additive_layer(head, feat, scale):
    """``nn.Upsample(scale, bilinear)`` then ``Conv2d(reduce, 1, k, padding='same', replicate)`` (base_clipseg.py:58-71)."""
    wgt, b = head["additive_decoder_layer.1.weight"], head["additive_decoder_layer.1.bias"]
    up = F.interpolate(feat, scale_factor=float(scale), mode="bilinear")
    k = wgt.shape[-1]
    lo = (k - 1) // 2
    up = F.pad(up, (lo, k - 1 - lo, lo, k - 1 - lo), mode="replicate")
    return F.conv2d(up, wgt, b)


def decoder(w, spec: ClipSegSpec, taps, cond, n_strip: int, head, blend: str):
    """base_clipseg.py:82-172 / vpt_clipseg.py:237-319 / HF decoder :599-626.

    ``blend``: "none" (stock decoder, CoOp), "ratio" ((1-r)*a + r*b), "add" (a + b, VPT).
    """
    d = "decoder"
    out = None
    for i, act in enumerate(taps[::-1]):
        red = _lin(w, f"{d}.reduces.{i}", act)
        out = red if out is None else red + out
        if i == spec.conditional_layer:
            out = (_lin(w, f"{d}.film_mul", cond) * out.permute(1, 0, 2) + _lin(w, f"{d}.film_add", cond)).permute(1, 0, 2)
        out = decoder_layer(w, f"{d}.layers.{i}", out, spec.dec_heads, spec.ln_eps)
    out = out[:, 1: (-n_strip if n_strip else None)].permute(0, 2, 1)
    B, C, N = out.shape
    g = math.isqrt(N)
    feat = out.reshape(B, C, g, g)
    logits = F.conv_transpose2d(feat, w[f"{d}.transposed_convolution.weight"], w[f"{d}.transposed_convolution.bias"],
                                stride=spec.patch_size)
    if blend != "none" and head is not None:
        add = additive_layer(head, feat, spec.patch_size)
        if blend == "ratio":
            r = head["residual_ratio"]
            logits = (1 - r) * logits + r * add
        elif blend == "add":
            logits = logits + add
        else:
            raise ValueError(blend)
    return logits  # (B, 1, H, W)


# ------------------------------------------------------------------------------------------------
# whole nets: net(text_input, image_input) -> logits (B,1,H,W)
# ------------------------------------------------------------------------------------------------
def net_forward(w, spec: ClipSegSpec, st: L.LearnerState, head, input_ids, attention_mask, pixel_values):
    if st.kind in ("coop", "cocoop"):
        taps, feats = vision_tower_stock(w, spec, pixel_values)
        cond = text_tower(w, spec, st, input_ids, attention_mask, image_features=feats)
        return decoder(w, spec, taps, cond, 0, None, "none")
    if st.kind == "vpt":
        cond = text_tower(w, spec, None, input_ids, attention_mask)
        taps = vision_tower_prompted(w, spec, st, pixel_values)
        return decoder(w, spec, taps, cond, st.num_context, head, "add")
    taps = vision_tower_prompted(w, spec, st, pixel_values)
    cond = text_tower(w, spec, st, input_ids, attention_mask)
    return decoder(w, spec, taps, cond, st.num_context, head, "ratio")


# ------------------------------------------------------------------------------------------------
# random-init weights (for tests / bench; HF _init_weights stds, modeling_clipseg.py:451-494)
# ------------------------------------------------------------------------------------------------
def init_weights(spec: ClipSegSpec, seed: int = 0, perturb: bool = True) -> dict[str, torch.Tensor]:
    """Random weights with HF's initialiser stds.  ``perturb`` makes LayerNorm affine and Linear biases
    non-trivial (HF inits them to 1/0) so that parity tests exercise those terms."""
    g = torch.Generator().manual_seed(seed)

    def nrm(*shape, std):
        return torch.randn(*shape, generator=g) * std

    w: dict[str, torch.Tensor] = {}

    def ln(prefix, d):
        w[f"{prefix}.weight"] = 1 + (nrm(d, std=0.1) if perturb else torch.zeros(d))
        w[f"{prefix}.bias"] = nrm(d, std=0.05) if perturb else torch.zeros(d)

    def lin(prefix, out, inp, std, bias=True):
        w[f"{prefix}.weight"] = nrm(out, inp, std=std)
        if bias:
            w[f"{prefix}.bias"] = nrm(out, std=0.02) if perturb else torch.zeros(out)

    def tower(prefix, layers, d, mlp, act_layers):
        for i in range(layers):
            p = f"{prefix}.{i}"
            in_std = d ** -0.5 * (2 * act_layers) ** -0.5
            for n in "kvq":
                lin(f"{p}.self_attn.{n}_proj", d, d, in_std)
            lin(f"{p}.self_attn.out_proj", d, d, d ** -0.5)
            ln(f"{p}.layer_norm1", d)
            lin(f"{p}.mlp.fc1", mlp, d, (2 * d) ** -0.5)
            lin(f"{p}.mlp.fc2", d, mlp, in_std)
            ln(f"{p}.layer_norm2", d)

    w["clip.logit_scale"] = torch.tensor(2.6592)
    t = "clip.text_model"
    w[f"{t}.embeddings.token_embedding.weight"] = nrm(spec.vocab_size, spec.t_hidden, std=0.02)
    w[f"{t}.embeddings.position_embedding.weight"] = nrm(spec.max_position_embeddings, spec.t_hidden, std=0.02)
    tower(f"{t}.encoder.layers", spec.t_layers, spec.t_hidden, spec.t_mlp, spec.t_layers)
    ln(f"{t}.final_layer_norm", spec.t_hidden)
    v = "clip.vision_model"
    w[f"{v}.embeddings.class_embedding"] = nrm(spec.v_hidden, std=spec.v_hidden ** -0.5)
    w[f"{v}.embeddings.patch_embedding.weight"] = nrm(spec.v_hidden, 3, spec.patch_size, spec.patch_size, std=0.02)
    w[f"{v}.embeddings.position_embedding.weight"] = nrm(spec.grid ** 2 + 1, spec.v_hidden, std=0.02)
    ln(f"{v}.pre_layrnorm", spec.v_hidden)
    tower(f"{v}.encoder.layers", spec.v_layers, spec.v_hidden, spec.v_mlp, spec.v_layers)
    ln(f"{v}.post_layernorm", spec.v_hidden)
    lin("clip.visual_projection", spec.projection_dim, spec.v_hidden, spec.v_hidden ** -0.5, bias=False)
    lin("clip.text_projection", spec.projection_dim, spec.t_hidden, spec.t_hidden ** -0.5, bias=False)
    r = spec.reduce_dim
    lin("decoder.film_mul", r, spec.projection_dim, spec.projection_dim ** -0.5)
    lin("decoder.film_add", r, spec.projection_dim, spec.projection_dim ** -0.5)
    w["decoder.transposed_convolution.weight"] = nrm(r, 1, spec.patch_size, spec.patch_size, std=r ** -0.5)
    w["decoder.transposed_convolution.bias"] = nrm(1, std=0.02)
    for i in range(len(spec.extract_layers)):
        lin(f"decoder.reduces.{i}", r, spec.v_hidden, spec.v_hidden ** -0.5)
    tower("decoder.layers", len(spec.extract_layers), r, spec.dec_mlp, spec.v_layers)
    return w


def init_head(spec: ClipSegSpec, seed: int = 1, kernel_size: int = 5, residual_ratio: float = 0.5):
    g = torch.Generator().manual_seed(seed)
    fan_in = spec.reduce_dim * kernel_size * kernel_size
    return {
        "additive_decoder_layer.1.weight": (torch.rand(1, spec.reduce_dim, kernel_size, kernel_size, generator=g) * 2 - 1) * fan_in ** -0.5,
        "additive_decoder_layer.1.bias": (torch.rand(1, generator=g) * 2 - 1) * fan_in ** -0.5,
        "residual_ratio": torch.tensor(residual_ratio),
    }
