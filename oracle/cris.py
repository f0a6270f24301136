"""Oracle restatement of the CRIS (CLIP-RN50) prompt-tuning forward as pure fp32 torch functions on CPU.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``): imported by ``tests/`` and the golden generator, never by the
product.  Pinned against the reference's own ``COOPCRIS`` class by ``tests/golden/make_golden_cris.py`` (fixtures
``tests/golden/cris_*.npz``, checked in ``tests/test_oracle_golden.py``).

Weights are a plain ``dict[str, Tensor]`` under the reference's own ``state_dict`` key names
(``backbone.visual.*``, ``backbone.transformer.*``, ``neck.*``, ``decoder.*``, ``proj.*``).  Everything is
evaluated with eval-mode semantics (BatchNorm running statistics, dropout off): ``freeze_all`` puts the whole
model in ``eval()`` (coop_cris.py:66-68) and the build pins that mode.

Reference files followed (relative to /root/reference/src/models):
  components/cris_model/clip.py:18-75      Bottleneck (anti-aliased stride: AvgPool after conv2 / before downsample)
  components/cris_model/clip.py:78-182     AttentionPool2d, CRIS variant: no mean token, bicubic-resized positional
                                           embedding, 1x1-conv + BN residual, spatial map kept
  components/cris_model/clip.py:185-274    ModifiedResNet -> (C3, C4, C5)
  components/cris_model/clip.py:291-343    ResidualAttentionBlock / Transformer (causal bool mask + key_padding_mask)
  core_models/coop/coop_cris.py:115-183    encode_text with the learner: ctx insert, rows 1..n re-written after block
                                           idx < prompt_depth with ctx[idx] (0-based: block 0's ctx rows are re-written
                                           with the SAME ctx[0] that was inserted), EOS pooling at argmax + n
  core_models/coop/coop_cris.py:102-113    pad mask with n zeros (= attend) prepended, cut to max_length
  components/cris_model/layers.py:394-445  FPN neck;  :124-356 TransformerDecoder;  :69-119 Projector
  core_models/coop/coop_cris.py:213-242    forward tail: bicubic (align_corners) upsample, additive layer, blend
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.nn.functional as F

from . import learners as L

BN_EPS = 1e-5
LN_EPS = 1e-5


@dataclass(frozen=True)
class CrisSpec:
    image_size: int = 416
    input_resolution: int = 224            # the resolution CLIP-RN50 was trained at: attnpool pos-emb is (res/32)^2 + 1
    rn_layers: tuple = (3, 4, 6, 3)
    rn_width: int = 64
    embed_dim: int = 1024                  # CLIP joint dim = ModifiedResNet.output_dim = word_dim
    t_width: int = 512
    t_layers: int = 12
    context_length: int = 77
    vocab_size: int = 49408
    fpn_out: tuple = (256, 512, 1024)
    dec_layers: int = 3
    dec_heads: int = 8
    dec_ffn: int = 2048
    max_length: int = 77                   # CRIS.max_length (cris_model/__init__.py:21)

    @property
    def fpn_in(self):
        return (self.rn_width * 8, self.rn_width * 16, self.embed_dim)

    @property
    def vis_dim(self) -> int:              # decoder d_model = fq channels = text width
        return self.fpn_out[1]

    @property
    def rn_heads(self) -> int:             # clip.py:430
        return self.rn_width * 32 // 64

    @property
    def t_heads(self) -> int:              # clip.py:621
        return self.t_width // 64

    @property
    def proj_in(self) -> int:              # Projector(word_dim, vis_dim // 2, 3) (cris_model/__init__.py:61)
        return self.vis_dim // 2


# ------------------------------------------------------------------------------------------------------------------
# small pieces
# ------------------------------------------------------------------------------------------------------------------
def _bn(w, prefix, x):
    return F.batch_norm(x, w[f"{prefix}.running_mean"], w[f"{prefix}.running_var"], w[f"{prefix}.weight"], w[f"{prefix}.bias"],
                        False, 0.0, BN_EPS)


def _ln(w, prefix, x):
    return F.layer_norm(x, (x.shape[-1],), w[f"{prefix}.weight"], w[f"{prefix}.bias"], LN_EPS)


def _conv_bn_relu(w, prefix, x, padding):
    """layers.py:14-26 ``conv_layer``: Conv2d(bias=False) -> BatchNorm2d -> ReLU, children 0 / 1."""
    return F.relu(_bn(w, f"{prefix}.1", F.conv2d(x, w[f"{prefix}.0.weight"], None, 1, padding)))


def mha(w, prefix, q_in, k_in, v_in, heads, attn_mask=None, key_padding_mask=None):
    """``nn.MultiheadAttention`` (batch_first=False semantics restated batch-first): inputs (B, L, D) / (B, S, D);
    ``attn_mask`` bool (L, S) True = blocked; ``key_padding_mask`` bool (B, S) True = padding."""
    D = q_in.shape[-1]
    hd = D // heads
    wi, bi = w[f"{prefix}.in_proj_weight"], w[f"{prefix}.in_proj_bias"]
    q = F.linear(q_in, wi[:D], bi[:D])
    k = F.linear(k_in, wi[D:2 * D], bi[D:2 * D])
    v = F.linear(v_in, wi[2 * D:], bi[2 * D:])
    B, Lq, S = q.shape[0], q.shape[1], k.shape[1]

    def split(t, n):
        return t.reshape(B, n, heads, hd).transpose(1, 2)

    s = (split(q, Lq) * hd ** -0.5) @ split(k, S).transpose(-1, -2)              # (B, H, L, S)
    if attn_mask is not None:
        s = s.masked_fill(attn_mask[None, None], float("-inf"))
    if key_padding_mask is not None:
        s = s.masked_fill(key_padding_mask[:, None, None, :], float("-inf"))
    o = (torch.softmax(s, dim=-1) @ split(v, S)).transpose(1, 2).reshape(B, Lq, D)
    return F.linear(o, w[f"{prefix}.out_proj.weight"], w[f"{prefix}.out_proj.bias"])


def quick_gelu(x):
    return x * torch.sigmoid(1.702 * x)


# ------------------------------------------------------------------------------------------------------------------
# image encoder (clip.py:18-274)
# ------------------------------------------------------------------------------------------------------------------
def bottleneck(w, prefix, x, stride):
    out = F.relu(_bn(w, f"{prefix}.bn1", F.conv2d(x, w[f"{prefix}.conv1.weight"])))
    out = F.relu(_bn(w, f"{prefix}.bn2", F.conv2d(out, w[f"{prefix}.conv2.weight"], padding=1)))
    if stride > 1:
        out = F.avg_pool2d(out, stride)
    out = _bn(w, f"{prefix}.bn3", F.conv2d(out, w[f"{prefix}.conv3.weight"]))
    identity = x
    if f"{prefix}.downsample.0.weight" in w:
        identity = F.avg_pool2d(x, stride) if stride > 1 else x      # nn.AvgPool2d(1) is the identity
        identity = _bn(w, f"{prefix}.downsample.1", F.conv2d(identity, w[f"{prefix}.downsample.0.weight"]))
    return F.relu(out + identity)


def attention_pool(w, spec: CrisSpec, x):
    p = "backbone.visual.attnpool"
    res = _bn(w, f"{p}.connect.1", F.conv2d(x, w[f"{p}.connect.0.weight"]))
    B, C, H, W = x.shape
    sd = spec.input_resolution // 32
    pos = w[f"{p}.positional_embedding"][-sd * sd:].reshape(1, sd, sd, C).permute(0, 3, 1, 2)
    pos = F.interpolate(pos, size=(H, W), mode="bicubic", align_corners=False).flatten(2)          # (1, C, HW)
    t = (x.flatten(2) + pos).transpose(1, 2)                                                     # (B, HW, C)
    ww = {"a.in_proj_weight": torch.cat([w[f"{p}.q_proj.weight"], w[f"{p}.k_proj.weight"], w[f"{p}.v_proj.weight"]]),
          "a.in_proj_bias": torch.cat([w[f"{p}.q_proj.bias"], w[f"{p}.k_proj.bias"], w[f"{p}.v_proj.bias"]]),
          "a.out_proj.weight": w[f"{p}.c_proj.weight"], "a.out_proj.bias": w[f"{p}.c_proj.bias"]}
    # multi_head_attention_forward with separate projections; out_proj maps C -> output_dim, so it cannot share mha()'s
    # square reshape: do the projection by hand
    D, heads = C, spec.rn_heads
    hd = D // heads
    q = F.linear(t, ww["a.in_proj_weight"][:D], ww["a.in_proj_bias"][:D])
    k = F.linear(t, ww["a.in_proj_weight"][D:2 * D], ww["a.in_proj_bias"][D:2 * D])
    v = F.linear(t, ww["a.in_proj_weight"][2 * D:], ww["a.in_proj_bias"][2 * D:])
    S = t.shape[1]

    def split(u):
        return u.reshape(B, S, heads, hd).transpose(1, 2)

    o = torch.softmax((split(q) * hd ** -0.5) @ split(k).transpose(-1, -2), dim=-1) @ split(v)
    o = F.linear(o.transpose(1, 2).reshape(B, S, D), ww["a.out_proj.weight"], ww["a.out_proj.bias"])  # (B, HW, out)
    return F.relu(o.transpose(1, 2).reshape(B, -1, H, W) + res)


def encode_image(w, spec: CrisSpec, image):
    p = "backbone.visual"
    x = F.relu(_bn(w, f"{p}.bn1", F.conv2d(image, w[f"{p}.conv1.weight"], stride=2, padding=1)))
    x = F.relu(_bn(w, f"{p}.bn2", F.conv2d(x, w[f"{p}.conv2.weight"], padding=1)))
    x = F.relu(_bn(w, f"{p}.bn3", F.conv2d(x, w[f"{p}.conv3.weight"], padding=1)))
    x = F.avg_pool2d(x, 2)
    outs = []
    for li, blocks in enumerate(spec.rn_layers, start=1):
        for bi in range(blocks):
            x = bottleneck(w, f"{p}.layer{li}.{bi}", x, 2 if (bi == 0 and li > 1) else 1)
        outs.append(x)
    return outs[1], outs[2], attention_pool(w, spec, outs[3])


# ------------------------------------------------------------------------------------------------------------------
# text encoder with the learner (coop_cris.py:102-183)
# ------------------------------------------------------------------------------------------------------------------
def pad_mask_with_context(st: L.LearnerState, input_ids, attention_mask, max_length):
    pad = ~attention_mask.bool() if attention_mask is not None else input_ids == 0      # cris_model/__init__.py:79-86
    zeros = torch.zeros(pad.shape[0], st.num_context, dtype=pad.dtype)
    return torch.cat((zeros, pad), dim=1)[:, :max_length]


def encode_text(w, spec: CrisSpec, st: L.LearnerState, input_ids, pad_mask, image_features=None):
    p = "backbone"
    x = w[f"{p}.token_embedding.weight"][input_ids]
    x = L.insert_textual_context(st, x, spec.max_length, image_features)
    S = x.shape[1]
    x = x + w[f"{p}.positional_embedding"][:S]
    causal = torch.triu(torch.ones(S, S, dtype=torch.bool), diagonal=1)
    n = st.num_context
    for idx in range(spec.t_layers):
        b = f"{p}.transformer.resblocks.{idx}"
        h = _ln(w, f"{b}.ln_1", x)
        x = x + mha(w, f"{b}.attn", h, h, h, spec.t_heads, causal, pad_mask)
        h = _ln(w, f"{b}.ln_2", x)
        x = x + F.linear(quick_gelu(F.linear(h, w[f"{b}.mlp.c_fc.weight"], w[f"{b}.mlp.c_fc.bias"])),
                         w[f"{b}.mlp.c_proj.weight"], w[f"{b}.mlp.c_proj.bias"])
        if idx < st.prompt_depth:
            ctx = L.textual_context(st, idx, image_features)
            x = torch.cat((x[:, :1], ctx.expand(x.shape[0], -1, -1) if ctx.ndim == 2 else ctx, x[:, 1 + n:]), dim=1)
    x = _ln(w, f"{p}.ln_final", x)
    pool = torch.minimum(input_ids.argmax(dim=-1) + n, torch.tensor(spec.max_length - 1))
    state = x[torch.arange(x.shape[0]), pool] @ w[f"{p}.text_projection"]
    return x, state


# ------------------------------------------------------------------------------------------------------------------
# neck, decoder, projector (layers.py)
# ------------------------------------------------------------------------------------------------------------------
def fpn(w, vis, state):
    v3, v4, v5 = vis
    s = F.relu(_bn(w, "neck.txt_proj.1", F.linear(state, w["neck.txt_proj.0.weight"])))[:, :, None, None]
    f5 = _conv_bn_relu(w, "neck.f1_v_proj", v5, 0)
    f5 = F.relu(_bn(w, "neck.norm_layer.0", f5 * s))
    f4 = _conv_bn_relu(w, "neck.f2_v_proj", v4, 1)
    f4 = _conv_bn_relu(w, "neck.f2_cat", torch.cat([f4, F.interpolate(f5, scale_factor=2, mode="bilinear")], dim=1), 0)
    f3 = F.avg_pool2d(_conv_bn_relu(w, "neck.f3_v_proj", v3, 1), 2, 2)
    f3 = _conv_bn_relu(w, "neck.f3_cat", torch.cat([f3, f4], dim=1), 0)
    fq5 = F.interpolate(_conv_bn_relu(w, "neck.f4_proj5", f5, 1), scale_factor=2, mode="bilinear")
    fq4 = _conv_bn_relu(w, "neck.f4_proj4", f4, 1)
    fq3 = _conv_bn_relu(w, "neck.f4_proj3", f3, 1)
    fq = _conv_bn_relu(w, "neck.aggr", torch.cat([fq3, fq4, fq5], dim=1), 0)
    B, _, H, W = fq.shape
    yy, xx = torch.meshgrid(torch.linspace(-1, 1, H), torch.linspace(-1, 1, W), indexing="ij")
    coord = torch.stack([xx, yy])[None].expand(B, -1, -1, -1)                 # channel order: x then y (layers.py:64)
    fq = _conv_bn_relu(w, "neck.coordconv.0.conv1", torch.cat([fq, coord], dim=1), 1)
    return _conv_bn_relu(w, "neck.coordconv.1", fq, 1)


def pos1d(d_model, length):
    pe = torch.zeros(length, d_model)
    ang = torch.arange(length, dtype=torch.float32)[:, None] * 1e-4 ** (torch.arange(0, d_model, 2, dtype=torch.float32) / d_model)
    pe[:, 0::2], pe[:, 1::2] = torch.sin(ang), torch.cos(ang)
    return pe                                                                  # (L, D)


def pos2d(d_model, height, width):
    pe = torch.zeros(d_model, height, width)
    half = d_model // 2
    mul = 1e-4 ** (torch.arange(0, half, 2, dtype=torch.float32) / half)
    aw = torch.arange(width, dtype=torch.float32)[:, None] * mul               # (W, half/2)
    ah = torch.arange(height, dtype=torch.float32)[:, None] * mul
    pe[0:half:2] = torch.sin(aw).t()[:, None, :].expand(-1, height, -1)
    pe[1:half:2] = torch.cos(aw).t()[:, None, :].expand(-1, height, -1)
    pe[half::2] = torch.sin(ah).t()[:, :, None].expand(-1, -1, width)
    pe[half + 1::2] = torch.cos(ah).t()[:, :, None].expand(-1, -1, width)
    return pe.reshape(d_model, height * width).t()                             # (HW, D)


def transformer_decoder(w, spec: CrisSpec, fq, words, pad_mask):
    B, C, H, W = fq.shape
    vpos, tpos = pos2d(C, H, W), pos1d(words.shape[-1], words.shape[1])
    vis = fq.flatten(2).transpose(1, 2)                                        # (B, HW, C)
    for i in range(spec.dec_layers):
        p = f"decoder.layers.{i}"
        v2 = _ln(w, f"{p}.norm1", vis)
        qk = v2 + vpos
        vis = vis + _ln(w, f"{p}.self_attn_norm", mha(w, f"{p}.self_attn", qk, qk, v2, spec.dec_heads))
        v2 = _ln(w, f"{p}.norm2", vis)
        v2 = mha(w, f"{p}.multihead_attn", v2 + vpos, words + tpos, words, spec.dec_heads, None, pad_mask)
        vis = vis + _ln(w, f"{p}.cross_attn_norm", v2)
        v2 = _ln(w, f"{p}.norm3", vis)
        v2 = F.relu(F.linear(v2, w[f"{p}.ffn.0.weight"], w[f"{p}.ffn.0.bias"]))
        v2 = F.linear(_ln(w, f"{p}.ffn.3", v2), w[f"{p}.ffn.4.weight"], w[f"{p}.ffn.4.bias"])
        vis = vis + v2
    return _ln(w, "decoder.norm", vis).transpose(1, 2).reshape(B, C, H, W)


def projector(w, fq, state, kernel_size=3):
    x = F.interpolate(fq, scale_factor=2, mode="bilinear")
    x = _conv_bn_relu(w, "proj.vis.1", x, 1)
    x = F.interpolate(x, scale_factor=2, mode="bilinear")
    x = _conv_bn_relu(w, "proj.vis.3", x, 1)
    x = F.conv2d(x, w["proj.vis.4.weight"], w["proj.vis.4.bias"])
    B, C, H, W = x.shape
    word = F.linear(state, w["proj.txt.weight"], w["proj.txt.bias"])
    weight, bias = word[:, :-1].reshape(B, C, kernel_size, kernel_size), word[:, -1]
    out = F.conv2d(x.reshape(1, B * C, H, W), weight, bias, padding=kernel_size // 2, groups=B)
    return out.transpose(0, 1)                                                 # (B, 1, H, W)


def additive_layer(head, fq, image_size):
    """coop_cris.py:72-88: Conv2d(1x1, no bias) -> bilinear Upsample(size=img) -> Conv2d(k, 'same', replicate)."""
    x = F.conv2d(fq, head["additive_decoder_layer.0.weight"])
    x = F.interpolate(x, size=image_size, mode="bilinear")
    k = head["additive_decoder_layer.2.weight"].shape[-1]
    x = F.pad(x, (k // 2,) * 4, mode="replicate")
    return F.conv2d(x, head["additive_decoder_layer.2.weight"], head["additive_decoder_layer.2.bias"])


def net_forward(w, spec: CrisSpec, st: L.LearnerState, head, input_ids, attention_mask, image, return_parts=False):
    """``COOPCRIS.forward`` (coop_cris.py:213-242) -> logits (B, 1, H, W)."""
    pad_mask = pad_mask_with_context(st, input_ids, attention_mask, spec.max_length)
    with torch.no_grad():                                                      # frozen, nothing upstream needs grad
        vis = encode_image(w, spec, image)
    feats = vis[-1].mean((2, 3)) if st.kind == "cocoop" else None              # coop_cris.py:90-93
    words, state = encode_text(w, spec, st, input_ids, pad_mask, feats)
    fq = fpn(w, vis, state)
    fq = transformer_decoder(w, spec, fq, words, pad_mask)
    pred = projector(w, fq, state)
    logits = F.interpolate(pred, spec.image_size, mode="bicubic", align_corners=True)
    if head is not None:
        r = head["residual_ratio"]
        logits = (1 - r) * logits + r * additive_layer(head, fq, spec.image_size)
    if return_parts:
        return logits, dict(vis=vis, words=words, state=state, fq=fq, pred=pred)
    return logits


# ------------------------------------------------------------------------------------------------------------------
# random weights of the reference's shapes
# ------------------------------------------------------------------------------------------------------------------
def init_weights(spec: CrisSpec, seed: int = 0) -> dict[str, torch.Tensor]:
    """Random frozen weights, all terms non-trivial (BN statistics, biases), scaled so activations stay O(1)."""
    g = torch.Generator().manual_seed(seed)
    w: dict[str, torch.Tensor] = {}

    def nrm(*shape, std=1.0):
        return torch.randn(*shape, generator=g) * std

    def conv(key, cout, cin, k):
        w[key] = nrm(cout, cin, k, k, std=(2.0 / (cin * k * k)) ** 0.5)

    def bn(prefix, c, gain=1.0):
        w[f"{prefix}.weight"] = gain * (1 + 0.1 * nrm(c))
        w[f"{prefix}.bias"] = 0.1 * nrm(c)
        w[f"{prefix}.running_mean"] = 0.1 * nrm(c)
        w[f"{prefix}.running_var"] = 1 + 0.2 * torch.rand(c, generator=g)

    def ln(prefix, d):
        w[f"{prefix}.weight"] = 1 + 0.05 * nrm(d)
        w[f"{prefix}.bias"] = 0.02 * nrm(d)

    def lin(prefix, out, inp, std=None, bias=True):
        w[f"{prefix}.weight"] = nrm(out, inp, std=std if std is not None else inp ** -0.5)
        if bias:
            w[f"{prefix}.bias"] = 0.02 * nrm(out)

    def attn(prefix, d, std=None):
        w[f"{prefix}.in_proj_weight"] = nrm(3 * d, d, std=std if std is not None else d ** -0.5)
        w[f"{prefix}.in_proj_bias"] = 0.02 * nrm(3 * d)
        lin(f"{prefix}.out_proj", d, d, std)

    def conv_layer(prefix, cin, cout, k):
        conv(f"{prefix}.0.weight", cout, cin, k)
        bn(f"{prefix}.1", cout)

    v, wd = "backbone.visual", spec.rn_width
    conv(f"{v}.conv1.weight", wd // 2, 3, 3); bn(f"{v}.bn1", wd // 2)
    conv(f"{v}.conv2.weight", wd // 2, wd // 2, 3); bn(f"{v}.bn2", wd // 2)
    conv(f"{v}.conv3.weight", wd, wd // 2, 3); bn(f"{v}.bn3", wd)
    inpl = wd
    for li, blocks in enumerate(spec.rn_layers, start=1):
        planes = wd * 2 ** (li - 1)
        for bi in range(blocks):
            p = f"{v}.layer{li}.{bi}"
            stride = 2 if (bi == 0 and li > 1) else 1
            conv(f"{p}.conv1.weight", planes, inpl, 1); bn(f"{p}.bn1", planes)
            conv(f"{p}.conv2.weight", planes, planes, 3); bn(f"{p}.bn2", planes)
            conv(f"{p}.conv3.weight", planes * 4, planes, 1); bn(f"{p}.bn3", planes * 4, gain=0.5)
            if stride > 1 or inpl != planes * 4:
                conv(f"{p}.downsample.0.weight", planes * 4, inpl, 1); bn(f"{p}.downsample.1", planes * 4, gain=0.7)
            inpl = planes * 4
    ed, sd = wd * 32, spec.input_resolution // 32
    a = f"{v}.attnpool"
    w[f"{a}.positional_embedding"] = nrm(sd * sd + 1, ed, std=ed ** -0.5)
    for nm in ("q_proj", "k_proj", "v_proj"):
        lin(f"{a}.{nm}", ed, ed)
    lin(f"{a}.c_proj", spec.embed_dim, ed)
    conv(f"{a}.connect.0.weight", spec.embed_dim, ed, 1); bn(f"{a}.connect.1", spec.embed_dim)

    D = spec.t_width
    w["backbone.token_embedding.weight"] = nrm(spec.vocab_size, D, std=0.02)
    w["backbone.positional_embedding"] = nrm(spec.context_length, D, std=0.01)
    for i in range(spec.t_layers):
        b = f"backbone.transformer.resblocks.{i}"
        attn(f"{b}.attn", D)
        ln(f"{b}.ln_1", D); ln(f"{b}.ln_2", D)
        lin(f"{b}.mlp.c_fc", 4 * D, D); lin(f"{b}.mlp.c_proj", D, 4 * D, std=(4 * D) ** -0.5 * 0.5)
    ln("backbone.ln_final", D)
    w["backbone.text_projection"] = nrm(D, spec.embed_dim, std=D ** -0.5)

    fi, fo = spec.fpn_in, spec.fpn_out
    lin("neck.txt_proj.0", fo[2], fi[2], bias=False); bn("neck.txt_proj.1", fo[2])
    conv_layer("neck.f1_v_proj", fi[2], fo[2], 1); bn("neck.norm_layer.0", fo[2])
    conv_layer("neck.f2_v_proj", fi[1], fo[1], 3); conv_layer("neck.f2_cat", fo[2] + fo[1], fo[1], 1)
    conv_layer("neck.f3_v_proj", fi[0], fo[0], 3); conv_layer("neck.f3_cat", fo[0] + fo[1], fo[1], 1)
    conv_layer("neck.f4_proj5", fo[2], fo[1], 3); conv_layer("neck.f4_proj4", fo[1], fo[1], 3)
    conv_layer("neck.f4_proj3", fo[1], fo[1], 3)
    conv_layer("neck.aggr", 3 * fo[1], fo[1], 1)
    conv_layer("neck.coordconv.0.conv1", fo[1] + 2, fo[1], 3); conv_layer("neck.coordconv.1", fo[1], fo[1], 3)

    C = spec.vis_dim
    for i in range(spec.dec_layers):
        p = f"decoder.layers.{i}"
        attn(f"{p}.self_attn", C); attn(f"{p}.multihead_attn", C)
        for nm in ("self_attn_norm", "cross_attn_norm", "norm1", "norm2", "norm3"):
            ln(f"{p}.{nm}", C)
        lin(f"{p}.ffn.0", spec.dec_ffn, C); ln(f"{p}.ffn.3", spec.dec_ffn); lin(f"{p}.ffn.4", C, spec.dec_ffn)
    ln("decoder.norm", C)

    pi = spec.proj_in
    conv_layer("proj.vis.1", 2 * pi, 2 * pi, 3); conv_layer("proj.vis.3", 2 * pi, pi, 3)
    conv("proj.vis.4.weight", pi, pi, 1); w["proj.vis.4.bias"] = 0.02 * nrm(pi)
    lin("proj.txt", pi * 9 + 1, spec.embed_dim, std=(spec.embed_dim * pi * 9) ** -0.5 * 3)
    return w


def init_head(spec: CrisSpec, seed: int = 1, kernel_size: int = 5, residual_ratio: float = 0.5, mid: int = 64):
    g = torch.Generator().manual_seed(seed)
    c = spec.proj_in * 2
    return {"additive_decoder_layer.0.weight": torch.randn(mid, c, 1, 1, generator=g) * c ** -0.5,
            "additive_decoder_layer.2.weight": torch.randn(1, mid, kernel_size, kernel_size, generator=g) * (mid * kernel_size ** 2) ** -0.5,
            "additive_decoder_layer.2.bias": torch.randn(1, generator=g) * 0.02,
            "residual_ratio": torch.tensor(residual_ratio)}
