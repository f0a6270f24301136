"""Oracle restatement of the six prompt learners as pure functions.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Each learner is described by a ``LearnerState``: a ``kind`` string, the
hyper-parameters that change arithmetic, and a plain ``dict[str, Tensor]`` that
uses the *reference's own ``state_dict`` key names*, so that a state dict taken
from a reference learner, from the product's learner or from a fixture can be
dropped in unchanged.

Reference files followed (relative to /root/reference/src/models/core_models/coop/context_learner):
  base_unimodal_learner.py:17-99   context_vectors parameter (depth, n, dim)
  coop_context_learner.py:82-180   mask helpers, row overwrite 1..n, ctx insertion with truncation
  base_projector_learner.py:57-139 per-depth projector (MLP / LoRA style)
  cocoop_context_learner.py:33-77  ctx_b = meta_net(img_feat_b)[:, None] + ctx
  maple_context_learner.py:7-20    visual ctx = projector_k(ctx_k)
  vpt_context_learner.py:46-64     ctx appended at the END of the vision sequence
  base_visual_learner.py:18-23     h[:, -n:] = visual ctx
  shared_attn_learner.py:43-104    shared ctx -> TransformerEncoderLayer -> split text | visual
  shared_separate_learner.py:81-98 shared ctx -> two projector lists
"""
from __future__ import annotations

from dataclasses import dataclass, field

import torch
import torch.nn.functional as F

KINDS = ("coop", "cocoop", "vpt", "maple", "shared_separate", "shared_attn")


@dataclass
class LearnerState:
    kind: str
    prompt_depth: int
    num_context: int
    params: dict[str, torch.Tensor]
    # projector structure: "mlp" (Linear+ReLU ... Linear [+LN]) or "lora" (Linear, [Linear], [LN]) or None
    proj_style: str | None = "mlp"
    norm_image_features: bool = False      # CoCoOp (cocoop_context_learner.py:26-31)
    textual_dim: int = 512                 # shared_attn split point (shared_attn_learner.py:78-87)
    nhead: int = 16                        # shared_attn TransformerEncoderLayer heads
    ln_eps: float = 1e-5
    extra: dict = field(default_factory=dict)

    def __post_init__(self) -> None:
        if self.kind not in KINDS:
            raise ValueError(f"unknown learner kind {self.kind!r}")

    # -- which branches does this learner touch -------------------------------------------------
    @property
    def is_visual(self) -> bool:           # isinstance(learner, BaseVisualLearner) in the reference
        return self.kind in ("vpt", "maple", "shared_separate", "shared_attn")

    @property
    def is_textual(self) -> bool:
        return self.kind != "vpt"


def _sequential(params: dict[str, torch.Tensor], prefix: str, x: torch.Tensor, style: str | None, eps: float):
    """Apply a projector stored under ``prefix`` (base_projector_learner.py:65-139).

    The module is either a bare ``nn.Linear`` (keys ``prefix.weight``) or an ``nn.Sequential`` whose
    parametrised children sit at integer sub-keys.  For the MLP style a ReLU follows every Linear
    except the last one; the LoRA style has no non-linearity.  A 1-D ``weight`` is a LayerNorm.
    """
    if f"{prefix}.weight" in params:       # intermediate_dim=None -> single Linear
        return F.linear(x, params[f"{prefix}.weight"], params.get(f"{prefix}.bias"))
    idxs = sorted({int(k[len(prefix) + 1:].split(".")[0]) for k in params if k.startswith(prefix + ".")})
    linear_idxs = [i for i in idxs if params[f"{prefix}.{i}.weight"].ndim == 2]
    for i in idxs:
        w = params[f"{prefix}.{i}.weight"]
        b = params.get(f"{prefix}.{i}.bias")
        if w.ndim == 2:
            x = F.linear(x, w, b)
            if style == "mlp" and i != linear_idxs[-1]:
                x = F.relu(x)
        else:
            x = F.layer_norm(x, (w.shape[0],), w, b, eps)
    return x


def _encoder_layer_norm_first(params, prefix, x, nhead, eps):
    """``nn.TransformerEncoderLayer(norm_first=True, activation=relu, batch_first=False)`` in eval mode.

    ``x`` is (seq, batch, d).  The reference feeds ``ctx[index].unsqueeze(0)`` = (1, n, d), i.e. a sequence
    of length ONE with the n context tokens in the *batch* slot (shared_attn_learner.py:66-76), so every
    token attends only to itself.  Restated generally anyway.
    """
    S, B, D = x.shape
    hd = D // nhead
    h = F.layer_norm(x, (D,), params[f"{prefix}.norm1.weight"], params[f"{prefix}.norm1.bias"], eps)
    qkv = F.linear(h, params[f"{prefix}.self_attn.in_proj_weight"], params[f"{prefix}.self_attn.in_proj_bias"])
    q, k, v = qkv.chunk(3, dim=-1)

    def heads(t):  # (S,B,D) -> (B*nhead, S, hd)
        return t.reshape(S, B * nhead, hd).transpose(0, 1)

    q, k, v = heads(q), heads(k), heads(v)
    att = torch.softmax((q * hd ** -0.5) @ k.transpose(1, 2), dim=-1) @ v
    att = att.transpose(0, 1).reshape(S, B, D)
    x = x + F.linear(att, params[f"{prefix}.self_attn.out_proj.weight"], params[f"{prefix}.self_attn.out_proj.bias"])
    h = F.layer_norm(x, (D,), params[f"{prefix}.norm2.weight"], params[f"{prefix}.norm2.bias"], eps)
    h = F.linear(F.relu(F.linear(h, params[f"{prefix}.linear1.weight"], params[f"{prefix}.linear1.bias"])),
                 params[f"{prefix}.linear2.weight"], params[f"{prefix}.linear2.bias"])
    return x + h


def _shared_attn_both(st: LearnerState, index: int):
    ctx = st.params["context_vectors"][index].unsqueeze(0)
    out = _encoder_layer_norm_first(st.params, f"projection_layers.{index}", ctx, st.nhead, st.ln_eps).squeeze(0)
    return out[:, : st.textual_dim], out[:, st.textual_dim:]


def textual_context(st: LearnerState, index: int = 0, image_features: torch.Tensor | None = None) -> torch.Tensor:
    """``learner.get_textual_context(index=, image_features=)`` -> (n, Dt) or (B, n, Dt) for CoCoOp."""
    p = st.params
    if st.kind == "coop" or st.kind == "maple":
        return p["context_vectors"][index]
    if st.kind == "cocoop":
        if image_features is None:
            raise ValueError("`image_features` must be provided for CoCoOp")
        f = image_features
        if st.norm_image_features:
            f = f / f.norm(dim=-1, keepdim=True)
        bias = _sequential(p, f"projection_layers.{index}", f, st.proj_style, st.ln_eps)
        return bias.unsqueeze(1) + p["context_vectors"][index]
    if st.kind == "shared_separate":
        return _sequential(p, f"textual_projection_layers.{index}", p["context_vectors"][index], st.proj_style, st.ln_eps)
    if st.kind == "shared_attn":
        return _shared_attn_both(st, index)[0]
    raise ValueError(f"{st.kind} has no textual branch")


def visual_context(st: LearnerState, index: int = 0) -> torch.Tensor:
    """``learner.get_visual_context(index=)`` -> (n, Dv)."""
    p = st.params
    if st.kind == "vpt":
        return p["context_vectors"][index]
    if st.kind == "maple":
        return _sequential(p, f"projection_layers.{index}", p["context_vectors"][index], st.proj_style, st.ln_eps)
    if st.kind == "shared_separate":
        return _sequential(p, f"visual_projection_layers.{index}", p["context_vectors"][index], st.proj_style, st.ln_eps)
    if st.kind == "shared_attn":
        return _shared_attn_both(st, index)[1]
    raise ValueError(f"{st.kind} has no visual branch")


def insert_textual_context(st: LearnerState, input_embeddings: torch.Tensor, max_length: int | None,
                           image_features: torch.Tensor | None = None) -> torch.Tensor:
    """``CoOpContextLearner.forward`` (coop_context_learner.py:136-180): [BOS, ctx*n, mid, last].

    With ``max_length`` the middle is cut so the total is ``min(L + n, max_length)`` while the final
    token (normally EOS) is preserved.
    """
    n = st.num_context
    L = input_embeddings.size(1)
    mid_last = -1 if max_length is None else min(max_length - n, L) - 1
    ctx = textual_context(st, 0, image_features)
    if ctx.ndim == 2:
        ctx = ctx.expand(input_embeddings.size(0), -1, -1)
    return torch.cat((input_embeddings[:, :1], ctx, input_embeddings[:, 1:mid_last], input_embeddings[:, -1:]), dim=1)


def attention_mask_for_context(st: LearnerState, attention_mask: torch.Tensor, max_length: int | None) -> torch.Tensor:
    """``update_attention_mask_for_context`` (coop_context_learner.py:82-107): n ones PREPENDED, cut to max_length."""
    ones = torch.ones(attention_mask.shape[0], st.num_context, dtype=attention_mask.dtype)
    return torch.cat((ones, attention_mask), dim=1)[:, :max_length]
