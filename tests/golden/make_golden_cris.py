"""Generate the CRIS golden fixtures (tests/golden/cris_*.npz) by running the REAL reference ``COOPCRIS``.

Run once in the build container (needs /root/reference; it does not exist on the GPU box):

    python tests/golden/make_golden_cris.py

The reference loads ``pretrain/RN50.pt`` with ``torch.jit.load`` (cris_model/__init__.py:74-77); there is no
checkpoint here, so ``CRIS.get_backbone`` is replaced (test side only) by ``build_model(<random state dict>)`` - the
same ``build_model`` the reference calls on the TorchScript archive's state dict (it infers every shape from the
keys, clip.py:578-639).  ``oracle.cris.init_weights(TINY, seed=2025)`` is then loaded over the whole net
(``load_state_dict(strict=False)``: only learner / additive-layer / BN-counter keys are missing), so the frozen weights
are bit-exact fp32 copies of what the oracle regenerates from the same seed - nothing has to be stored.

The context learner is seeded from an initializer string through a stub tokenizer, because COOPCRIS passes
``context_dim=word_dim`` (1024-d; coop_cris.py:45) while the text transformer is 512-d: random initialisation would
not even concatenate, so the reference's configs always use ``context_initializer: "a photo of a"``.
"""
from __future__ import annotations

import os
import sys
from functools import partial
from types import SimpleNamespace

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from oracle import cris as OCR  # noqa: E402

TINY = OCR.CrisSpec(image_size=64, input_resolution=96, rn_layers=(1, 2, 1, 1), rn_width=8, embed_dim=160,
                    t_width=128, t_layers=3, context_length=77, vocab_size=600, fpn_out=(64, 128, 192),
                    dec_layers=2, dec_heads=2, dec_ffn=256)


class StubTokenizer:
    """'a photo of a' -> 4 fixed ids (what a CLIP BPE tokenizer returns with add_special_tokens=False)."""

    def __call__(self, text, **kw):
        texts = [text] if isinstance(text, str) else list(text)
        ids = [[320 + 7 * i + 13 * j for j, _ in enumerate(t.split())] for i, t in enumerate(texts)]
        return SimpleNamespace(input_ids=torch.tensor(ids))


def make_inputs(B, Ltxt, seed, pad_from=None, use_mask=True):
    g = torch.Generator().manual_seed(seed)
    img = torch.randn(B, 3, TINY.image_size, TINY.image_size, generator=g)
    ids = torch.randint(1, TINY.vocab_size - 10, (B, Ltxt), generator=g)
    ids[:, 0] = TINY.vocab_size - 2
    am = torch.ones(B, Ltxt, dtype=torch.long)
    for b in range(B):
        eos = Ltxt - 1 if pad_from is None else max(2, pad_from - b)
        ids[b, eos] = TINY.vocab_size - 1
        ids[b, eos + 1:] = 0
        am[b, eos + 1:] = 0
    return img, ids, (am if use_mask else None)


def run_case(name, COOPCRIS, learner_partial, weights, B, Ltxt, pad_from, seed, use_mask=True):
    torch.manual_seed(seed)
    s = TINY
    net = COOPCRIS(
        model_cfg=dict(clip_pretrain="unused", fpn_in=list(s.fpn_in), fpn_out=list(s.fpn_out), vis_dim=s.vis_dim,
                       word_dim=s.embed_dim, num_layers=s.dec_layers, num_head=s.dec_heads, dim_ffn=s.dec_ffn,
                       dropout=0.2, return_intermediate=False, img_size=s.image_size, freeze_encoder=True,
                       cris_pretrain=None),
        context_learner=learner_partial, freeze_all=True, no_freeze_last_layer=False, use_new_last_layer=True,
        new_last_layer_kernel_size=5, residual_ratio=0.35)
    res = net.load_state_dict(weights, strict=False)
    assert not res.unexpected_keys, res.unexpected_keys
    assert all(k.startswith(("context_learner.", "additive_decoder_layer.", "residual_ratio")) or
               k.endswith("num_batches_tracked") or k == "backbone.logit_scale" for k in res.missing_keys), res.missing_keys
    assert not net.training or not net.decoder.training        # freeze_all -> eval()
    with torch.no_grad():
        for p in net.context_learner.parameters():
            p.add_(torch.randn_like(p) * 0.05)
    img, ids, am = make_inputs(B, Ltxt, seed + 100, pad_from, use_mask)
    ti = {"input_ids": ids} if am is None else {"input_ids": ids, "attention_mask": am}
    logits = net(text_input=ti, image_input=img)
    gw = torch.randn(logits.shape, generator=torch.Generator().manual_seed(seed + 7))
    (logits * gw).sum().backward()
    out = {"image": img.numpy(), "input_ids": ids.numpy(), "logits": logits.detach().numpy(), "grad_weight": gw.numpy(),
           "has_mask": np.array(am is not None)}
    if am is not None:
        out["attention_mask"] = am.numpy()
    for k, v in net.context_learner.state_dict().items():
        out[f"learner/{k}"] = v.detach().numpy()
    for k, p in net.context_learner.named_parameters():
        out[f"learner_grad/{k}"] = (p.grad if p.grad is not None else torch.zeros_like(p)).numpy()
        out[f"learner_hasgrad/{k}"] = np.array(p.grad is not None)
    named = dict(net.named_parameters())
    for k in ("additive_decoder_layer.0.weight", "additive_decoder_layer.2.weight", "additive_decoder_layer.2.bias", "residual_ratio"):
        p = named[k]
        out[f"head/{k}"] = p.detach().numpy()
        out[f"head_grad/{k}"] = (p.grad if p.grad is not None else torch.zeros_like(p)).numpy()
        out[f"head_hasgrad/{k}"] = np.array(p.grad is not None)
    np.savez(os.path.join(HERE, f"{name}.npz"), **out)
    print(f"{name}: logits {tuple(logits.shape)} mean|x|={logits.abs().mean():.4f} max|x|={logits.abs().max():.3f}")
    return net


def main():
    sys.path.insert(0, HERE)
    from make_golden import install_shim     # the package __init__ imports the CLIPSeg wrappers (transformers 4.x API)

    install_shim()
    from src.models.components import cris_model as ref_cris
    from src.models.components.cris_model.clip import build_model
    from src.models.core_models.coop.context_learner import CoCoOpContextLearner, CoOpContextLearner
    from src.models.core_models.coop.coop_cris import COOPCRIS

    torch.set_float32_matmul_precision("highest")      # src/models/__init__.py sets "medium" at import

    raw = OCR.init_weights(TINY, seed=2025)
    backbone_sd = {k[len("backbone."):]: v.clone() for k, v in raw.items() if k.startswith("backbone.")}
    ref_cris.CRIS.get_backbone = staticmethod(lambda _path: build_model(dict(backbone_sd)).float())
    tok = StubTokenizer()

    net = run_case("cris_coop_d1_n4", COOPCRIS,
                   partial(CoOpContextLearner, prompt_depth=1, num_context=4, context_initializer="a photo of a", tokenizer=tok),
                   raw, B=2, Ltxt=8, pad_from=6, seed=21)
    sd = {k: v for k, v in net.state_dict().items()
          if not k.startswith(("context_learner.", "additive_decoder_layer.", "residual_ratio")) and
          not k.endswith("num_batches_tracked") and k != "backbone.logit_scale"}
    assert set(sd) == set(raw) and all(torch.equal(sd[k], raw[k]) for k in raw)      # the net holds exactly `raw`

    run_case("cris_coop_d3_n4_nomask", COOPCRIS,
             partial(CoOpContextLearner, prompt_depth=3, num_context=4, context_initializer="a photo of a", tokenizer=tok),
             raw, B=2, Ltxt=10, pad_from=7, seed=22, use_mask=False)
    run_case("cris_cocoop_d1_n4", COOPCRIS,
             partial(CoCoOpContextLearner, prompt_depth=1, num_context=4, intermediate_dim=8, use_proj_norm=True,
                     use_unified_projection=False, use_lora_proj=False, norm_image_features=False,
                     context_initializer="a photo of a", tokenizer=tok),
             raw, B=3, Ltxt=8, pad_from=6, seed=23)
    run_case("cris_cocoop_d2_n4_long", COOPCRIS,
             partial(CoCoOpContextLearner, prompt_depth=2, num_context=4, intermediate_dim=8, use_proj_norm=True,
                     use_unified_projection=False, use_lora_proj=False, norm_image_features=True,
                     context_initializer="a photo of a", tokenizer=tok),
             raw, B=2, Ltxt=76, pad_from=None, seed=24)


if __name__ == "__main__":
    main()
