"""Shared builders for the tests: the same random-init weights go into the oracle (dict) and the product (HF container)."""
from __future__ import annotations

from functools import partial

import torch

from oracle import clipseg as OC
from oracle import learners as OL

# small but kernel-compatible geometry: tower head dim 64, decoder head dim 16
SMALL = OC.ClipSegSpec(image_size=64, patch_size=16, v_hidden=128, v_heads=2, v_layers=12, v_mlp=256,
                       t_hidden=128, t_heads=2, t_layers=12, t_mlp=256, vocab_size=1000, max_position_embeddings=77,
                       projection_dim=64, reduce_dim=64, dec_heads=4, dec_mlp=128, eos_token_id=999)
FULL = OC.ClipSegSpec()   # CIDAS/clipseg-rd64 geometry: ViT-B/16 @ 352


def hf_config(spec: OC.ClipSegSpec):
    from transformers import CLIPSegConfig

    return CLIPSegConfig(
        vision_config=dict(image_size=spec.image_size, patch_size=spec.patch_size, hidden_size=spec.v_hidden,
                           num_attention_heads=spec.v_heads, num_hidden_layers=spec.v_layers, intermediate_size=spec.v_mlp),
        text_config=dict(hidden_size=spec.t_hidden, num_attention_heads=spec.t_heads, num_hidden_layers=spec.t_layers,
                         intermediate_size=spec.t_mlp, vocab_size=spec.vocab_size, eos_token_id=spec.eos_token_id,
                         bos_token_id=spec.vocab_size - 2, pad_token_id=0),
        projection_dim=spec.projection_dim, reduce_dim=spec.reduce_dim, decoder_num_attention_heads=spec.dec_heads,
        decoder_intermediate_size=spec.dec_mlp, extract_layers=list(spec.extract_layers))


def hf_model(spec: OC.ClipSegSpec, weights: dict):
    from transformers import CLIPSegForImageSegmentation

    with torch.device("meta"):
        model = CLIPSegForImageSegmentation(hf_config(spec))
    model = model.to_empty(device="cpu")
    res = model.load_state_dict(weights, strict=False)
    assert not res.unexpected_keys, res.unexpected_keys
    assert all("position_ids" in k for k in res.missing_keys), res.missing_keys
    for m in model.modules():       # non-persistent position_ids buffers are not in the state dict
        if hasattr(m, "position_ids"):
            m.position_ids = torch.arange(m.position_ids.shape[-1]).expand((1, -1))
    return model.eval()


LEARNER_CASES = {
    "maple": dict(cls="MapleCLIPSeg", learner="MapleContextLearner",
                  kw=dict(prompt_depth=9, num_context=4, intermediate_dim=64, use_proj_norm=True, use_unified_projection=False,
                          use_lora_proj=False, context_initializer=None),
                  oracle=dict(kind="maple", proj_style="mlp")),
    "vpt": dict(cls="VPTCLIPSeg", learner="VPTContextLearner", kw=dict(prompt_depth=12, num_context=8),
                oracle=dict(kind="vpt")),
    "coop": dict(cls="COOPCLIPSeg", learner="CoOpContextLearner", kw=dict(prompt_depth=1, num_context=4, context_initializer=None),
                 oracle=dict(kind="coop")),
    "coop_deep": dict(cls="COOPCLIPSeg", learner="CoOpContextLearner", kw=dict(prompt_depth=5, num_context=4, context_initializer=None),
                      oracle=dict(kind="coop")),
    "cocoop": dict(cls="COOPCLIPSeg", learner="CoCoOpContextLearner",
                   kw=dict(prompt_depth=2, num_context=4, intermediate_dim=64, use_proj_norm=True, use_unified_projection=False,
                           use_lora_proj=False, norm_image_features=False, context_initializer=None),
                   oracle=dict(kind="cocoop", proj_style="mlp", norm_image_features=False)),
    "shared_separate": dict(cls="SharedSeparateCLIPSeg", learner="SharedSeparateLearner",
                            kw=dict(shared_dim=64, prompt_depth=9, num_context=4, intermediate_dim=None, use_proj_norm=True,
                                    use_unified_projection=False, use_lora_proj=False),
                            oracle=dict(kind="shared_separate", proj_style="mlp")),
    "shared_attn": dict(cls="SharedAttnCLIPSeg", learner="SharedAttnLearner",
                        kw=dict(prompt_depth=3, num_context=4, use_unified_projection=False),
                        oracle=dict(kind="shared_attn", nhead=4)),
}


def build_net(case: str, spec: OC.ClipSegSpec, weights: dict, seed: int = 0, residual_ratio: float = 0.35):
    """Product net (on CPU; move it with .cuda()) with perturbed learner / head parameters."""
    import tunevlseg_b200.models.core_models.coop as nets
    import tunevlseg_b200.models.core_models.coop.context_learner as learners

    c = LEARNER_CASES[case]
    kw = dict(c["kw"])
    if c["learner"] == "SharedAttnLearner":
        kw["unified_projector"] = partial(torch.nn.TransformerEncoderLayer, nhead=4, dim_feedforward=96, dropout=0.25, norm_first=True)
    torch.manual_seed(seed)
    net = getattr(nets, c["cls"])(
        model_cfg=dict(pretrained_model_name_or_path=hf_model(spec, weights), freeze_encoder=False, freeze_decoder=False),
        context_learner=partial(getattr(learners, c["learner"]), **kw), freeze_all=True, no_freeze_last_layer=False,
        use_new_last_layer=True, new_last_layer_kernel_size=5, residual_ratio=residual_ratio)
    net.context_learner.eval()      # shared-attn dropout off: the oracle states eval semantics
    with torch.no_grad():
        for p in net.context_learner.parameters():
            p.add_(torch.randn_like(p) * 0.05)
    return net


def oracle_state(case: str, net, spec: OC.ClipSegSpec) -> OL.LearnerState:
    c = LEARNER_CASES[case]
    params = {k: v.detach().cpu().clone().float().requires_grad_(v.is_floating_point())
              for k, v in net.context_learner.state_dict().items()}
    return OL.LearnerState(params=params, prompt_depth=c["kw"]["prompt_depth"], num_context=c["kw"]["num_context"],
                           textual_dim=spec.t_hidden, **c["oracle"])


def oracle_head(net):
    return {k: dict(net.named_parameters())[k].detach().cpu().clone().float().requires_grad_(True)
            for k in ("additive_decoder_layer.1.weight", "additive_decoder_layer.1.bias", "residual_ratio")}


def make_batch(spec: OC.ClipSegSpec, B: int, L: int, seed: int, pad: bool = True):
    g = torch.Generator().manual_seed(seed)
    img = torch.randn(B, 3, spec.image_size, spec.image_size, generator=g)
    ids = torch.randint(1, spec.vocab_size - 10, (B, L), generator=g)
    ids[:, 0] = spec.vocab_size - 2
    am = torch.ones(B, L, dtype=torch.long)
    for b in range(B):
        eos = L - 1 if not pad else max(2, L - 1 - 2 * b)
        ids[b, eos] = spec.eos_token_id
        ids[b, eos + 1:] = 0
        am[b, eos + 1:] = 0
    mask = (torch.rand(B, 1, spec.image_size, spec.image_size, generator=g) < 0.3).float()
    return img, ids, am, mask


# ---- CRIS ------------------------------------------------------------------------------------------------------------
from types import SimpleNamespace  # noqa: E402

from oracle import cris as OCR  # noqa: E402

CRIS_SMALL = OCR.CrisSpec(image_size=64, input_resolution=96, rn_layers=(1, 2, 1, 1), rn_width=8, embed_dim=160,
                          t_width=128, t_layers=3, context_length=77, vocab_size=600, fpn_out=(64, 128, 192),
                          dec_layers=2, dec_heads=2, dec_ffn=256)
CRIS_FULL = OCR.CrisSpec()        # CLIP-RN50 @ 416x416 (configs/model/coop/cris.yaml)

CRIS_CASES = {
    "coop": dict(learner="CoOpContextLearner", kw=dict(prompt_depth=1, num_context=4), oracle=dict(kind="coop")),
    "coop_d3": dict(learner="CoOpContextLearner", kw=dict(prompt_depth=3, num_context=4), oracle=dict(kind="coop")),
    "cocoop": dict(learner="CoCoOpContextLearner",
                   kw=dict(prompt_depth=1, num_context=4, intermediate_dim=64, use_proj_norm=True, use_unified_projection=False,
                           use_lora_proj=False, norm_image_features=False),
                   oracle=dict(kind="cocoop", proj_style="mlp", norm_image_features=False)),
    "cocoop_d2_norm": dict(learner="CoCoOpContextLearner",
                           kw=dict(prompt_depth=2, num_context=4, intermediate_dim=8, use_proj_norm=True, use_unified_projection=False,
                                   use_lora_proj=False, norm_image_features=True),
                           oracle=dict(kind="cocoop", proj_style="mlp", norm_image_features=True)),
}


class StubTokenizer:
    """'a photo of a' -> 4 fixed ids: stands in for the CLIP BPE tokenizer of configs/model/coop/cris.yaml:22-24."""

    def __call__(self, text, **kw):
        texts = [text] if isinstance(text, str) else list(text)
        return SimpleNamespace(input_ids=torch.tensor([[320 + 7 * i + 13 * j for j, _ in enumerate(t.split())] for i, t in enumerate(texts)]))


def cris_model_cfg(spec: OCR.CrisSpec, weights: dict):
    backbone = {k[len("backbone."):]: v for k, v in weights.items() if k.startswith("backbone.")}
    return dict(clip_pretrain=backbone, fpn_in=list(spec.fpn_in), fpn_out=list(spec.fpn_out), vis_dim=spec.vis_dim,
                word_dim=spec.embed_dim, num_layers=spec.dec_layers, num_head=spec.dec_heads, dim_ffn=spec.dec_ffn, dropout=0.2,
                return_intermediate=False, img_size=spec.image_size, freeze_encoder=True, cris_pretrain=None)


def build_cris_net(case: str, spec: OCR.CrisSpec, weights: dict, seed: int = 0, residual_ratio: float = 0.35, new_last_layer: bool = True):
    """Product COOPCRIS on CPU holding exactly ``weights`` (the oracle's dict) as frozen parameters."""
    import tunevlseg_b200.models.core_models.coop as nets
    import tunevlseg_b200.models.core_models.coop.context_learner as learners

    c = CRIS_CASES[case]
    torch.manual_seed(seed)
    net = nets.COOPCRIS(model_cfg=cris_model_cfg(spec, weights),
                        context_learner=partial(getattr(learners, c["learner"]), context_initializer="a photo of a",
                                                tokenizer=StubTokenizer(), **c["kw"]),
                        freeze_all=True, no_freeze_last_layer=False, use_new_last_layer=new_last_layer, new_last_layer_kernel_size=5,
                        residual_ratio=residual_ratio)
    res = net.load_state_dict(weights, strict=False)       # undo build_model's fp16 round trip: hold the oracle's exact values
    assert not res.unexpected_keys, res.unexpected_keys
    with torch.no_grad():
        for p in net.context_learner.parameters():
            p.add_(torch.randn_like(p) * 0.05)
    return net


def cris_oracle_state(case: str, net) -> OL.LearnerState:
    c = CRIS_CASES[case]
    params = {k: v.detach().cpu().clone().float().requires_grad_(v.is_floating_point())
              for k, v in net.context_learner.state_dict().items()}
    return OL.LearnerState(params=params, prompt_depth=c["kw"]["prompt_depth"], num_context=c["kw"]["num_context"], **c["oracle"])


CRIS_HEAD_KEYS = ("additive_decoder_layer.0.weight", "additive_decoder_layer.2.weight", "additive_decoder_layer.2.bias", "residual_ratio")


def cris_oracle_head(net):
    named = dict(net.named_parameters())
    if "residual_ratio" not in named:
        return None
    return {k: named[k].detach().cpu().clone().float().requires_grad_(True) for k in CRIS_HEAD_KEYS}


def make_cris_batch(spec: OCR.CrisSpec, B: int, L: int, seed: int, pad: bool = True):
    g = torch.Generator().manual_seed(seed)
    img = torch.randn(B, 3, spec.image_size, spec.image_size, generator=g)
    ids = torch.randint(1, spec.vocab_size - 10, (B, L), generator=g)
    ids[:, 0] = spec.vocab_size - 2
    am = torch.ones(B, L, dtype=torch.long)
    for b in range(B):
        eos = L - 1 if not pad else max(2, L - 1 - 2 * b)
        ids[b, eos] = spec.vocab_size - 1
        ids[b, eos + 1:] = 0
        am[b, eos + 1:] = 0
    mask = (torch.rand(B, 1, spec.image_size, spec.image_size, generator=g) < 0.3).float()
    return img, ids, am, mask
