"""Generate the golden fixtures under tests/golden/ by running the REAL reference classes.

Run once in the build container (it needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

What it does
  * imports ``src.models.core_models.coop`` from /root/reference unmodified,
  * applies a TEST-SIDE compatibility shim, because the reference targets the transformers 4.3x API
    and this image ships 5.5.0: the two 4-D mask helpers are re-exported into
    ``modeling_clipseg`` and the encoder/decoder layer ``forward`` accepts the 4.x positional
    ``(hidden, attention_mask, causal_attention_mask, output_attentions=)`` call and returns a 1-tuple,
  * builds a tiny random-init ``CLIPSegForImageSegmentation`` (weights from ``oracle.clipseg.init_weights``),
  * runs MapleCLIPSeg / VPTCLIPSeg / SharedSeparateCLIPSeg / SharedAttnCLIPSeg / COOPCLIPSeg(CoOp, CoCoOp)
    forward + backward and stores inputs, logits and learner gradients as ``.npz``.

The fixtures are what ``tests/test_oracle_golden.py`` pins the oracle to (<=1e-5).
"""
from __future__ import annotations

import functools
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from oracle import clipseg as OC  # noqa: E402

TINY = OC.ClipSegSpec(image_size=64, patch_size=16, v_hidden=32, v_heads=4, v_layers=12, v_mlp=64,
                      t_hidden=32, t_heads=4, t_layers=12, t_mlp=64, vocab_size=600, max_position_embeddings=77,
                      projection_dim=32, reduce_dim=16, dec_heads=4, dec_mlp=64, eos_token_id=599)


from oracle.ref_shim import install_shim  # noqa: E402


def hf_model_dir(weights) -> str:
    from transformers import CLIPSegConfig, CLIPSegForImageSegmentation

    s = TINY
    cfg = CLIPSegConfig(
        vision_config=dict(image_size=s.image_size, patch_size=s.patch_size, hidden_size=s.v_hidden,
                           num_attention_heads=s.v_heads, num_hidden_layers=s.v_layers, intermediate_size=s.v_mlp),
        text_config=dict(hidden_size=s.t_hidden, num_attention_heads=s.t_heads, num_hidden_layers=s.t_layers,
                         intermediate_size=s.t_mlp, vocab_size=s.vocab_size, eos_token_id=s.eos_token_id,
                         bos_token_id=s.vocab_size - 2, pad_token_id=0),
        projection_dim=s.projection_dim, reduce_dim=s.reduce_dim, decoder_num_attention_heads=s.dec_heads,
        decoder_intermediate_size=s.dec_mlp)
    model = CLIPSegForImageSegmentation(cfg)
    missing = model.load_state_dict(weights, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    d = tempfile.mkdtemp(prefix="tvs_golden_hf_")
    model.save_pretrained(d)
    return d


def make_inputs(B, Ltxt, seed, pad_from=None):
    g = torch.Generator().manual_seed(seed)
    s = TINY
    img = torch.randn(B, 3, s.image_size, s.image_size, generator=g)
    ids = torch.randint(1, s.vocab_size - 10, (B, Ltxt), generator=g)
    ids[:, 0] = s.vocab_size - 2
    am = torch.ones(B, Ltxt, dtype=torch.long)
    for b in range(B):
        eos = Ltxt - 1 if pad_from is None else max(2, pad_from - b)
        ids[b, eos] = s.eos_token_id
        ids[b, eos + 1:] = 0
        am[b, eos + 1:] = 0
    return img, ids, am


def run_case(name, net_cls, learner_partial, model_dir, B, Ltxt, pad_from, seed, eval_learner=False):
    torch.manual_seed(seed)
    net = net_cls(
        model_cfg=dict(pretrained_model_name_or_path=model_dir, freeze_encoder=False, freeze_decoder=False),
        context_learner=learner_partial, freeze_all=True, no_freeze_last_layer=False, use_new_last_layer=True,
        new_last_layer_kernel_size=5, residual_ratio=0.35)
    if eval_learner:
        net.context_learner.eval()
    # make the learner parameters non-degenerate (LayerNorm affine etc.)
    with torch.no_grad():
        for p in net.context_learner.parameters():
            p.add_(torch.randn_like(p) * 0.05)
    img, ids, am = make_inputs(B, Ltxt, seed + 100, pad_from)
    logits = net(text_input={"input_ids": ids, "attention_mask": am}, image_input=img)
    gw = torch.randn(logits.shape, generator=torch.Generator().manual_seed(seed + 7))
    (logits * gw).sum().backward()
    out = {"image": img.numpy(), "input_ids": ids.numpy(), "attention_mask": am.numpy(),
           "logits": logits.detach().numpy(), "grad_weight": gw.numpy()}
    for k, v in net.context_learner.state_dict().items():
        out[f"learner/{k}"] = v.detach().numpy()
    for k, p in net.context_learner.named_parameters():
        out[f"learner_grad/{k}"] = (p.grad if p.grad is not None else torch.zeros_like(p)).numpy()
        out[f"learner_hasgrad/{k}"] = np.array(p.grad is not None)
    for k in ("additive_decoder_layer.1.weight", "additive_decoder_layer.1.bias", "residual_ratio"):
        p = dict(net.named_parameters())[k]
        out[f"head/{k}"] = p.detach().numpy()
        out[f"head_grad/{k}"] = (p.grad if p.grad is not None else torch.zeros_like(p)).numpy()
        out[f"head_hasgrad/{k}"] = np.array(p.grad is not None)
    np.savez(os.path.join(HERE, f"{name}.npz"), **out)
    print(f"{name}: logits {tuple(logits.shape)} mean|x|={logits.abs().mean():.4f}")


def main():
    install_shim()
    from functools import partial

    from src.models.core_models.coop import (COOPCLIPSeg, MapleCLIPSeg, SharedAttnCLIPSeg, SharedSeparateCLIPSeg,
                                             VPTCLIPSeg)
    from src.models.core_models.coop.context_learner import (CoCoOpContextLearner, CoOpContextLearner,
                                                             MapleContextLearner, SharedAttnLearner,
                                                             SharedSeparateLearner, VPTContextLearner)

    # /root/reference/src/models/__init__.py:6 sets float32 matmul precision to "medium" at import, which lets
    # the CPU GEMMs run with bf16 internals (errors ~5e-3).  The fixtures pin the reference's fp32 path.
    torch.set_float32_matmul_precision("highest")

    weights = OC.init_weights(TINY, seed=2024)
    np.savez(os.path.join(HERE, "clipseg_tiny_weights.npz"), **{k: v.numpy() for k, v in weights.items()})
    d = hf_model_dir(weights)

    run_case("maple_d9_n4", MapleCLIPSeg,
             partial(MapleContextLearner, prompt_depth=9, num_context=4, intermediate_dim=8, use_proj_norm=True,
                     use_unified_projection=False, use_lora_proj=False, context_initializer=None),
             d, B=2, Ltxt=8, pad_from=None, seed=11)
    run_case("maple_d3_n2_padded_unified_lora", MapleCLIPSeg,
             partial(MapleContextLearner, prompt_depth=3, num_context=2, intermediate_dim=8, use_proj_norm=False,
                     use_unified_projection=True, use_lora_proj=True, context_initializer=None),
             d, B=3, Ltxt=12, pad_from=7, seed=12)
    run_case("vpt_d12_n3", VPTCLIPSeg, partial(VPTContextLearner, prompt_depth=12, num_context=3),
             d, B=2, Ltxt=8, pad_from=6, seed=13)
    run_case("shared_separate_d9_n4", SharedSeparateCLIPSeg,
             partial(SharedSeparateLearner, shared_dim=16, prompt_depth=9, num_context=4, intermediate_dim=None,
                     use_proj_norm=True, use_unified_projection=False, use_lora_proj=False),
             d, B=2, Ltxt=8, pad_from=None, seed=14)
    run_case("shared_attn_d3_n4", SharedAttnCLIPSeg,
             partial(SharedAttnLearner, prompt_depth=3, num_context=4, use_unified_projection=False,
                     unified_projector=partial(torch.nn.TransformerEncoderLayer, nhead=4, dim_feedforward=48,
                                               dropout=0.25, norm_first=True)),
             d, B=2, Ltxt=8, pad_from=None, seed=15, eval_learner=True)
    run_case("coop_d1_n4", COOPCLIPSeg,
             partial(CoOpContextLearner, prompt_depth=1, num_context=4, context_initializer=None),
             d, B=2, Ltxt=8, pad_from=6, seed=16)
    run_case("coop_d5_n4_long", COOPCLIPSeg,
             partial(CoOpContextLearner, prompt_depth=5, num_context=4, context_initializer=None),
             d, B=2, Ltxt=76, pad_from=None, seed=17)
    run_case("cocoop_d2_n4", COOPCLIPSeg,
             partial(CoCoOpContextLearner, prompt_depth=2, num_context=4, intermediate_dim=8, use_proj_norm=True,
                     use_unified_projection=False, use_lora_proj=False, norm_image_features=False,
                     context_initializer=None),
             d, B=2, Ltxt=8, pad_from=None, seed=18)


if __name__ == "__main__":
    main()
