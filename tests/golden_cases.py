"""Shared description of the golden fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py)."""
from __future__ import annotations

import os

import numpy as np
import torch

from oracle import clipseg as OC
from oracle import cris as OCR
from oracle import learners as OL

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

TINY = OC.ClipSegSpec(image_size=64, patch_size=16, v_hidden=32, v_heads=4, v_layers=12, v_mlp=64,
                      t_hidden=32, t_heads=4, t_layers=12, t_mlp=64, vocab_size=600, max_position_embeddings=77,
                      projection_dim=32, reduce_dim=16, dec_heads=4, dec_mlp=64, eos_token_id=599)

# name -> LearnerState keyword arguments (what make_golden.py passed to the reference learner)
CASES = {
    "maple_d9_n4": dict(kind="maple", prompt_depth=9, num_context=4, proj_style="mlp"),
    "maple_d3_n2_padded_unified_lora": dict(kind="maple", prompt_depth=3, num_context=2, proj_style="lora"),
    "vpt_d12_n3": dict(kind="vpt", prompt_depth=12, num_context=3),
    "shared_separate_d9_n4": dict(kind="shared_separate", prompt_depth=9, num_context=4, proj_style="mlp"),
    "shared_attn_d3_n4": dict(kind="shared_attn", prompt_depth=3, num_context=4, textual_dim=32, nhead=4),
    "coop_d1_n4": dict(kind="coop", prompt_depth=1, num_context=4),
    "coop_d5_n4_long": dict(kind="coop", prompt_depth=5, num_context=4),
    "cocoop_d2_n4": dict(kind="cocoop", prompt_depth=2, num_context=4, proj_style="mlp", norm_image_features=False),
}


def load_weights():
    z = np.load(os.path.join(GOLDEN, "clipseg_tiny_weights.npz"))
    return {k: torch.from_numpy(z[k]) for k in z.files}


def load_case(name):
    z = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    d = {k: torch.from_numpy(np.asarray(z[k])) for k in z.files}
    learner = {k[len("learner/"):]: v for k, v in d.items() if k.startswith("learner/")}
    head = {k[len("head/"):]: v for k, v in d.items() if k.startswith("head/")}
    return d, learner, head


def learner_state(name, params):
    return OL.LearnerState(params=params, **CASES[name])


# ---- CRIS (tests/golden/make_golden_cris.py) -----------------------------------------------------------------------
CRIS_TINY = OCR.CrisSpec(image_size=64, input_resolution=96, rn_layers=(1, 2, 1, 1), rn_width=8, embed_dim=160,
                         t_width=128, t_layers=3, context_length=77, vocab_size=600, fpn_out=(64, 128, 192),
                         dec_layers=2, dec_heads=2, dec_ffn=256)

CRIS_CASES = {
    "cris_coop_d1_n4": dict(kind="coop", prompt_depth=1, num_context=4),
    "cris_coop_d3_n4_nomask": dict(kind="coop", prompt_depth=3, num_context=4),
    "cris_cocoop_d1_n4": dict(kind="cocoop", prompt_depth=1, num_context=4, proj_style="mlp", norm_image_features=False),
    "cris_cocoop_d2_n4_long": dict(kind="cocoop", prompt_depth=2, num_context=4, proj_style="mlp", norm_image_features=True),
}


def load_cris_weights():
    """The generator loads ``init_weights(CRIS_TINY, seed=2025)`` over the whole reference net (bit-exact fp32), so the
    frozen weights are regenerated from the seed instead of being stored (12 MB)."""
    return OCR.init_weights(CRIS_TINY, seed=2025)


def cris_learner_state(name, params):
    return OL.LearnerState(params=params, **CRIS_CASES[name])
