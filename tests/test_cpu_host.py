"""CPU-side checks (no GPU): the C ABI library loads and exports every declared symbol, the product refuses to run
without a CUDA device (no fallback), the host-side learners agree with the oracle learners on the reference's own
state dicts (golden fixtures), module plumbing (optimizer groups, metric formulas)."""
import ctypes
import os
import re
from functools import partial

import pytest
import torch

from oracle import learners as OL
from oracle import loss_metrics as OLM
from tests.golden_cases import CASES, TINY, learner_state, load_case

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from tunevlseg_b200 import abi

    assert os.path.exists(abi.lib_path()), "run `python -m tunevlseg_b200.build` (or __graft_entry__.build())"
    lib = ctypes.CDLL(abi.lib_path())
    hdr = open(os.path.join(ROOT, "include", "tvs_b200.h")).read()
    names = sorted(set(re.findall(r"\b(tvs_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 25
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    lib.tvs_version.restype = ctypes.c_int
    assert lib.tvs_version() == 1
    abi.load()      # argtypes for every entry resolve


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    from tunevlseg_b200 import abi, engine

    with pytest.raises(abi.TvsError):
        abi.require_device()
    x = torch.zeros(2, 1, 8, 8)
    with pytest.raises(abi.TvsError):
        engine.DiceBceFn.apply(x, x, 0.5, 1.0, 0.2, None)


def test_product_does_not_import_oracle():
    import subprocess
    import sys

    code = ("import sys; import tunevlseg_b200, tunevlseg_b200.engine, tunevlseg_b200.models.core_models.coop, "
            "tunevlseg_b200.models.image_text_mask_module, tunevlseg_b200.optim; "
            "bad=[m for m in sys.modules if m == 'oracle' or m.startswith('oracle.')]; sys.exit(1 if bad else 0)")
    assert subprocess.run([sys.executable, "-c", code], cwd=ROOT).returncode == 0


GOLDEN_LEARNERS = {
    "maple_d9_n4": ("MapleContextLearner", dict(visual_dim=32, context_dim=32, prompt_depth=9, num_context=4, intermediate_dim=8,
                                                 use_proj_norm=True, use_unified_projection=False, use_lora_proj=False)),
    "maple_d3_n2_padded_unified_lora": ("MapleContextLearner", dict(visual_dim=32, context_dim=32, prompt_depth=3, num_context=2,
                                                                     intermediate_dim=8, use_proj_norm=False,
                                                                     use_unified_projection=True, use_lora_proj=True)),
    "vpt_d12_n3": ("VPTContextLearner", dict(context_dim=32, prompt_depth=12, num_context=3)),
    "shared_separate_d9_n4": ("SharedSeparateLearner", dict(textual_dim=32, visual_dim=32, shared_dim=16, prompt_depth=9, num_context=4,
                                                            intermediate_dim=None, use_proj_norm=True, use_unified_projection=False)),
    "shared_attn_d3_n4": ("SharedAttnLearner", dict(textual_dim=32, visual_dim=32, prompt_depth=3, num_context=4,
                                                    use_unified_projection=False,
                                                    unified_projector=partial(torch.nn.TransformerEncoderLayer, nhead=4,
                                                                              dim_feedforward=48, dropout=0.25, norm_first=True))),
    "cocoop_d2_n4": ("CoCoOpContextLearner", dict(visual_dim=32, context_dim=32, prompt_depth=2, num_context=4, intermediate_dim=8,
                                                  use_proj_norm=True, use_unified_projection=False, use_lora_proj=False,
                                                  norm_image_features=False)),
    "coop_d5_n4_long": ("CoOpContextLearner", dict(context_dim=32, prompt_depth=5, num_context=4)),
}


@pytest.mark.parametrize("name", list(GOLDEN_LEARNERS))
def test_learners_match_oracle_on_reference_state(name):
    """The reference learner's own state_dict (fixture) loads strictly into the product learner, and every context the
    towers consume equals the oracle's restatement."""
    import tunevlseg_b200.models.core_models.coop.context_learner as L

    cls_name, kw = GOLDEN_LEARNERS[name]
    _, params, _ = load_case(name)
    learner = getattr(L, cls_name)(max_network_depth=12, **kw).eval()
    learner.load_state_dict(params, strict=True)
    st = learner_state(name, params)
    feats = torch.randn(3, 32, generator=torch.Generator().manual_seed(1))
    if st.is_visual:
        stack = learner.visual_stack(10)
        for i in range(min(st.prompt_depth, 11)):
            torch.testing.assert_close(stack[i], OL.visual_context(st, i), rtol=1e-5, atol=1e-6)
        if hasattr(learner, "_computed_textual_context_cache"):
            assert set(learner._computed_textual_context_cache) == set(range(min(st.prompt_depth, 11)))
    if st.is_textual:
        f = feats if st.kind == "cocoop" else None
        emb = torch.randn(3, 9, 32, generator=torch.Generator().manual_seed(2))
        torch.testing.assert_close(learner(input_embeddings=emb, max_length=77, image_features=f),
                                   OL.insert_textual_context(st, emb, 77, f), rtol=1e-5, atol=1e-6)
        deep = learner.textual_deep_stack(12, image_features=f)
        if st.prompt_depth == 1:
            assert deep is None
        else:
            for i in range(1, st.prompt_depth):
                torch.testing.assert_close(deep[i - 1], OL.textual_context(st, i, f), rtol=1e-5, atol=1e-6)
        long = torch.randn(2, 76, 32)
        assert learner(input_embeddings=long, max_length=77, image_features=None if f is None else f[:2]).shape[1] == 77
        am = torch.ones(2, 76, dtype=torch.long)
        assert learner.update_attention_mask_for_context(am, 77).shape == (2, 77)


def test_learner_argument_errors():
    import tunevlseg_b200.models.core_models.coop.context_learner as L

    with pytest.raises(ValueError):
        L.VPTContextLearner(max_network_depth=12, prompt_depth=13, num_context=4, context_dim=8)
    with pytest.raises(ValueError):
        L.VPTContextLearner(max_network_depth=12, prompt_depth=0, num_context=4, context_dim=8)
    with pytest.raises(ValueError):
        L.CoOpContextLearner(max_network_depth=12, prompt_depth=1)
    with pytest.raises(NotImplementedError):
        L.SharedAttnLearner(max_network_depth=12, textual_dim=8, visual_dim=8, unified_projector=None, num_context=2)
    with pytest.raises(ValueError):
        L.CoCoOpContextLearner(max_network_depth=12, visual_dim=8, context_dim=8, num_context=2).get_textual_context()


def test_metric_formulas_known_answers():
    from tunevlseg_b200.metrics import Dice, JaccardIndex

    # all-zero prediction and target -> Dice 1, IoU 1 (zero_division=1)
    counts = torch.zeros(2, 3, dtype=torch.int64)
    assert Dice.score(counts, 1.0).item() == 1.0
    assert JaccardIndex.score(torch.tensor([10, 0, 0, 0]), 1.0).item() == 1.0
    # tp=3 fp=1 fn=2 -> dice 6/9, iou 3/6
    assert abs(Dice.score(torch.tensor([[3, 1, 2]]), 1.0).item() - 6 / 9) < 1e-7
    assert abs(JaccardIndex.score(torch.tensor([4, 1, 2, 3]), 1.0).item() - 0.5) < 1e-7
    # product formulas == oracle formulas on random counters
    c = torch.randint(0, 50, (7, 3))
    assert torch.allclose(Dice.score(c, 1.0), OLM.dice_from_counts(c, 1.0))
    cf = torch.randint(0, 50, (4,))
    assert torch.allclose(JaccardIndex.score(cf, 1.0), OLM.iou_from_confmat(cf.view(2, 2), 1.0))


def test_module_optimizer_groups_and_setup():
    from tunevlseg_b200.losses import DiceCELoss
    from tunevlseg_b200.models.image_text_mask_module import ImageTextMaskModule

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.lin = torch.nn.Linear(4, 4)
            self.norm = torch.nn.LayerNorm(4)
            self.context_vectors = torch.nn.Parameter(torch.zeros(2, 4))
            self.residual_ratio = torch.nn.Parameter(torch.tensor(0.5))

    mod = ImageTextMaskModule(net=Net(), loss_fn=DiceCELoss(sigmoid=True, lambda_dice=1, lambda_ce=0.2),
                              optimizer=partial(torch.optim.AdamW, lr=2e-4), scheduler=None, compile=False, task="binary",
                              threshold=0.5, weight_decay=0.01)
    groups = mod.get_optim_groups()
    decay = {id(p) for p in groups[0]["params"]}
    assert id(mod.net.lin.weight) in decay and id(mod.net.norm.weight) not in decay
    assert id(mod.net.context_vectors) not in decay and id(mod.net.residual_ratio) not in decay
    assert groups[1]["weight_decay"] == 0.0
    mod.setup("fit")
    assert {"train_dice", "train_iou", "val_dice", "val_iou"} <= set(mod.registered_metric_names)
    assert "optimizer" in mod.configure_optimizers()
    with pytest.raises(NotImplementedError):
        DiceCELoss(sigmoid=False)


def test_install_as_src_resolves_reference_targets():
    import importlib

    import tunevlseg_b200

    tunevlseg_b200.install_as_src()
    for target in ("src.models.image_text_mask_module.ImageTextMaskModule", "src.models.core_models.coop.MapleCLIPSeg",
                   "src.models.core_models.coop.COOPCLIPSeg", "src.models.core_models.coop.VPTCLIPSeg",
                   "src.models.core_models.coop.SharedAttnCLIPSeg", "src.models.core_models.coop.SharedSeparateCLIPSeg",
                   "src.models.core_models.coop.context_learner.MapleContextLearner",
                   "src.models.core_models.coop.context_learner.CoCoOpContextLearner",
                   "src.models.core_models.coop.context_learner.SharedSeparateLearner",
                   "src.models.components.hf_clipseg_wrapper.HFCLIPSegWrapper"):
        mod, _, attr = target.rpartition(".")
        assert hasattr(importlib.import_module(mod), attr), target
