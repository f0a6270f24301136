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
from tests.golden_cases import TINY, learner_state, load_case

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
    assert lib.tvs_version() == 3          # TVS_ABI_VERSION (3: train-time augmentation entries)
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
                   "src.models.components.hf_clipseg_wrapper.HFCLIPSegWrapper",
                   "src.data.components.data_collator.CustomDataCollatorWithPadding"):
        mod, _, attr = target.rpartition(".")
        assert hasattr(importlib.import_module(mod), attr), target


# ---- CRIS host side ----------------------------------------------------------------------------------------------------
def test_cris_state_dict_keys_match_reference():
    """COOPCRIS holds its parameters under the reference's state_dict names: frozen keys == the oracle's weight dict
    (whose key set the golden generator asserted against the real reference net), plus BN counters / logit_scale."""
    from oracle import cris as OCR
    from tests.helpers import CRIS_HEAD_KEYS, CRIS_SMALL, build_cris_net

    w = OCR.init_weights(CRIS_SMALL, seed=3)
    net = build_cris_net("cocoop", CRIS_SMALL, w)
    sd = net.state_dict()
    frozen = {k for k in sd if not k.startswith(("context_learner.", "additive_decoder_layer.", "residual_ratio"))
              and not k.endswith("num_batches_tracked") and k != "backbone.logit_scale"}
    assert frozen == set(w)
    assert all(torch.equal(sd[k], w[k]) for k in w)
    assert all(k in sd for k in CRIS_HEAD_KEYS) and "backbone.logit_scale" in sd
    assert "context_learner.context_vectors" in sd and sd["context_learner.context_vectors"].shape == (1, 4, CRIS_SMALL.t_width)
    trainable = {k for k, p in net.named_parameters() if p.requires_grad}
    assert trainable == {k for k in sd if k.startswith(("context_learner.", "additive_decoder_layer."))
                         and not k.endswith(("running_mean", "running_var", "num_batches_tracked"))} | {"residual_ratio"}
    net.train()
    assert not net.backbone.training and not net.decoder.training and not net.neck.training and not net.proj.training


def test_cris_build_model_fp16_round_trip_and_errors():
    from oracle import cris as OCR
    from tests.helpers import CRIS_SMALL, cris_model_cfg
    from tunevlseg_b200.models.components.cris_model import CRIS, build_model

    w = OCR.init_weights(CRIS_SMALL, seed=4)
    bb = {k[len("backbone."):]: v for k, v in w.items() if k.startswith("backbone.")}
    m = build_model(dict(bb, input_resolution=torch.tensor(96), context_length=torch.tensor(77), vocab_size=torch.tensor(600)))
    sd = m.state_dict()
    # clip.py:556-575: conv / linear / attention tensors go through fp16, norms and embeddings do not
    assert torch.equal(sd["visual.layer1.0.conv2.weight"], bb["visual.layer1.0.conv2.weight"].half().float())
    assert torch.equal(sd["transformer.resblocks.0.attn.in_proj_weight"], bb["transformer.resblocks.0.attn.in_proj_weight"].half().float())
    assert torch.equal(sd["text_projection"], bb["text_projection"].half().float())
    assert torch.equal(sd["visual.bn1.weight"], bb["visual.bn1.weight"]) and torch.equal(sd["token_embedding.weight"], bb["token_embedding.weight"])
    assert m.visual.layers == CRIS_SMALL.rn_layers and m.visual.input_resolution == 96 and not m.training
    with pytest.raises(NotImplementedError):
        build_model({"visual.proj": torch.zeros(1)})
    net = CRIS(**cris_model_cfg(CRIS_SMALL, w))
    with pytest.raises(NotImplementedError):
        net(None, None)
    with pytest.raises(ValueError):
        from tunevlseg_b200.models.components.cris_model import FPN
        FPN(in_channels=(1, 2), out_channels=(1, 2, 3))


def test_cris_bicubic_tables_match_torch():
    """The tap tables driving tvs_resample2d reproduce F.interpolate(bicubic, align_corners=True) and its adjoint."""
    from tunevlseg_b200.engine_cris import _bicubic_tables

    for n_in, n_out in ((16, 64), (104, 416), (5, 5)):
        idx, wt, t_idx, t_w, cnt, mt = _bicubic_tables(n_in, n_out, "cpu")
        R = torch.zeros(n_out, n_in)
        for o in range(n_out):
            for a in range(4):
                R[o, idx[o, a]] += wt[o, a]
        x = torch.randn(1, 1, n_in, n_in)
        ref = torch.nn.functional.interpolate(x, (n_out, n_out), mode="bicubic", align_corners=True)[0, 0]
        assert (R @ x[0, 0] @ R.t() - ref).abs().max() < 2e-5
        Rt = torch.zeros(n_in, n_out)
        for i in range(n_in):
            for j in range(int(cnt[i])):
                Rt[i, t_idx[i, j]] += t_w[i, j]
        assert torch.allclose(Rt, R.t(), atol=1e-7)


@pytest.mark.parametrize("case", ["coop_d3", "cocoop"])
def test_cris_engine_composition_against_oracle(case, monkeypatch):
    """The host-side composition of the CRIS engine (operand packing with folded BatchNorm, layouts, strides, flags,
    autograd wiring) run over a CPU emulation of the C ABI (tests/fake_abi.py) must reproduce the oracle to fp32
    round-off - forward and every learner / additive-layer gradient."""
    from oracle import cris as OCR
    from tests import fake_abi
    from tests.helpers import CRIS_SMALL, build_cris_net, cris_oracle_head, cris_oracle_state, make_cris_batch

    fake_abi.install(monkeypatch)
    w = OCR.init_weights(CRIS_SMALL, seed=7)
    net = build_cris_net(case, CRIS_SMALL, w, seed=31)
    st, head = cris_oracle_state(case, net), cris_oracle_head(net)
    img, ids, am, mask = make_cris_batch(CRIS_SMALL, 2, 8, 32)
    logits = net(text_input={"input_ids": ids, "attention_mask": am}, image_input=img)
    ref = OCR.net_forward(w, CRIS_SMALL, st, head, ids, am, img)
    # the emulation rounds GEMM operands to tf32 and the flash-attention operands to bf16, as the device does
    assert (logits - ref).abs().max().item() < 2e-2, (logits - ref).abs().max().item()
    gw = torch.randn(ref.shape, generator=torch.Generator().manual_seed(5))
    (logits * gw).sum().backward()
    (ref * gw).sum().backward()
    named = dict(net.named_parameters())
    for k, p_ref in list(st.params.items()) + list(head.items()):
        pk = k if k in head else f"context_learner.{k}"
        g, g_ref = named[pk].grad, p_ref.grad
        assert g is not None and g_ref is not None, pk
        err = ((g - g_ref).norm() / g_ref.norm()).item()
        assert err < 5e-2, f"{pk}: {err}"


def test_predict_tail_tables_and_metric_formulas():
    """align_corners=False bicubic tap tables == F.interpolate / TF.resize(antialias=False); eval-script formulas."""
    from tunevlseg_b200.engine_cris import _bicubic_tables
    from tunevlseg_b200.scripts.eval_metrics import metrics_from_counts

    for n_in, n_out in ((64, 301), (352, 200), (9, 9), (352, 480)):
        idx, wt, *_ = _bicubic_tables(n_in, n_out, "cpu", align_corners=False)
        R = torch.zeros(n_out, n_in)
        for o in range(n_out):
            for a in range(4):
                R[o, idx[o, a]] += wt[o, a]
        x = torch.randn(1, 1, n_in, n_in)
        ref = torch.nn.functional.interpolate(x, (n_out, n_out), mode="bicubic", align_corners=False)[0, 0]
        assert (R @ x[0, 0] @ R.t() - ref).abs().max() < 1e-4
    m = metrics_from_counts(tp=30, fp=10, fn=20, n_pos_gt=50, n=200)
    assert abs(m["dice"] - 100 * 60 / 90) < 1e-9 and abs(m["iou"] - 50.0) < 1e-9
    assert abs(m["ones_dice_diff"] - (100 * 60 / 90 - 100 * 100 / 250)) < 1e-9
    assert metrics_from_counts(0, 0, 0, 0, 100)["dice"] == 100.0 and metrics_from_counts(0, 5, 0, 0, 100)["iou"] == 0.0


def test_stacked_parameter_gradients_go_straight_into_preallocated_buffers():
    """``_StackParams`` (the per-depth projector parameters stacked for the batched projection) must leave exactly the
    gradients ordinary autograd leaves: into preallocated ``.grad`` views of a flat buffer (FusedAdamW's layout) with
    accumulation semantics, through ordinary AccumulateGrad when there is no buffer, and through ordinary autograd when
    a unified projector lists the same parameter once per depth."""
    import copy

    from torch import nn

    from tunevlseg_b200.models.core_models.coop.context_learner.learners import _batched_projection

    torch.manual_seed(0)

    def mk():
        return nn.Sequential(nn.Linear(8, 4), nn.ReLU(), nn.Linear(4, 6, bias=False), nn.LayerNorm(6))

    x = torch.randn(3, 5, 8)
    plain = nn.ModuleList(mk() for _ in range(3))
    direct = copy.deepcopy(plain)
    (_batched_projection(plain, x) ** 2).sum().backward()           # no .grad yet: falls back to AccumulateGrad
    flat = torch.zeros(sum(p.numel() for p in direct.parameters()))
    o = 0
    for p in direct.parameters():
        p.grad = flat[o:o + p.numel()].view_as(p)
        o += p.numel()
    for rep in (1, 2):
        (_batched_projection(direct, x) ** 2).sum().backward()
        for a, b in zip(direct.parameters(), plain.parameters()):
            assert a.grad.data_ptr() >= flat.data_ptr() and torch.allclose(a.grad, rep * b.grad, rtol=1e-6, atol=1e-7)
    assert flat.abs().sum() > 0

    shared = mk()
    shared_ref = copy.deepcopy(shared)
    for p in shared.parameters():
        p.grad = torch.zeros_like(p)
    (_batched_projection(nn.ModuleList((shared,) * 3), x) ** 2).sum().backward()
    (torch.stack([shared_ref(x[i]) for i in range(3)]) ** 2).sum().backward()
    for a, b in zip(shared.parameters(), shared_ref.parameters()):
        assert torch.allclose(a.grad, b.grad, rtol=2e-2, atol=1e-5)


def test_lightning_checkpoint_shaped_state_dict_round_trip():
    """a reference Lightning checkpoint stores ``state_dict`` keys ``net.model.clip.*``,
    ``net.model.decoder.*``, ``net.context_learner.*``, ``net.additive_decoder_layer.*`` and ``net.residual_ratio``
    (callbacks/default.yaml:9-15 saves the LightningModule, whose only sub-module is ``net``).  The product module must
    produce exactly that key space and load it strictly, so that reference runs drop onto the kernels."""
    from oracle import clipseg as OC
    from tests.golden_cases import TINY
    from tests.helpers import build_net
    from tunevlseg_b200.losses import DiceCELoss
    from tunevlseg_b200.models.image_text_mask_module import ImageTextMaskModule

    def module(seed):
        net = build_net("maple", TINY, OC.init_weights(TINY, seed=7), seed=seed)
        return ImageTextMaskModule(net=net, loss_fn=DiceCELoss(sigmoid=True, lambda_dice=1, lambda_ce=0.2),
                                   optimizer=partial(torch.optim.AdamW, lr=2e-4), scheduler=None, compile=False, task="binary",
                                   threshold=0.5, weight_decay=0.01)

    src, dst = module(1), module(2)
    sd = src.state_dict()
    prefixes = ("net.model.clip.", "net.model.decoder.", "net.context_learner.", "net.additive_decoder_layer.", "net.residual_ratio")
    assert all(k.startswith(prefixes) for k in sd), [k for k in sd if not k.startswith(prefixes)][:5]
    for p in prefixes:
        assert any(k.startswith(p) for k in sd), p
    assert "net.context_learner.context_vectors" in sd and "net.additive_decoder_layer.1.weight" in sd
    assert not torch.equal(dst.state_dict()["net.context_learner.context_vectors"], sd["net.context_learner.context_vectors"])
    ckpt = {"state_dict": {k: v.clone() for k, v in sd.items()}}            # what torch.load(<lightning .ckpt>) returns
    res = dst.load_state_dict(ckpt["state_dict"], strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    for k, v in dst.state_dict().items():
        assert torch.equal(v, sd[k]), k


@pytest.mark.parametrize("case", ["maple", "vpt", "coop", "coop_deep", "cocoop", "shared_attn", "shared_separate"])
def test_clipseg_engine_composition_against_oracle(case, monkeypatch):
    """The host-side composition of the CLIPSeg engine - operand packing (fused QKV, folded scale, split-bf16 FFN
    weights), layouts, flags, the prompt table overwrite / gradient routing, the bottom-block backward restricted to the
    prompt rows (``tvs_attn_bwd_tail`` leaves the other rows unwritten: the emulation poisons them with NaN), the
    stacked-parameter gradient accumulation and the autograd wiring of the three nodes - run over a CPU emulation of the
    C ABI (tests/fake_abi.py) must reproduce the oracle: logits to the emulated operand precision, every learner / head
    gradient, and exactly-zero gradients for the parameters the reference never uses."""
    import contextlib

    from oracle import clipseg as OC
    from tests import fake_abi
    from tests.helpers import SMALL, build_net, make_batch, oracle_head, oracle_state

    fake_abi.install(monkeypatch)

    class _NoStream:                                  # there is no CUDA here: both towers run in line
        def wait_stream(self, other): ...

    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a, **k: _NoStream())
    monkeypatch.setattr(torch.cuda, "stream", lambda s: contextlib.nullcontext())
    monkeypatch.setattr(torch.Tensor, "record_stream", lambda self, s: None, raising=False)
    monkeypatch.setenv("TVS_TEXT_STREAM", "0")

    weights = OC.init_weights(SMALL, seed=7)
    net = build_net(case, SMALL, weights, seed=5)
    st, head = oracle_state(case, net, SMALL), oracle_head(net)
    img, ids, am, mask = make_batch(SMALL, 2, 8, 9)
    logits = net(text_input={"input_ids": ids, "attention_mask": am}, image_input=img)
    ref = OC.net_forward(weights, SMALL, st, head, ids, am, img)
    assert logits.shape == ref.shape
    assert (logits - ref).abs().max().item() < 2e-2, (logits - ref).abs().max().item()
    gw = torch.randn(ref.shape, generator=torch.Generator().manual_seed(5))
    (logits * gw).sum().backward()
    (ref * gw).sum().backward()
    named = dict(net.named_parameters())
    checked = 0
    for k, p_ref in list(st.params.items()) + list(head.items()):
        pk = k if k in head else f"context_learner.{k}"
        if pk not in named:
            continue
        g, g_ref = named[pk].grad, p_ref.grad
        if g_ref is None or g_ref.abs().max() == 0:
            assert g is None or g.abs().max().item() == 0, pk
            continue
        assert g is not None and torch.isfinite(g).all(), pk
        err = ((g - g_ref).norm() / g_ref.norm()).item()
        assert err < 5e-2, f"{pk}: {err}"
        checked += 1
    assert checked > 0


def test_module_training_and_eval_steps_over_cpu_abi_emulation(monkeypatch):
    """``ImageTextMaskModule`` end to end on the CPU ABI emulation: ``training_step`` (fused loss + counters, backward,
    torch AdamW on the learner parameters), ``validation_step`` / ``test_step`` metric accumulation and ``predict_step``,
    against the oracle's loss, Dice (average="samples", >=) and IoU (global confusion matrix, >) on the same logits."""
    import contextlib

    from oracle import clipseg as OC
    from oracle import loss_metrics as OLM
    from tests import fake_abi
    from tests.helpers import SMALL, build_net, make_batch, oracle_head, oracle_state
    from tunevlseg_b200.losses import DiceCELoss
    from tunevlseg_b200.models.image_text_mask_module import ImageTextMaskModule

    fake_abi.install(monkeypatch)

    class _NoStream:
        def wait_stream(self, other): ...

    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a, **k: _NoStream())
    monkeypatch.setattr(torch.cuda, "stream", lambda s: contextlib.nullcontext())
    monkeypatch.setattr(torch.Tensor, "record_stream", lambda self, s: None, raising=False)
    monkeypatch.setenv("TVS_TEXT_STREAM", "0")

    weights = OC.init_weights(SMALL, seed=7)
    net = build_net("maple", SMALL, weights, seed=5)
    st, head = oracle_state("maple", net, SMALL), oracle_head(net)
    module = ImageTextMaskModule(net=net, loss_fn=DiceCELoss(sigmoid=True, lambda_dice=1, lambda_ce=0.2),
                                 optimizer=partial(torch.optim.AdamW, lr=1e-3), scheduler=None, compile=False, task="binary",
                                 threshold=0.5, weight_decay=0.01)
    module.setup("fit")
    module.setup("test")
    opt = module.configure_optimizers()["optimizer"]
    img, ids, am, mask = make_batch(SMALL, 3, 8, 21)
    batch = {"image": img, "mask": mask, "input_ids": ids, "attention_mask": am, "mask_name": ["a", "b", "c"],
             "mask_shape": torch.tensor([[64, 64]] * 3)}

    ref = OC.net_forward(weights, SMALL, st, head, ids, am, img)
    ref_loss = OLM.dice_ce_loss(ref, mask)
    loss = module.training_step(batch, 0)
    assert abs(loss.item() - ref_loss.item()) < 5e-3
    before = {k: p.detach().clone() for k, p in net.named_parameters() if p.requires_grad}
    loss.backward()
    opt.step()
    moved = [k for k, p in net.named_parameters() if p.requires_grad and p.grad is not None and not torch.equal(p, before[k])]
    assert any(k.startswith("context_learner.") for k in moved) and "residual_ratio" in moved

    # metric accumulation: counters of the batch == the oracle's counters on the module's own logits
    module.test_dice.reset(); module.test_iou.reset()
    with torch.no_grad():
        logits = module.get_logits(batch)
        module.test_step(batch, 0)
    _, c_counts, c_conf = OLM.c_dicebce_metrics(logits, mask)
    tp, fp, fn = (c_counts[:, k].double() for k in range(3))
    dice_ref = torch.where(2 * tp + fp + fn > 0, 2 * tp / (2 * tp + fp + fn), torch.ones_like(tp)).mean()
    tn_, fp_, fn_, tp_ = (c_conf.view(-1)[k].double() for k in range(4))
    iou_ref = tp_ / (tp_ + fp_ + fn_) if (tp_ + fp_ + fn_) > 0 else torch.tensor(1.0)
    assert abs(module.test_dice.compute().item() - dice_ref.item()) < 1e-6
    assert abs(module.test_iou.compute().item() - iou_ref.item()) < 1e-6
    with torch.no_grad():
        out = module.predict_step(batch, 0)
    assert out["preds"].shape == (3, 1, SMALL.image_size, SMALL.image_size) and out["mask_name"] == ["a", "b", "c"]
    assert float(out["preds"].min()) >= 0.0 and float(out["preds"].max()) <= 1.0


def test_cris_distributed_checkpoint_conversion_and_strict_load(tmp_path):
    """scripts/process_cris_checkpoint.py (reference :5-25) -> ``CRIS(cris_pretrain=...)`` strict load (reference
    cris_model/__init__.py:64-69): a DDP-saved ``{"state_dict": {"module.<key>": ...}}`` file becomes the single-process
    state dict and restores every CRIS tensor."""
    from oracle import cris as OCR
    from tests.helpers import CRIS_SMALL, cris_model_cfg
    from tunevlseg_b200.models.components.cris_model import CRIS
    from tunevlseg_b200.scripts import process_cris_checkpoint as P

    w = OCR.init_weights(CRIS_SMALL, seed=11)
    src = CRIS(**cris_model_cfg(CRIS_SMALL, w))
    with torch.no_grad():
        for p in src.neck.parameters():
            p.add_(0.25)
    ddp_path, single_path = tmp_path / "cris_best.pth", tmp_path / "cris_best_single.pth"
    torch.save({"state_dict": {"module." + k: v for k, v in src.state_dict().items()}, "epoch": 3}, ddp_path)
    P.main(ddp_path, single_path, prefix="model.", pickle_protocol=5)
    from tunevlseg_b200.models.components.cris_model import load_checkpoint

    converted = load_checkpoint(single_path)             # protocol-5 file: the weights-only unpickler alone rejects it
    assert set(converted) == set(src.state_dict())
    dst = CRIS(**dict(cris_model_cfg(CRIS_SMALL, OCR.init_weights(CRIS_SMALL, seed=12)), cris_pretrain=str(single_path)))
    for k, v in src.state_dict().items():
        assert torch.equal(dst.state_dict()[k], v), k
    with pytest.raises(ValueError):
        P.convert({"short": torch.zeros(1)})


# ---- round-2 drop-in fixes -------------------------------------------------------------------------------------------
def test_epoch_metrics_are_computed_over_the_accumulated_state_and_reset():
    """The reference logs the torchmetrics objects (image_text_mask_module.py:118-125,144-170): Lightning reports
    ``compute()`` over the epoch's accumulated state and resets afterwards.  Two unequal batches: the epoch IoU is the IoU
    of the SUMMED confusion matrix (not a mean of per-batch IoUs), Dice the mean over all 5 samples; the next epoch starts
    from an empty state."""
    from tunevlseg_b200.losses import DiceCELoss
    from tunevlseg_b200.metrics import Dice, JaccardIndex
    from tunevlseg_b200.models.image_text_mask_module import ImageTextMaskModule

    module = ImageTextMaskModule(net=torch.nn.Linear(1, 1), loss_fn=DiceCELoss(sigmoid=True, lambda_dice=1, lambda_ce=0.2),
                                 optimizer=None, scheduler=None, compile=False, task="binary", threshold=0.5)
    module.setup("fit")
    c1 = torch.tensor([[10, 2, 3], [0, 0, 0], [5, 5, 5]])
    c2 = torch.tensor([[1, 9, 0], [7, 0, 1]])
    f1, f2 = torch.tensor([100, 7, 8, 15]), torch.tensor([10, 9, 1, 8])
    for stage, keys in (("val", ("val_dice", "val_iou")), ("train", ("train_dice_epoch", "train_iou_epoch"))):
        dice, iou = getattr(module, f"{stage}_dice"), getattr(module, f"{stage}_iou")
        for c, f in ((c1, f1), (c2, f2)):
            dice.update_from_counts(c)
            iou.update_from_confmat(f)
        getattr(module, f"on_{'validation' if stage == 'val' else stage}_epoch_end")()
        want_dice = Dice.score(torch.cat((c1, c2)), 1.0)
        want_iou = JaccardIndex.score(f1 + f2, 1.0)
        assert abs(float(module.logged[keys[0]]) - float(want_dice)) < 1e-6
        assert abs(float(module.logged[keys[1]]) - float(want_iou)) < 1e-6
        mean_of_batches = (JaccardIndex.score(f1, 1.0) + JaccardIndex.score(f2, 1.0)) / 2
        assert abs(float(want_iou) - float(mean_of_batches)) > 1e-3          # the two notions really differ on this data
        assert not dice.has_updates and int(iou.confmat.sum()) == 0          # reset for the next epoch
        module.logged.clear()
        getattr(module, f"on_{'validation' if stage == 'val' else stage}_epoch_end")()   # nothing accumulated: logs nothing, no error
        assert not module.logged


def test_install_as_src_aliases_monai_dice_ce_loss():
    """configs/model/maple_clipseg.yaml:29-33 ``_target_: monai.losses.DiceCELoss`` must resolve to the fused loss, so the
    module takes the one-pass loss + counters branch with the stock YAML."""
    import importlib

    import tunevlseg_b200
    from tunevlseg_b200.losses import DiceCELoss

    tunevlseg_b200.install_as_src()
    cls = getattr(importlib.import_module("monai.losses"), "DiceCELoss")
    assert cls is DiceCELoss
    loss = cls(sigmoid=True, lambda_dice=1, lambda_ce=0.2)       # the reference's kwargs
    assert isinstance(loss, DiceCELoss)


def test_load_checkpoint_rejects_pickles_that_name_foreign_globals(tmp_path, monkeypatch):
    """A protocol-5 checkpoint is re-read with an unpickler restricted to torch's weights-only allow-list; a file that
    references any other callable (the malicious case) must fail instead of being fully unpickled."""
    import pickle

    from tunevlseg_b200.models.components.cris_model import load_checkpoint

    class Evil:
        def __reduce__(self):
            return (print, ("arbitrary code ran",))

    good, bad = tmp_path / "good.pth", tmp_path / "bad.pth"
    torch.save({"w": torch.arange(6.0).view(2, 3)}, good, pickle_protocol=5)
    torch.save({"w": Evil()}, bad, pickle_protocol=5)
    assert torch.equal(load_checkpoint(good)["w"], torch.arange(6.0).view(2, 3))
    monkeypatch.delenv("TVS_ALLOW_UNSAFE_PICKLE", raising=False)
    with pytest.raises(pickle.UnpicklingError):
        load_checkpoint(bad)
    assert load_checkpoint(bad, allow_unsafe_pickle=True)["w"] is None      # explicit opt-in: print(...) ran and returned None


@pytest.mark.parametrize("case", ["maple", "coop"])
def test_no_freeze_last_layer_composition_over_cpu_abi_emulation(case, monkeypatch):
    """``no_freeze_last_layer=True`` without the additive layer (base_clipseg.py:73-80): only the transposed convolution of the
    frozen model trains; its weight / bias gradients (one TN GEMM over the patches in DecoderFn.backward) and the learner
    gradients against autograd through the oracle - over the CPU emulation of the C ABI."""
    import contextlib
    from functools import partial

    import tunevlseg_b200.models.core_models.coop as nets
    import tunevlseg_b200.models.core_models.coop.context_learner as learners
    from oracle import clipseg as OC
    from tests import fake_abi
    from tests.helpers import LEARNER_CASES, SMALL, hf_model, make_batch, oracle_state

    fake_abi.install(monkeypatch)

    class _NoStream:
        def wait_stream(self, other): ...

    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a, **k: _NoStream())
    monkeypatch.setattr(torch.cuda, "stream", lambda s: contextlib.nullcontext())
    monkeypatch.setattr(torch.Tensor, "record_stream", lambda self, s: None, raising=False)
    monkeypatch.setenv("TVS_TEXT_STREAM", "0")

    weights = OC.init_weights(SMALL, seed=7)
    c = LEARNER_CASES[case]
    torch.manual_seed(3)
    net = getattr(nets, c["cls"])(
        model_cfg=dict(pretrained_model_name_or_path=hf_model(SMALL, weights), freeze_encoder=False, freeze_decoder=False),
        context_learner=partial(getattr(learners, c["learner"]), **dict(c["kw"])), freeze_all=True, no_freeze_last_layer=True,
        use_new_last_layer=False, new_last_layer_kernel_size=5, residual_ratio=0.5)
    net.context_learner.eval()
    trainable = {k for k, p in net.named_parameters() if p.requires_grad}
    assert {"model.decoder.transposed_convolution.weight", "model.decoder.transposed_convolution.bias"} <= trainable
    assert not any(k.startswith("model.") and "transposed_convolution" not in k for k in trainable)
    st = oracle_state(case, net, SMALL)
    w = dict(weights)
    for k in ("decoder.transposed_convolution.weight", "decoder.transposed_convolution.bias"):
        w[k] = weights[k].detach().clone().requires_grad_(True)
    img, ids, am, _ = make_batch(SMALL, 2, 8, 9)
    logits = net(text_input={"input_ids": ids, "attention_mask": am}, image_input=img)
    ref = OC.net_forward(w, SMALL, st, None, ids, am, img)
    assert (logits - ref).abs().max().item() < 2e-2
    gw = torch.randn(ref.shape, generator=torch.Generator().manual_seed(5))
    (logits * gw).sum().backward()
    (ref * gw).sum().backward()
    tc = net.model.decoder.transposed_convolution
    for g, g_ref in ((tc.weight.grad, w["decoder.transposed_convolution.weight"].grad), (tc.bias.grad, w["decoder.transposed_convolution.bias"].grad)):
        assert g is not None and g.shape == g_ref.shape
        assert ((g - g_ref).norm() / g_ref.norm()).item() < 5e-2, ((g - g_ref).norm() / g_ref.norm()).item()
    named = dict(net.named_parameters())
    for k, p_ref in st.params.items():
        pk = f"context_learner.{k}"
        if pk in named and p_ref.grad is not None and p_ref.grad.abs().max() > 0:
            assert ((named[pk].grad - p_ref.grad).norm() / p_ref.grad.norm()).item() < 5e-2, pk


def test_nvtx_ranges_bracket_library_calls_and_phases():
    """TVS_NVTX=1: every library call and every step phase is bracketed by an NVTX range (pushed and popped even when the call
    raises - here it does, there is no GPU).  Read once at import: exercised in a fresh interpreter."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = r"""
import sys, torch
sys.path.insert(0, %r)
log = []
torch.cuda.nvtx.range_push = lambda name: log.append(("push", name))
torch.cuda.nvtx.range_pop = lambda: log.append(("pop", None))
from tunevlseg_b200 import abi, engine
assert abi.NVTX
try:
    abi.cast_bf16(torch.zeros(4, 4), torch.zeros(4, 4, dtype=torch.bfloat16))
except abi.TvsError:
    pass
with abi.nvtx_range("vision_tower.fwd"):
    pass
assert log == [("push", "tvs.cast_bf16"), ("pop", None), ("push", "vision_tower.fwd"), ("pop", None)], log
assert engine.VisionTowerFn.forward.__name__ == "forward"
print("ok")
""" % root
    env = dict(os.environ, TVS_NVTX="1")
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
    assert res.returncode == 0 and "ok" in res.stdout, res.stderr[-2000:]
