"""Pin the oracle to outputs of the REAL reference classes (fixtures made by tests/golden/make_golden.py)."""
import pytest
import torch

from oracle import clipseg as OC
from oracle import cris as OCR
from tests.golden_cases import (CASES, CRIS_CASES, CRIS_TINY, TINY, cris_learner_state, learner_state, load_case,
                                load_cris_weights, load_weights)

TOL = 1e-5


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_matches_reference(name):
    w = load_weights()
    d, learner, head = load_case(name)
    learner = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in learner.items()}
    head = {k: v.clone().requires_grad_(True) for k, v in head.items()}
    st = learner_state(name, learner)
    logits = OC.net_forward(w, TINY, st, head, d["input_ids"], d["attention_mask"], d["image"])
    assert logits.shape == d["logits"].shape
    err = (logits - d["logits"]).abs().max().item()
    assert err <= TOL, f"logits max-abs {err}"
    (logits * d["grad_weight"]).sum().backward()
    # reference state_dict lists shared (unified) projector tensors once per depth; parameters() dedups.
    for k, v in d.items():
        if k.startswith("learner_grad/"):
            pk = k[len("learner_grad/"):]
            has = bool(d[f"learner_hasgrad/{pk}"])
            g = learner[pk].grad
            if not has:
                assert g is None or g.abs().max() == 0, f"{pk}: reference gives no grad"
                continue
            # unified projection: the reference accumulates all depths into the one shared tensor
            if CASES[name].get("proj_style") == "lora" and "projection_layers." in pk:
                tail = pk.split(".", 2)[2]
                g = sum(learner[k2].grad for k2 in learner if k2.startswith("projection_layers.") and
                        k2.split(".", 2)[2] == tail and learner[k2].grad is not None)
            scale = max(1.0, v.abs().max().item())
            gerr = (g - v).abs().max().item() / scale
            assert gerr <= 5e-5, f"{pk}: grad rel err {gerr}"
        if k.startswith("head_grad/"):
            pk = k[len("head_grad/"):]
            has = bool(d[f"head_hasgrad/{pk}"])
            g = head[pk].grad
            if not has:
                assert g is None or g.abs().max() == 0, f"{pk}: reference gives no grad"
            else:
                scale = max(1.0, v.abs().max().item())
                assert (g - v).abs().max().item() / scale <= 5e-5, pk


def test_known_quirks():
    """Reference control-flow quirks the oracle (and the product) must keep."""
    d, _, _ = load_case("vpt_d12_n3")
    # VPT depth 12 with the 10-layer early exit: ctx[0] is concatenated, ctx[1..10] overwrite after layers
    # 1..10 (the write after layer 10 still reaches the decoder through the last tap) -> only ctx[11] is dead.
    g = d["learner_grad/context_vectors"]
    assert g[11:].abs().max() == 0 and all(g[i].abs().max() > 0 for i in range(11))
    # VPT ignores residual_ratio; CoOp ignores the whole additive layer
    assert not bool(d["head_hasgrad/residual_ratio"])
    d, _, _ = load_case("coop_d1_n4")
    assert not bool(d["head_hasgrad/additive_decoder_layer.1.weight"])
    assert not bool(d["head_hasgrad/residual_ratio"])


@pytest.mark.parametrize("name", list(CRIS_CASES))
def test_cris_oracle_matches_reference(name):
    """oracle/cris.py against the reference's own COOPCRIS (fixtures from tests/golden/make_golden_cris.py)."""
    w = load_cris_weights()
    d, learner, head = load_case(name)
    learner = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in learner.items()}
    head = {k: v.clone().requires_grad_(True) for k, v in head.items()}
    st = cris_learner_state(name, learner)
    am = d["attention_mask"] if bool(d["has_mask"]) else None
    logits = OCR.net_forward(w, CRIS_TINY, st, head, d["input_ids"], am, d["image"])
    assert logits.shape == d["logits"].shape
    err = (logits - d["logits"]).abs().max().item()
    assert err <= 2e-5, f"logits max-abs {err}"
    (logits * d["grad_weight"]).sum().backward()
    checked = 0
    for k, v in d.items():
        for kind, store in (("learner_grad/", learner), ("head_grad/", head)):
            if not k.startswith(kind):
                continue
            pk = k[len(kind):]
            has = bool(d[f"{kind[:-6]}_hasgrad/{pk}"])
            g = store[pk].grad
            if not has:
                assert g is None or g.abs().max() == 0, f"{pk}: reference gives no grad"
                continue
            scale = max(1.0, v.abs().max().item())
            assert (g - v).abs().max().item() / scale <= 5e-5, f"{pk}: grad rel err {(g - v).abs().max().item() / scale}"
            checked += 1
    assert checked >= 5


def test_cris_known_quirks():
    # depth 1: the rows inserted from ctx[0] are re-written with the same ctx[0] after block 0 (0-based overwrite,
    # coop_cris.py:129-141), so ctx[0] gets gradient through both uses; every additive-layer tensor is live.
    d, _, _ = load_case("cris_coop_d1_n4")
    assert bool(d["learner_hasgrad/context_vectors"]) and d["learner_grad/context_vectors"].abs().max() > 0
    for k in ("additive_decoder_layer.0.weight", "additive_decoder_layer.2.weight", "additive_decoder_layer.2.bias", "residual_ratio"):
        assert bool(d[f"head_hasgrad/{k}"])


def test_built_reference_archive_imports_without_sources():
    """bench.py's reference arm on the GPU box: /root/reference is absent there and file-sync tools may drop *.pyc, so
    oracle/build_ref.py also packs the byte-compiled reference into oracle/_ref/reference_src.zip; the reference's own
    MapleCLIPSeg must import from that archive alone (zipimport, sourceless members, packages without __init__)."""
    import os
    import subprocess
    import sys

    import pytest

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    archive = os.path.join(root, "oracle", "_ref", "reference_src.zip")
    if not os.path.exists(archive):
        pytest.skip("oracle/_ref not built (python oracle/build_ref.py needs /root/reference)")
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r); from oracle import ref_shim; ref_shim.install_shim(); "
            "import src.models.core_models.coop as m; print(m.MapleCLIPSeg.__name__, m.__spec__.origin)") % (root, archive)
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd="/tmp", timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    assert "MapleCLIPSeg" in res.stdout and "reference_src.zip" in res.stdout, res.stdout
