"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.

Bars (BASELINE.json north_star): logits within 2e-2 max-abs of the fp32 oracle (bf16 compute, fp32 accumulate);
integer TP/FP/FN / confusion-matrix counters bit-exact; prompt / head gradients within bf16 tolerance; parameters
the reference never uses get exactly no gradient.
"""
import pytest
import torch

from oracle import clipseg as OC
from oracle import loss_metrics as OLM
from tests.helpers import FULL, LEARNER_CASES, SMALL, build_net, make_batch, oracle_head, oracle_state

pytestmark = pytest.mark.gpu

LOGIT_TOL = 2e-2        # max-abs, stated by north_star
GRAD_TOL = 8e-2         # max-abs error relative to the largest reference gradient entry (bf16 operands, 10-12 layers)
GRAD_L2_TOL = 4e-2      # ||g - g_ref||_2 / ||g_ref||_2


def _run_case(case, spec, B, L, seed, weights=None, logit_tol=LOGIT_TOL):
    from tunevlseg_b200.losses import DiceCELoss

    weights = weights or OC.init_weights(spec, seed=7)
    net = build_net(case, spec, weights, seed=seed)
    st, head = oracle_state(case, net, spec), oracle_head(net)
    img, ids, am, mask = make_batch(spec, B, L, seed + 1)

    net = net.cuda()
    loss_fn = DiceCELoss(sigmoid=True, lambda_dice=1, lambda_ce=0.2)
    logits = net(text_input={"input_ids": ids.cuda(), "attention_mask": am.cuda()}, image_input=img.cuda())
    conf = torch.zeros(4, dtype=torch.int64, device="cuda")
    loss, counts = loss_fn.forward_with_metrics(logits, mask.cuda(), 0.5, conf)
    loss.backward()
    torch.cuda.synchronize()

    ref = OC.net_forward(weights, spec, st, head, ids, am, img)
    ref_loss = OLM.dice_ce_loss(ref, mask)
    ref_loss.backward()

    assert logits.shape == ref.shape == (B, 1, spec.image_size, spec.image_size)
    err = (logits.detach().cpu() - ref.detach()).abs().max().item()
    ref_max = ref.detach().abs().max().item()
    print(f"PARITY clipseg {case} B={B} {spec.image_size}px: logits max-abs err {err:.5f} (tol {logit_tol:.4f}, |logit|max {ref_max:.2f})")
    assert err <= logit_tol, f"{case}: logits max-abs err {err:.4f} > {logit_tol} (|logit|max {ref_max:.2f})"
    assert abs(loss.item() - ref_loss.item()) <= 5e-3, f"{case}: loss {loss.item()} vs {ref_loss.item()}"

    # integer counters: bit-exact against the C oracle evaluated on the SAME (GPU) logits
    _, c_counts, c_conf = OLM.c_dicebce_metrics(logits.detach().cpu(), mask)
    assert torch.equal(counts.cpu(), c_counts), f"{case}: per-sample tp/fp/fn differ"
    assert torch.equal(conf.cpu().view(2, 2), c_conf), f"{case}: confusion matrix differs"

    # gradients of every learner / head parameter
    named = dict(net.named_parameters())
    checked = 0
    for k, p_ref in list(st.params.items()) + list(head.items()):
        pk = k if k in head else f"context_learner.{k}"
        if pk not in named:     # unified projections list one tensor under several keys
            continue
        g = named[pk].grad
        g_ref = p_ref.grad
        if g_ref is None or g_ref.abs().max() == 0:
            assert g is None or g.abs().max().item() == 0, f"{case}: {pk} must get no gradient (reference quirk)"
            continue
        assert g is not None, f"{case}: {pk} got no gradient"
        scale = g_ref.abs().max().item()
        gerr = (g.detach().cpu() - g_ref).abs().max().item() / scale
        assert gerr <= GRAD_TOL, f"{case}: grad {pk} rel err {gerr:.4f}"
        l2 = ((g.detach().cpu() - g_ref).norm() / g_ref.norm()).item()
        assert l2 <= GRAD_L2_TOL, f"{case}: grad {pk} L2 rel err {l2:.4f}"
        checked += 1
    assert checked > 0
    return err


@pytest.mark.parametrize("case", list(LEARNER_CASES))
def test_small_geometry_all_learners(case):
    _run_case(case, SMALL, B=3, L=9, seed=11)


def test_small_long_prompt_truncation():
    # L + n > 77: the learner truncates to 77 positions and the EOS pooling index is clamped to 76
    _run_case("coop_deep", SMALL, B=2, L=76, seed=5)


@pytest.mark.parametrize("case", ["maple", "vpt", "coop", "shared_separate", "shared_attn", "cocoop", "coop_deep"])
def test_full_geometry(case):
    """ViT-B/16 @ 352^2 (the BASELINE.json geometry), B=2 - the oracle still finishes in seconds.  All learner families."""
    _run_case(case, FULL, B=2, L=8, seed=3)


def test_eval_counters_large_batch():
    """cfg5-style eval: batch 256 @ 416^2 through the fused loss/metric kernel; size-independent properties:
    tp+fn == sum(target), counts sum to N, IoU/Dice of identical pred/target == 1."""
    from tunevlseg_b200 import engine

    B, N = 256, 416 * 416
    g = torch.Generator(device="cuda").manual_seed(0)
    logits = torch.randn(B, 1, 416, 416, device="cuda", generator=g) * 3
    mask = (torch.rand(B, 1, 416, 416, device="cuda", generator=g) < 0.3).float()
    conf = torch.zeros(4, dtype=torch.int64, device="cuda")
    loss, counts = engine.DiceBceFn.apply(logits, mask, 0.5, 1.0, 0.2, conf)
    tgt = mask.flatten(1).sum(1).long()
    assert torch.equal(counts[:, 0] + counts[:, 2], tgt)
    assert int(conf.sum()) == B * N and int(conf[2] + conf[3]) == int(tgt.sum())
    # Dice thresholds with >=, IoU with >: they differ exactly by the pixels whose fp32 sigmoid is exactly 0.5
    p = torch.sigmoid(logits)
    ties_pos = int(((p == 0.5) & (mask > 0)).sum())
    assert int(counts[:, 0].sum()) - int(conf[3]) == ties_pos
    perfect = (mask * 2 - 1) * 20
    conf2 = torch.zeros(4, dtype=torch.int64, device="cuda")
    _, c2 = engine.DiceBceFn.apply(perfect, mask, 0.5, 1.0, 0.2, conf2)
    assert int(c2[:, 1].sum()) == 0 and int(c2[:, 2].sum()) == 0 and int(conf2[1] + conf2[2]) == 0
    assert torch.isfinite(loss)


@pytest.mark.parametrize("case", ["maple", "vpt"])
def test_bottom_block_pruning_matches_full_backward(case, monkeypatch):
    """Below the first block only the prompt rows carry a gradient (the patch / class embeddings are frozen), so the
    engine runs that block's attention backward, QKV dgrad and LayerNorm backward on those rows only
    (tvs_attn_bwd_tail).  Row-wise arithmetic is unchanged up to the tile shape of the small QKV dgrad GEMM (its fp32
    accumulation order may differ in the last bit before the bf16 rounding), so every learner gradient must agree with the
    full-row backward to ~1e-3 of its largest entry - a dropped row or tile would show up as an O(1) difference."""
    from tunevlseg_b200 import engine
    from tunevlseg_b200.losses import DiceCELoss

    if case not in LEARNER_CASES:
        pytest.skip(f"{case} not among the learner cases")
    spec, B, L = FULL, 2, 8
    weights = OC.init_weights(spec, seed=7)
    img, ids, am, mask = make_batch(spec, B, L, 11)
    grads = []
    for prune in (True, False):
        monkeypatch.setattr(engine, "TAIL_PRUNE", prune)
        net = build_net(case, spec, weights, seed=3).cuda()
        logits = net(text_input={"input_ids": ids.cuda(), "attention_mask": am.cuda()}, image_input=img.cuda())
        loss = DiceCELoss(sigmoid=True, lambda_dice=1, lambda_ce=0.2)(logits, mask.cuda())
        loss.backward()
        torch.cuda.synchronize()
        grads.append({k: p.grad.detach().clone() for k, p in net.named_parameters() if p.grad is not None})
    assert grads[0].keys() == grads[1].keys() and len(grads[0]) > 0
    worst = 0.0
    for k in grads[0]:
        scale = grads[1][k].abs().max().item()
        diff = (grads[0][k] - grads[1][k]).abs().max().item()
        worst = max(worst, diff / max(scale, 1e-30))
        assert diff <= 2e-3 * scale + 1e-12, f"{case}: {k} differs between pruned and full bottom-block backward ({diff:.3e} of {scale:.3e})"
    print(f"PARITY pruning {case}: worst relative difference {worst:.2e}")


@pytest.mark.parametrize("case", ["maple", "coop", "vpt"])
def test_no_freeze_last_layer_trains_the_transposed_convolution(case):
    """base_clipseg.py:73-80 (``no_freeze_last_layer=True`` without the additive layer): the decoder's transposed convolution
    trains with the prompts.  Logits against the oracle, and the weight / bias gradients of the transposed convolution (one
    small GEMM in DecoderFn.backward) plus the learner gradients against autograd through the oracle."""
    import tunevlseg_b200.models.core_models.coop as nets
    import tunevlseg_b200.models.core_models.coop.context_learner as learners
    from functools import partial

    from tests.helpers import hf_model
    from tunevlseg_b200.losses import DiceCELoss

    spec, B, L, seed = SMALL, 3, 9, 21
    weights = OC.init_weights(spec, seed=7)
    c = LEARNER_CASES[case]
    torch.manual_seed(seed)
    net = getattr(nets, c["cls"])(
        model_cfg=dict(pretrained_model_name_or_path=hf_model(spec, weights), freeze_encoder=False, freeze_decoder=False),
        context_learner=partial(getattr(learners, c["learner"]), **dict(c["kw"])), freeze_all=True, no_freeze_last_layer=True,
        use_new_last_layer=False, new_last_layer_kernel_size=5, residual_ratio=0.5)
    net.context_learner.eval()
    with torch.no_grad():
        for p in net.context_learner.parameters():
            p.add_(torch.randn_like(p) * 0.05)
    trainable = {k for k, p in net.named_parameters() if p.requires_grad}
    assert {"model.decoder.transposed_convolution.weight", "model.decoder.transposed_convolution.bias"} <= trainable
    assert not any(k.startswith("model.") and "transposed_convolution" not in k for k in trainable)
    st = oracle_state(case, net, spec)
    w = dict(weights)
    for k in ("decoder.transposed_convolution.weight", "decoder.transposed_convolution.bias"):
        w[k] = weights[k].detach().clone().requires_grad_(True)
    img, ids, am, mask = make_batch(spec, B, L, seed + 1)

    net = net.cuda()
    logits = net(text_input={"input_ids": ids.cuda(), "attention_mask": am.cuda()}, image_input=img.cuda())
    conf = torch.zeros(4, dtype=torch.int64, device="cuda")
    loss, _ = DiceCELoss(sigmoid=True, lambda_dice=1, lambda_ce=0.2).forward_with_metrics(logits, mask.cuda(), 0.5, conf)
    loss.backward()
    torch.cuda.synchronize()
    ref = OC.net_forward(w, spec, st, None, ids, am, img)
    OLM.dice_ce_loss(ref, mask).backward()
    err = (logits.detach().cpu() - ref.detach()).abs().max().item()
    assert err <= LOGIT_TOL, f"{case}: logits max-abs err {err:.4f}"
    tc = net.model.decoder.transposed_convolution
    for name, g, g_ref in (("weight", tc.weight.grad, w["decoder.transposed_convolution.weight"].grad),
                           ("bias", tc.bias.grad, w["decoder.transposed_convolution.bias"].grad)):
        assert g is not None and g.shape == g_ref.shape, name
        rel = (g.detach().cpu() - g_ref).abs().max().item() / g_ref.abs().max().item()
        l2 = ((g.detach().cpu() - g_ref).norm() / g_ref.norm()).item()
        print(f"PARITY no_freeze_last_layer {case}: transposed_convolution.{name} grad rel err {rel:.4f} (L2 {l2:.4f})")
        assert rel <= GRAD_TOL and l2 <= GRAD_L2_TOL, f"{case}: transposed_convolution.{name} grad rel err {rel:.4f} / L2 {l2:.4f}"
    named = dict(net.named_parameters())
    for k, p_ref in st.params.items():
        pk = f"context_learner.{k}"
        if pk not in named or p_ref.grad is None or p_ref.grad.abs().max() == 0:
            continue
        l2 = ((named[pk].grad.detach().cpu() - p_ref.grad).norm() / p_ref.grad.norm()).item()
        assert l2 <= GRAD_L2_TOL, f"{case}: grad {pk} L2 rel err {l2:.4f}"
