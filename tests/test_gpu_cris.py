"""GPU parity of the CRIS path (COOPCRIS: CLIP-RN50 + CoOp / CoCoOp prompts): kernels against fp32 torch references,
then the whole net against the CPU oracle (oracle/cris.py, pinned to the reference's COOPCRIS by tests/golden).

Bars as for CLIPSeg (BASELINE.json north_star): logits within 2e-2 max-abs (or one bf16 ulp of the largest logit) of
the fp32 oracle, TP/FP/FN counters bit-exact on the same logits, prompt / meta-net / additive-layer gradients within
bf16 tolerance."""

import pytest
import torch
import torch.nn.functional as F

from oracle import cris as OCR
from oracle import loss_metrics as OLM
from tests.helpers import (CRIS_CASES, CRIS_FULL, CRIS_SMALL, build_cris_net, cris_oracle_head, cris_oracle_state,
                           make_cris_batch)

pytestmark = pytest.mark.gpu

LOGIT_TOL = 2e-2
GRAD_TOL = 8e-2
GRAD_L2_TOL = 4e-2


@pytest.fixture(autouse=True)
def _fp32_torch_reference():
    """The torch references below must be true fp32: cuDNN / cuBLAS default to TF32 for convolutions."""
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def _rel(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()


# ---- kernels against torch -------------------------------------------------------------------------------------------
def _nhwc(x):           # (B,C,H,W) -> [B*H*W, C]
    B, C, H, W = x.shape
    return x.permute(0, 2, 3, 1).reshape(B * H * W, C).contiguous()


def _nchw(x2, B, H, W):
    return x2.view(B, H, W, -1).permute(0, 3, 1, 2)


@pytest.mark.parametrize("cin,cout,k,hw", [(64, 128, 3, 13), (514, 64, 3, 8), (96, 64, 1, 9), (128, 256, 3, 26)])
def test_conv_fwd_dgrad_tf32(cin, cout, k, hw):
    from tunevlseg_b200 import engine_cris as E

    g = torch.Generator(device="cuda").manual_seed(cin + k)
    B = 2
    sd = {"c.0.weight": torch.randn(cout, cin, k, k, device="cuda", generator=g) * (cin * k * k) ** -0.5,
          "c.1.weight": 1 + 0.1 * torch.randn(cout, device="cuda", generator=g), "c.1.bias": 0.1 * torch.randn(cout, device="cuda", generator=g),
          "c.1.running_mean": 0.1 * torch.randn(cout, device="cuda", generator=g), "c.1.running_var": 1 + torch.rand(cout, device="cuda", generator=g)}
    op = E.ConvOp(sd, "c.0.weight", "c.1", torch.float32, True)
    x = torch.randn(B, cin, hw, hw, device="cuda", generator=g)
    x2 = _nhwc(x).requires_grad_(True)
    y2 = E.conv(x2, op, B, hw, hw)
    xr = x.clone().requires_grad_(True)
    pre = F.batch_norm(F.conv2d(xr, sd["c.0.weight"], padding=k // 2), sd["c.1.running_mean"], sd["c.1.running_var"],
                       sd["c.1.weight"], sd["c.1.bias"], False, 0.0, 1e-5)
    ref = F.relu(pre)
    assert _rel(_nchw(y2, B, hw, hw), ref) < 2e-3                 # tf32 operands
    gy = torch.randn(ref.shape, device="cuda", generator=g)
    # ReLU mask taken from OUR output: pre-activations within tf32 round-off of zero may legitimately flip, and one
    # flipped tap moves dx by ~1/sqrt(cin k^2) of its scale
    (pre * (_nchw(y2, B, hw, hw).detach() > 0)).backward(gy)
    y2.backward(_nhwc(gy))
    cx = cin // 4 * 4
    assert _rel(_nchw(x2.grad, B, hw, hw)[:, :cx], xr.grad[:, :cx]) < 3e-3


def test_pool_and_upsample():
    from tunevlseg_b200 import abi
    from tunevlseg_b200 import engine_cris as E

    g = torch.Generator(device="cuda").manual_seed(5)
    B, C, H, W = 3, 24, 10, 6
    x = torch.randn(B, C, H, W, device="cuda", generator=g)
    for dt, tol in ((torch.float32, 1e-6), (torch.bfloat16, 1e-2)):
        xi = _nhwc(x).to(dt)
        y = torch.empty(B * (H // 2) * (W // 2), C, device="cuda", dtype=dt)
        abi.avgpool2_nhwc(xi, B, H, W, C, y)
        assert _rel(_nchw(y.float(), B, H // 2, W // 2), F.avg_pool2d(xi.float().view(B, H, W, C).permute(0, 3, 1, 2), 2)) < tol
    x2 = _nhwc(x).requires_grad_(True)
    up = E.upsample2x(x2, B, H, W)
    xr = x.clone().requires_grad_(True)
    ref = F.interpolate(xr, scale_factor=2, mode="bilinear")
    assert _rel(_nchw(up, B, 2 * H, 2 * W), ref) < 1e-6
    gy = torch.randn(ref.shape, device="cuda", generator=g)
    ref.backward(gy)
    # gradient arriving as a column slice of a wider (concatenated) buffer
    wide = torch.zeros(B * 4 * H * W, C + 8, device="cuda")
    wide[:, 8:] = _nhwc(gy)
    up.backward(wide[:, 8:])
    assert _rel(_nchw(x2.grad, B, H, W), xr.grad) < 1e-5


@pytest.mark.parametrize("Sq,Sk,H,causal", [(676, 12, 8, False), (16, 77, 2, False), (130, 80, 1, False), (77, 77, 8, True), (12, 12, 2, True)])
def test_short_key_attention(Sq, Sk, H, causal):
    from tunevlseg_b200 import abi

    g = torch.Generator(device="cuda").manual_seed(Sq)
    B, hd = 2, 64
    D = H * hd
    q = torch.randn(B * Sq, D, device="cuda", generator=g) * 0.3
    kv = torch.randn(B * Sk, 2 * D, device="cuda", generator=g)
    km = torch.ones(B, Sk, dtype=torch.uint8, device="cuda")
    km[1, Sk // 2:] = 0
    out = torch.empty(B * Sq, D, device="cuda")
    lse = torch.empty(B, H, Sq, device="cuda")
    abi.cross_attn_fwd(q, kv[:, :D], kv[:, D:], km, B, Sq, Sk, H, hd, out, lse, causal=causal)
    qr, kr, vr = (t.clone().requires_grad_(True) for t in (q, kv[:, :D].contiguous(), kv[:, D:].contiguous()))
    s = qr.view(B, Sq, H, hd).transpose(1, 2) @ kr.view(B, Sk, H, hd).transpose(1, 2).transpose(-1, -2)
    s = s.masked_fill(km[:, None, None, :] == 0, float("-inf"))
    if causal:
        s = s.masked_fill(torch.triu(torch.ones(Sq, Sk, dtype=torch.bool, device="cuda"), 1), float("-inf"))
    ref = (torch.softmax(s, -1) @ vr.view(B, Sk, H, hd).transpose(1, 2)).transpose(1, 2).reshape(B * Sq, D)
    assert _rel(out, ref) < 6e-4                       # the output is rounded to nearest tf32 (it feeds a tf32 GEMM)
    assert _rel(lse, torch.logsumexp(s, -1)) < 1e-5
    go = torch.randn(B * Sq, D, device="cuda", generator=g)
    ref.backward(go)
    dq, dkv = torch.empty_like(q), torch.empty_like(kv)
    delta = torch.empty_like(lse)
    abi.cross_attn_bwd(q, kv[:, :D], kv[:, D:], km, out, go, lse, B, Sq, Sk, H, hd, dq, dkv[:, :D], dkv[:, D:], delta, causal=causal)
    assert _rel(dq, qr.grad) < 1e-3 and _rel(dkv[:, :D], kr.grad) < 1e-3 and _rel(dkv[:, D:], vr.grad) < 1e-3
    assert dkv[Sk + Sk // 2:, :].abs().max() == 0        # padded keys of sample 1 get exactly no gradient


def test_dynamic_conv():
    from tunevlseg_b200 import engine_cris as E

    g = torch.Generator(device="cuda").manual_seed(9)
    B, C, H, W = 3, 64, 20, 12
    x = torch.randn(B, C, H, W, device="cuda", generator=g)
    word = torch.randn(B, C * 9 + 1, device="cuda", generator=g) * 0.1
    x2, wd = _nhwc(x).requires_grad_(True), word.clone().requires_grad_(True)
    out = E.DynConvFn.apply(x2, wd, B, H, W)
    xr, wr = x.clone().requires_grad_(True), word.clone().requires_grad_(True)
    ref = F.conv2d(xr.reshape(1, B * C, H, W), wr[:, :-1].reshape(B, C, 3, 3), wr[:, -1], padding=1, groups=B).transpose(0, 1)
    assert _rel(out, ref) < 1e-5
    go = torch.randn(ref.shape, device="cuda", generator=g)
    ref.backward(go)
    out.backward(go)
    assert _rel(_nchw(x2.grad, B, H, W), xr.grad) < 1e-5 and _rel(wd.grad, wr.grad) < 1e-4


def test_tail_bicubic_additive_blend():
    from tunevlseg_b200 import engine_cris as E

    g = torch.Generator(device="cuda").manual_seed(11)
    B, G, C, mid, img = 2, 4, 128, 64, 64
    pk = E.PackedCris.__new__(E.PackedCris)
    pk.image_size, pk._tables, pk.device = img, {}, torch.device("cuda")
    pred = torch.randn(B, 1, 4 * G, 4 * G, device="cuda", generator=g)
    fq = torch.randn(B, C, G, G, device="cuda", generator=g)
    w0 = torch.randn(mid, C, 1, 1, device="cuda", generator=g) * C ** -0.5
    w2 = torch.randn(1, mid, 5, 5, device="cuda", generator=g) * 0.05
    b2, r = torch.randn(1, device="cuda", generator=g), torch.tensor(0.35, device="cuda")
    ours = [t.clone().requires_grad_(True) for t in (pred, _nhwc(fq), w0, w2, b2, r)]
    logits = E.TailFn.apply(*ours, pk, B, 4 * G, 4 * G, G)
    refs = [t.clone().requires_grad_(True) for t in (pred, fq, w0, w2, b2, r)]
    add = F.interpolate(F.conv2d(refs[1], refs[2]), size=img, mode="bilinear")
    add = F.conv2d(F.pad(add, (2, 2, 2, 2), mode="replicate"), refs[3], refs[4])
    ref = (1 - refs[5]) * F.interpolate(refs[0], img, mode="bicubic", align_corners=True) + refs[5] * add
    assert (logits - ref).abs().max().item() < 5e-3                    # tf32 contraction of the additive map
    go = torch.randn(ref.shape, device="cuda", generator=g)
    ref.backward(go)
    logits.backward(go)
    assert _rel(ours[0].grad, refs[0].grad) < 1e-2                     # the head kernel hands dlogits on in bf16
    assert _rel(_nchw(ours[1].grad, B, G, G), refs[1].grad) < 2e-2
    for i in (2, 3, 4, 5):
        assert _rel(ours[i].grad, refs[i].grad) < 2e-2, i
    # no additive layer: plain bicubic upsampling
    p2 = pred.clone().requires_grad_(True)
    plain = E.TailFn.apply(p2, None, None, None, None, None, pk, B, 4 * G, 4 * G, G)
    assert (plain - F.interpolate(pred, img, mode="bicubic", align_corners=True)).abs().max().item() < 1e-4


def test_layernorm_wide_rows():
    from tunevlseg_b200 import engine_cris as E

    g = torch.Generator(device="cuda").manual_seed(13)
    x = torch.randn(37, 2048, device="cuda", generator=g)
    op = E.LnOp({"n.weight": 1 + 0.1 * torch.randn(2048, device="cuda", generator=g), "n.bias": 0.1 * torch.randn(2048, device="cuda", generator=g)}, "n")
    xo, xr = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    y, ref = E.layer_norm(xo, op), F.layer_norm(xr, (2048,), op.g, op.b, 1e-5)
    assert _rel(y, ref) < 1e-5
    go = torch.randn(37, 2048, device="cuda", generator=g)
    y.backward(go); ref.backward(go)
    assert _rel(xo.grad, xr.grad) < 1e-4


# ---- whole net against the oracle --------------------------------------------------------------------------------------
def _run_cris(case, spec, B, L, seed, pad=True, use_mask=True, new_last_layer=True):
    from tunevlseg_b200.losses import DiceCELoss

    weights = OCR.init_weights(spec, seed=7)
    net = build_cris_net(case, spec, weights, seed=seed, new_last_layer=new_last_layer)
    st, head = cris_oracle_state(case, net), cris_oracle_head(net)
    img, ids, am, mask = make_cris_batch(spec, B, L, seed + 1, pad)
    net = net.cuda()
    ti = {"input_ids": ids.cuda()}
    if use_mask:
        ti["attention_mask"] = am.cuda()
    logits = net(text_input=ti, image_input=img.cuda())
    conf = torch.zeros(4, dtype=torch.int64, device="cuda")
    loss, counts = DiceCELoss(sigmoid=True, lambda_dice=1, lambda_ce=0.2).forward_with_metrics(logits, mask.cuda(), 0.5, conf)
    loss.backward()
    torch.cuda.synchronize()

    ref = OCR.net_forward(weights, spec, st, head, ids, am if use_mask else None, img)
    ref_loss = OLM.dice_ce_loss(ref, mask)
    ref_loss.backward()
    assert logits.shape == ref.shape == (B, 1, spec.image_size, spec.image_size)
    err = (logits.detach().cpu() - ref.detach()).abs().max().item()
    ref_max = ref.detach().abs().max().item()
    tol = LOGIT_TOL          # the plain north_star bar (2e-2 max-abs): no ulp relaxation
    print(f"PARITY cris {case} B={B} {spec.image_size}px: logits max-abs err {err:.5f} (tol {tol:.4f}, |logit|max {ref_max:.2f})")
    assert err <= tol, f"{case}: logits max-abs err {err:.4f} > {tol} (|logit|max {ref_max:.2f})"
    assert abs(loss.item() - ref_loss.item()) <= 5e-3
    _, c_counts, c_conf = OLM.c_dicebce_metrics(logits.detach().cpu(), mask)
    assert torch.equal(counts.cpu(), c_counts) and torch.equal(conf.cpu().view(2, 2), c_conf)

    named = dict(net.named_parameters())
    checked = 0
    for k, p_ref in list(st.params.items()) + list((head or {}).items()):
        pk = k if (head and k in head) else f"context_learner.{k}"
        g, g_ref = named[pk].grad, p_ref.grad
        assert g_ref is not None and g is not None, pk
        scale = g_ref.abs().max().item()
        gerr = (g.detach().cpu() - g_ref).abs().max().item() / scale
        l2 = ((g.detach().cpu() - g_ref).norm() / g_ref.norm()).item()
        assert gerr <= GRAD_TOL and l2 <= GRAD_L2_TOL, f"{case}: grad {pk} max rel {gerr:.4f} l2 rel {l2:.4f}"
        checked += 1
    assert checked >= (2 if head else 1)
    return err


@pytest.mark.parametrize("case", list(CRIS_CASES))
def test_cris_small_geometry(case):
    _run_cris(case, CRIS_SMALL, B=3, L=8, seed=31)


def test_cris_small_no_mask_long_prompt_and_plain_head():
    _run_cris("coop_d3", CRIS_SMALL, B=2, L=76, seed=33, pad=False, use_mask=False)       # truncated to 77 tokens
    _run_cris("coop", CRIS_SMALL, B=2, L=8, seed=35, new_last_layer=False)                # plain bicubic tail


@pytest.mark.parametrize("case", ["coop", "cocoop"])
def test_cris_full_geometry(case):
    """CLIP-RN50 @ 416x416 (configs/model/coop/cris.yaml, cocoop/cris.yaml), batch 2."""
    _run_cris(case, CRIS_FULL, B=2, L=10, seed=41)


def test_cris_train_step_graph_matches_eager():
    """COOPCRIS inside the reference-shaped LightningModule: five AdamW steps lower the loss, and the whole-step CUDA
    graph replays to the same parameters as the eagerly driven step."""
    from functools import partial

    from tunevlseg_b200.graph import GraphedTrainStep
    from tunevlseg_b200.losses import DiceCELoss
    from tunevlseg_b200.models.image_text_mask_module import ImageTextMaskModule
    from tunevlseg_b200.optim import FusedAdamW

    def make():
        net = build_cris_net("cocoop", CRIS_SMALL, OCR.init_weights(CRIS_SMALL, seed=7), seed=3)
        m = ImageTextMaskModule(net=net, loss_fn=DiceCELoss(sigmoid=True, lambda_dice=1, lambda_ce=0.2),
                                optimizer=partial(FusedAdamW, lr=2e-3, weight_decay=0.0), scheduler=None, compile=False,
                                task="binary", threshold=0.5, weight_decay=0.0).to("cuda")
        m.setup("fit")
        m.train()
        return m, m.configure_optimizers()["optimizer"]

    img, ids, am, mask = make_cris_batch(CRIS_SMALL, 4, 8, 12)
    batch = {"image": img.cuda(), "mask": mask.cuda(), "input_ids": ids.cuda(), "attention_mask": am.cuda()}
    m_e, o_e = make()
    losses = []
    for _ in range(5):
        o_e.zero_grad()
        loss = m_e.training_step(batch, 0)
        loss.backward()
        o_e.step()
        losses.append(loss.item())
    assert losses[-1] < losses[0]
    m_g, o_g = make()
    step = GraphedTrainStep(m_g, o_g, batch, warmup=3)
    l4, l5 = step(batch).item(), step(batch).item()
    assert abs(l4 - losses[3]) <= 1e-4 and abs(l5 - losses[4]) <= 1e-4, (l4, l5, losses)
    pe, pg = dict(m_e.named_parameters()), dict(m_g.named_parameters())
    for k, p in pe.items():
        if p.requires_grad:
            assert torch.allclose(p, pg[k], rtol=1e-4, atol=1e-5), k
