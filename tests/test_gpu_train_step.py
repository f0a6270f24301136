"""GPU: the pieces around the forward/backward parity of test_gpu_parity.py that make up a whole training step -
the flat-buffer AdamW kernel against torch.optim.AdamW (configs/model/maple_clipseg.yaml:36-39), the CUDA-graph replay
against the same step driven eagerly, the no-grad eval step, and "no silent fallback" behaviour."""
from functools import partial

import pytest
import torch

from oracle import clipseg as OC
from tests.helpers import SMALL, build_net, make_batch

pytestmark = pytest.mark.gpu


def _module(case="maple", seed=3, lr=2e-3):
    from tunevlseg_b200.losses import DiceCELoss
    from tunevlseg_b200.models.image_text_mask_module import ImageTextMaskModule
    from tunevlseg_b200.optim import FusedAdamW

    net = build_net(case, SMALL, OC.init_weights(SMALL, seed=7), seed=seed)
    module = ImageTextMaskModule(net=net, loss_fn=DiceCELoss(sigmoid=True, lambda_dice=1, lambda_ce=0.2),
                                 optimizer=partial(FusedAdamW, lr=lr, weight_decay=0.01), scheduler=None, compile=False,
                                 task="binary", threshold=0.5, weight_decay=0.01).to("cuda")
    module.setup("fit")
    return module, module.configure_optimizers()["optimizer"]


def _batch(B=4, L=8, seed=11):
    img, ids, am, mask = make_batch(SMALL, B, L, seed)
    return {"image": img.cuda(), "mask": mask.cuda(), "input_ids": ids.cuda(), "attention_mask": am.cuda()}


@pytest.mark.parametrize("wd", [0.0, 0.05])
def test_fused_adamw_matches_torch(wd):
    from tunevlseg_b200.optim import FusedAdamW

    g = torch.Generator(device="cuda").manual_seed(0)
    shapes = [(9, 4, 512), (64, 512), (64,), (1,), (5, 5, 3)]      # 5*5*3 = 75: the flat buffer gets padded to 4
    ours = [torch.nn.Parameter(torch.randn(s, device="cuda", generator=g)) for s in shapes]
    theirs = [torch.nn.Parameter(p.detach().clone()) for p in ours]
    a = FusedAdamW(ours, lr=3e-3, betas=(0.9, 0.98), eps=1e-8, weight_decay=wd)
    b = torch.optim.AdamW(theirs, lr=3e-3, betas=(0.9, 0.98), eps=1e-8, weight_decay=wd)
    for step in range(6):
        a.zero_grad()
        for p, q in zip(ours, theirs):
            gr = torch.randn(p.shape, device="cuda", generator=g) * (0.1 + step)
            p.grad.copy_(gr)
            q.grad = gr.clone()
        if step == 3:       # a scheduler changing the group lr must reach the kernel - param_groups is all it touches
            a.param_groups[0]["lr"] = b.param_groups[0]["lr"] = 1e-3
        a.step()
        b.step()
    for p, q in zip(ours, theirs):
        assert torch.allclose(p, q, rtol=2e-6, atol=2e-7), (p - q).abs().max().item()


def _twin_params(shapes, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    ours = [torch.nn.Parameter(torch.randn(s, device="cuda", generator=g)) for s in shapes]
    return ours, [torch.nn.Parameter(p.detach().clone()) for p in ours], g


def _drive(opts, params, g, steps, scale0=0):
    for step in range(steps):
        grads = [torch.randn(p.shape, device="cuda", generator=g) * (0.1 + scale0 + step) for p in params[0]]
        for opt, ps in zip(opts, params):
            opt.zero_grad()
            for p, gr in zip(ps, grads):
                if p.grad is None:
                    p.grad = gr.clone()
                else:
                    p.grad.copy_(gr)
            opt.step()


def test_fused_adamw_follows_reduce_lr_on_plateau():
    """The reference's default scheduler (maple_clipseg.yaml:50-55) only edits ``param_groups[i]["lr"]``; the kernel reads
    a device scalar.  Driving a REAL ReduceLROnPlateau on both optimizers (no sync_lr call anywhere) must keep them equal."""
    from tunevlseg_b200.optim import FusedAdamW

    ours, theirs, g = _twin_params([(9, 4, 64), (33,), (1,)])
    a = FusedAdamW(ours, lr=5e-3, weight_decay=0.01)
    b = torch.optim.AdamW(theirs, lr=5e-3, weight_decay=0.01)
    sa = torch.optim.lr_scheduler.ReduceLROnPlateau(a, mode="min", factor=0.1, patience=0)
    sb = torch.optim.lr_scheduler.ReduceLROnPlateau(b, mode="min", factor=0.1, patience=0)
    for epoch, val_loss in enumerate([1.0, 1.1, 1.2, 0.5, 0.6]):      # two plateaus -> lr 5e-3 -> 5e-4 -> 5e-5 -> ...
        _drive([a, b], [ours, theirs], g, 2, scale0=epoch)
        sa.step(val_loss)
        sb.step(val_loss)
    assert a.param_groups[0]["lr"] == b.param_groups[0]["lr"] < 5e-3
    _drive([a, b], [ours, theirs], g, 2)
    for p, q in zip(ours, theirs):
        assert torch.allclose(p, q, rtol=2e-6, atol=2e-7), (p - q).abs().max().item()


def test_fused_adamw_state_dict_round_trip_and_torch_interchange():
    """Lightning checkpoints ``optimizer.state_dict()``: moments and step count must be in it (torch.optim.AdamW layout),
    a resumed optimizer must continue bit for bit, and a torch.optim.AdamW state dict must load."""
    from tunevlseg_b200.optim import FusedAdamW

    shapes = [(9, 4, 64), (33,), (5, 5, 3)]
    ours, theirs, g = _twin_params(shapes)
    a = FusedAdamW(ours, lr=3e-3, weight_decay=0.02)
    b = torch.optim.AdamW(theirs, lr=3e-3, weight_decay=0.02)
    _drive([a, b], [ours, theirs], g, 3)
    sd = a.state_dict()
    sd_t = b.state_dict()
    assert set(sd["state"]) == set(sd_t["state"]) == {0, 1, 2}
    for i in range(3):
        assert float(sd["state"][i]["step"]) == float(sd_t["state"][i]["step"]) == 3.0
        for k in ("exp_avg", "exp_avg_sq"):
            assert sd["state"][i][k].shape == sd_t["state"][i][k].shape
            assert torch.allclose(sd["state"][i][k], sd_t["state"][i][k], rtol=1e-4, atol=1e-7)       # torch lerps, the kernel fmas
    # resume into a fresh optimizer over copies of the parameters, with a DIFFERENT constructor lr: the checkpoint wins
    resumed = [torch.nn.Parameter(p.detach().clone()) for p in ours]
    a2 = FusedAdamW(resumed, lr=1.0, weight_decay=0.02)
    a2.load_state_dict(sd)
    assert a2.param_groups[0]["lr"] == 3e-3
    # and a torch.optim.AdamW state dict into a third one
    from_torch = [torch.nn.Parameter(q.detach().clone()) for q in theirs]
    a3 = FusedAdamW(from_torch, lr=3e-3, weight_decay=0.02)
    a3.load_state_dict(sd_t)
    g2 = torch.Generator(device="cuda").manual_seed(5)
    for step in range(3):
        grads = [torch.randn(s, device="cuda", generator=g2) for s in shapes]
        for opt, ps in ((a, ours), (a2, resumed), (a3, from_torch), (b, theirs)):
            opt.zero_grad()
            for p, gr in zip(ps, grads):
                if p.grad is None:
                    p.grad = gr.clone()
                else:
                    p.grad.copy_(gr)
            opt.step()
    for p, r, t, q in zip(ours, resumed, from_torch, theirs):
        assert torch.equal(p, r), "resumed FusedAdamW diverged from the uninterrupted one"
        assert torch.allclose(t, q, rtol=2e-6, atol=2e-7)
    assert float(a2.state_dict()["state"][0]["step"]) == 6.0


def test_fused_adamw_rejects_parameters_moved_off_the_flat_buffer():
    from tunevlseg_b200.abi import TvsError
    from tunevlseg_b200.optim import FusedAdamW

    lin = torch.nn.Linear(8, 8).cuda()
    opt = FusedAdamW(lin.parameters(), lr=1e-3)
    lin.weight.data = lin.weight.data.clone()          # what model.to()/float() does: new storage
    with pytest.raises(TvsError, match="storage no longer aliases"):
        opt.step()


def test_graph_replay_sees_scheduler_lr_changes():
    """Inside a captured step the kernel reads the lr from a device scalar; GraphedTrainStep refreshes it from
    param_groups before every replay, so a scheduler step between replays takes effect (and equals the eager run)."""
    from tunevlseg_b200.graph import GraphedTrainStep

    batch = _batch()
    m_e, o_e = _module()
    m_g, o_g = _module()
    step = GraphedTrainStep(m_g, o_g, batch, warmup=3)
    for _ in range(3):
        o_e.zero_grad()
        m_e.training_step(batch, 0).backward()
        o_e.step()
    for opt in (o_e, o_g):
        for grp in opt.param_groups:
            grp["lr"] = grp["lr"] * 0.05
    o_e.zero_grad()
    le = m_e.training_step(batch, 0)
    le.backward()
    o_e.step()
    lg = step(batch)
    assert abs(le.item() - lg.item()) <= 1e-5
    pe, pg = dict(m_e.named_parameters()), dict(m_g.named_parameters())
    for k, p in pe.items():
        if p.requires_grad:
            assert torch.allclose(p, pg[k], rtol=1e-5, atol=1e-6), f"{k}: {(p - pg[k]).abs().max().item()}"
    assert m_g.train_dice.streaming        # GraphedTrainStep switched the Dice metrics to device-side accumulation


def test_fused_adamw_rejects_detached_grads():
    from tunevlseg_b200.abi import TvsError
    from tunevlseg_b200.optim import FusedAdamW

    p = torch.nn.Parameter(torch.ones(8, device="cuda"))
    opt = FusedAdamW([p], lr=1e-3)
    p.grad = None
    with pytest.raises(TvsError):
        opt.step()
    with pytest.raises(TvsError):
        FusedAdamW([torch.nn.Parameter(torch.ones(8))], lr=1e-3)       # CPU parameters: no CPU fallback


def test_graph_replay_equals_eager_steps():
    from tunevlseg_b200.graph import GraphedTrainStep

    batch = _batch()
    m_e, o_e = _module()
    m_g, o_g = _module()
    eager_losses = []
    for _ in range(5):
        o_e.zero_grad()
        loss = m_e.training_step(batch, 0)
        loss.backward()
        o_e.step()
        eager_losses.append(loss.item())
    step = GraphedTrainStep(m_g, o_g, batch, warmup=3)      # 3 eager steps inside, then capture (not executed)
    l4 = step(batch).item()
    l5 = step(batch).item()
    assert eager_losses[0] > eager_losses[-1], "five AdamW steps on one batch must lower the loss"
    assert abs(l4 - eager_losses[3]) <= 1e-5 and abs(l5 - eager_losses[4]) <= 1e-5, (l4, l5, eager_losses)
    pe, pg = dict(m_e.named_parameters()), dict(m_g.named_parameters())
    for k, p in pe.items():
        if p.requires_grad:
            assert torch.allclose(p, pg[k], rtol=1e-5, atol=1e-6), f"{k}: {(p - pg[k]).abs().max().item()}"
    # a different batch through the same graph: static inputs are really re-read
    other = _batch(seed=29)
    l6 = step(other).item()
    o_e.zero_grad()
    ref = m_e.training_step(other, 0).item()
    assert abs(l6 - ref) <= 1e-4, (l6, ref)


def test_graph_prefetch_path_matches_direct_load():
    from tunevlseg_b200.graph import GraphedTrainStep

    batch = _batch()
    m_a, o_a = _module()
    m_b, o_b = _module()
    s_a = GraphedTrainStep(m_a, o_a, batch, warmup=3)
    s_b = GraphedTrainStep(m_b, o_b, batch, warmup=3)
    pinned = [{k: v.cpu().pin_memory() for k, v in _batch(seed=40 + i).items()} for i in range(3)]
    s_b.prefetch(pinned[0])
    for i in range(3):
        la = s_a({k: v.cuda() for k, v in pinned[i].items()}).item()
        out = s_b.step_prefetched()
        if i + 1 < 3:
            s_b.prefetch(pinned[i + 1])
        assert la == out.item()


def test_eval_step_no_grad_matches_train_forward():
    module, _ = _module("vpt")
    batch = _batch()
    module.eval()
    with torch.no_grad():
        out_eval = module.net(text_input={"input_ids": batch["input_ids"], "attention_mask": batch["attention_mask"]},
                              image_input=batch["image"])
    module.train()
    out_train = module.net(text_input={"input_ids": batch["input_ids"], "attention_mask": batch["attention_mask"]},
                           image_input=batch["image"])
    assert not out_eval.requires_grad and out_train.requires_grad
    assert torch.equal(out_eval, out_train.detach())
    module.eval()
    module.setup("test")
    with torch.no_grad():
        module.validation_step(batch, 0)
        module.test_step(batch, 0)


def test_cpu_tensors_fail_loudly():
    """The product path has no CPU route: handing the C ABI host tensors is an error, not a silent torch fallback."""
    from tunevlseg_b200 import abi

    with pytest.raises((abi.TvsError, ValueError, RuntimeError)):
        abi.layernorm_fwd(torch.randn(4, 64), torch.ones(64), torch.zeros(64), 1e-5)
