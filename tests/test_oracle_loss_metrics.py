"""Known-answer cases for the loss / metric oracle (monai and torchmetrics are not installed: these hand-derived cases,
plus the python<->C twin agreement, are the only pin - "parity unpinned" for those third-party boundaries)."""
import math

import torch

from oracle import loss_metrics as OLM


def test_dice_ce_hand_computed():
    # one sample, 4 pixels: logits 0 -> p = 0.5 everywhere; mask = [1,1,0,0]
    x = torch.zeros(1, 1, 2, 2)
    y = torch.tensor([1.0, 1.0, 0.0, 0.0]).view(1, 1, 2, 2)
    inter, P, G = 1.0, 2.0, 2.0
    dice = 1 - (2 * inter + 1e-5) / (P + G + 1e-5)
    bce = math.log(2.0)
    assert abs(OLM.dice_ce_loss(x, y).item() - (dice + 0.2 * bce)) < 1e-6


def test_threshold_conventions():
    # p == 0.5 exactly: positive for Dice (>=), negative for IoU (>)
    x = torch.zeros(1, 1, 1, 4)
    y = torch.tensor([1.0, 0.0, 1.0, 0.0]).view(1, 1, 1, 4)
    counts, conf = OLM.metric_counts(torch.sigmoid(x), y)
    assert counts.tolist() == [[2, 2, 0]]
    assert conf.tolist() == [[2, 0], [2, 0]]
    # soft masks: the metrics see mask.long() (0.7 -> 0), the loss sees the float mask
    y2 = torch.tensor([0.7, 1.0, 0.2, 0.0]).view(1, 1, 1, 4)
    counts, _ = OLM.metric_counts(torch.full((1, 1, 1, 4), 0.9), y2)
    assert counts.tolist() == [[1, 3, 0]]
    # empty prediction and target -> Dice 1 and IoU 1 with zero_division=1
    c, cf = OLM.metric_counts(torch.zeros(2, 1, 2, 2), torch.zeros(2, 1, 2, 2))
    assert OLM.dice_from_counts(c).item() == 1.0 and OLM.iou_from_confmat(cf).item() == 1.0


def test_c_twin_agrees_with_python_and_torch_sigmoid_near_half():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(3, 1, 40, 40, generator=g) * 2
    # adversarial logits around the p == 0.5 decision (|x| down to 2^-30) and exact zeros
    k = torch.randint(-40, 41, (3 * 1600,), generator=g).float()
    e = torch.randint(22, 31, (3 * 1600,), generator=g).float()
    adv = (k * torch.pow(2.0, -e)).view(3, 1, 40, 40)
    x[:, :, ::2] = adv[:, :, ::2]
    y = (torch.rand(3, 1, 40, 40, generator=g) < 0.4).float()
    parts, counts, conf = OLM.c_dicebce_metrics(x, y)
    pc, pconf = OLM.metric_counts(torch.sigmoid(x), y)
    assert torch.equal(counts, pc) and torch.equal(conf, pconf), "C twin and torch's fp32 sigmoid disagree on a threshold decision"
    ref = torch.stack(OLM.dice_ce_parts(x, y), dim=1)
    assert torch.allclose(parts, ref, rtol=1e-6, atol=1e-6)
    # loss recomposed from the C partial sums == torch formula
    B, N = 3, 1600
    loss_c = sum(1 - (2 * parts[b, 0] + 1e-5) / (parts[b, 1] + parts[b, 2] + 1e-5) for b in range(B)) / B + 0.2 * parts[:, 3].sum() / (B * N)
    assert abs(float(loss_c) - OLM.dice_ce_loss(x, y).item()) < 1e-5
