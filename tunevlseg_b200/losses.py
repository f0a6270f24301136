"""Drop-in for ``monai.losses.DiceCELoss`` as the reference configures it
(/root/reference/configs/model/maple_clipseg.yaml:29-33: ``sigmoid=True, lambda_dice=1, lambda_ce=0.2``), computed by
the fused sm_100a kernel together with the Dice / IoU integer counters."""
from __future__ import annotations

import torch
from torch import nn

from . import engine


class DiceCELoss(nn.Module):
    def __init__(self, sigmoid: bool = True, lambda_dice: float = 1.0, lambda_ce: float = 1.0, include_background: bool = True,
                 to_onehot_y: bool = False, softmax: bool = False, squared_pred: bool = False, jaccard: bool = False,
                 reduction: str = "mean", smooth_nr: float = 1e-5, smooth_dr: float = 1e-5, batch: bool = False, **kwargs) -> None:
        super().__init__()
        unsupported = dict(sigmoid=(sigmoid, True), include_background=(include_background, True), to_onehot_y=(to_onehot_y, False),
                           softmax=(softmax, False), squared_pred=(squared_pred, False), jaccard=(jaccard, False),
                           reduction=(reduction, "mean"), smooth_nr=(smooth_nr, 1e-5), smooth_dr=(smooth_dr, 1e-5), batch=(batch, False))
        bad = {k: v[0] for k, v in unsupported.items() if v[0] != v[1]}
        if bad or kwargs:
            raise NotImplementedError(f"fused DiceCELoss covers the reference's configuration only; unsupported: {bad or kwargs}")
        if lambda_dice < 0.0 or lambda_ce < 0.0:
            raise ValueError("lambda_dice and lambda_ce should be no less than 0.0.")
        self.lambda_dice, self.lambda_ce = float(lambda_dice), float(lambda_ce)
        self.last_counts: torch.Tensor | None = None

    def forward_with_metrics(self, input: torch.Tensor, target: torch.Tensor, threshold: float = 0.5, confmat: torch.Tensor | None = None):
        """-> (loss, per-sample int64 (B,3) tp/fp/fn with p >= thr); ``confmat`` int64[4] (tn,fp,fn,tp with p > thr) is
        accumulated in place when given."""
        if input.shape != target.shape:
            raise ValueError(f"the number of dimensions for input and target should be the same, got {input.shape} and {target.shape}")
        if input.dim() < 3 or input.shape[1] != 1:
            raise ValueError("fused DiceCELoss expects single-channel logits (B, 1, H, W)")
        loss, counts = engine.DiceBceFn.apply(input, target, threshold, self.lambda_dice, self.lambda_ce, confmat)
        self.last_counts = counts
        return loss, counts

    def forward(self, input: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        return self.forward_with_metrics(input, target)[0]
