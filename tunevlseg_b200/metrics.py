"""``Dice`` and ``JaccardIndex`` with the interface the reference uses from torchmetrics
(/root/reference/src/models/image_text_mask_module.py:284-302: ``Dice(threshold, zero_division=1, average="samples")``,
``JaccardIndex(task="binary", threshold, zero_division=1)``), holding the same integer states - per-sample tp/fp/fn
lists (cat-reduced) and a 2x2 confusion matrix (sum-reduced) - but filled by the fused loss/metric kernel."""
from __future__ import annotations

import torch
import torch.distributed as dist
from torch import nn

from . import abi


def _scratch(B, N, device):
    return torch.empty(abi.dicebce_scratch_bytes(B, N), dtype=torch.uint8, device=device)


def _counts_from_probs(preds, target, threshold):
    B = preds.shape[0]
    p = preds.detach().to(torch.float32).contiguous()
    t = target.detach().to(torch.float32).contiguous()
    counts = torch.empty((B, 3), dtype=torch.int64, device=p.device)
    conf = torch.zeros(4, dtype=torch.int64, device=p.device)
    abi.metrics_from_probs(p, t, float(threshold), counts, conf, _scratch(B, p.numel() // B, p.device))
    return counts, conf


class _Metric(nn.Module):
    def reset(self) -> None: ...
    def compute(self) -> torch.Tensor: ...

    def forward(self, *args, **kwargs) -> torch.Tensor:
        return self.update(*args, **kwargs)


class Dice(_Metric):
    """Per-sample Dice averaged over samples; predictions are positive when ``p >= threshold``."""

    def __init__(self, threshold: float = 0.5, zero_division: float = 0, average: str = "micro", **kwargs) -> None:
        super().__init__()
        if average != "samples":
            raise NotImplementedError("only average='samples' (the reference's setting) is implemented")
        self.threshold, self.zero_division = float(threshold), float(zero_division)
        self._counts: list[torch.Tensor] = []
        # streaming state for CUDA-graph replay (a Python list cannot grow inside a replayed graph): running sum of the
        # per-sample scores and the sample count, updated by device ops; same value as the cat-reduced lists
        self.streaming = False
        self.register_buffer("_score_sum", torch.zeros((), dtype=torch.float64), persistent=False)
        self.register_buffer("_n", torch.zeros((), dtype=torch.int64), persistent=False)

    def reset(self) -> None:
        self._counts = []
        self._score_sum.zero_()
        self._n.zero_()

    @staticmethod
    def score(counts: torch.Tensor, zero_division: float) -> torch.Tensor:
        tp, fp, fn = (counts[:, i].to(torch.float32) for i in range(3))
        num, den = 2 * tp, 2 * tp + fp + fn
        zero = den == 0
        return (torch.where(zero, torch.full_like(num, zero_division), num) / torch.where(zero, torch.ones_like(den), den)).mean()

    def update_from_counts(self, counts: torch.Tensor) -> torch.Tensor:
        """counts: int64 (B,3) tp/fp/fn from the fused kernel.  Returns this batch's value (torchmetrics ``forward``)."""
        value = self.score(counts, self.zero_division)
        if self.streaming:
            self._score_sum += value.to(torch.float64) * counts.shape[0]
            self._n += counts.shape[0]
        else:
            self._counts.append(counts)
        return value

    def update(self, preds: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        if preds.shape != target.shape:
            raise ValueError("preds and target must have the same shape")
        return self.update_from_counts(_counts_from_probs(preds, target, self.threshold)[0])

    @property
    def has_updates(self) -> bool:
        return bool(self._counts) or (self.streaming and int(self._n) > 0)

    def compute(self) -> torch.Tensor:
        """Mean of the per-sample scores over every sample seen (on all ranks).  The cross-rank reduction is a SUM of
        (score sum, sample count): ranks may hold different numbers of samples (last validation batch not dropped,
        image_text_mask_datamodule.py:40-47), which a fixed-shape all_gather of the count lists cannot express."""
        tot = torch.stack((self._score_sum, self._n.to(torch.float64)))        # streaming part (graph replays)
        if self._counts:
            counts = torch.cat(self._counts)
            tp, fp, fn = (counts[:, i].to(torch.float32) for i in range(3))
            num, den = 2 * tp, 2 * tp + fp + fn
            zero = den == 0
            scores = torch.where(zero, torch.full_like(num, self.zero_division), num) / torch.where(zero, torch.ones_like(den), den)
            tot = tot + torch.stack((scores.to(torch.float64).sum(), torch.tensor(float(counts.shape[0]), dtype=torch.float64, device=scores.device))).to(tot.device)
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        if float(tot[1]) == 0:
            raise RuntimeError("Dice.compute() called before any update")
        return (tot[0] / tot[1]).to(torch.float32)


class JaccardIndex(_Metric):
    """Binary IoU over all pixels seen; predictions are positive when ``p > threshold`` (strict)."""

    def __init__(self, task: str = "binary", threshold: float = 0.5, zero_division: float = 0, **kwargs) -> None:
        super().__init__()
        if task != "binary":
            raise NotImplementedError("only task='binary' (the reference's setting) is implemented")
        self.threshold, self.zero_division = float(threshold), float(zero_division)
        self.register_buffer("confmat", torch.zeros(4, dtype=torch.int64), persistent=False)   # tn, fp, fn, tp

    def reset(self) -> None:
        self.confmat.zero_()

    @property
    def has_updates(self) -> bool:
        return int(self.confmat.sum()) > 0

    @staticmethod
    def score(conf: torch.Tensor, zero_division: float) -> torch.Tensor:
        fp, fn, tp = conf[1].to(torch.float32), conf[2].to(torch.float32), conf[3].to(torch.float32)
        den = tp + fp + fn
        return torch.where(den == 0, torch.full_like(den, zero_division), tp / torch.where(den == 0, torch.ones_like(den), den))

    def update_from_confmat(self, batch_conf: torch.Tensor) -> torch.Tensor:
        self.confmat += batch_conf.to(self.confmat.device)
        return self.score(batch_conf, self.zero_division)

    def update(self, preds: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        return self.update_from_confmat(_counts_from_probs(preds, target, self.threshold)[1])

    def compute(self) -> torch.Tensor:
        conf = self.confmat.clone()
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(conf, op=dist.ReduceOp.SUM)
        return self.score(conf, self.zero_division)
