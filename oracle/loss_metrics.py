"""Oracle restatement of the loss and the metric counters (fp32 torch on the CPU + a plain-C twin).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Neither ``monai`` nor ``torchmetrics`` is installed in this image, so the formulas
are restated from the libraries' sources at the versions the reference pins, and anchored on the
reference's call sites.  **Parity unpinned** for these two third-party boundaries: the only checks
available are the hand-derived known-answer cases in ``tests/test_oracle_loss_metrics.py``.

* ``monai.losses.DiceCELoss(sigmoid=True, lambda_dice=1, lambda_ce=0.2)`` - monai>=1.3.0
  (requirements.txt:28); call site /root/reference/configs/model/maple_clipseg.yaml:29-33.
  ``p = sigmoid(x)``; per (b, c): ``I = sum p*y, P = sum p, G = sum y`` over H, W;
  ``dice_b = 1 - (2I + 1e-5) / (P + G + 1e-5)``; ``loss = mean_b dice_b + 0.2 * mean BCEWithLogits(x, y)``
  (single output channel -> the BCE branch).
* ``torchmetrics.Dice(threshold=0.5, zero_division=1, average="samples")`` - torchmetrics>=1.4.0
  (requirements.txt:3); call site /root/reference/src/models/image_text_mask_module.py:284-298.
  Float preds with an int target of the same shape are "multilabel": ``phat = p >= threshold``, flattened
  per sample; int tp/fp/fn per sample; value = mean_b 2tp / (2tp + fp + fn), 0-denominator -> 1.
* ``torchmetrics.JaccardIndex(task="binary", threshold=0.5, zero_division=1)`` - same call site :289-298.
  ``phat = p > threshold`` (STRICT); global int64 confusion matrix [[tn, fp], [fn, tp]] over all pixels;
  value = tp / (tp + fp + fn), 0-denominator -> 1.
* targets are ``mask.long()`` (image_text_mask_module.py:107): truncation; the loss sees the float mask.

The probability that is thresholded is ``p = fl32(1 / fl32(1 + fl32(exp(-x))))`` - torch's fp32 sigmoid
formula.  The C twin (and the CUDA kernel it checks) computes ``exp`` in double and rounds once, i.e.
a correctly-rounded fp32 ``exp``; torch's own CPU ``exp`` (Sleef, <=1 ulp) can differ from that only for
|x| of a few 1e-8, which ``tests/test_oracle_loss_metrics.py`` sweeps.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np
import torch
import torch.nn.functional as F

SMOOTH = 1e-5
_HERE = os.path.dirname(os.path.abspath(__file__))


def dice_ce_loss(logits: torch.Tensor, mask: torch.Tensor, lambda_dice: float = 1.0, lambda_ce: float = 0.2) -> torch.Tensor:
    """(B,1,H,W) fp32 logits and float mask -> scalar (differentiable)."""
    p = torch.sigmoid(logits)
    y = mask.to(p.dtype)
    dims = tuple(range(2, logits.ndim))
    inter = (p * y).sum(dims)
    denom = y.sum(dims) + p.sum(dims)
    dice = (1.0 - (2.0 * inter + SMOOTH) / (denom + SMOOTH)).mean()
    ce = F.binary_cross_entropy_with_logits(logits, y)
    return lambda_dice * dice + lambda_ce * ce


def dice_ce_parts(logits: torch.Tensor, mask: torch.Tensor):
    """Per-sample (I, P, G, bce_sum) in float64 - what the fused kernel's partial sums are checked against."""
    x = logits.double().flatten(1)
    y = mask.double().flatten(1)
    p = torch.sigmoid(x)
    bce = torch.clamp(x, min=0) - x * y + torch.log1p(torch.exp(-x.abs()))
    return (p * y).sum(1), p.sum(1), y.sum(1), bce.sum(1)


def metric_counts(preds: torch.Tensor, mask: torch.Tensor, threshold: float = 0.5):
    """Integer counters from probabilities ``preds`` (B,1,H,W) and the float mask.

    Returns (dice_counts int64 (B,3) = tp,fp,fn with ``>=``; confmat int64 (2,2) = [[tn,fp],[fn,tp]] with ``>``).
    """
    t = mask.long().flatten(1)
    if t.numel() and (t.max() > 1 or t.min() < 0):
        raise ValueError("target should be binary")  # torchmetrics _check_shape_and_type_consistency
    p = preds.flatten(1)
    ge = (p >= threshold).long()
    tp = (ge * t).sum(1)
    fp = (ge * (1 - t)).sum(1)
    fn = ((1 - ge) * t).sum(1)
    gt = (p > threshold).long()
    idx = (t * 2 + gt).flatten()
    conf = torch.bincount(idx, minlength=4).reshape(2, 2)
    return torch.stack((tp, fp, fn), dim=1), conf


def dice_from_counts(counts: torch.Tensor, zero_division: float = 1.0) -> torch.Tensor:
    tp, fp, fn = counts[:, 0].float(), counts[:, 1].float(), counts[:, 2].float()
    num, den = 2 * tp, 2 * tp + fp + fn
    zero = den == 0
    score = torch.where(zero, torch.tensor(zero_division), num) / torch.where(zero, torch.tensor(1.0), den)
    return score.mean()


def iou_from_confmat(conf: torch.Tensor, zero_division: float = 1.0) -> torch.Tensor:
    tp, fp, fn = conf[1, 1].float(), conf[0, 1].float(), conf[1, 0].float()
    den = tp + fp + fn
    return torch.where(den == 0, torch.tensor(zero_division), tp / torch.where(den == 0, torch.tensor(1.0), den))


# ------------------------------------------------------------------------------------------------
# plain-C twin (oracle/loss_metrics.c), built by __graft_entry__.build() / on demand by the tests
# ------------------------------------------------------------------------------------------------
def build_c(force: bool = False) -> str:
    src = os.path.join(_HERE, "loss_metrics.c")
    out = os.path.join(_HERE, "_build", "liboracle_loss.so")
    if force or not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-fno-fast-math", "-shared", "-fPIC", "-o", out, src, "-lm"])
    return out


_lib = None


def _c():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build_c())
        _lib.oracle_dicebce_metrics.restype = None
        _lib.oracle_dicebce_metrics.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_longlong, ctypes.c_longlong,
                                                ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    return _lib


def c_dicebce_metrics(logits: torch.Tensor, mask: torch.Tensor, threshold: float = 0.5):
    """Run the C twin.  Returns (parts float64 (B,4) = I,P,G,bce_sum ; dice_counts int64 (B,3) ; confmat int64 (2,2))."""
    x = np.ascontiguousarray(logits.detach().float().cpu().numpy()).reshape(logits.shape[0], -1)
    y = np.ascontiguousarray(mask.detach().float().cpu().numpy()).reshape(mask.shape[0], -1)
    B, N = x.shape
    parts = np.zeros((B, 4), np.float64)
    counts = np.zeros((B, 3), np.int64)
    conf = np.zeros(4, np.int64)
    _c().oracle_dicebce_metrics(x.ctypes.data, y.ctypes.data, B, N, ctypes.c_float(threshold),
                                parts.ctypes.data, counts.ctypes.data, conf.ctypes.data)
    return torch.from_numpy(parts), torch.from_numpy(counts), torch.from_numpy(conf.reshape(2, 2))
