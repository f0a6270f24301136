"""Offline metric script (scripts/eval_metrics.py:1-150 of the reference): per-image IoU / Dice of saved masks against
ground-truth PNGs, written to a csv.

The reference thresholds both images and calls monai ``compute_iou`` / ``compute_dice`` (``ignore_empty=False``) per
file on the CPU; here the integer TP / FP / FN counts come from the fused counter kernel (``tvs_metrics_from_probs``,
the same kernel the training step uses - so this script is also the cross-check of the in-training counters), batched
over all files of one size.  ``dice = 2TP / (2TP + FP + FN)`` and ``iou = TP / (TP + FP + FN)``, both defined as 1 when
the denominator is 0 (empty prediction and ground truth: what monai returns with ``ignore_empty=False``).

    python -m tunevlseg_b200.scripts.eval_metrics --seg_path <dir> --gt_path <dir> --csv_path out.csv [--threshold 127]
"""
from __future__ import annotations

import csv
from argparse import ArgumentParser
from collections import defaultdict
from pathlib import Path

import numpy as np
import torch

from .. import abi


def load_image(image_path: str) -> np.ndarray:
    import cv2

    image = cv2.imread(str(image_path), cv2.IMREAD_GRAYSCALE)
    if image is None:
        raise ValueError(f"Image Not found: {image_path}")
    return image


def counts_for_batch(pred_bin: torch.Tensor, gt_bin: torch.Tensor) -> torch.Tensor:
    """pred_bin, gt_bin: (B, H, W) {0,1} float CUDA tensors -> int64 (B, 3) = tp, fp, fn (threshold 0.5, '>=')."""
    B, N = pred_bin.shape[0], pred_bin[0].numel()
    counts = torch.empty((B, 3), dtype=torch.int64, device=pred_bin.device)
    conf = torch.zeros(4, dtype=torch.int64, device=pred_bin.device)
    scratch = torch.empty(abi.dicebce_scratch_bytes(B, N), dtype=torch.uint8, device=pred_bin.device)
    abi.metrics_from_probs(pred_bin.contiguous(), gt_bin.contiguous(), 0.5, counts, conf, scratch)
    return counts


def metrics_from_counts(tp: int, fp: int, fn: int, n_pos_gt: int, n: int) -> dict:
    den = 2 * tp + fp + fn
    dice = 100.0 * (2 * tp / den if den else 1.0)
    iou = 100.0 * (tp / (tp + fp + fn) if (tp + fp + fn) else 1.0)
    ones_den = 2 * n_pos_gt + (n - n_pos_gt)                  # all-ones prediction: tp = |gt|, fp = n - |gt|, fn = 0
    ones_dice = 100.0 * (2 * n_pos_gt / ones_den if ones_den else 1.0)
    return {"iou": iou, "dice": dice, "ones_dice_diff": dice - ones_dice}


def evaluate(seg_path: Path, gt_path: Path, threshold: int = 127, device="cuda") -> list[dict]:
    names = sorted(p.name for p in Path(seg_path).glob("*.png"))           # the reference walks the predictions
    groups = defaultdict(list)
    for name in names:
        gt, pred = load_image(Path(gt_path) / name), load_image(Path(seg_path) / name)
        assert gt.shape == pred.shape, f"Images {name} are of different sizes"
        groups[gt.shape].append((name, gt > 127, pred > threshold))
    rows = []
    for shape, items in groups.items():
        gt = torch.from_numpy(np.stack([g for _, g, _ in items])).to(device).float()
        pr = torch.from_numpy(np.stack([p for _, _, p in items])).to(device).float()
        counts = counts_for_batch(pr, gt).cpu()
        npos = gt.flatten(1).sum(1).long().cpu()
        for (name, _, _), c, g in zip(items, counts, npos):
            rows.append({"filename": name, **metrics_from_counts(int(c[0]), int(c[1]), int(c[2]), int(g), shape[0] * shape[1])})
    return sorted(rows, key=lambda r: r["filename"])


def main() -> None:
    ap = ArgumentParser()
    ap.add_argument("--seg_path", type=Path, required=True)
    ap.add_argument("--gt_path", type=Path, required=True)
    ap.add_argument("--csv_path", type=Path, required=True)
    ap.add_argument("--threshold", type=int, default=127)
    a = ap.parse_args()
    rows = evaluate(a.seg_path, a.gt_path, a.threshold)
    with open(a.csv_path, "w", newline="") as f:
        w = csv.DictWriter(f, fieldnames=["filename", "iou", "dice", "ones_dice_diff"])
        w.writeheader()
        w.writerows({k: (f"{v:.4f}" if isinstance(v, float) else v) for k, v in r.items()} for r in rows)
    for k in ("iou", "dice", "ones_dice_diff"):
        print(f"{k}: {np.mean([r[k] for r in rows]):.5f}")


if __name__ == "__main__":
    main()
