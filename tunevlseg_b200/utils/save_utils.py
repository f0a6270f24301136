"""Predict tail: write the predicted masks at their original size (src/utils/save_utils.py:19-112).

The reference does, per sample, ``TF.resize(pred.float(), mask_shape, BICUBIC, antialias=False)`` and
``torchvision.utils.save_image`` (``mul(255).add(0.5).clamp(0, 255).to(uint8)``, one grey value replicated to RGB).
Here the resize and the quantisation are ONE sm_100a kernel per sample (``tvs_resample2d_u8``: table-driven separable
bicubic, the probabilities are read once and only the u8 image leaves the GPU - 1 byte per output pixel instead of a
4-byte float map); PNG encoding stays on the host (PIL).  The bytes are IDENTICAL to the reference's: weights and
accumulation order follow ATen's CPU kernel operation for operation (``engine_cris._bicubic_tables(aten_cpu=True)``,
``oracle/resize_u8.py``), since Lightning hands ``save_predictions`` host tensors.
"""
from __future__ import annotations

from pathlib import Path

import torch

from .. import abi
from ..engine_cris import resample_tables

_TABLES: dict = {}


def _tables(hi, wi, ho, wo, device):
    key = (hi, wi, ho, wo, str(device))
    if key not in _TABLES:
        if len(_TABLES) > 256:
            _TABLES.clear()
        _TABLES[key] = resample_tables(hi, wi, ho, wo, device, align_corners=False, aten_cpu=True)
    return _TABLES[key]


def resize_to_png_array(pred: torch.Tensor, mask_shape) -> torch.Tensor:
    """pred: (1, h, w) or (h, w) CUDA probabilities -> uint8 (H, W) CUDA tensor, exactly what save_image would write."""
    abi.check_cuda_input(pred)
    p = pred.detach().to(torch.float32).reshape(pred.shape[-2], pred.shape[-1]).contiguous()
    ho, wo = (int(v) for v in mask_shape)
    out = torch.empty((ho, wo), dtype=torch.uint8, device=p.device)
    abi.resample2d_u8(p, 1, p.shape[0], p.shape[1], ho, wo, _tables(p.shape[0], p.shape[1], ho, wo, p.device), out)
    return out


def save_predictions(cfg, log, trainer, model, dataloaders, ckpt_path) -> None:
    """Same contract as the reference's ``save_predictions``: ``trainer.predict`` yields dicts with ``preds``,
    ``mask_name`` and ``mask_shape``; one PNG per sample lands under ``cfg['output_masks_dir']``."""
    from PIL import Image

    output_masks_dir = cfg.get("output_masks_dir")
    if output_masks_dir is None:
        output_masks_dir = "output_masks"
        log.warning(f"`output_masks_dir` was not passed in the config.Defaulting to {output_masks_dir}")
    output_masks_dir = Path(output_masks_dir)
    if output_masks_dir.exists():
        log.warning(f"{output_masks_dir} exists.The output masks may override the previous ones.")
        if not cfg.get("overwrite_outputs"):
            log.info("`overwrite_outputs` was not passed or if passed as False.So stopping the prediction instead of overwriting.")
            return
    interp = cfg.get("output_interpolation")
    if interp is not None and str(getattr(interp, "value", interp)).lower() != "bicubic":
        raise NotImplementedError("the fused resize + quantise kernel implements the reference's default (bicubic) interpolation")
    log.info("Generating prediction masks of test dataset")
    pred_outputs = list(trainer.predict(model=model, dataloaders=dataloaders, ckpt_path=ckpt_path))
    log.info(f"Saving the generated masks in directory {output_masks_dir}")
    total = 0
    for p in pred_outputs:
        for pred, mask_name, mask_shape in zip(p["preds"], p["mask_name"], p["mask_shape"], strict=True):
            file_path = output_masks_dir / mask_name
            file_path.parent.mkdir(parents=True, exist_ok=True)
            shape = mask_shape.tolist() if isinstance(mask_shape, torch.Tensor) else list(mask_shape)
            grey = resize_to_png_array(pred, shape).cpu().numpy()
            Image.fromarray(grey).convert("RGB").save(file_path)          # save_image writes the grey value to 3 channels
            total += 1
    log.info(f"Logged {total} masks to {output_masks_dir} using bicubic interpolation.")
