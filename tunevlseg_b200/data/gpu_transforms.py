"""The reference's EVAL transforms on the GPU.

Replaces, for the validation / test / predict loaders (configs/experiment/coop/clipseg.yaml:113-131):

    albumentations.Compose([Resize(S, S, interpolation=cv2.INTER_CUBIC), Normalize(mean, std), ToTensorV2(transpose_mask=True)])

as called by ``ImageTextMaskDataset.__getitem__`` (src/data/core_datasets/image_text_mask_dataset.py:52-84) with the uint8
RGB image from ``cv2.imread`` and the float32 ``mask / 255``.  At ~3 000 images/s per GPU the eight-worker cv2 loader of the
reference (clipseg.yaml:44) is two orders of magnitude too slow; here the host only decodes and uploads the raw bytes, and one
kernel per image (``tvs_preproc_image_u8``) resizes in OpenCV's 8-bit fixed-point arithmetic, normalises and transposes.

The tap tables are computed on the host exactly as ``cv::resize`` computes them (float32 cubic with A = -0.75 at
``fx = (d + 0.5) * scale - 0.5``, ``saturate_cast<short>(w * 2048)``); they are cached per (source size, target size).
The train-time augmentations (Affine / RandomBrightnessContrast with p = 0.2, clipseg.yaml:84-103) are random and stay on
the host side of the reference's pipeline; this class covers the deterministic transforms.
"""
from __future__ import annotations

from functools import lru_cache

import numpy as np
import torch

from .. import abi

_COEF_SCALE = np.float32(2048)       # INTER_RESIZE_COEF_SCALE


def _cubic_coeffs(x: np.float32) -> np.ndarray:
    """cv::interpolateCubic (imgproc/src/resize.cpp), float32, same operation order."""
    a, one = np.float32(-0.75), np.float32(1)
    x = np.float32(x)
    c = np.empty(4, np.float32)
    c[0] = ((a * (x + one) - np.float32(5) * a) * (x + one) + np.float32(8) * a) * (x + one) - np.float32(4) * a
    c[1] = ((a + np.float32(2)) * x - (a + np.float32(3))) * x * x + one
    c[2] = ((a + np.float32(2)) * (one - x) - (a + np.float32(3))) * (one - x) * (one - x) + one
    c[3] = one - c[0] - c[1] - c[2]
    return c


@lru_cache(maxsize=4096)
def cubic_tables(n_in: int, n_out: int) -> tuple[np.ndarray, np.ndarray]:
    """(ofs int32 [n_out], coef int32 [n_out, 4]); tap k of output d reads source index clamp(ofs[d] + k - 1)."""
    scale = np.float64(1.0) / (np.float64(n_out) / np.float64(n_in))          # cv::resize: scale_x = 1. / inv_scale_x
    ofs = np.empty(n_out, np.int32)
    coef = np.empty((n_out, 4), np.int32)
    for d in range(n_out):
        fx = np.float32((d + 0.5) * scale - 0.5)
        sx = int(np.floor(fx))
        fx = np.float32(fx - np.float32(sx))
        coef[d] = np.clip(np.rint(_cubic_coeffs(fx) * _COEF_SCALE), -32768, 32767).astype(np.int32)
        ofs[d] = sx
    return ofs, coef


@lru_cache(maxsize=4096)
def nearest_table(n_in: int, n_out: int) -> np.ndarray:
    """cv::resizeNN source indices: min(floor(d * (1 / inv_scale)), n_in - 1)."""
    scale = np.float64(1.0) / (np.float64(n_out) / np.float64(n_in))
    return np.minimum(np.floor(np.arange(n_out) * scale).astype(np.int64), n_in - 1).astype(np.int32)


class GpuEvalTransforms:
    """``transforms(image=uint8 HWC, mask=float32 HW[1]) -> {"image": f32 [3,S,S], "mask": f32 [1,S,S]}`` on ``device``.

    Same call signature and output keys as the albumentations ``Compose`` it replaces (keyword arguments, dict result),
    so it can be handed to the reference's dataset classes as ``transforms``; the tensors it returns live on the GPU.
    """

    def __init__(self, img_size: int, mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225), max_pixel_value: float = 255.0,
                 device: str | torch.device = "cuda") -> None:
        self.size = int(img_size)
        self.device = torch.device(device)
        # albumentations.functional.normalize: mean * max_pixel_value and 1 / (std * max_pixel_value), all float32
        self.mean255 = (np.array(mean, dtype=np.float32) * np.float32(max_pixel_value)).tolist()
        self.inv_std255 = np.reciprocal(np.array(std, dtype=np.float32) * np.float32(max_pixel_value), dtype=np.float32).tolist()
        self._dev_tables: dict = {}

    def _cubic(self, n_in: int):
        key = ("c", n_in)
        if key not in self._dev_tables:
            ofs, coef = cubic_tables(n_in, self.size)
            self._dev_tables[key] = (torch.from_numpy(ofs).to(self.device), torch.from_numpy(coef).to(self.device).contiguous())
        return self._dev_tables[key]

    def _nearest(self, n_in: int):
        key = ("n", n_in)
        if key not in self._dev_tables:
            self._dev_tables[key] = torch.from_numpy(nearest_table(n_in, self.size)).to(self.device)
        return self._dev_tables[key]

    def _to_device(self, a, dtype):
        t = torch.from_numpy(np.ascontiguousarray(a)) if isinstance(a, np.ndarray) else a
        if t.dtype != dtype:
            raise TypeError(f"expected {dtype}, got {t.dtype}")
        return t.to(self.device, non_blocking=True)

    def image(self, image_u8, out: torch.Tensor | None = None) -> torch.Tensor:
        img = self._to_device(image_u8, torch.uint8)
        if img.dim() != 3 or img.shape[2] != 3:
            raise ValueError(f"image must be uint8 [H, W, 3], got {tuple(img.shape)}")
        img = img.contiguous()
        out = torch.empty((3, self.size, self.size), dtype=torch.float32, device=self.device) if out is None else out
        xo, xc = self._cubic(img.shape[1])
        yo, yc = self._cubic(img.shape[0])
        abi.preproc_image_u8(img, xo, xc, yo, yc, self.mean255, self.inv_std255, out_chw=out)
        return out

    def mask(self, mask_f32, out: torch.Tensor | None = None) -> torch.Tensor:
        m = self._to_device(mask_f32, torch.float32)
        if m.dim() == 3 and m.shape[2] == 1:
            m = m[..., 0]
        if m.dim() != 2:
            raise ValueError(f"mask must be float32 [H, W] or [H, W, 1], got {tuple(m.shape)}")
        m = m.contiguous()
        out = torch.empty((1, self.size, self.size), dtype=torch.float32, device=self.device) if out is None else out
        abi.resize_nearest_f32(m, self._nearest(m.shape[1]), self._nearest(m.shape[0]), out[0])
        return out

    def __call__(self, *, image, mask=None, **kwargs) -> dict:
        res = {"image": self.image(image)}
        if mask is not None:
            res["mask"] = self.mask(mask)
        return res
