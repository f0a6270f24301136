"""The reference's EVAL transforms on the GPU.

Replaces, for the validation / test / predict loaders (configs/experiment/coop/clipseg.yaml:113-131):

    albumentations.Compose([Resize(S, S, interpolation=cv2.INTER_CUBIC), Normalize(mean, std), ToTensorV2(transpose_mask=True)])

as called by ``ImageTextMaskDataset.__getitem__`` (src/data/core_datasets/image_text_mask_dataset.py:52-84) with the uint8
RGB image from ``cv2.imread`` and the float32 ``mask / 255``.  At ~3 000 images/s per GPU the eight-worker cv2 loader of the
reference (clipseg.yaml:44) is two orders of magnitude too slow; here the host only decodes and uploads the raw bytes, and one
kernel per image (``tvs_preproc_image_u8``) resizes in OpenCV's 8-bit fixed-point arithmetic, normalises and transposes.

The tap tables are computed on the host exactly as ``cv::resize`` computes them (float32 cubic with A = -0.75 at
``fx = (d + 0.5) * scale - 0.5``, ``saturate_cast<short>(w * 2048)``); they are cached per (source size, target size).
``GpuTrainTransforms`` adds the train-time augmentations (clipseg.yaml:80-111): ``Affine`` (cv2.warpAffine, INTER_CUBIC,
BORDER_REPLICATE, p = 0.2) and ``RandomBrightnessContrast`` (a uint8 look-up table, p = 0.2).  The host draws the parameters
and builds cv::warpAffine's fixed-point walk tables in double, exactly as OpenCV does; the pixels are computed on the device in
integer arithmetic and equal cv2's bit for bit (oracle/augment.py, pinned to cv2 4.13).  ``PadIfNeeded`` /
``CropNonEmptyMaskIfExists`` are identities on an S x S image and are skipped.
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


# ------------------------------------------------------------------------------------------------------------------
# train-time augmentations (clipseg.yaml:80-111)
# ------------------------------------------------------------------------------------------------------------------
_AB_BITS, _INTER_BITS, _REMAP_COEF_SCALE = 10, 5, np.float32(1 << 15)


@lru_cache(maxsize=1)
def warp_cubic_table() -> np.ndarray:
    """int16 [32, 32, 4, 4]: cv::initInterTab2D(INTER_CUBIC, fixed point) - outer products of the float cubic at k / 32 scaled by
    2^15, saturated to short, and corrected to sum to 2^15 (the C loop looks for the extreme tap among taps [2, 4) x [2, 4))."""
    t1 = np.stack([_cubic_coeffs(np.float32(i) * np.float32(1.0 / 32)) for i in range(32)])
    tab = np.empty((32, 32, 4, 4), np.int32)
    for i in range(32):
        for j in range(32):
            it = np.clip(np.rint((t1[i][:, None] * t1[j][None, :]).astype(np.float32) * _REMAP_COEF_SCALE), -32768, 32767).astype(np.int32)
            diff = int(it.sum()) - (1 << 15)
            if diff:
                sub = it[2:4, 2:4]
                # first minimum / first maximum in scan order, as the strict comparisons of the C loop select them
                k = np.unravel_index(np.argmax(sub) if diff < 0 else np.argmin(sub), sub.shape)
                it[2 + k[0], 2 + k[1]] -= diff
            tab[i, j] = it
    return tab.astype(np.int16)


def affine_walk_tables(matrix, dsize, nearest: bool):
    """(adelta, bdelta, x0, y0) int32: cv::warpAffine's inverse-matrix walk in 1 / 1024 pixel (imgwarp.cpp), rounding term
    (16 for INTER_CUBIC, 512 for INTER_NEAREST) folded into x0 / y0.  ``matrix``: the forward 2x3 (or 3x3) matrix."""
    dw, dh = dsize
    m = np.array(matrix, np.float64)[:2].reshape(-1).copy()
    D = m[0] * m[4] - m[1] * m[3]
    D = 1.0 / D if D != 0 else 0.0
    a11, a22 = m[4] * D, m[0] * D
    m[0] = a11
    m[1] *= -D
    m[3] *= -D
    m[4] = a22
    b1 = -m[0] * m[2] - m[1] * m[5]
    b2 = -m[3] * m[2] - m[4] * m[5]
    m[2], m[5] = b1, b2
    scale = float(1 << _AB_BITS)
    xs, ys = np.arange(dw, dtype=np.float64), np.arange(dh, dtype=np.float64)
    rd = (1 << _AB_BITS) // 2 if nearest else (1 << _AB_BITS) // 32 // 2

    def sat(v):          # saturate_cast<int>(double) = cvRound, clamped
        return np.clip(np.rint(v), -2 ** 31, 2 ** 31 - 1).astype(np.int64)

    tabs = (sat(m[0] * xs * scale), sat(m[3] * xs * scale), sat((m[1] * ys + m[2]) * scale) + rd, sat((m[4] * ys + m[5]) * scale) + rd)
    if any(np.abs(t).max() >= 2 ** 30 for t in tabs):
        raise ValueError("affine matrix maps the output more than 2^20 pixels away from the source: outside cv::warpAffine's int32 walk")
    return tuple(t.astype(np.int32) for t in tabs)


def affine_matrix(h: int, w: int, scale_x: float, scale_y: float, translate_x: float, translate_y: float, rotate_deg: float) -> np.ndarray:
    """albumentations 1.3 ``Affine.get_params_dependent_on_targets`` (shear = 0): move the image centre (w / 2 - 0.5,
    h / 2 - 0.5) to the origin, scale / rotate by MINUS the drawn angle / translate (pixels), move back.  3x3 float64."""
    r = np.deg2rad(-rotate_deg)
    sx, sy = w / 2 - 0.5, h / 2 - 0.5

    def aff(a=1.0, b=1.0, rot=0.0, tx=0.0, ty=0.0):
        return np.array([[a * np.cos(rot), -b * np.sin(rot), tx], [a * np.sin(rot), b * np.cos(rot), ty], [0.0, 0.0, 1.0]], np.float64)

    # the two shear-axis rotations of the original chain (by -pi/2 and +pi/2 around a zero shear) are kept: they are not
    # exactly the identity in floating point and the reference's matrix carries their ~1e-16 residue
    chain = [aff(tx=-sx, ty=-sy), aff(rot=-np.pi / 2), aff(), aff(rot=np.pi / 2), aff(scale_x, scale_y, r, translate_x, translate_y), aff(tx=sx, ty=sy)]
    m = chain[0]
    for nxt in chain[1:]:
        m = nxt @ m
    return m


def brightness_contrast_lut(alpha: float, beta: float) -> np.ndarray:
    """albumentations ``_brightness_contrast_adjust_uint`` (brightness_by_max=True): uint8 [256], float32 arithmetic, truncating cast."""
    lut = np.arange(0, 256).astype(np.float32)
    if alpha != 1:
        lut *= alpha
    if beta != 0:
        lut += beta * 255
    return np.clip(lut, 0, 255).astype(np.uint8)


class GpuTrainTransforms(GpuEvalTransforms):
    """The reference's ``train_transforms`` (clipseg.yaml:80-111) with the pixel work on the GPU; same call signature as the
    albumentations ``Compose``.  Parameters are drawn on the host from ``rng`` (``numpy.random.Generator``); the draw
    order / generator of albumentations itself is not reproduced (it is not installed here) - ``apply`` takes explicit
    parameters and is what the parity tests pin against cv2.
    """

    def __init__(self, img_size: int, mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225), max_pixel_value: float = 255.0,
                 device: str | torch.device = "cuda", scale=(0.98, 1.02), translate_percent=(-0.02, 0.02), rotate=(-5.0, 5.0),
                 affine_p: float = 0.2, brightness_limit: float = 0.1, contrast_limit: float = 0.1, brightness_contrast_p: float = 0.2,
                 keep_ratio: bool = False, seed: int | None = None) -> None:
        super().__init__(img_size, mean, std, max_pixel_value, device)
        self.scale, self.translate_percent, self.rotate = tuple(scale), tuple(translate_percent), tuple(rotate)
        self.affine_p, self.bc_p = float(affine_p), float(brightness_contrast_p)
        self.brightness_limit, self.contrast_limit = float(brightness_limit), float(contrast_limit)
        self.keep_ratio = keep_ratio
        self.rng = np.random.default_rng(seed)
        self._tab = None

    def draw(self) -> dict:
        """One sample's augmentation parameters: {"matrix": 3x3 float64 or None, "lut": uint8 [256] or None}."""
        S = self.size
        matrix = lut = None
        if self.rng.random() < self.affine_p:
            sx = self.rng.uniform(*self.scale)
            sy = sx if self.keep_ratio else self.rng.uniform(*self.scale)
            tx, ty = self.rng.uniform(*self.translate_percent) * S, self.rng.uniform(*self.translate_percent) * S
            matrix = affine_matrix(S, S, sx, sy, tx, ty, self.rng.uniform(*self.rotate))
        if self.rng.random() < self.bc_p:
            alpha = 1.0 + self.rng.uniform(-self.contrast_limit, self.contrast_limit)
            beta = 0.0 + self.rng.uniform(-self.brightness_limit, self.brightness_limit)
            lut = brightness_contrast_lut(alpha, beta)
        return {"matrix": matrix, "lut": lut}

    def _walk(self, matrix, nearest: bool):
        return tuple(torch.from_numpy(t).to(self.device) for t in affine_walk_tables(matrix, (self.size, self.size), nearest))

    def apply(self, image_u8, mask_f32=None, matrix=None, lut=None) -> dict:
        """Resize -> [Affine(matrix)] -> [LUT] -> Normalize -> CHW for the image; resize(nearest) -> [Affine nearest] for the mask."""
        S = self.size
        if matrix is not None and np.allclose(np.asarray(matrix, np.float64)[:2], np.eye(3)[:2]):
            matrix = None               # albumentations.functional.warp_affine: `if is_identity_matrix(matrix): return image` (np.allclose)
        lut_d = None if lut is None else torch.from_numpy(np.ascontiguousarray(lut, dtype=np.uint8)).to(self.device)
        out = torch.empty((3, S, S), dtype=torch.float32, device=self.device)
        if matrix is None and lut is None:
            res = {"image": self.image(image_u8, out)}
        else:
            img = self._to_device(image_u8, torch.uint8)
            if img.dim() != 3 or img.shape[2] != 3:
                raise ValueError(f"image must be uint8 [H, W, 3], got {tuple(img.shape)}")
            resized = torch.empty((S, S, 3), dtype=torch.uint8, device=self.device)
            xo, xc = self._cubic(img.shape[1])
            yo, yc = self._cubic(img.shape[0])
            abi.preproc_image_u8(img.contiguous(), xo, xc, yo, yc, self.mean255, self.inv_std255, out_u8=resized)
            if matrix is None:
                abi.lut_normalize_u8(resized, lut_d, self.mean255, self.inv_std255, out)
            else:
                if self._tab is None:
                    self._tab = torch.from_numpy(warp_cubic_table()).to(self.device).contiguous()
                abi.warp_affine_u8(resized, self._walk(matrix, False), self._tab, self.mean255, self.inv_std255, lut=lut_d, out_chw=out)
            res = {"image": out}
        if mask_f32 is not None:
            m = self.mask(mask_f32)
            if matrix is not None:
                warped = torch.empty_like(m)
                abi.warp_affine_nearest_f32(m[0], self._walk(matrix, True), warped[0])
                m = warped
            res["mask"] = m
        return res

    def __call__(self, *, image, mask=None, **kwargs) -> dict:
        p = self.draw()
        return self.apply(image, mask, p["matrix"], p["lut"])
