"""CPU restatement of the reference's EVAL input transforms (test infrastructure only).

Reference: configs/experiment/coop/clipseg.yaml:113-127 (`_eval_transforms`):
    albumentations.Resize(img_size, img_size, interpolation=cv2.INTER_CUBIC)   -> cv2.resize on the uint8 HWC image,
                                                                                  masks with cv2.INTER_NEAREST
    albumentations.Normalize(mean, std)                                        -> float32 (x - 255 mean) * (1 / (255 std))
    albumentations.pytorch.ToTensorV2(transpose_mask=True)                     -> CHW
applied by src/data/core_datasets/image_text_mask_dataset.py:52-84 to `cv2.imread` output (uint8 RGB) and to the
float32 mask / 255.

Third-party arithmetic not under /root/reference:
  * OpenCV (`opencv-python`, requirements.txt, unpinned; 4.13.0 in this image).  cv2.resize(INTER_CUBIC) on 8-bit images is
    fixed point: per output column / row four taps `saturate_cast<short>(w * 2048)` of the a = -0.75 cubic at
    fx = (d + 0.5) * (n_in / n_out) - 0.5 (float), replicated borders per tap, horizontal pass in int32, vertical pass
    `(sum + 2^21) >> 22`, saturated (modules/imgproc/src/resize.cpp: interpolateCubic, HResizeCubic, VResizeCubic with
    FixedPtCast<int, uchar, INTER_RESIZE_COEF_BITS * 2>).  This file restates exactly that integer definition.
    PINNED against cv2 itself (tests/test_oracle_preprocess.py): identical to OpenCV's own code path (IPP disabled) except
    where its SIMD vertical pass, which works in float, rounds the other way (< 2e-4 of the values, 1 LSB); the IPP-enabled
    build that pip ships differs from OpenCV's own code in ~4.5 % of the values by 1 LSB, so the reference's output is
    build dependent at that level and "1 LSB, >= 95 % identical" is the strongest statement that holds for every build.
    INTER_NEAREST: source index min(floor(d * n_in / n_out), n_in - 1) (resize.cpp resizeNN) - pinned bit-exact.
  * albumentations (requirements.txt, unpinned) is NOT installed here: Normalize / ToTensorV2 are restated from its
    published source (functional.normalize: `img.astype(float32); img -= mean * max_pixel_value; img *= 1 / (std *
    max_pixel_value)`) - parity unpinned for these two steps.
"""
from __future__ import annotations

import numpy as np

COEF_BITS = 11
COEF_SCALE = 1 << COEF_BITS


def cubic_coeffs(x: np.float32) -> np.ndarray:
    """OpenCV interpolateCubic (A = -0.75) in float32, operation order as in the C source."""
    A = np.float32(-0.75)
    x = np.float32(x)
    one = np.float32(1)
    c = np.empty(4, np.float32)
    c[0] = ((A * (x + one) - np.float32(5) * A) * (x + one) + np.float32(8) * A) * (x + one) - np.float32(4) * A
    c[1] = ((A + np.float32(2)) * x - (A + np.float32(3))) * x * x + one
    c[2] = ((A + np.float32(2)) * (one - x) - (A + np.float32(3))) * (one - x) * (one - x) + one
    c[3] = one - c[0] - c[1] - c[2]
    return c


def cubic_tables(n_in: int, n_out: int) -> tuple[np.ndarray, np.ndarray]:
    """(ofs int32 [n_out], coef int16-valued int32 [n_out, 4]): tap k of output d reads clamp(ofs[d] + k - 1)."""
    scale = np.float64(1.0) / (np.float64(n_out) / np.float64(n_in))     # cv::resize: scale_x = 1. / inv_scale_x
    ofs = np.empty(n_out, np.int32)
    coef = np.empty((n_out, 4), np.int32)
    for d in range(n_out):
        fx = np.float32((d + 0.5) * scale - 0.5)
        sx = int(np.floor(fx))
        fx = np.float32(fx - np.float32(sx))
        coef[d] = np.clip(np.rint(cubic_coeffs(fx) * np.float32(COEF_SCALE)), -32768, 32767).astype(np.int32)
        ofs[d] = sx
    return ofs, coef


def resize_cubic_u8(img: np.ndarray, h_out: int, w_out: int) -> np.ndarray:
    """cv2.resize(img, (w_out, h_out), interpolation=cv2.INTER_CUBIC) for uint8 HWC (or HW) images, integer definition."""
    squeeze = img.ndim == 2
    if squeeze:
        img = img[..., None]
    h_in, w_in, _ = img.shape
    if (h_in, w_in) == (h_out, w_out):
        return img[..., 0].copy() if squeeze else img.copy()
    xo, xc = cubic_tables(w_in, w_out)
    yo, yc = cubic_tables(h_in, h_out)
    src = img.astype(np.int64)
    tmp = np.zeros((h_in, w_out, img.shape[2]), np.int64)
    for k in range(4):
        tmp += src[:, np.clip(xo + k - 1, 0, w_in - 1), :] * xc[:, k][None, :, None]
    out = np.zeros((h_out, w_out, img.shape[2]), np.int64)
    for k in range(4):
        out += tmp[np.clip(yo + k - 1, 0, h_in - 1)] * yc[:, k][:, None, None]
    out = np.clip((out + (1 << (2 * COEF_BITS - 1))) >> (2 * COEF_BITS), 0, 255).astype(np.uint8)
    return out[..., 0] if squeeze else out


def nearest_table(n_in: int, n_out: int) -> np.ndarray:
    scale = np.float64(1.0) / (np.float64(n_out) / np.float64(n_in))     # resizeNN: ifx = 1. / inv_scale_x
    return np.minimum(np.floor(np.arange(n_out) * scale).astype(np.int64), n_in - 1).astype(np.int32)


def resize_nearest(img: np.ndarray, h_out: int, w_out: int) -> np.ndarray:
    """cv2.resize(..., interpolation=cv2.INTER_NEAREST) (any dtype, HW or HWC)."""
    yo, xo = nearest_table(img.shape[0], h_out), nearest_table(img.shape[1], w_out)
    return img[yo][:, xo].copy()


def normalize_chw(img_u8: np.ndarray, mean, std, max_pixel_value: float = 255.0) -> np.ndarray:
    """albumentations.Normalize followed by ToTensorV2: float32 CHW."""
    m = np.array(mean, dtype=np.float32) * np.float32(max_pixel_value)
    d = np.reciprocal(np.array(std, dtype=np.float32) * np.float32(max_pixel_value), dtype=np.float32)
    x = img_u8.astype(np.float32)
    x -= m
    x *= d
    return np.ascontiguousarray(x.transpose(2, 0, 1))


def eval_transform(image_u8: np.ndarray, mask_f32: np.ndarray | None, img_size: int, mean, std):
    """The whole `_eval_transforms` pipeline: (image float32 [3, S, S], mask float32 [1, S, S] or None)."""
    img = normalize_chw(resize_cubic_u8(image_u8, img_size, img_size), mean, std)
    if mask_f32 is None:
        return img, None
    m = mask_f32 if mask_f32.ndim == 3 else mask_f32[..., None]
    return img, np.ascontiguousarray(resize_nearest(m, img_size, img_size).transpose(2, 0, 1))
