"""CPU restatement of the reference's TRAIN-time augmentations (test infrastructure only).

Reference: configs/experiment/coop/clipseg.yaml:80-111 (`train_transforms`), applied by
src/data/core_datasets/image_text_mask_dataset.py:52-84 to the uint8 RGB image and the float32 mask / 255:

    Resize(S, S, INTER_CUBIC)                                               -> oracle/preprocess.py
    Affine(scale=[0.98, 1.02], translate_percent=[-0.02, 0.02], rotate=[-5, 5], interpolation=INTER_CUBIC,
           mode=BORDER_REPLICATE, p=0.2)                                    -> affine_matrix + warp_affine_cubic_u8 / _nearest
    PadIfNeeded(S, S) / CropNonEmptyMaskIfExists(S, S)                      -> identities on an S x S image
    RandomBrightnessContrast(0.1, 0.1, p=0.2)                               -> brightness_contrast_lut + LUT
    Normalize / ToTensorV2                                                  -> oracle/preprocess.py

Third-party arithmetic not under /root/reference:
  * OpenCV 4.13.0 (`cv2.warpAffine`, `cv2.LUT`): PINNED - tests/test_oracle_augment.py compares every function below with
    cv2 itself, bit for bit.  warpAffine (modules/imgproc/src/imgwarp.cpp): the 2x3 matrix is inverted in double; source
    coordinates are fixed point with AB_BITS = 10 (per-column `adelta = cvRound(M0 x 1024)`, per-row
    `X0 = cvRound((M1 y + M2) 1024) + 16`), reduced to INTER_BITS = 5 fractional bits; the 4x4 weights come from the 32 x 32
    table `initInterTab2D(INTER_CUBIC, fixpt)` = outer products of the float a = -0.75 cubic at k / 32, scaled by 2^15,
    rounded, and corrected so that they sum to 2^15 (the correction looks for the extreme tap among taps [2, 4) x [2, 4) -
    as the C source does; restated as is); the result is `(sum + 2^14) >> 15`, saturated.  BORDER_REPLICATE clamps each tap.
    INTER_NEAREST: `X = (X0' + adelta) >> 10` with the rounding term 512.
  * albumentations (>= 1.2.1, requirements.txt:27) and scikit-image (its `AffineTransform` composes the matrix) are NOT
    installed: `affine_matrix` and `brightness_contrast_lut` restate albumentations 1.3's published source
    (augmentations/geometric/transforms.py `Affine.get_params_dependent_on_targets`, augmentations/functional.py
    `_brightness_contrast_adjust_uint`) - parity unpinned for the matrix composition and the LUT definition; the pixel work
    given a matrix / LUT is pinned to cv2.
"""
from __future__ import annotations

from functools import lru_cache

import numpy as np

from .preprocess import cubic_coeffs

AB_BITS = 10
AB_SCALE = 1 << AB_BITS
INTER_BITS = 5
INTER_TAB_SIZE = 1 << INTER_BITS
REMAP_COEF_BITS = 15
REMAP_COEF_SCALE = 1 << REMAP_COEF_BITS


@lru_cache(maxsize=1)
def cubic_tab2d() -> np.ndarray:
    """int32 [32 (fy), 32 (fx), 4 (ky), 4 (kx)]: cv::initInterTab2D(INTER_CUBIC, fixpt = true)."""
    t1 = np.stack([cubic_coeffs(np.float32(i) * np.float32(1.0 / INTER_TAB_SIZE)) for i in range(INTER_TAB_SIZE)])
    tab = np.empty((INTER_TAB_SIZE, INTER_TAB_SIZE, 4, 4), np.int32)
    for i in range(INTER_TAB_SIZE):
        for j in range(INTER_TAB_SIZE):
            v = (t1[i][:, None] * t1[j][None, :]).astype(np.float32)
            it = np.clip(np.rint(v * np.float32(REMAP_COEF_SCALE)), -32768, 32767).astype(np.int32)
            diff = int(it.sum()) - REMAP_COEF_SCALE
            if diff != 0:
                big = small = (2, 2)                       # ksize / 2 = 2: the C loop scans taps [2, 4) x [2, 4)
                for k1 in range(2, 4):
                    for k2 in range(2, 4):
                        if it[k1, k2] < it[small]:
                            small = (k1, k2)
                        elif it[k1, k2] > it[big]:
                            big = (k1, k2)
                if diff < 0:
                    it[big] -= diff
                else:
                    it[small] -= diff
            tab[i, j] = it
    return tab


def _cv_round(x):
    """cvRound(double): round half to even."""
    return np.rint(x).astype(np.int64)


def invert_affine(M) -> np.ndarray:
    """cv::warpAffine's in-place inversion of the 2x3 forward matrix (double), operation order as in the C source."""
    m = np.array(M, np.float64).reshape(-1).copy()
    D = m[0] * m[4] - m[1] * m[3]
    D = 1.0 / D if D != 0 else 0.0
    A11, A22 = m[4] * D, m[0] * D
    m[0] = A11
    m[1] *= -D
    m[3] *= -D
    m[4] = A22
    b1 = -m[0] * m[2] - m[1] * m[5]
    b2 = -m[3] * m[2] - m[4] * m[5]
    m[2], m[5] = b1, b2
    return m


def warp_tables(M, dsize, nearest: bool):
    """The integer tables cv::warpAffine walks: (adelta [dw], bdelta [dw], X0 [dh], Y0 [dh]) as int64; the rounding term is in
    X0 / Y0.  Source position of (x, y): X = X0[y] + adelta[x], Y = Y0[y] + bdelta[x] in 1 / 1024 pixel."""
    dw, dh = dsize
    m = invert_affine(M)
    xs, ys = np.arange(dw, dtype=np.float64), np.arange(dh, dtype=np.float64)
    adelta = _cv_round(m[0] * xs * AB_SCALE)
    bdelta = _cv_round(m[3] * xs * AB_SCALE)
    rd = AB_SCALE // 2 if nearest else AB_SCALE // INTER_TAB_SIZE // 2
    X0 = _cv_round((m[1] * ys + m[2]) * AB_SCALE) + rd
    Y0 = _cv_round((m[4] * ys + m[5]) * AB_SCALE) + rd
    return adelta, bdelta, X0, Y0


def warp_affine_cubic_u8(src: np.ndarray, M, dsize) -> np.ndarray:
    """cv2.warpAffine(src uint8 [H, W(, C)], M, dsize=(w, h), flags=INTER_CUBIC, borderMode=BORDER_REPLICATE)."""
    H, W = src.shape[:2]
    dw, dh = dsize
    tab = cubic_tab2d()
    adelta, bdelta, X0, Y0 = warp_tables(M, dsize, nearest=False)
    X = (X0[:, None] + adelta[None, :]) >> (AB_BITS - INTER_BITS)
    Y = (Y0[:, None] + bdelta[None, :]) >> (AB_BITS - INTER_BITS)
    sx = np.clip(X >> INTER_BITS, -32768, 32767) - 1
    sy = np.clip(Y >> INTER_BITS, -32768, 32767) - 1
    w = tab[Y & (INTER_TAB_SIZE - 1), X & (INTER_TAB_SIZE - 1)].astype(np.int64)          # [dh, dw, 4, 4]
    acc = np.zeros((dh, dw) + src.shape[2:], np.int64)
    for k1 in range(4):
        yy = np.clip(sy + k1, 0, H - 1)
        for k2 in range(4):
            xx = np.clip(sx + k2, 0, W - 1)
            wk = w[:, :, k1, k2]
            acc += src[yy, xx].astype(np.int64) * (wk[..., None] if src.ndim == 3 else wk)
    return np.clip((acc + (1 << (REMAP_COEF_BITS - 1))) >> REMAP_COEF_BITS, 0, 255).astype(np.uint8)


def warp_affine_nearest(src: np.ndarray, M, dsize) -> np.ndarray:
    """cv2.warpAffine(src [H, W(, C)] any dtype, M, dsize, flags=INTER_NEAREST, borderMode=BORDER_REPLICATE) - the mask path."""
    H, W = src.shape[:2]
    adelta, bdelta, X0, Y0 = warp_tables(M, dsize, nearest=True)
    sx = np.clip((X0[:, None] + adelta[None, :]) >> AB_BITS, -32768, 32767)
    sy = np.clip((Y0[:, None] + bdelta[None, :]) >> AB_BITS, -32768, 32767)
    return src[np.clip(sy, 0, H - 1), np.clip(sx, 0, W - 1)]


def affine_matrix(h: int, w: int, scale_x: float, scale_y: float, translate_x: float, translate_y: float, rotate_deg: float,
                  shear_x_deg: float = 0.0, shear_y_deg: float = 0.0) -> np.ndarray:
    """The 3x3 matrix albumentations 1.3 `Affine` hands to cv2.warpAffine (rows [:2]).  ``translate_*`` in pixels
    (translate_percent x size), ``rotate_deg`` / ``shear_*_deg`` the DRAWN values (the transform negates them itself).
    skimage semantics: `a + b` applies a first, and AffineTransform(scale, rotation, shear, translation) has
    [[sx cos r, -sy sin(r + s), tx], [sx sin r, sy cos(r + s), ty]]."""
    def aff(sx=1.0, sy=1.0, rot=0.0, shear=0.0, tx=0.0, ty=0.0):
        return np.array([[sx * np.cos(rot), -sy * np.sin(rot + shear), tx],
                         [sx * np.sin(rot), sy * np.cos(rot + shear), ty],
                         [0.0, 0.0, 1.0]], np.float64)

    rot = np.deg2rad(-rotate_deg)
    shx, shy = np.deg2rad(-shear_x_deg), np.deg2rad(-shear_y_deg)
    shift_x, shift_y = w / 2 - 0.5, h / 2 - 0.5
    chain = [aff(tx=-shift_x, ty=-shift_y), aff(rot=-np.pi / 2), aff(shear=shy), aff(rot=np.pi / 2),
             aff(sx=scale_x, sy=scale_y, rot=rot, shear=shx, tx=translate_x, ty=translate_y), aff(tx=shift_x, ty=shift_y)]
    m = chain[0]
    for nxt in chain[1:]:
        m = nxt @ m
    return m


def brightness_contrast_lut(alpha: float, beta: float, max_value: int = 255) -> np.ndarray:
    """albumentations `_brightness_contrast_adjust_uint(img, alpha, beta, beta_by_max=True)`: a 256-entry uint8 table
    (alpha = 1 + contrast draw, beta = brightness draw; float32 arithmetic, clip, truncating cast), applied with cv2.LUT."""
    lut = np.arange(0, max_value + 1).astype(np.float32)
    if alpha != 1:
        lut *= alpha
    if beta != 0:
        lut += beta * max_value
    return np.clip(lut, 0, max_value).astype(np.uint8)


def apply_lut(img_u8: np.ndarray, lut: np.ndarray) -> np.ndarray:
    """cv2.LUT on uint8."""
    return lut[img_u8]
