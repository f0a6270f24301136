"""ORACLE (test infrastructure only - never imported by the product): the predict tail's resize + PNG quantisation,
restated with numpy fp32 scalars and rationals, operation for operation as the reference evaluates it.

Reference: /root/reference/src/utils/save_utils.py:74-104 - per sample ``TF.resize(pred.float(), mask_shape, BICUBIC,
antialias=False)`` on the HOST tensor Lightning's predict loop returns, then ``torchvision.utils.save_image``
(``mul(255).add_(0.5).clamp_(0, 255).to(uint8)``).  ``TF.resize`` is ``F.interpolate(mode="bicubic",
align_corners=False)``, i.e. ATen's CPU kernel (aten/src/ATen/native/cpu/UpSampleKernel.cpp, ``HelperInterpCubic`` +
``Interpolate<2, float, float, int64_t, 4>``) as compiled for x86 with FMA contraction:

    scale = fl(n_in / n_out);  src = fma(scale, o + 0.5, -0.5);  t = clamp(src - floor(src), 0, 1)
    conv1(x) = fl(fl(fl(fma(1.25, x, -2.25) * x) * x) + 1)          (A = -0.75:  A + 2 = 1.25, A + 3 = 2.25)
    conv2(x) = fl(fl(fma(fma(-0.75, x, 3.75), x, -6) * x) + 3)      (-5A = 3.75, 8A = -6, -4A = 3)
    w = (conv2(t + 1), conv1(t), conv1(1 - t), conv2((1 - t) + 1)),  taps floor(src) - 1 .. + 2 clamped to the border
    sum4(v, w) = fma(v3, w3, fma(v2, w2, fma(v0, w0, fl(v1 * w1))));  out = sum4_y(sum4_x(...))

Pinned: tests/test_oracle_resize_u8.py compares this file with ``F.interpolate`` itself (bit for bit, several shapes).
"""
from __future__ import annotations

from fractions import Fraction

import numpy as np

f32 = np.float32


def fma32(a, b, c) -> np.float32:
    """fl32(a*b + c), one rounding (exact rational arithmetic, ties to even)."""
    exact = Fraction(float(a)) * Fraction(float(b)) + Fraction(float(c))
    guess = f32(float(exact))
    cands = [np.nextafter(guess, f32(-np.inf)), guess, np.nextafter(guess, f32(np.inf))]
    return f32(min(cands, key=lambda v: (abs(Fraction(float(v)) - exact), int(f32(v).view(np.uint32)) & 1)))


def _conv1(x):
    return f32(f32(f32(fma32(f32(1.25), x, f32(-2.25)) * x) * x) + f32(1))


def _conv2(x):
    return f32(f32(fma32(fma32(f32(-0.75), x, f32(3.75)), x, f32(-6)) * x) + f32(3))


def taps(n_in: int, n_out: int):
    """-> (idx int64 [n_out, 4], w float32 [n_out, 4])"""
    scale = f32(n_in) / f32(n_out)
    idx = np.zeros((n_out, 4), np.int64)
    w = np.zeros((n_out, 4), np.float32)
    for o in range(n_out):
        src = fma32(scale, f32(o) + f32(0.5), f32(-0.5))
        fl = int(np.floor(src))
        t = min(max(f32(src - f32(fl)), f32(0)), f32(1))
        ws = (_conv2(f32(t + f32(1))), _conv1(t), _conv1(f32(f32(1) - t)), _conv2(f32(f32(f32(1) - t) + f32(1))))
        for a in range(4):
            idx[o, a] = min(max(fl - 1 + a, 0), n_in - 1)
            w[o, a] = ws[a]
    return idx, w


def _fma_vec(a, b, c):
    # fl32(a*b + c) for arrays: the product of two fp32 is exact in fp64; the fp64 sum is then rounded twice (53 -> 24 bits).
    # A double-rounding slip needs the 53-bit result to land exactly on a 24-bit midpoint (~2^-29 per element); the
    # rational fix-up below removes even that for the (vanishingly few) suspicious elements.
    r64 = a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)
    out = r64.astype(np.float32)
    # elements whose fp64 value sits exactly half way between two fp32 neighbours: redo exactly
    lo = np.nextafter(out, f32(-np.inf)).astype(np.float64)
    hi = np.nextafter(out, f32(np.inf)).astype(np.float64)
    o64 = out.astype(np.float64)
    sus = (r64 - o64 == (hi - o64) / 2) | (r64 - o64 == (lo - o64) / 2)
    if sus.any():
        a_, b_, c_ = (np.broadcast_to(v, out.shape) for v in (a, b, c))
        for i in zip(*np.nonzero(sus)):
            out[i] = fma32(a_[i], b_[i], c_[i])
    return out


def _sum4(v, w):
    out = _fma_vec(v[..., 0], w[..., 0], (v[..., 1] * w[..., 1]).astype(np.float32))
    out = _fma_vec(v[..., 2], w[..., 2], out)
    return _fma_vec(v[..., 3], w[..., 3], out)


def resize_bicubic(x: np.ndarray, ho: int, wo: int) -> np.ndarray:
    """x: float32 (h, w) -> float32 (ho, wo), bit for bit F.interpolate(mode='bicubic', align_corners=False) on the CPU."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    iy, wy = taps(x.shape[0], ho)
    ix, wx = taps(x.shape[1], wo)
    g = x[iy][:, :, ix]                              # (ho, 4y, wo, 4x)
    inner = _sum4(g, wx[None, None])                 # (ho, 4y, wo)
    return _sum4(np.moveaxis(inner, 1, -1), wy[:, None, :])


def quantise_u8(v: np.ndarray) -> np.ndarray:
    """torchvision.utils.save_image: mul(255).add_(0.5).clamp_(0, 255).to(uint8) (two fp32 roundings, then truncation)."""
    q = ((v.astype(np.float32) * f32(255)).astype(np.float32) + f32(0.5)).astype(np.float32)
    return np.clip(q, 0, 255).astype(np.uint8)


def resize_to_png_array(pred: np.ndarray, mask_shape) -> np.ndarray:
    ho, wo = (int(v) for v in mask_shape)
    return quantise_u8(resize_bicubic(pred.reshape(pred.shape[-2], pred.shape[-1]), ho, wo))
