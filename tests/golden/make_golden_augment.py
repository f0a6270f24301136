"""Golden vectors for the train-time augmentations: cv2.warpAffine (INTER_CUBIC / INTER_NEAREST, BORDER_REPLICATE) and cv2.LUT
outputs on seeded inputs, generated with the cv2 of this image (4.13.0).  python tests/golden/make_golden_augment.py"""
import os
import sys

import cv2
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import augment as OA  # noqa: E402  (matrix composition + LUT definition only; the pixels below come from cv2)

rng = np.random.default_rng(2024)
out = {}
cases = [((48, 64), False), ((64, 64), False), ((37, 29), True), ((80, 56), True)]
for i, ((h, w), wide) in enumerate(cases):
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    mask = (rng.random((h, w)) < 0.35).astype(np.float32)
    if wide:
        M = OA.affine_matrix(h, w, rng.uniform(0.6, 1.5), rng.uniform(0.6, 1.5), rng.uniform(-8, 8), rng.uniform(-8, 8), rng.uniform(-170, 170))
    else:
        M = OA.affine_matrix(h, w, rng.uniform(0.98, 1.02), rng.uniform(0.98, 1.02), rng.uniform(-0.02, 0.02) * w, rng.uniform(-0.02, 0.02) * h,
                             rng.uniform(-5, 5))
    lut = OA.brightness_contrast_lut(1.0 + rng.uniform(-0.1, 0.1), rng.uniform(-0.1, 0.1))
    out[f"{i}/image"], out[f"{i}/mask"], out[f"{i}/matrix"], out[f"{i}/lut"] = img, mask, M, lut
    out[f"{i}/cubic"] = cv2.warpAffine(img, M[:2], (w, h), flags=cv2.INTER_CUBIC, borderMode=cv2.BORDER_REPLICATE)
    out[f"{i}/nearest"] = cv2.warpAffine(mask, M[:2], (w, h), flags=cv2.INTER_NEAREST, borderMode=cv2.BORDER_REPLICATE)
    out[f"{i}/lut_out"] = cv2.LUT(img, lut)
out["n"] = np.int64(len(cases))
p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "augment_cv2.npz")
np.savez_compressed(p, **out)
print(p, os.path.getsize(p), "bytes, cv2", cv2.__version__)
