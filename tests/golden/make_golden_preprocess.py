"""Golden vectors for the eval input transforms, produced by the reference's own third-party code path: cv2.resize with
INTER_CUBIC / INTER_NEAREST (OpenCV's own implementation: IPP disabled, so that the fixture does not depend on the vendor
library a particular wheel ships) on seeded random images.  Run from the repo root:

    python tests/golden/make_golden_preprocess.py

Writes tests/golden/preprocess_cv2.npz (a few KB)."""
import os

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def main() -> None:
    cv2.ipp.setUseIPP(False)
    rng = np.random.default_rng(20240)
    out = {"cv2_version": np.array(cv2.__version__)}
    for name, (hi, wi, s) in {"up": (23, 31, 48), "down": (75, 61, 40), "mixed": (30, 90, 44)}.items():
        img = rng.integers(0, 256, (hi, wi, 3), dtype=np.uint8)
        mask = (rng.random((hi, wi)) < 0.4).astype(np.float32)
        out[f"{name}/image"] = img
        out[f"{name}/mask"] = mask
        out[f"{name}/size"] = np.array(s)
        out[f"{name}/cubic"] = cv2.resize(img, (s, s), interpolation=cv2.INTER_CUBIC)
        out[f"{name}/nearest"] = cv2.resize(mask, (s, s), interpolation=cv2.INTER_NEAREST)
    np.savez_compressed(os.path.join(HERE, "preprocess_cv2.npz"), **out)
    print("wrote", os.path.join(HERE, "preprocess_cv2.npz"))


if __name__ == "__main__":
    main()
