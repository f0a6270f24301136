"""The eval-transform oracle (oracle/preprocess.py) against the reference's own third-party code (cv2) and its golden
vectors, and the product's host-side tap tables against the oracle's.  No GPU needed."""
import os

import numpy as np
import pytest

from oracle import preprocess as OP

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "preprocess_cv2.npz")
MEAN, STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)


def test_oracle_matches_committed_cv2_vectors():
    g = np.load(GOLDEN)
    for name in ("up", "down", "mixed"):
        s = int(g[f"{name}/size"])
        mine = OP.resize_cubic_u8(g[f"{name}/image"], s, s)
        diff = np.abs(mine.astype(int) - g[f"{name}/cubic"].astype(int))
        # OpenCV's SIMD vertical pass works in float and may round the other way on isolated values (oracle header)
        assert diff.max() <= 1 and (diff == 0).mean() >= 0.999, (name, diff.max(), (diff == 0).mean())
        assert np.array_equal(OP.resize_nearest(g[f"{name}/mask"], s, s), g[f"{name}/nearest"]), name


def test_oracle_against_cv2_if_present():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(7)
    had_ipp = cv2.ipp.useIPP()
    try:
        for ipp, min_equal in ((False, 0.999), (True, 0.94)):
            cv2.ipp.setUseIPP(ipp)
            for hi, wi, s in ((480, 640, 352), (333, 517, 352), (100, 90, 416), (352, 352, 352), (37, 53, 64)):
                img = rng.integers(0, 256, (hi, wi, 3), dtype=np.uint8)
                diff = np.abs(cv2.resize(img, (s, s), interpolation=cv2.INTER_CUBIC).astype(int) - OP.resize_cubic_u8(img, s, s).astype(int))
                assert diff.max() <= 1 and (diff == 0).mean() >= min_equal, (ipp, hi, wi, s, diff.max(), (diff == 0).mean())
                m = rng.random((hi, wi)).astype(np.float32)
                assert np.array_equal(cv2.resize(m, (s, s), interpolation=cv2.INTER_NEAREST), OP.resize_nearest(m, s, s))
    finally:
        cv2.ipp.setUseIPP(had_ipp)


def test_identity_size_is_a_copy_and_normalize_formula():
    rng = np.random.default_rng(1)
    img = rng.integers(0, 256, (16, 16, 3), dtype=np.uint8)
    assert np.array_equal(OP.resize_cubic_u8(img, 16, 16), img)
    ofs, coef = OP.cubic_tables(16, 16)
    assert np.array_equal(ofs, np.arange(16)) and np.array_equal(coef, np.tile([0, 2048, 0, 0], (16, 1)))
    x = OP.normalize_chw(img, MEAN, STD)
    assert x.shape == (3, 16, 16) and x.dtype == np.float32
    ref = (img.astype(np.float64) / 255.0 - np.array(MEAN)) / np.array(STD)
    assert np.abs(x - ref.transpose(2, 0, 1)).max() < 1e-5
    im, mk = OP.eval_transform(img, (rng.random((16, 16)) < 0.5).astype(np.float32), 8, MEAN, STD)
    assert im.shape == (3, 8, 8) and mk.shape == (1, 8, 8)


def test_product_tables_equal_oracle_tables():
    from tunevlseg_b200.data import cubic_tables, nearest_table

    for n_in, n_out in ((640, 352), (480, 352), (90, 416), (352, 352), (1000, 352), (53, 64), (1, 8), (7, 3)):
        o1, c1 = cubic_tables(n_in, n_out)
        o2, c2 = OP.cubic_tables(n_in, n_out)
        assert np.array_equal(o1, o2) and np.array_equal(c1, c2)
        assert np.array_equal(nearest_table(n_in, n_out), OP.nearest_table(n_in, n_out))
        assert np.abs(c1.sum(1) - 2048).max() <= 2          # the taps of a cubic sum to one (up to the short rounding)
