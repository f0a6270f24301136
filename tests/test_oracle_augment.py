"""The train-augmentation oracle (oracle/augment.py) against the reference's own third-party code (cv2.warpAffine, cv2.LUT)
and committed golden vectors.  No GPU needed."""
import os

import numpy as np
import pytest

from oracle import augment as OA

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "augment_cv2.npz")


def _matrices(rng, h, w, n, wide=False):
    out = []
    for _ in range(n):
        if wide:          # far outside the reference's ranges: large rotations, anisotropic scale, shear, big shifts
            out.append(OA.affine_matrix(h, w, rng.uniform(0.5, 1.7), rng.uniform(0.5, 1.7), rng.uniform(-0.3, 0.3) * w,
                                        rng.uniform(-0.3, 0.3) * h, rng.uniform(-180, 180), rng.uniform(-20, 20), rng.uniform(-20, 20)))
        else:             # clipseg.yaml:84-91: scale [0.98, 1.02], translate_percent [-0.02, 0.02], rotate [-5, 5]
            out.append(OA.affine_matrix(h, w, rng.uniform(0.98, 1.02), rng.uniform(0.98, 1.02), rng.uniform(-0.02, 0.02) * w,
                                        rng.uniform(-0.02, 0.02) * h, rng.uniform(-5, 5)))
    return out


def test_oracle_matches_committed_cv2_vectors():
    g = np.load(GOLDEN)
    n = int(g["n"])
    for i in range(n):
        img, mask, M = g[f"{i}/image"], g[f"{i}/mask"], g[f"{i}/matrix"]
        h, w = img.shape[:2]
        assert np.array_equal(OA.warp_affine_cubic_u8(img, M[:2], (w, h)), g[f"{i}/cubic"]), i
        assert np.array_equal(OA.warp_affine_nearest(mask, M[:2], (w, h)), g[f"{i}/nearest"]), i
        assert np.array_equal(OA.apply_lut(img, g[f"{i}/lut"]), g[f"{i}/lut_out"]), i


def test_warp_affine_against_cv2_bit_exact():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(11)
    for (h, w), wide in (((352, 352), False), ((416, 416), False), ((97, 131), True), ((64, 40), True), ((5, 7), True), ((352, 352), True)):
        for M in _matrices(rng, h, w, 3, wide):
            img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
            ref = cv2.warpAffine(img, M[:2], (w, h), flags=cv2.INTER_CUBIC, borderMode=cv2.BORDER_REPLICATE)
            assert np.array_equal(OA.warp_affine_cubic_u8(img, M[:2], (w, h)), ref), (h, w, wide)
            gray = img[..., 0].copy()
            ref = cv2.warpAffine(gray, M[:2], (w, h), flags=cv2.INTER_CUBIC, borderMode=cv2.BORDER_REPLICATE)
            assert np.array_equal(OA.warp_affine_cubic_u8(gray, M[:2], (w, h)), ref)
            mask = (rng.random((h, w)) < 0.4).astype(np.float32)
            ref = cv2.warpAffine(mask, M[:2], (w, h), flags=cv2.INTER_NEAREST, borderMode=cv2.BORDER_REPLICATE)
            assert np.array_equal(OA.warp_affine_nearest(mask, M[:2], (w, h)), ref), (h, w, wide)


def test_warp_affine_non_square_output_and_saturation():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(3)
    img = np.where(rng.random((60, 80, 3)) < 0.5, 0, 255).astype(np.uint8)          # cubic overshoot must saturate, not wrap
    M = OA.affine_matrix(60, 80, 1.3, 0.8, 3.5, -2.25, 33.0)
    ref = cv2.warpAffine(img, M[:2], (50, 90), flags=cv2.INTER_CUBIC, borderMode=cv2.BORDER_REPLICATE)
    got = OA.warp_affine_cubic_u8(img, M[:2], (50, 90))
    assert got.shape == (90, 50, 3) and np.array_equal(got, ref)
    assert got.min() == 0 and got.max() == 255


def test_interpolation_table_properties():
    tab = OA.cubic_tab2d()
    assert tab.shape == (32, 32, 4, 4) and (tab.reshape(32, 32, 16).sum(-1) == OA.REMAP_COEF_SCALE).all()
    # zero fraction: the pixel itself - 2^15 saturates to a short (32767) and the sum correction puts the missing 1 on tap (2, 2)
    assert tab[0, 0, 1, 1] == 32767 and tab[0, 0, 2, 2] == 1 and np.count_nonzero(tab[0, 0]) == 2
    assert tab.min() >= -32768 and tab.max() <= 32767


def test_identity_matrix_is_a_copy_and_matrix_composition():
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, (33, 47, 3), dtype=np.uint8)
    I = OA.affine_matrix(33, 47, 1.0, 1.0, 0.0, 0.0, 0.0)
    assert np.allclose(I, np.eye(3), atol=1e-12)
    assert np.array_equal(OA.warp_affine_cubic_u8(img, np.eye(3)[:2], (47, 33)), img)
    # a pure rotation keeps the image centre (w / 2 - 0.5, h / 2 - 0.5) fixed; the transform negates the drawn angle
    R = OA.affine_matrix(33, 47, 1.0, 1.0, 0.0, 0.0, 30.0)
    c = np.array([47 / 2 - 0.5, 33 / 2 - 0.5, 1.0])
    assert np.allclose(R @ c, c, atol=1e-9)
    assert np.allclose(R[:2, :2], [[np.cos(np.deg2rad(-30)), -np.sin(np.deg2rad(-30))], [np.sin(np.deg2rad(-30)), np.cos(np.deg2rad(-30))]])
    T = OA.affine_matrix(33, 47, 1.0, 1.0, 2.5, -1.5, 0.0)
    assert np.allclose(T @ np.array([0.0, 0.0, 1.0]), [2.5, -1.5, 1.0], atol=1e-9)


def test_brightness_contrast_lut_against_cv2_and_known_answers():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(9)
    img = rng.integers(0, 256, (40, 30, 3), dtype=np.uint8)
    for alpha, beta in ((1.0, 0.0), (1.1, 0.0), (0.9, 0.1), (1.05, -0.1), (1.0, 0.07)):
        lut = OA.brightness_contrast_lut(alpha, beta)
        assert lut.dtype == np.uint8 and lut.shape == (256,)
        assert np.array_equal(OA.apply_lut(img, lut), cv2.LUT(img, lut))
        x = np.arange(256, dtype=np.float64) * alpha + beta * 255
        assert np.abs(lut.astype(np.float64) - np.clip(np.floor(x + 1e-9), 0, 255)).max() <= 1          # truncating cast, float32 rounding
    assert np.array_equal(OA.brightness_contrast_lut(1.0, 0.0), np.arange(256, dtype=np.uint8))
    assert OA.brightness_contrast_lut(1.1, 0.1)[255] == 255 and OA.brightness_contrast_lut(0.9, -0.1)[0] == 0


def test_product_tables_equal_oracle_tables():
    """The host-side tables the product uploads (tunevlseg_b200/data/gpu_transforms.py) against the cv2-pinned oracle's."""
    from tunevlseg_b200.data import affine_matrix, affine_walk_tables, brightness_contrast_lut, warp_cubic_table

    assert np.array_equal(warp_cubic_table().astype(np.int32), OA.cubic_tab2d())
    rng = np.random.default_rng(21)
    for h, w in ((352, 352), (416, 416), (90, 47)):
        for _ in range(4):
            p = (rng.uniform(0.9, 1.1), rng.uniform(0.9, 1.1), rng.uniform(-9, 9), rng.uniform(-9, 9), rng.uniform(-30, 30))
            M = affine_matrix(h, w, *p)
            assert np.array_equal(M, OA.affine_matrix(h, w, *p))
            for nearest in (False, True):
                mine, ref = affine_walk_tables(M, (w, h), nearest), OA.warp_tables(M[:2], (w, h), nearest)
                assert all(np.array_equal(a.astype(np.int64), b) for a, b in zip(mine, ref))
    for alpha, beta in ((1.0, 0.0), (1.07, -0.03), (0.92, 0.1)):
        assert np.array_equal(brightness_contrast_lut(alpha, beta), OA.brightness_contrast_lut(alpha, beta))
    with pytest.raises(ValueError):
        affine_walk_tables(np.array([[1e-7, 0, 0], [0, 1e-7, 0]]), (8, 8), False)
