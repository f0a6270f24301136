"""GPU parity of the eval input transforms (csrc/preprocess.cu through the C ABI) against oracle/preprocess.py:
uint8 resize bit-exact (integer fixed-point arithmetic), normalised float image bit-exact (two float32 roundings, no
FMA), nearest-neighbour mask bit-exact; the committed cv2 vectors within 1 LSB."""
import os

import numpy as np
import pytest
import torch

from oracle import preprocess as OP

pytestmark = pytest.mark.gpu

MEAN, STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "preprocess_cv2.npz")


@pytest.mark.parametrize("hi,wi,size", [(480, 640, 352), (333, 517, 352), (100, 90, 416), (352, 352, 352), (37, 53, 64),
                                        (1080, 1920, 352), (5, 3, 16), (1, 1, 8)])
def test_eval_transforms_match_oracle_bit_exact(hi, wi, size):
    from tunevlseg_b200 import abi
    from tunevlseg_b200.data import GpuEvalTransforms, cubic_tables

    rng = np.random.default_rng(hi * 7919 + wi)
    img = rng.integers(0, 256, (hi, wi, 3), dtype=np.uint8)
    mask = (rng.random((hi, wi, 1)) < 0.3).astype(np.float32)
    t = GpuEvalTransforms(size, MEAN, STD)
    out = t(image=img, mask=mask)
    ref_img, ref_mask = OP.eval_transform(img, mask, size, MEAN, STD)
    assert out["image"].shape == (3, size, size) and out["mask"].shape == (1, size, size)
    assert np.array_equal(out["image"].cpu().numpy(), ref_img)
    assert np.array_equal(out["mask"].cpu().numpy(), ref_mask)
    # the resized uint8 image itself
    xo, xc = (torch.from_numpy(a).cuda() for a in cubic_tables(wi, size))
    yo, yc = (torch.from_numpy(a).cuda() for a in cubic_tables(hi, size))
    u8 = torch.empty((size, size, 3), dtype=torch.uint8, device="cuda")
    abi.preproc_image_u8(torch.from_numpy(img).cuda(), xo, xc.contiguous(), yo, yc.contiguous(), t.mean255, t.inv_std255, out_u8=u8)
    assert np.array_equal(u8.cpu().numpy(), OP.resize_cubic_u8(img, size, size))


def test_eval_transforms_against_committed_cv2_vectors():
    from tunevlseg_b200.data import GpuEvalTransforms

    g = np.load(GOLDEN)
    for name in ("up", "down", "mixed"):
        s = int(g[f"{name}/size"])
        t = GpuEvalTransforms(s, MEAN, STD)
        out = t(image=g[f"{name}/image"], mask=g[f"{name}/mask"])
        ref = OP.normalize_chw(g[f"{name}/cubic"], MEAN, STD)              # cv2's resize, then the Normalize restatement
        lsb = (1.0 / (255.0 * np.array(STD, dtype=np.float32)))[:, None, None]
        assert (np.abs(out["image"].cpu().numpy() - ref) <= lsb * 1.0001).all(), name
        assert np.array_equal(out["mask"].cpu().numpy()[0], g[f"{name}/nearest"]), name


def test_eval_transforms_reject_bad_inputs():
    from tunevlseg_b200.data import GpuEvalTransforms

    t = GpuEvalTransforms(32)
    with pytest.raises(ValueError):
        t.image(np.zeros((8, 8), dtype=np.uint8))
    with pytest.raises(TypeError):
        t.image(np.zeros((8, 8, 3), dtype=np.float32))
    with pytest.raises(ValueError):
        t.mask(np.zeros((8, 8, 2), dtype=np.float32))


# ---------------------------------------------------------------------------------------------------------------------
# train-time augmentations (clipseg.yaml:80-111): cv2.warpAffine / LUT on the device, bit-exact against the cv2-pinned oracle
# ---------------------------------------------------------------------------------------------------------------------
AUG_GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "augment_cv2.npz")


@pytest.mark.parametrize("hi,wi,size,wide", [(480, 640, 352, False), (352, 352, 352, False), (333, 517, 416, False), (97, 131, 64, True),
                                             (37, 53, 48, True), (5, 3, 16, True)])
def test_train_transforms_match_oracle_bit_exact(hi, wi, size, wide):
    from oracle import augment as OA
    from tunevlseg_b200.data import GpuTrainTransforms

    rng = np.random.default_rng(hi * 31 + wi)
    img = rng.integers(0, 256, (hi, wi, 3), dtype=np.uint8)
    mask = (rng.random((hi, wi)) < 0.3).astype(np.float32)
    t = GpuTrainTransforms(size, MEAN, STD, seed=1)
    resized, rmask = OP.resize_cubic_u8(img, size, size), OP.resize_nearest(mask, size, size)
    for trial in range(3):
        if wide:
            M = OA.affine_matrix(size, size, rng.uniform(0.6, 1.5), rng.uniform(0.6, 1.5), rng.uniform(-0.2, 0.2) * size, rng.uniform(-0.2, 0.2) * size,
                                 rng.uniform(-180, 180))
        else:
            M = OA.affine_matrix(size, size, rng.uniform(0.98, 1.02), rng.uniform(0.98, 1.02), rng.uniform(-0.02, 0.02) * size,
                                 rng.uniform(-0.02, 0.02) * size, rng.uniform(-5, 5))
        lut = OA.brightness_contrast_lut(1.0 + rng.uniform(-0.1, 0.1), rng.uniform(-0.1, 0.1))
        for use_m, use_l in ((True, True), (True, False), (False, True), (False, False)):
            out = t.apply(img, mask, M if use_m else None, lut if use_l else None)
            u8 = OA.warp_affine_cubic_u8(resized, M[:2], (size, size)) if use_m else resized
            u8 = OA.apply_lut(u8, lut) if use_l else u8
            ref_mask = OA.warp_affine_nearest(rmask, M[:2], (size, size)) if use_m else rmask
            assert np.array_equal(out["image"].cpu().numpy(), OP.normalize_chw(u8, MEAN, STD)), (trial, use_m, use_l)
            assert np.array_equal(out["mask"].cpu().numpy()[0], ref_mask), (trial, use_m, use_l)


def test_warp_affine_against_committed_cv2_vectors_and_u8_output():
    from tunevlseg_b200 import abi
    from tunevlseg_b200.data import affine_walk_tables, warp_cubic_table

    g = np.load(AUG_GOLDEN)
    tab = torch.from_numpy(warp_cubic_table()).cuda()
    for i in range(int(g["n"])):
        img, mask, M, lut = g[f"{i}/image"], g[f"{i}/mask"], g[f"{i}/matrix"], g[f"{i}/lut"]
        h, w = img.shape[:2]
        walk = tuple(torch.from_numpy(a).cuda() for a in affine_walk_tables(M, (w, h), False))
        u8 = torch.empty((h, w, 3), dtype=torch.uint8, device="cuda")
        abi.warp_affine_u8(torch.from_numpy(img).cuda(), walk, tab, out_u8=u8)
        assert np.array_equal(u8.cpu().numpy(), g[f"{i}/cubic"]), i                     # cv2.warpAffine's own bytes
        abi.warp_affine_u8(torch.from_numpy(img).cuda(), walk, tab, lut=torch.from_numpy(lut).cuda(), out_u8=u8)
        assert np.array_equal(u8.cpu().numpy(), g[f"{i}/lut"][g[f"{i}/cubic"]]), i
        walk_n = tuple(torch.from_numpy(a).cuda() for a in affine_walk_tables(M, (w, h), True))
        out = torch.empty((h, w), dtype=torch.float32, device="cuda")
        abi.warp_affine_nearest_f32(torch.from_numpy(mask).cuda(), walk_n, out)
        assert np.array_equal(out.cpu().numpy(), g[f"{i}/nearest"]), i


def test_train_transforms_draw_ranges_and_call_signature():
    from tunevlseg_b200.data import GpuTrainTransforms

    t = GpuTrainTransforms(64, MEAN, STD, seed=3, affine_p=1.0, brightness_contrast_p=1.0)
    p = t.draw()
    assert p["matrix"].shape == (3, 3) and p["lut"].shape == (256,) and p["lut"].dtype == np.uint8
    s = np.linalg.svd(p["matrix"][:2, :2], compute_uv=False)
    assert 0.979 <= s.min() and s.max() <= 1.021 and abs(p["matrix"][0, 2]) < 64 * 0.2
    rng = np.random.default_rng(0)
    out = t(image=rng.integers(0, 256, (80, 90, 3), dtype=np.uint8), mask=(rng.random((80, 90, 1)) < 0.5).astype(np.float32))
    assert out["image"].shape == (3, 64, 64) and out["mask"].shape == (1, 64, 64) and out["image"].is_cuda
    assert set(np.unique(out["mask"].cpu().numpy())) <= {0.0, 1.0}
    never = GpuTrainTransforms(64, MEAN, STD, seed=3, affine_p=0.0, brightness_contrast_p=0.0).draw()
    assert never["matrix"] is None and never["lut"] is None
