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
