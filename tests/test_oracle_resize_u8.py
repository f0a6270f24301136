"""CPU: pin oracle/resize_u8.py (the predict tail's resize + PNG quantisation) to torch itself - ``F.interpolate`` on a
host tensor is exactly what the reference's save_predictions runs (src/utils/save_utils.py:74-104) - and pin the product's
tap tables (``engine_cris._bicubic_tables(aten_cpu=True)``), which the CUDA kernel consumes, to the oracle's."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import resize_u8 as OR

CASES = [((352, 352), (480, 640)), ((352, 352), (200, 133)), ((416, 416), (416, 416)), ((64, 64), (301, 7)), ((37, 50), (111, 64))]


@pytest.mark.parametrize("src,dst", CASES)
def test_oracle_resize_is_aten_cpu_bit_for_bit(src, dst):
    g = torch.Generator().manual_seed(src[0] + dst[1])
    pred = torch.sigmoid(torch.randn(1, 1, *src, generator=g) * 3)
    ref = F.interpolate(pred, size=list(dst), mode="bicubic", align_corners=False)[0, 0]
    ours = OR.resize_bicubic(pred[0, 0].numpy(), *dst)
    assert np.array_equal(ours, ref.numpy()), f"{(ours != ref.numpy()).sum()} of {ours.size} values differ"
    ref_u8 = ref.mul(255).add_(0.5).clamp_(0, 255).to(torch.uint8).numpy()          # torchvision.utils.save_image
    assert np.array_equal(OR.quantise_u8(ours), ref_u8)


def test_oracle_matches_torchvision_resize():
    tv = pytest.importorskip("torchvision.transforms.functional")
    g = torch.Generator().manual_seed(5)
    pred = torch.sigmoid(torch.randn(1, 64, 48, generator=g) * 3)
    ref = tv.resize(pred, size=[99, 131], interpolation=tv.InterpolationMode.BICUBIC, antialias=False)
    assert np.array_equal(OR.resize_bicubic(pred[0].numpy(), 99, 131), ref[0].numpy())


@pytest.mark.parametrize("n_in,n_out", [(352, 480), (352, 133), (64, 301), (416, 416), (500, 1333), (37, 1000)])
def test_product_tap_tables_equal_the_oracle(n_in, n_out):
    from tunevlseg_b200.engine_cris import _bicubic_tables

    idx, w = OR.taps(n_in, n_out)
    p_idx, p_w, *_ = _bicubic_tables(n_in, n_out, "cpu", align_corners=False, aten_cpu=True)
    assert np.array_equal(p_idx.numpy().astype(np.int64), idx)
    assert np.array_equal(p_w.numpy(), w)
    # and the weights ARE aten's: interpolating the identity along one axis exposes them
    ref = F.interpolate(torch.eye(n_in)[None, None], size=(n_out, n_in), mode="bicubic", align_corners=False)[0, 0].numpy()
    dense = np.zeros((n_out, n_in), np.float32)
    for a in range(4):                                  # border taps that clamp to the same row add up in aten's order
        np.add.at(dense, (np.arange(n_out), idx[:, a]), 0)
    rows = [o for o in range(n_out) if len(set(idx[o])) == 4]
    for a in range(4):
        assert np.array_equal(ref[rows, idx[rows, a]], w[rows, a])
