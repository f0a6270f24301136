"""GPU parity of the fused decoder feed-forward kernel (csrc/ffn_sm100.cu) against an fp64 torch evaluation of
CLIPSegMLP inside CLIPSegDecoderLayer (transformers modeling_clipseg.py:341-354, :421-431): forward with the residual,
and the dgrad against autograd.  Called through the C ABI (tvs_ffn64_fwd / tvs_ffn64_bwd)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

FWD_TOL = 2e-4          # split (head + tail) operands: ~2^-17 relative on values of magnitude ~5
BF16_TOL = 5e-2         # heads only


def _case(M, F, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    r = lambda *s, std=1.0: torch.randn(*s, generator=g, device="cuda") * std
    return dict(x=r(M, 64), g=r(M, 64), w1=r(F, 64, std=0.125), w2=r(64, F, std=0.03), b1=r(F, std=0.3), b2=r(64, std=0.3))


def _reference(c):
    x = c["x"].double().requires_grad_(True)
    pre = x @ c["w1"].double().t() + c["b1"].double()
    out = x + torch.relu(pre) @ c["w2"].double().t() + c["b2"].double()
    (dx,) = torch.autograd.grad(out, x, c["g"].double())
    # rows whose ReLU mask is decided within rounding distance of zero are not comparable element by element
    safe = (pre.abs() > 1e-4).all(dim=1)
    return out.detach(), dx, safe


@pytest.mark.parametrize("M,F", [(128, 64), (300, 256), (15648, 2048)])
@pytest.mark.parametrize("head_only", [False, True])
def test_ffn64_matches_fp64(M, F, head_only):
    from tunevlseg_b200 import abi, engine

    c = _case(M, F, 7 + M)
    w1 = engine.split_bf16(c["w1"], head_only)
    w2t = engine.split_bf16(c["w2"].t(), head_only)
    out, dx = torch.empty_like(c["x"]), torch.empty_like(c["x"])
    abi.ffn64_fwd(c["x"], w1, w2t, c["b1"], c["b2"], out)
    abi.ffn64_bwd(c["x"], c["g"], w1, w2t, c["b1"], dx)
    ref_out, ref_dx, safe = _reference(c)
    tol = BF16_TOL if head_only else FWD_TOL
    assert (out.double() - ref_out).abs().max().item() < tol
    if not head_only:       # with bf16 heads only the recomputed mask itself differs near zero
        assert safe.any()
        assert (dx.double() - ref_dx)[safe].abs().max().item() < tol


def test_ffn64_rejects_bad_shapes():
    from tunevlseg_b200 import abi, engine

    c = _case(64, 96, 3)          # F not a multiple of 64
    w1, w2t = engine.split_bf16(c["w1"]), engine.split_bf16(c["w2"].t())
    with pytest.raises(abi.TvsError):
        abi.ffn64_fwd(c["x"], w1, w2t, c["b1"], c["b2"], torch.empty_like(c["x"]))
