"""CPU emulation of the C ABI (TEST INFRASTRUCTURE ONLY) so that the host-side composition of the engines - which
kernel is called with which operand, layout, stride and flag - is exercised by the ``-m "not gpu"`` suite.

Every function restates, with plain fp32 torch ops, the contract documented for its entry point in
``include/tvs_b200.h``; ``install(monkeypatch)`` swaps them into ``tunevlseg_b200.abi``.  The product never imports
this module: without it (and without a GPU) every abi call raises ``TvsError``.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from tunevlseg_b200 import abi

F32, BF16 = torch.float32, torch.bfloat16


def _qg(x):
    return x * torch.sigmoid(1.702 * x)


def _qg_grad(x):
    s = torch.sigmoid(1.702 * x)
    return s * (1 + 1.702 * x * (1 - s))


def _tf32(x):
    """What the kind::tf32 MMA does to an fp32 operand: the low 13 mantissa bits are TRUNCATED (measured on B200)."""
    return (x.contiguous().view(torch.int32) & ~0x1FFF).view(torch.float32)


def _tf32_rn(x):
    """cvt.rna.tf32.f32: round to nearest (ties away from zero) on the 10-bit mantissa."""
    return ((x.contiguous().view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)


def round_tf32(x, y):
    y.copy_(_tf32_rn(x))


def gemm(A, W, *, bias=None, residual=None, out_f32=None, out_bf16=None, pre_bf16=None, aux_bf16=None, act=abi.ACT_NONE, tile_n=0, round_out=False, conv_hw=None,
         overwrite=None, stream_k=False, pre_is_grad=False):
    H16 = (BF16, torch.float16)          # kind::f16 takes two bf16 or two fp16 operands (mixing is illegal on sm_100); kind::tf32 two fp32
    assert A.dtype == W.dtype and A.dtype in (BF16, torch.float16, F32), (A.dtype, W.dtype)
    assert A.dim() == 2 and W.dim() == 2, (A.shape, W.shape)
    assert A.stride(1) == 1 and W.stride(1) == 1 and A.shape[1] % (8 if A.dtype in H16 else 4) == 0
    Ar, Wr = (A.float(), W.float()) if A.dtype in H16 else (_tf32(A), _tf32(W))
    if conv_hw is not None:          # implicit-GEMM 3x3 conv: A = zero-bordered image [B*(H+2)*(W+2), C], W = [N, (ky, kx, c)]
        H_, W_ = conv_hw
        C, N = A.shape[1], W.shape[0]
        assert A.is_contiguous() and W.shape[1] == 9 * C and C % (64 if A.dtype in H16 else 32) == 0 and N % 32 == 0
        B_ = A.shape[0] // ((H_ + 2) * (W_ + 2))
        img = Ar.view(B_, H_ + 2, W_ + 2, C)
        assert img[:, 0].abs().max() == 0 and img[:, -1].abs().max() == 0 and img[:, :, 0].abs().max() == 0 and img[:, :, -1].abs().max() == 0
        v = F.conv2d(img.permute(0, 3, 1, 2), Wr.view(N, 3, 3, C).permute(0, 3, 1, 2))        # valid conv over the padded image
        v = v.permute(0, 2, 3, 1).reshape(B_ * H_ * W_, N)
    else:
        assert A.shape[1] == W.shape[1]
        v = Ar @ Wr.t()
    if bias is not None:
        assert bias.numel() == W.shape[0]
        v = v + bias
    if pre_bf16 is not None:
        if pre_is_grad:          # TVS_GEMM_PRE_DGELU: the derivative of the activation is saved instead of the pre-activation
            assert act == abi.ACT_QGELU
            pre_bf16.copy_(_qg_grad(v))
        else:
            pre_bf16.copy_(v)
    if act == abi.ACT_QGELU:
        v = _qg(v)
    elif act == abi.ACT_RELU:
        v = torch.relu(v)
    elif act == abi.ACT_DQGELU:
        v = v * _qg_grad(aux_bf16.float())
    elif act == abi.ACT_DRELU:
        v = v * (aux_bf16.float() > 0)
    elif act == abi.ACT_MULAUX:
        v = v * aux_bf16.float()
    if residual is not None:
        assert residual.shape == v.shape
        v = v + residual
    if act == abi.ACT_RES_RELU:
        v = torch.relu(v)
    if overwrite is not None:        # fused deep-prompt overwrite: only the 16-bit bias + f32-residual -> f32 GEMM implements it
        ctx, S_, row0, n_ = overwrite
        assert A.dtype in H16 and residual is not None and out_f32 is not None and out_bf16 is None and pre_bf16 is None and act == abi.ACT_NONE
        assert v.shape[0] % S_ == 0 and ctx.shape[-2:] == (n_, v.shape[1]) and v.shape[1] % 32 == 0
        v = v.clone().view(-1, S_, v.shape[1])
        v[:, row0:row0 + n_] = ctx
        v = v.view(-1, v.shape[-1])
    for o in (out_f32, out_bf16):
        if o is not None:
            assert o.shape == v.shape and o.stride(1) == 1
            o.copy_(_tf32_rn(v) if (round_out and o is out_f32) else v)


def layernorm_fwd(x, gamma, beta, eps, *, y_f32=None, y_bf16=None, mean=None, rstd=None, round_tf32=False):
    D = x.shape[-1]
    x2 = x.reshape(-1, D)
    mu = x2.mean(1)
    var = x2.var(1, unbiased=False)
    rs = torch.rsqrt(var + eps)
    y = (x2 - mu[:, None]) * rs[:, None] * gamma + beta
    for o in (y_f32, y_bf16):
        if o is not None:
            o.view(-1, D).copy_(_tf32_rn(y) if (round_tf32 and o is y_f32) else y)
    if mean is not None:
        mean.copy_(mu)
    if rstd is not None:
        rstd.copy_(rs)


def layernorm_bwd(dy, x, gamma, mean, rstd, *, dx_add=None, dx_f32=None, dx_bf16=None):
    D = x.shape[-1]
    x2, g = x.reshape(-1, D), dy.reshape(-1, D).float() * gamma
    xh = (x2 - mean[:, None]) * rstd[:, None]
    dx = rstd[:, None] * (g - g.mean(1, keepdim=True) - xh * (g * xh).mean(1, keepdim=True))
    if dx_add is not None:
        dx = dx + dx_add.reshape(-1, D)
    for o in (dx_f32, dx_bf16):
        if o is not None:
            o.view(-1, D).copy_(dx)


def _attn_ref(qkv, B, S, H, hd, causal, key_mask):
    D = H * hd
    q, k, v = (qkv[:, i * D:(i + 1) * D].reshape(B, S, H, hd).transpose(1, 2) for i in range(3))
    s = q @ k.transpose(-1, -2)
    if causal:
        s = s.masked_fill(torch.triu(torch.ones(S, S, dtype=torch.bool), 1), float("-inf"))
    if key_mask is not None:
        s = s.masked_fill(key_mask.view(B, 1, 1, S) == 0, float("-inf"))
    return (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(B * S, D), torch.logsumexp(s, -1)


def attn_fwd(qkv, B, S, H, hd, causal, key_mask, out, lse, out_f32=None):
    assert qkv.dtype == BF16 and qkv.is_contiguous() and qkv.shape == (B * S, 3 * H * hd)
    assert out.dtype == BF16 or (out.dtype == torch.float16 and hd == 64 and not causal and key_mask is None)   # TVS_ATTN_O_F16: tcgen05 path only
    o, l = _attn_ref(qkv.float(), B, S, H, hd, causal, key_mask)
    out.copy_(o)
    lse.copy_(l)
    if out_f32 is not None:
        out_f32.copy_(_tf32_rn(o))


def attn_bwd(qkv, out, dout, lse, B, S, H, hd, causal, key_mask, delta, dqkv, row_begin=0):
    q = qkv.float().requires_grad_(True)
    with torch.enable_grad():
        o, _ = _attn_ref(q, B, S, H, hd, causal, key_mask)
    (g,) = torch.autograd.grad(o, q, dout.float())
    dqkv.copy_(g)
    if row_begin:       # tvs_attn_bwd_tail: rows below the 128-row tile of row_begin are NOT produced - poison them
        first = (row_begin // 128) * 128
        dqkv.view(B, S, -1)[:, :first] = float("nan")


def prompt_overwrite(x, row0, n, ctx, x_bf16=None):
    x[:, row0:row0 + n] = ctx
    if x_bf16 is not None:
        x_bf16[:, row0:row0 + n] = ctx


def prompt_grad(dx, row0, n, dctx, zero_rows=True, dx_bf16=None):
    g = dx[:, row0:row0 + n]
    dctx += g.sum(0) if dctx.dim() == 2 else g
    if zero_rows:
        dx[:, row0:row0 + n] = 0
        if dx_bf16 is not None:
            dx_bf16[:, row0:row0 + n] = 0


def wgrad_small(dy, x, dw):
    dw += dy.t() @ x


def cast_bf16(x, y):
    y.copy_(x)


def pad_nhwc(x, B, H, W, C, xp, round_tf32=False):
    assert x.is_contiguous() and xp.is_contiguous() and xp.shape == (B * (H + 2) * (W + 2), C)
    v = _tf32_rn(x) if round_tf32 else x
    xp.copy_(F.pad(v.view(B, H, W, C), (0, 0, 1, 1, 1, 1)).reshape(-1, C))


def im2col_nhwc(x, B, H, W, C, ksize, stride, pad, col, round_tf32=False):
    assert x.is_contiguous() and x.shape == (B * H * W, C) and col.dtype == x.dtype and col.is_contiguous()
    cols = F.unfold(x.float().view(B, H, W, C).permute(0, 3, 1, 2), ksize, padding=pad, stride=stride)      # (B, C*k*k, L), (c,ky,kx)
    L = cols.shape[-1]
    cols = cols.view(B, C, ksize, ksize, L).permute(0, 4, 2, 3, 1).reshape(B * L, ksize * ksize * C)
    assert col.shape[0] == B * L and col.shape[1] >= cols.shape[1]
    col.zero_()
    col[:, : cols.shape[1]] = _tf32_rn(cols) if round_tf32 else cols


def col2im_nhwc(dcol, B, H, W, Ccol, Cx, ksize, dx, relu_mask=None):
    K = ksize * ksize * Ccol
    cols = dcol[:, :K].reshape(B, H * W, ksize, ksize, Ccol).permute(0, 4, 2, 3, 1).reshape(B, Ccol * ksize * ksize, H * W)
    g = F.fold(cols, (H, W), ksize, padding=ksize // 2)                                                      # (B, Ccol, H, W)
    g = g.permute(0, 2, 3, 1).reshape(B * H * W, Ccol)[:, :Cx]
    if relu_mask is not None:
        g = g * (relu_mask[:, :Cx] > 0)
    assert dx.shape == (B * H * W, Cx)
    dx.copy_(g)


def relu_mask(dy, y, out):
    out.copy_(dy * (y > 0))


def avgpool2_nhwc(x, B, H, W, C, y, round_tf32=False):
    p = F.avg_pool2d(x.float().view(B, H, W, C).permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1).reshape(-1, C)
    y.copy_(_tf32_rn(p) if round_tf32 else p)


def upsample2x_fwd(x, B, H, W, C, y):
    u = F.interpolate(x.view(B, H, W, C).permute(0, 3, 1, 2), scale_factor=2, mode="bilinear")
    y.copy_(u.permute(0, 2, 3, 1).reshape(-1, C))


def upsample2x_bwd(dy, B, H, W, C, dx):
    x = torch.zeros(B, C, H, W, requires_grad=True)
    with torch.enable_grad():
        u = F.interpolate(x, scale_factor=2, mode="bilinear")
    (g,) = torch.autograd.grad(u, x, dy.reshape(B, 2 * H, 2 * W, C).permute(0, 3, 1, 2))
    dx.copy_(g.permute(0, 2, 3, 1).reshape(-1, C))


def _xattn_ref(q, k, v, key_mask, B, Sq, Sk, H, hd, causal=False):
    s = q.reshape(B, Sq, H, hd).transpose(1, 2) @ k.reshape(B, Sk, H, hd).transpose(1, 2).transpose(-1, -2)
    if causal:
        s = s.masked_fill(torch.triu(torch.ones(Sq, Sk, dtype=torch.bool), 1), float("-inf"))
    if key_mask is not None:
        s = s.masked_fill(key_mask.view(B, 1, 1, Sk) == 0, float("-inf"))
    o = (torch.softmax(s, -1) @ v.reshape(B, Sk, H, hd).transpose(1, 2)).transpose(1, 2).reshape(B * Sq, H * hd)
    return o, torch.logsumexp(s, -1)


def cross_attn_fwd(q, k, v, key_mask, B, Sq, Sk, H, hd, out, lse, causal=False):
    assert hd == 64 and Sk <= 80
    o, l = _xattn_ref(q, k, v, key_mask, B, Sq, Sk, H, hd, causal)
    out.copy_(_tf32_rn(o))
    lse.copy_(l)


def cross_attn_bwd(q, k, v, key_mask, out, dout, lse, B, Sq, Sk, H, hd, dq, dk, dv, delta, causal=False):
    qq, kk, vv = (t.detach().clone().requires_grad_(True) for t in (q, k, v))
    with torch.enable_grad():
        o, _ = _xattn_ref(qq, kk, vv, key_mask, B, Sq, Sk, H, hd, causal)
    gq, gk, gv = torch.autograd.grad(o, (qq, kk, vv), dout)
    dq.copy_(gq); dk.copy_(gk); dv.copy_(gv)


def _dyn_ref(x, w, B, H, W, C):
    weight, bias = w[:, : C * 9].reshape(B, C, 3, 3), w[:, C * 9]
    xi = x.view(B, H, W, C).permute(0, 3, 1, 2).reshape(1, B * C, H, W)
    return F.conv2d(xi, weight, bias, padding=1, groups=B).transpose(0, 1)


def dynconv_fwd(x, w, bias, B, H, W, C, taps, out):
    assert bias.data_ptr() == w[:, C * 9:].data_ptr()
    out.copy_(_dyn_ref(x, w, B, H, W, C))


def dynconv_bwd(dout, x, w, B, H, W, C, dx, dw_part):
    xx, ww = x.detach().clone().requires_grad_(True), w.detach().clone().requires_grad_(True)
    with torch.enable_grad():
        o = _dyn_ref(xx, ww, B, H, W, C)
    gx, gw = torch.autograd.grad(o, (xx, ww), dout)
    dx.copy_(gx)
    dw_part.zero_()
    dw_part[0] = gw[:, : C * 9]


def _dense(idx, wt, n_in):
    R = torch.zeros(idx.shape[0], n_in)
    for a in range(idx.shape[1]):
        R.scatter_add_(1, idx[:, a:a + 1].long(), wt[:, a:a + 1])
    return R


def _untile(t, B, Ho, Wo, tile):
    if tile <= 0:
        return t.reshape(B, Ho, Wo)
    return t.reshape(B, Ho // tile, Wo // tile, tile, tile).permute(0, 1, 3, 2, 4).reshape(B, Ho, Wo)


def _tile(t, B, Ho, Wo, tile):
    if tile <= 0:
        return t.reshape(B, Ho, Wo)
    return t.reshape(B, Ho // tile, tile, Wo // tile, tile).permute(0, 1, 3, 2, 4).reshape(B * (Ho // tile) * (Wo // tile), tile * tile)


def resample2d_fwd(inp, B, Hi, Wi, Ho, Wo, tab, tile, out):
    Ry, Rx = _dense(tab["iy"], tab["wy"], Hi), _dense(tab["ix"], tab["wx"], Wi)
    big = Ry @ inp.reshape(B, Hi, Wi) @ Rx.t()
    out.copy_(_tile(big, B, Ho, Wo, tile).reshape(out.shape))


def resample2d_bwd(dout, B, Hi, Wi, Ho, Wo, tab, tile, din):
    Ry, Rx = _dense(tab["iy"], tab["wy"], Hi), _dense(tab["ix"], tab["wx"], Wi)
    g = _untile(dout.float(), B, Ho, Wo, tile)
    din.copy_((Ry.t() @ g @ Rx).reshape(din.shape))


def _head_ref(tconv, addmap, bias_t, bias_a, ratio, blend, B, G, P, ks):
    img = G * P
    base = _untile(tconv, B, img, img, P) + bias_t
    if blend == abi.BLEND_NONE:
        return base, None
    KK = ks * ks
    up = F.interpolate(addmap.reshape(B, G, G, KK).permute(0, 3, 1, 2), scale_factor=P, mode="bilinear")       # (B, KK, img, img)
    up = F.pad(up, (ks // 2,) * 4, mode="replicate")
    add = sum(up[:, ky * ks + kx, ky:ky + img, kx:kx + img] for ky in range(ks) for kx in range(ks)) + bias_a
    if blend == abi.BLEND_RATIO:
        return (1 - ratio) * base + ratio * add, add
    return base + add, add


def head_fwd(tconv, addmap, bias_t, bias_a, ratio, blend, B, G, P, ksize, logits, add_out=None):
    lg, add = _head_ref(tconv, addmap, bias_t, bias_a, ratio, blend, B, G, P, ksize)
    logits.copy_(lg.reshape(logits.shape))
    if add_out is not None:
        add_out.copy_(add)


def head_bwd(dlogits, tconv, add_out, bias_t, ratio, blend, B, G, P, ksize, dtconv_bf16, daddmap, dbias_a, dratio):
    img = G * P
    g = dlogits.reshape(B, img, img)
    if blend == abi.BLEND_NONE:
        dtconv_bf16.copy_(_tile(g, B, img, img, P))
        return
    KK = ksize * ksize
    tc = tconv.detach().clone().requires_grad_(True)
    am = torch.zeros(B * G * G, KK, requires_grad=True)
    ba = torch.zeros(1, requires_grad=True)
    rr = ratio.detach().clone().requires_grad_(True) if ratio is not None else None
    with torch.enable_grad():
        lg, add = _head_ref(tc, am, bias_t, ba, rr, blend, B, G, P, ksize)
        # the additive map is linear in addmap, so the gradient does not depend on its value; dratio does:
        if blend == abi.BLEND_RATIO:
            lg = (1 - rr) * (_untile(tc, B, img, img, P) + bias_t) + rr * (add - add.detach() + add_out)
    outs = torch.autograd.grad(lg, [tc, am, ba] + ([rr] if rr is not None else []), g)
    dtconv_bf16.copy_(outs[0])
    if daddmap is not None:          # None: the caller computes the tap-map gradient itself (two GEMMs, engine.DecoderFn.backward)
        daddmap.copy_(outs[1])
    dbias_a += outs[2]
    if rr is not None and dratio is not None:
        dratio += outs[3]


# ---- CLIPSeg-only entries -------------------------------------------------------------------------------------------------
def im2col_patches(image, P, out_bf16):
    B, C, H, W = image.shape
    g = H // P
    cols = image.view(B, C, g, P, g, P).permute(0, 2, 4, 1, 3, 5).reshape(B * g * g, C * P * P)      # column order (c, py, px)
    out_bf16.copy_(cols)


def vision_assemble(patches, cls, pos, ctx, B, G2, n, D, h):
    hv = h.view(B, 1 + G2 + n, D)
    hv[:, 0] = cls.view(-1) + pos[0]
    hv[:, 1:1 + G2] = patches.view(B, G2, D) + pos[1:1 + G2]
    if n:
        hv[:, 1 + G2:] = ctx                      # (n, D) shared or (B, n, D) per sample; no position embedding on prompts


def slice_rows(x, row0, nrows, y_f32=None, y_bf16=None):
    if y_f32 is not None:
        y_f32.copy_(x[:, row0:row0 + nrows])
    if y_bf16 is not None:
        y_bf16.copy_(x[:, row0:row0 + nrows])


def unslice_rows(dy, S, row0, dx):
    dx.zero_()
    dx.view(dy.shape[0], S, -1)[:, row0:row0 + dy.shape[1]] = dy


def add_f32(y, x):
    y += x


def film_fwd(x, mul, add, y=None, y_bf16=None):
    o = mul.unsqueeze(1) * x + add.unsqueeze(1)
    if y is not None:
        y.copy_(o)
    if y_bf16 is not None:
        y_bf16.copy_(o)


def film_bwd(dy, x, mul, dx, dmul, dadd):
    dx.copy_(mul.unsqueeze(1) * dy)
    dmul.copy_((dy * x).sum(1))
    dadd.copy_(dy.sum(1))


def _ffn_w(pair):
    hi, lo = pair
    return hi.float() if lo is None else hi.float() + lo.float()


def ffn64_fwd(x, w1, w2t, b1, b2, out):
    assert x.shape[1] == 64 and w1[0].shape == w2t[0].shape and w1[0].dtype == BF16
    out.copy_(x + torch.relu(x @ _ffn_w(w1).t() + b1) @ _ffn_w(w2t) + b2)


def ffn64_bwd(x, g, w1, w2t, b1, dx):
    mask = (x @ _ffn_w(w1).t() + b1) > 0
    dx.copy_(g + ((g @ _ffn_w(w2t).t()) * mask) @ _ffn_w(w1))


# ---- loss + metric counters (loss_metrics.cu) -------------------------------------------------------------------------------
def dicebce_scratch_bytes(B, N):
    return 16


def _counts(p, y, thr):
    t = y.long() & 1
    ge, gt = (p >= thr).long(), (p > thr).long()
    counts = torch.stack(((ge * t).sum(1), (ge * (1 - t)).sum(1), ((1 - ge) * t).sum(1)), dim=1)
    conf = torch.stack((((1 - t) * (1 - gt)).sum(), ((1 - t) * gt).sum(), (t * (1 - gt)).sum(), (t * gt).sum()))
    return counts, conf


def dicebce_metrics_fwd(logits, mask, threshold, lambda_dice, lambda_ce, parts, counts, confmat, loss, scratch):
    B = logits.shape[0]
    x, y = logits.reshape(B, -1).float(), mask.reshape(B, -1).float()
    p = 1.0 / (1.0 + torch.exp(-x))            # fl32(1 / fl32(1 + fl32(exp(-x)))): the definition the counters are bit-exact against
    bce = torch.clamp(x, min=0) - x * y + torch.log1p(torch.exp(-x.abs()))
    pr = torch.stack(((p * y).sum(1), p.sum(1), y.sum(1), bce.sum(1)), dim=1).double()
    if parts is not None:
        parts.copy_(pr)
    c, cf = _counts(p, y, threshold)
    if counts is not None:
        counts.copy_(c)
    if confmat is not None:
        confmat += cf
    if loss is not None:
        dice = 1.0 - (2.0 * pr[:, 0] + 1e-5) / (pr[:, 1] + pr[:, 2] + 1e-5)
        loss.fill_(float(lambda_dice * dice.mean() + lambda_ce * pr[:, 3].sum() / (B * x.shape[1])))


def dicebce_bwd(logits, mask, parts, gscale, lambda_dice, lambda_ce, dlogits):
    B = logits.shape[0]
    x, y = logits.reshape(B, -1).float(), mask.reshape(B, -1).float()
    N = x.shape[1]
    p = torch.sigmoid(x)
    I, P, G = (parts[:, k].unsqueeze(1) for k in range(3))
    den = P + G + 1e-5
    ddice = (-(2.0 * y * den - (2.0 * I + 1e-5)) / den ** 2).float()
    gs = 1.0 if gscale is None else float(gscale)
    dlogits.copy_((gs * (lambda_dice / B * ddice * p * (1 - p) + lambda_ce / (B * N) * (p - y))).view_as(dlogits))


def metrics_from_probs(preds, mask, threshold, counts, confmat, scratch):
    B = preds.shape[0]
    c, cf = _counts(preds.reshape(B, -1).float(), mask.reshape(B, -1).float(), threshold)
    if counts is not None:
        counts.copy_(c)
    if confmat is not None:
        confmat += cf


def install(monkeypatch):
    for name in ("gemm", "layernorm_fwd", "layernorm_bwd", "attn_fwd", "attn_bwd", "prompt_overwrite", "prompt_grad", "wgrad_small",
                 "cast_bf16", "round_tf32", "pad_nhwc", "im2col_nhwc", "col2im_nhwc", "relu_mask", "avgpool2_nhwc", "upsample2x_fwd", "upsample2x_bwd",
                 "cross_attn_fwd", "cross_attn_bwd", "dynconv_fwd", "dynconv_bwd", "resample2d_fwd", "resample2d_bwd", "head_fwd",
                 "head_bwd", "im2col_patches", "vision_assemble", "slice_rows", "unslice_rows", "add_f32", "film_fwd", "film_bwd", "ffn64_fwd",
                 "ffn64_bwd", "dicebce_scratch_bytes", "dicebce_metrics_fwd", "dicebce_bwd", "metrics_from_probs"):
        monkeypatch.setattr(abi, name, globals()[name])
    monkeypatch.setattr(abi, "require_device", lambda: None)
    monkeypatch.setattr(abi, "check_cuda_input", lambda t: None)
