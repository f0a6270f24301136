"""ctypes binding of libtvs_b200.so (the C ABI in include/tvs_b200.h) for torch tensors.

PyTorch is plumbing here: it owns device memory and the stream; every compute call below hands raw device
pointers and ``torch.cuda.current_stream().cuda_stream`` to the hand-written sm_100a kernels.  There is NO
fallback: if the shared library is missing, cannot be loaded, or the device is not sm_100, a ``TvsError`` is
raised - the product never routes through torch ops or the CPU oracle for the hot path.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, byref, c_char_p, c_float, c_int32, c_int64, c_void_p

import torch

_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libtvs_b200.so")

ACT_NONE, ACT_QGELU, ACT_RELU, ACT_DQGELU, ACT_DRELU, ACT_RES_RELU, ACT_MULAUX = range(7)
AB_BF16, AB_TF32, AB_F16 = range(3)
H16 = (torch.bfloat16, torch.float16)       # 16-bit activation formats of the kind::f16 MMA (per operand)
BLEND_NONE, BLEND_RATIO, BLEND_ADD = range(3)


class TvsError(RuntimeError):
    pass


class GemmArgs(Structure):
    _fields_ = [
        ("A", c_void_p), ("lda", c_int64),
        ("W", c_void_p), ("ldw", c_int64),
        ("M", c_int32), ("N", c_int32), ("K", c_int32),
        ("bias", c_void_p),
        ("residual", c_void_p), ("ldr", c_int64),
        ("out_f32", c_void_p), ("ldo32", c_int64),
        ("out_bf16", c_void_p), ("ldo16", c_int64),
        ("pre_bf16", c_void_p), ("ldpre", c_int64),
        ("aux_bf16", c_void_p), ("ldaux", c_int64),
        ("act", c_int32), ("tile_n", c_int32), ("ab_dtype", c_int32), ("reserved", c_int32),
        ("conv_h", c_int32), ("conv_w", c_int32),
        ("ovr_ctx", c_void_p), ("ovr_batch_stride", c_int64),
        ("ovr_S", c_int32), ("ovr_row0", c_int32), ("ovr_n", c_int32), ("ovr_reserved", c_int32),
    ]


_lib = None


def lib_path() -> str:
    return _LIB_PATH


def load():
    """Load the shared library (once).  Raises TvsError when it is absent - build it with
    ``python -m tunevlseg_b200.build`` (``__graft_entry__.build()`` does)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise TvsError(f"{_LIB_PATH} not found: run `python -m tunevlseg_b200.build` (no CPU / torch fallback exists)")
    try:
        lib = ctypes.CDLL(_LIB_PATH)
    except OSError as e:  # pragma: no cover
        raise TvsError(f"cannot load {_LIB_PATH}: {e}") from e
    lib.tvs_last_error.restype = c_char_p
    lib.tvs_launch_count.restype = c_int64
    lib.tvs_gemm_last_variant.restype = c_int32
    lib.tvs_dicebce_scratch_bytes.restype = c_int64
    lib.tvs_dicebce_scratch_bytes.argtypes = [c_int32, c_int64]
    lib.tvs_gemm_bf16.argtypes = [POINTER(GemmArgs), c_void_p]
    P, I32, I64, F = c_void_p, c_int32, c_int64, c_float
    sig = {
        "tvs_layernorm_fwd": [P, P, P, F, I64, I32, P, P, P, P, I32, P],
        "tvs_layernorm_bwd": [P, P, P, P, P, P, P, I64, I32, P, P, P],
        "tvs_attn_fwd": [P, I32, I32, I32, I32, I32, P, P, P, P, I32, P],
        "tvs_attn_bwd": [P, P, P, P, I32, I32, I32, I32, I32, P, P, P, I32, P],
        "tvs_attn_bwd_tail": [P, P, P, P, I32, I32, I32, I32, I32, P, P, P, I32, I32, P],
        "tvs_im2col_patches": [P, I32, I32, I32, I32, I32, P, I32, P],
        "tvs_vision_assemble": [P, P, P, P, I64, I32, I32, I32, I32, P, P],
        "tvs_prompt_overwrite": [P, P, I32, I32, I32, I32, I32, P, I64, P],
        "tvs_prompt_grad": [P, P, I32, I32, I32, I32, I32, P, I64, I32, P],
        "tvs_slice_rows": [P, I32, I32, I32, I32, I32, P, P, P],
        "tvs_unslice_rows": [P, I32, I32, I32, I32, I32, P, P],
        "tvs_wgrad_small": [P, I64, P, I64, I64, I32, I32, P, P],
        "tvs_cast_bf16": [P, P, I64, P],
        "tvs_add_f32": [P, P, I64, P],
        "tvs_film_fwd": [P, P, P, I32, I32, I32, P, P, P],
        "tvs_film_bwd": [P, P, P, I32, I32, I32, P, P, P, P],
        "tvs_head_fwd": [P, I64, P, I64, P, P, P, I32, I32, I32, I32, I32, P, P, P],
        "tvs_head_bwd": [P, P, I64, P, P, P, I32, I32, I32, I32, I32, P, I64, P, I64, P, P, P],
        "tvs_dicebce_metrics_fwd": [P, P, I32, I64, F, F, F, P, P, P, P, P, P],
        "tvs_dicebce_bwd": [P, P, P, P, I32, I64, F, F, P, P],
        "tvs_metrics_from_probs": [P, P, I32, I64, F, P, P, P, P],
        "tvs_adamw_flat": [P, P, P, P, I64, F, F, F, F, F, I32, F, P, P, P],
        "tvs_counter_inc": [P, P],
        "tvs_im2col_nhwc": [P, I32, I32, I32, I32, I32, I32, I32, I32, P, I64, I32, P],
        "tvs_round_tf32": [P, I64, I64, I32, P, I64, P],
        "tvs_pad_nhwc": [P, I32, I32, I32, I32, I32, P, I32, P],
        "tvs_col2im_nhwc": [P, I64, I32, I32, I32, I32, I32, I32, P, I64, P, I64, P],
        "tvs_relu_mask": [P, I64, P, I64, I64, I32, P, I64, P],
        "tvs_avgpool2_nhwc": [P, I32, I32, I32, I32, I32, P, I64, P],
        "tvs_upsample2x_fwd": [P, I32, I32, I32, I32, P, I64, P],
        "tvs_upsample2x_bwd": [P, I64, I32, I32, I32, I32, P, P],
        "tvs_cross_attn_fwd": [P, I64, P, P, I64, P, I32, I32, I32, I32, I32, I32, P, I64, P, P],
        "tvs_cross_attn_bwd": [P, I64, P, P, I64, P, P, P, I64, P, I32, I32, I32, I32, I32, I32, P, I64, P, P, I64, P, P],
        "tvs_dynconv_fwd": [P, P, I64, P, I64, I32, I32, I32, I32, P, P, P],
        "tvs_dynconv_bwd": [P, P, P, I64, I32, I32, I32, I32, P, P, I32, P],
        "tvs_ffn64_fwd": [P, P, P, P, P, P, P, I64, I32, I32, P, P],
        "tvs_ffn64_bwd": [P, P, P, P, P, P, P, I64, I32, I32, P, P],
        "tvs_preproc_image_u8": [P, I32, I32, I64, P, P, P, P, P, P, I32, I32, P, P, P],
        "tvs_resize_nearest_f32": [P, I32, I32, I64, P, P, I32, I32, P, P],
        "tvs_warp_affine_u8": [P, I32, I32, I64, P, P, P, P, P, P, P, P, I32, I32, P, P, P],
        "tvs_warp_affine_nearest_f32": [P, I32, I32, I64, P, P, P, P, I32, I32, P, P],
        "tvs_lut_normalize_u8": [P, I32, I32, I64, P, P, P, P, P],
        "tvs_resample2d_fwd": [P, I32, I32, I32, I32, I32, P, P, P, P, I32, I32, P, P],
        "tvs_resample2d_u8": [P, I32, I32, I32, I32, I32, P, P, P, P, I32, P, P],
        "tvs_resample2d_bwd": [P, I32, I32, I32, I32, I32, I32, P, P, P, P, P, P, I32, I32, P, P],
    }
    for name, argtypes in sig.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = c_int32
    _lib = lib
    return lib


_device_ok = False


def require_device() -> None:
    """Fail loudly unless the current CUDA device is an sm_100 part the library can drive."""
    global _device_ok
    if _device_ok:
        return
    lib = load()
    if not torch.cuda.is_available():
        raise TvsError("tunevlseg_b200 needs a CUDA (sm_100a) device; there is no CPU fallback")
    torch.cuda.current_device()  # make sure a context exists
    if lib.tvs_device_check() != 0:
        raise TvsError(lib.tvs_last_error().decode())
    _device_ok = True


def check_cuda_input(t: torch.Tensor) -> None:
    if not t.is_cuda:
        raise TvsError("tunevlseg_b200 runs on a CUDA (sm_100a) device only; inputs must be CUDA tensors")


def launch_count() -> int:
    return int(load().tvs_launch_count())


def _ck(rc: int, what: str) -> None:
    if rc != 0:
        raise TvsError(f"{what}: {load().tvs_last_error().decode()} (rc={rc})")


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t: torch.Tensor | None) -> int | None:
    return None if t is None else t.data_ptr()


def _chk(t: torch.Tensor | None, dtype, name: str, dim2: bool = False) -> None:
    if t is None:
        return
    if not t.is_cuda:
        raise TvsError(f"{name} must be a CUDA tensor")
    if t.dtype != dtype and not (isinstance(dtype, tuple) and t.dtype in dtype):
        raise TvsError(f"{name} must be {dtype}, got {t.dtype}")
    if dim2:
        if t.dim() != 2 or t.stride(1) != 1:
            raise TvsError(f"{name} must be 2-D with unit inner stride, got shape {tuple(t.shape)} strides {t.stride()}")
    elif not t.is_contiguous():
        raise TvsError(f"{name} must be contiguous")


# ------------------------------------------------------------------------------------------------------------------
def gemm(A, W, *, bias=None, residual=None, out_f32=None, out_bf16=None, pre_bf16=None, aux_bf16=None, act=ACT_NONE,
         tile_n=0, round_out=False, conv_hw=None, overwrite=None, stream_k=False, pre_is_grad=False):
    """C[M,N] = epilogue(A[M,K] @ W[N,K]^T); see tvs_gemm_bf16 in include/tvs_b200.h.  2-D views with a row stride
    are accepted (ld = stride(0)).  ``round_out``: out_f32 is rounded to nearest tf32 (it only feeds further tf32 GEMMs).
    ``conv_hw=(H, W)``: implicit-GEMM 3x3 convolution - A is the zero-bordered image [B*(H+2)*(W+2), C] from ``pad_nhwc``,
    W is [N, 9*C]; outputs / residual are unpadded [B*H*W, N].
    ``overwrite=(ctx, S, row0, n)``: deep-prompt overwrite fused into the epilogue - output rows at positions row0 .. row0+n-1 of
    every S-row sample receive ctx ((n, N) shared or (B, n, N) per sample, f32) instead of the result (fc2 of a vision block).
    ``stream_k``: opt-in stream-K schedule of the specialised pair kernels (TVS_GEMM_STREAM_K; measured slower - off by default).
    ``pre_is_grad`` (with ACT_QGELU and pre_bf16): pre_bf16 receives QuickGELU'(pre-activation) instead of the pre-activation
    (TVS_GEMM_PRE_DGELU); the dgrad through the activation is then ``act=ACT_MULAUX`` with that tensor as ``aux_bf16``."""
    require_device()
    ab = {(torch.bfloat16, torch.bfloat16): AB_BF16, (torch.float32, torch.float32): AB_TF32, (torch.float16, torch.float16): AB_F16}.get((A.dtype, W.dtype))
    if ab is None:      # a mixed fp16 x bf16 kind::f16 MMA is an illegal instruction on sm_100 (measured)
        raise TvsError(f"gemm: A and W must both be f32 (tf32 MMA), both bf16 or both fp16 (kind::f16 MMA), got {A.dtype} / {W.dtype}")
    _chk(A, A.dtype, "A", True); _chk(W, W.dtype, "W", True)
    _chk(bias, torch.float32, "bias"); _chk(residual, torch.float32, "residual", True)
    _chk(out_f32, torch.float32, "out_f32", True); _chk(out_bf16, H16, "out_bf16", True)
    _chk(pre_bf16, torch.bfloat16, "pre_bf16", True); _chk(aux_bf16, torch.bfloat16, "aux_bf16", True)
    M, K = A.shape
    N, K2 = W.shape
    Mo = M
    if conv_hw is not None:
        H_, W_ = conv_hw
        if K2 != 9 * K or M % ((H_ + 2) * (W_ + 2)) or not A.is_contiguous():
            raise TvsError(f"gemm(conv): A {tuple(A.shape)} must be the contiguous padded image and W {tuple(W.shape)} = [N, 9*C]")
        Mo, K = M // ((H_ + 2) * (W_ + 2)) * H_ * W_, K2
    elif K != K2:
        raise TvsError(f"gemm: A is {tuple(A.shape)} but W is {tuple(W.shape)}")
    for name, t in (("residual", residual), ("out_f32", out_f32), ("out_bf16", out_bf16), ("pre_bf16", pre_bf16),
                    ("aux_bf16", aux_bf16)):
        if t is not None and tuple(t.shape) != (Mo, N):
            raise TvsError(f"gemm: {name} must be {(Mo, N)}, got {tuple(t.shape)}")
    if bias is not None and bias.numel() != N:
        raise TvsError("gemm: bias length")
    g = GemmArgs()
    g.A, g.lda, g.W, g.ldw = A.data_ptr(), A.stride(0), W.data_ptr(), W.stride(0)
    g.M, g.N, g.K = M, N, K
    g.bias = _p(bias)
    g.residual, g.ldr = _p(residual), (residual.stride(0) if residual is not None else 0)
    g.out_f32, g.ldo32 = _p(out_f32), (out_f32.stride(0) if out_f32 is not None else 0)
    g.out_bf16, g.ldo16 = _p(out_bf16), (out_bf16.stride(0) if out_bf16 is not None else 0)
    g.pre_bf16, g.ldpre = _p(pre_bf16), (pre_bf16.stride(0) if pre_bf16 is not None else 0)
    g.aux_bf16, g.ldaux = _p(aux_bf16), (aux_bf16.stride(0) if aux_bf16 is not None else 0)
    g.act, g.tile_n = act, tile_n
    g.ab_dtype = ab
    g.reserved = (1 if round_out else 0) | (2 if (out_bf16 is not None and out_bf16.dtype == torch.float16) else 0) | (4 if stream_k else 0) | (8 if pre_is_grad else 0)
    g.conv_h, g.conv_w = conv_hw if conv_hw is not None else (0, 0)
    if overwrite is not None:
        ctx, S_, row0, n_ = overwrite
        _chk(ctx, torch.float32, "overwrite ctx")
        if ctx.shape[-1] != N or ctx.shape[-2] != n_:
            raise TvsError(f"gemm: overwrite ctx must be (..., {n_}, {N}), got {tuple(ctx.shape)}")
        g.ovr_ctx, g.ovr_batch_stride = ctx.data_ptr(), (0 if ctx.dim() == 2 else n_ * N)
        g.ovr_S, g.ovr_row0, g.ovr_n = S_, row0, n_
    _ck(load().tvs_gemm_bf16(byref(g), _stream()), "tvs_gemm_bf16")


def gemm_last_variant() -> str:
    """Template instance the calling thread's last ``gemm`` launched, e.g. ``"256x6 bf16 cta_group::2"``."""
    v = load().tvs_gemm_last_variant()
    epi = ("generic", "out_bf16", "res_f32", "fc1", "dqgelu")[(v >> 5) & 7]
    sk = " stream-k" if (v >> 28) & 1 else ""          # the (tile, k-block) space cut into one contiguous range per CTA pair
    return f"{(v >> 16) & 0xFFF}x{(v >> 8) & 0xFF} {'tf32' if (v >> 4) & 1 else 'bf16'} epi:{epi} cta_group::{v & 0xF}{sk}"


def layernorm_fwd(x, gamma, beta, eps, *, y_f32=None, y_bf16=None, mean=None, rstd=None, round_tf32=False):
    require_device()
    _chk(x, torch.float32, "x"); _chk(gamma, torch.float32, "gamma"); _chk(beta, torch.float32, "beta")
    _chk(y_f32, torch.float32, "y_f32"); _chk(y_bf16, H16, "y_bf16")
    D = x.shape[-1]
    M = x.numel() // D
    flags = int(bool(round_tf32)) | (2 if (y_bf16 is not None and y_bf16.dtype == torch.float16) else 0)
    _ck(load().tvs_layernorm_fwd(x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), eps, M, D, _p(y_f32), _p(y_bf16),
                                 _p(mean), _p(rstd), flags, _stream()), "tvs_layernorm_fwd")


def layernorm_bwd(dy, x, gamma, mean, rstd, *, dx_add=None, dx_f32=None, dx_bf16=None):
    require_device()
    _chk(x, torch.float32, "x"); _chk(dx_add, torch.float32, "dx_add"); _chk(dx_f32, torch.float32, "dx_f32")
    _chk(dx_bf16, torch.bfloat16, "dx_bf16")
    if not dy.is_contiguous():
        raise TvsError("dy must be contiguous")
    D = x.shape[-1]
    M = x.numel() // D
    d16 = dy.data_ptr() if dy.dtype == torch.bfloat16 else None
    d32 = dy.data_ptr() if dy.dtype == torch.float32 else None
    if d16 is None and d32 is None:
        raise TvsError("dy must be bf16 or f32")
    _ck(load().tvs_layernorm_bwd(d16, d32, x.data_ptr(), gamma.data_ptr(), mean.data_ptr(), rstd.data_ptr(), _p(dx_add),
                                 M, D, _p(dx_f32), _p(dx_bf16), _stream()), "tvs_layernorm_bwd")


def ffn64_fwd(x, w1, w2t, b1, b2, out):
    """out = x + relu(x W1^T + b1) W2^T + b2 in one kernel (decoder FFN, D = 64).  ``w1`` / ``w2t``: (hi, lo) pairs of
    bf16 [F, 64] tensors (lo may be None for single bf16 products)."""
    require_device()
    _chk(x, torch.float32, "x"); _chk(out, torch.float32, "out"); _chk(b1, torch.float32, "b1"); _chk(b2, torch.float32, "b2")
    for t in (*w1, *w2t):
        _chk(t, torch.bfloat16, "ffn weight")
    M, D = x.shape
    F = w1[0].shape[0]
    if tuple(w1[0].shape) != (F, D) or tuple(w2t[0].shape) != (F, D) or b1.numel() != F or b2.numel() != D or out.shape != x.shape:
        raise TvsError("ffn64_fwd: shape mismatch")
    _ck(load().tvs_ffn64_fwd(x.data_ptr(), _p(w1[0]), _p(w1[1]), _p(w2t[0]), _p(w2t[1]), b1.data_ptr(), b2.data_ptr(), M, D, F,
                             out.data_ptr(), _stream()), "tvs_ffn64_fwd")


def ffn64_bwd(x, g, w1, w2t, b1, dx):
    """dx = g + ((g W2) o [x W1^T + b1 > 0]) W1 (dgrad of ffn64_fwd including its residual)."""
    require_device()
    _chk(x, torch.float32, "x"); _chk(g, torch.float32, "g"); _chk(dx, torch.float32, "dx"); _chk(b1, torch.float32, "b1")
    for t in (*w1, *w2t):
        _chk(t, torch.bfloat16, "ffn weight")
    M, D = x.shape
    F = w1[0].shape[0]
    if tuple(w1[0].shape) != (F, D) or tuple(w2t[0].shape) != (F, D) or b1.numel() != F or g.shape != x.shape or dx.shape != x.shape:
        raise TvsError("ffn64_bwd: shape mismatch")
    _ck(load().tvs_ffn64_bwd(x.data_ptr(), g.data_ptr(), _p(w1[0]), _p(w1[1]), _p(w2t[0]), _p(w2t[1]), b1.data_ptr(), M, D, F,
                             dx.data_ptr(), _stream()), "tvs_ffn64_bwd")


def attn_fwd(qkv, B, S, H, hd, causal, key_mask, out, lse, out_f32=None):
    require_device()
    _chk(qkv, torch.bfloat16, "qkv"); _chk(out, H16, "out"); _chk(lse, torch.float32, "lse")
    _chk(key_mask, torch.uint8, "key_mask"); _chk(out_f32, torch.float32, "out_f32")
    _ck(load().tvs_attn_fwd(qkv.data_ptr(), B, S, H, hd, int(causal), _p(key_mask), out.data_ptr(), _p(out_f32),
                            lse.data_ptr(), int(out.dtype == torch.float16), _stream()), "tvs_attn_fwd")


def attn_bwd(qkv, out, dout, lse, B, S, H, hd, causal, key_mask, delta, dqkv, row_begin=0):
    """``row_begin`` > 0 (tvs_attn_bwd_tail): only rows >= row_begin of dqkv are needed; rows below the 128-row tile of
    row_begin may be left unwritten."""
    require_device()
    _chk(qkv, torch.bfloat16, "qkv"); _chk(out, H16, "out"); _chk(dout, torch.bfloat16, "dout")
    _chk(dqkv, torch.bfloat16, "dqkv"); _chk(lse, torch.float32, "lse"); _chk(delta, torch.float32, "delta")
    flags = int(out.dtype == torch.float16)
    if row_begin:
        _ck(load().tvs_attn_bwd_tail(qkv.data_ptr(), out.data_ptr(), dout.data_ptr(), lse.data_ptr(), B, S, H, hd, int(causal),
                                     _p(key_mask), delta.data_ptr(), dqkv.data_ptr(), int(row_begin), flags, _stream()), "tvs_attn_bwd_tail")
        return
    _ck(load().tvs_attn_bwd(qkv.data_ptr(), out.data_ptr(), dout.data_ptr(), lse.data_ptr(), B, S, H, hd, int(causal),
                            _p(key_mask), delta.data_ptr(), dqkv.data_ptr(), flags, _stream()), "tvs_attn_bwd")


def im2col_patches(image, P, out_bf16):
    require_device()
    _chk(image, torch.float32, "image"); _chk(out_bf16, H16, "out")
    B, C, H, W = image.shape
    _ck(load().tvs_im2col_patches(image.data_ptr(), B, C, H, W, P, out_bf16.data_ptr(), int(out_bf16.dtype == torch.float16), _stream()),
        "tvs_im2col_patches")


def vision_assemble(patches, cls, pos, ctx, B, G2, n, D, h):
    require_device()
    _chk(patches, torch.float32, "patches"); _chk(cls, torch.float32, "cls"); _chk(pos, torch.float32, "pos")
    _chk(ctx, torch.float32, "ctx"); _chk(h, torch.float32, "h")
    bs = 0 if (ctx is None or ctx.dim() == 2) else n * D
    _ck(load().tvs_vision_assemble(patches.data_ptr(), cls.data_ptr(), pos.data_ptr(), _p(ctx), bs, B, G2, n, D,
                                   h.data_ptr(), _stream()), "tvs_vision_assemble")


def prompt_overwrite(x, row0, n, ctx, x_bf16=None):
    require_device()
    _chk(x, torch.float32, "x"); _chk(ctx, torch.float32, "ctx"); _chk(x_bf16, torch.bfloat16, "x_bf16")
    B, S, D = x.shape
    bs = 0 if ctx.dim() == 2 else n * D
    _ck(load().tvs_prompt_overwrite(x.data_ptr(), _p(x_bf16), B, S, D, row0, n, ctx.data_ptr(), bs, _stream()),
        "tvs_prompt_overwrite")


def prompt_grad(dx, row0, n, dctx, zero_rows=True, dx_bf16=None):
    """dctx (+)= sum_b dx[:, row0:row0+n]; dctx is (n, D) [batch-reduced] or (B, n, D) [per sample]."""
    require_device()
    _chk(dx, torch.float32, "dx"); _chk(dctx, torch.float32, "dctx"); _chk(dx_bf16, torch.bfloat16, "dx_bf16")
    B, S, D = dx.shape
    bs = 0 if dctx.dim() == 2 else n * D
    _ck(load().tvs_prompt_grad(dx.data_ptr(), _p(dx_bf16), B, S, D, row0, n, dctx.data_ptr(), bs, int(zero_rows),
                               _stream()), "tvs_prompt_grad")


def slice_rows(x, row0, nrows, y_f32=None, y_bf16=None):
    require_device()
    _chk(x, torch.float32, "x"); _chk(y_f32, torch.float32, "y_f32"); _chk(y_bf16, torch.bfloat16, "y_bf16")
    B, S, D = x.shape
    _ck(load().tvs_slice_rows(x.data_ptr(), B, S, D, row0, nrows, _p(y_f32), _p(y_bf16), _stream()), "tvs_slice_rows")


def unslice_rows(dy, S, row0, dx):
    require_device()
    _chk(dy, torch.float32, "dy"); _chk(dx, torch.float32, "dx")
    B, nrows, D = dy.shape
    _ck(load().tvs_unslice_rows(dy.data_ptr(), B, S, D, row0, nrows, dx.data_ptr(), _stream()), "tvs_unslice_rows")


def wgrad_small(dy, x, dw):
    """dw[n,k] += sum_m dy[m,n] x[m,k]; dy (M,N) and x (M,K) are f32 2-D views."""
    require_device()
    _chk(dy, torch.float32, "dy", True); _chk(x, torch.float32, "x", True); _chk(dw, torch.float32, "dw")
    M, N = dy.shape
    K = x.shape[1]
    _ck(load().tvs_wgrad_small(dy.data_ptr(), dy.stride(0), x.data_ptr(), x.stride(0), M, N, K, dw.data_ptr(), _stream()),
        "tvs_wgrad_small")


def cast_bf16(x, y):
    require_device()
    _chk(x, torch.float32, "x"); _chk(y, torch.bfloat16, "y")
    _ck(load().tvs_cast_bf16(x.data_ptr(), y.data_ptr(), x.numel(), _stream()), "tvs_cast_bf16")


def add_f32(y, x):
    require_device()
    _chk(x, torch.float32, "x"); _chk(y, torch.float32, "y")
    _ck(load().tvs_add_f32(y.data_ptr(), x.data_ptr(), x.numel(), _stream()), "tvs_add_f32")


def film_fwd(x, mul, add, y=None, y_bf16=None):
    require_device()
    _chk(x, torch.float32, "x"); _chk(mul, torch.float32, "mul"); _chk(add, torch.float32, "add")
    B, S, D = x.shape
    _ck(load().tvs_film_fwd(x.data_ptr(), mul.data_ptr(), add.data_ptr(), B, S, D, _p(y), _p(y_bf16), _stream()),
        "tvs_film_fwd")


def film_bwd(dy, x, mul, dx, dmul, dadd):
    require_device()
    for n_, t in (("dy", dy), ("x", x), ("mul", mul), ("dx", dx), ("dmul", dmul), ("dadd", dadd)):
        _chk(t, torch.float32, n_)
    B, S, D = x.shape
    _ck(load().tvs_film_bwd(dy.data_ptr(), x.data_ptr(), mul.data_ptr(), B, S, D, dx.data_ptr(), dmul.data_ptr(),
                            dadd.data_ptr(), _stream()), "tvs_film_bwd")


def head_fwd(tconv, addmap, bias_t, bias_a, ratio, blend, B, G, P, ksize, logits, add_out=None):
    require_device()
    _chk(tconv, torch.float32, "tconv", True); _chk(addmap, torch.float32, "addmap", True)
    _chk(logits, torch.float32, "logits"); _chk(add_out, torch.float32, "add_out")
    _ck(load().tvs_head_fwd(tconv.data_ptr(), tconv.stride(0), _p(addmap), addmap.stride(0) if addmap is not None else 0,
                            _p(bias_t), _p(bias_a), _p(ratio), blend, B, G, P, ksize, logits.data_ptr(), _p(add_out),
                            _stream()), "tvs_head_fwd")


def head_bwd(dlogits, tconv, add_out, bias_t, ratio, blend, B, G, P, ksize, dtconv_bf16, daddmap, dbias_a, dratio):
    require_device()
    _chk(dlogits, torch.float32, "dlogits"); _chk(dtconv_bf16, torch.bfloat16, "dtconv", True)
    _chk(daddmap, torch.float32, "daddmap", True)
    _ck(load().tvs_head_bwd(dlogits.data_ptr(), _p(tconv), tconv.stride(0) if tconv is not None else 0, _p(add_out),
                            _p(bias_t), _p(ratio), blend, B, G, P, ksize, dtconv_bf16.data_ptr(), dtconv_bf16.stride(0),
                            _p(daddmap), daddmap.stride(0) if daddmap is not None else 0, _p(dbias_a), _p(dratio),
                            _stream()), "tvs_head_bwd")


def dicebce_scratch_bytes(B: int, N: int) -> int:
    return int(load().tvs_dicebce_scratch_bytes(B, N))


def dicebce_metrics_fwd(logits, mask, threshold, lambda_dice, lambda_ce, parts, counts, confmat, loss, scratch):
    require_device()
    _chk(logits, torch.float32, "logits"); _chk(mask, torch.float32, "mask")
    _chk(parts, torch.float64, "parts"); _chk(counts, torch.int64, "counts"); _chk(confmat, torch.int64, "confmat")
    _chk(loss, torch.float32, "loss")
    B = logits.shape[0]
    N = logits.numel() // B
    _ck(load().tvs_dicebce_metrics_fwd(logits.data_ptr(), mask.data_ptr(), B, N, threshold, lambda_dice, lambda_ce,
                                       _p(parts), _p(counts), _p(confmat), _p(loss), scratch.data_ptr(), _stream()),
        "tvs_dicebce_metrics_fwd")


def metrics_from_probs(preds, mask, threshold, counts, confmat, scratch):
    require_device()
    _chk(preds, torch.float32, "preds"); _chk(mask, torch.float32, "mask")
    _chk(counts, torch.int64, "counts"); _chk(confmat, torch.int64, "confmat")
    B = preds.shape[0]
    N = preds.numel() // B
    _ck(load().tvs_metrics_from_probs(preds.data_ptr(), mask.data_ptr(), B, N, threshold, _p(counts), _p(confmat),
                                      scratch.data_ptr(), _stream()), "tvs_metrics_from_probs")


def dicebce_bwd(logits, mask, parts, gscale, lambda_dice, lambda_ce, dlogits):
    require_device()
    _chk(logits, torch.float32, "logits"); _chk(mask, torch.float32, "mask"); _chk(parts, torch.float64, "parts")
    _chk(gscale, torch.float32, "gscale"); _chk(dlogits, torch.float32, "dlogits")
    B = logits.shape[0]
    N = logits.numel() // B
    _ck(load().tvs_dicebce_bwd(logits.data_ptr(), mask.data_ptr(), parts.data_ptr(), _p(gscale), B, N, lambda_dice,
                               lambda_ce, dlogits.data_ptr(), _stream()), "tvs_dicebce_bwd")


def adamw_flat(param, grad, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, weight_decay, step, grad_scale=1.0,
               step_dev=None, lr_dev=None):
    require_device()
    for n_, t in (("param", param), ("grad", grad), ("exp_avg", exp_avg), ("exp_avg_sq", exp_avg_sq)):
        _chk(t, torch.float32, n_)
    _ck(load().tvs_adamw_flat(param.data_ptr(), grad.data_ptr(), exp_avg.data_ptr(), exp_avg_sq.data_ptr(), param.numel(),
                              lr, beta1, beta2, eps, weight_decay, step, grad_scale, _p(step_dev), _p(lr_dev), _stream()),
        "tvs_adamw_flat")


def counter_inc(counter):
    require_device()
    _chk(counter, torch.int32, "counter")
    _ck(load().tvs_counter_inc(counter.data_ptr(), _stream()), "tvs_counter_inc")


# ------------------------------------------------------------------------------------------------------------------
# optional per-call timing (bench.py's kernel-share / roofline pass): CUDA events on the launching stream
# ------------------------------------------------------------------------------------------------------------------
_prof: list | None = None


def set_profiler(records: list | None) -> None:
    """When ``records`` is a list every op below appends (name, key, start_event, end_event, algorithmic_flops,
    algorithmic_bytes, stream handle)."""
    global _prof
    _prof = records


def _flops(name, args, kwargs) -> tuple[str, float]:
    if name == "gemm":
        M, (N, K) = args[0].shape[0], args[1].shape
        conv = "conv3x3_" if kwargs.get("conv_hw") is not None else ""
        tag = "_tf32" if args[0].dtype == torch.float32 else ("_f16" if args[0].dtype == torch.float16 else "")
        return f"{conv}{M}x{N}x{K}{tag}", 2.0 * M * N * K
    if name == "attn_fwd":
        _, B, S, H, hd = args[:5]
        return f"B{B}S{S}H{H}d{hd}", 4.0 * B * H * S * S * hd
    if name == "attn_bwd":
        B, S, H, hd = args[4:8]
        rb = kwargs.get("row_begin", 0)
        frac = 1.0 if not rb else (-(-S // 128) - rb // 128) / -(-S // 128)
        return f"B{B}S{S}H{H}d{hd}" + (f"_from{rb}" if rb else ""), 10.0 * B * H * S * S * hd * frac
    if name in ("ffn64_fwd", "ffn64_bwd"):
        (M, D), F = args[0].shape, (args[1] if name == "ffn64_fwd" else args[2])[0].shape[0]
        return f"{M}x{D}x{F}", (4.0 if name == "ffn64_fwd" else 6.0) * M * D * F
    if name in ("cross_attn_fwd", "cross_attn_bwd"):
        i = 4 if name == "cross_attn_fwd" else 7
        B, Sq, Sk, H = args[i:i + 4]
        return f"B{B}Sq{Sq}Sk{Sk}H{H}", 0.0
    if name in ("layernorm_fwd", "layernorm_bwd"):          # per-shape rows: the [B*S, 768] tower launches are not pooled with the
        return "x".join(str(d) for d in args[0].shape), 0.0     # launch-bound [B*L, 512] text-tower ones in the roofline table
    if name in ("im2col_nhwc", "col2im_nhwc", "round_tf32", "relu_mask"):
        t = args[-1] if name == "im2col_nhwc" else args[0]
        return "x".join(str(d) for d in t.shape), 0.0
    return "", 0.0


def _bytes(name, args, kwargs) -> float:
    """Algorithmic HBM bytes of the HBM-bound kernels (each operand read / written once)."""
    def nb(t):
        return 0 if t is None else t.numel() * t.element_size()

    if name == "layernorm_fwd":
        return nb(args[0]) + nb(kwargs.get("y_f32")) + nb(kwargs.get("y_bf16"))
    if name == "layernorm_bwd":
        return nb(args[0]) + nb(args[1]) + sum(nb(kwargs.get(k)) for k in ("dx_add", "dx_f32", "dx_bf16"))
    if name in ("dicebce_metrics_fwd", "metrics_from_probs"):
        return nb(args[0]) + nb(args[1])
    if name == "dicebce_bwd":
        return nb(args[0]) + nb(args[1]) + nb(args[-1])
    if name == "adamw_flat":
        return 7.0 * nb(args[0])                    # p, g, m, v read; p, m, v written
    if name in ("cast_bf16", "add_f32"):
        return 1.5 * nb(args[0]) if name == "cast_bf16" else 3.0 * nb(args[0])
    if name == "head_fwd":
        return nb(args[0]) + nb(args[1]) + nb(args[-2]) + nb(args[-1])
    if name == "head_bwd":
        return sum(nb(a) for a in args if torch.is_tensor(a))
    if name in ("film_fwd", "film_bwd"):
        return sum(nb(a) for a in args if torch.is_tensor(a))
    return 0.0


# NVTX ranges (tracing): TVS_NVTX=1 brackets every library call ("tvs.<op> <shape>") and the phases the
# engines mark with ``nvtx_range`` (towers, decoder, loss, optimizer), so a timeline tool shows the step's structure.  Off by
# default: the ranges are host-side markers and cost a few hundred nanoseconds per call outside graph replay.
NVTX = os.environ.get("TVS_NVTX", "0") == "1"


class nvtx_range:
    """``with abi.nvtx_range("vision_tower.fwd"):`` - a no-op unless TVS_NVTX=1."""

    __slots__ = ("name",)

    def __init__(self, name: str):
        self.name = name

    def __enter__(self):
        if NVTX:
            torch.cuda.nvtx.range_push(self.name)
        return self

    def __exit__(self, *exc):
        if NVTX:
            torch.cuda.nvtx.range_pop()
        return False


def _wrap(fn, name):
    def op(*args, **kwargs):
        if _prof is None and not NVTX:
            return fn(*args, **kwargs)
        if NVTX:
            torch.cuda.nvtx.range_push(f"tvs.{name} {_flops(name, args, kwargs)[0]}".rstrip())
            try:
                return op_inner(*args, **kwargs)
            finally:
                torch.cuda.nvtx.range_pop()
        return op_inner(*args, **kwargs)

    def op_inner(*args, **kwargs):
        if _prof is None:
            return fn(*args, **kwargs)
        # inside a stream capture the events become EVENT-RECORD NODES of the graph (cudaEventRecordExternal): every replay
        # of that (instrumented) graph re-stamps them, which is how bench.py times kernels inside the graph itself
        ext = torch.cuda.is_current_stream_capturing()
        e0, e1 = torch.cuda.Event(enable_timing=True, external=ext), torch.cuda.Event(enable_timing=True, external=ext)
        e0.record()
        out = fn(*args, **kwargs)
        e1.record()
        key, fl = _flops(name, args, kwargs)
        if name == "gemm":
            key += f"|{gemm_last_variant()}"
        _prof.append((name, key, e0, e1, fl, _bytes(name, args, kwargs), torch.cuda.current_stream().cuda_stream))
        return out

    op.__name__, op.__doc__ = fn.__name__, fn.__doc__
    return op


for _n in ("gemm", "layernorm_fwd", "layernorm_bwd", "attn_fwd", "attn_bwd", "im2col_patches", "vision_assemble",
           "prompt_overwrite", "prompt_grad", "slice_rows", "unslice_rows", "wgrad_small", "cast_bf16", "add_f32", "film_fwd",
           "film_bwd", "head_fwd", "head_bwd", "dicebce_metrics_fwd", "dicebce_bwd", "metrics_from_probs", "adamw_flat",
           "counter_inc", "ffn64_fwd", "ffn64_bwd"):
    globals()[_n] = _wrap(globals()[_n], _n)


# ---- CRIS path (conv_ops.cu) -----------------------------------------------------------------------------------------
def _chk2(t, name, dtypes=(torch.float32,)):
    if t is None:
        return
    if not t.is_cuda:
        raise TvsError(f"{name} must be a CUDA tensor")
    if t.dtype not in dtypes:
        raise TvsError(f"{name} must be one of {dtypes}, got {t.dtype}")
    if t.dim() != 2 or t.stride(1) != 1:
        raise TvsError(f"{name} must be 2-D with unit inner stride, got shape {tuple(t.shape)} strides {t.stride()}")


def round_tf32(x, y):
    """y = x rounded to nearest tf32 (the tf32 MMA itself truncates); 2-D f32 views, row strides allowed."""
    require_device()
    _chk2(x, "x"); _chk2(y, "y")
    if x.shape != y.shape:
        raise TvsError("round_tf32: shape mismatch")
    _ck(load().tvs_round_tf32(x.data_ptr(), x.stride(0), x.shape[0], x.shape[1], y.data_ptr(), y.stride(0), _stream()), "tvs_round_tf32")


def pad_nhwc(x, B, H, W, C, xp, round_tf32=False):
    """x contiguous [B*H*W, C] -> xp contiguous [B*(H+2)*(W+2), C] with a zero border (implicit-GEMM conv operand)."""
    require_device()
    _chk2(x, "x", (torch.float32, torch.bfloat16)); _chk2(xp, "xp", (x.dtype,))
    if not x.is_contiguous() or not xp.is_contiguous() or tuple(x.shape) != (B * H * W, C) or tuple(xp.shape) != (B * (H + 2) * (W + 2), C):
        raise TvsError("pad_nhwc: x must be contiguous [B*H*W, C] and xp contiguous [B*(H+2)*(W+2), C]")
    _ck(load().tvs_pad_nhwc(x.data_ptr(), x.element_size(), B, H, W, C, xp.data_ptr(), int(round_tf32), _stream()), "tvs_pad_nhwc")


def im2col_nhwc(x, B, H, W, C, ksize, stride, pad, col, round_tf32=False):
    """x: contiguous [B*H*W, C] (bf16 or f32) -> col [B*Ho*Wo, ld >= k*k*C] of the same dtype (tail zero-filled)."""
    require_device()
    _chk2(x, "x", (torch.float32, torch.bfloat16)); _chk2(col, "col", (x.dtype,))
    if not x.is_contiguous() or tuple(x.shape) != (B * H * W, C):
        raise TvsError(f"im2col: x must be contiguous {(B * H * W, C)}, got {tuple(x.shape)}")
    Ho, Wo = (H + 2 * pad - ksize) // stride + 1, (W + 2 * pad - ksize) // stride + 1
    if col.shape[0] != B * Ho * Wo or col.stride(0) != col.shape[1]:
        raise TvsError("im2col: col must be a dense [B*Ho*Wo, ld] matrix")
    _ck(load().tvs_im2col_nhwc(x.data_ptr(), x.element_size(), B, H, W, C, ksize, stride, pad, col.data_ptr(), col.shape[1],
                               int(round_tf32), _stream()), "tvs_im2col_nhwc")


def col2im_nhwc(dcol, B, H, W, Ccol, Cx, ksize, dx, relu_mask=None):
    require_device()
    _chk2(dcol, "dcol"); _chk2(dx, "dx"); _chk2(relu_mask, "relu_mask")
    _ck(load().tvs_col2im_nhwc(dcol.data_ptr(), dcol.stride(0), B, H, W, Ccol, Cx, ksize, _p(relu_mask),
                               relu_mask.stride(0) if relu_mask is not None else 0, dx.data_ptr(), dx.stride(0), _stream()),
        "tvs_col2im_nhwc")


def relu_mask(dy, y, out):
    require_device()
    _chk2(dy, "dy"); _chk2(y, "y"); _chk2(out, "out")
    if dy.shape != y.shape or dy.shape != out.shape:
        raise TvsError("relu_mask: shape mismatch")
    _ck(load().tvs_relu_mask(dy.data_ptr(), dy.stride(0), y.data_ptr(), y.stride(0), dy.shape[0], dy.shape[1], out.data_ptr(),
                             out.stride(0), _stream()), "tvs_relu_mask")


def avgpool2_nhwc(x, B, H, W, C, y, round_tf32=False):
    require_device()
    _chk2(x, "x", (torch.float32, torch.bfloat16)); _chk2(y, "y", (x.dtype,))
    if not x.is_contiguous():
        raise TvsError("avgpool2: x must be contiguous")
    _ck(load().tvs_avgpool2_nhwc(x.data_ptr(), int(x.dtype == torch.bfloat16) | (2 if round_tf32 else 0), B, H, W, C, y.data_ptr(),
                                 y.stride(0), _stream()), "tvs_avgpool2_nhwc")


def upsample2x_fwd(x, B, H, W, C, y):
    require_device()
    _chk2(x, "x"); _chk2(y, "y")
    if not x.is_contiguous():
        raise TvsError("upsample2x: x must be contiguous")
    _ck(load().tvs_upsample2x_fwd(x.data_ptr(), B, H, W, C, y.data_ptr(), y.stride(0), _stream()), "tvs_upsample2x_fwd")


def upsample2x_bwd(dy, B, H, W, C, dx):
    require_device()
    _chk2(dy, "dy"); _chk2(dx, "dx")
    if not dx.is_contiguous():
        raise TvsError("upsample2x_bwd: dx must be contiguous")
    _ck(load().tvs_upsample2x_bwd(dy.data_ptr(), dy.stride(0), B, H, W, C, dx.data_ptr(), _stream()), "tvs_upsample2x_bwd")


def cross_attn_fwd(q, k, v, key_mask, B, Sq, Sk, H, hd, out, lse, causal=False):
    require_device()
    for n, t in (("q", q), ("k", k), ("v", v), ("out", out)):
        _chk2(t, n)
    if k.stride(0) != v.stride(0):
        raise TvsError("cross_attn: k and v must share a row stride")
    _chk(lse, torch.float32, "lse"); _chk(key_mask, torch.uint8, "key_mask")
    _ck(load().tvs_cross_attn_fwd(q.data_ptr(), q.stride(0), k.data_ptr(), v.data_ptr(), k.stride(0), _p(key_mask), B, Sq, Sk, H, hd,
                                  int(causal), out.data_ptr(), out.stride(0), lse.data_ptr(), _stream()), "tvs_cross_attn_fwd")


def cross_attn_bwd(q, k, v, key_mask, out, dout, lse, B, Sq, Sk, H, hd, dq, dk, dv, delta, causal=False):
    require_device()
    for n, t in (("q", q), ("k", k), ("v", v), ("out", out), ("dout", dout), ("dq", dq), ("dk", dk), ("dv", dv)):
        _chk2(t, n)
    if k.stride(0) != v.stride(0) or out.stride(0) != dout.stride(0) or dk.stride(0) != dv.stride(0):
        raise TvsError("cross_attn_bwd: paired operands must share row strides")
    _ck(load().tvs_cross_attn_bwd(q.data_ptr(), q.stride(0), k.data_ptr(), v.data_ptr(), k.stride(0), _p(key_mask), out.data_ptr(),
                                  dout.data_ptr(), out.stride(0), lse.data_ptr(), B, Sq, Sk, H, hd, int(causal), dq.data_ptr(), dq.stride(0),
                                  dk.data_ptr(), dv.data_ptr(), dk.stride(0), delta.data_ptr(), _stream()), "tvs_cross_attn_bwd")


def dynconv_fwd(x, w, bias, B, H, W, C, taps, out):
    """w: [B, >= C*9] rows (channel-major, tap-minor); bias: a column view [B, 1] of the same matrix."""
    require_device()
    _chk2(x, "x"); _chk2(w, "w"); _chk(taps, torch.float32, "taps"); _chk(out, torch.float32, "out")
    _ck(load().tvs_dynconv_fwd(x.data_ptr(), w.data_ptr(), w.stride(0), bias.data_ptr(), bias.stride(0), B, H, W, C, taps.data_ptr(),
                               out.data_ptr(), _stream()), "tvs_dynconv_fwd")


def dynconv_bwd(dout, x, w, B, H, W, C, dx, dw_part):
    require_device()
    _chk(dout, torch.float32, "dout"); _chk2(x, "x"); _chk2(w, "w"); _chk(dx, torch.float32, "dx"); _chk(dw_part, torch.float32, "dw_part")
    _ck(load().tvs_dynconv_bwd(dout.data_ptr(), x.data_ptr(), w.data_ptr(), w.stride(0), B, H, W, C, dx.data_ptr(), dw_part.data_ptr(),
                               dw_part.shape[0], _stream()), "tvs_dynconv_bwd")


def resample2d_fwd(inp, B, Hi, Wi, Ho, Wo, tab, tile, out):
    require_device()
    _chk(inp, torch.float32, "in"); _chk(out, torch.float32, "out")
    _ck(load().tvs_resample2d_fwd(inp.data_ptr(), B, Hi, Wi, Ho, Wo, tab["iy"].data_ptr(), tab["wy"].data_ptr(), tab["ix"].data_ptr(),
                                  tab["wx"].data_ptr(), tab["ntaps"], tile, out.data_ptr(), _stream()), "tvs_resample2d_fwd")


def resample2d_u8(inp, B, Hi, Wi, Ho, Wo, tab, out):
    """Bicubic resize fused with save_image's quantisation: f32 (B,Hi,Wi) -> u8 (B,Ho,Wo)."""
    require_device()
    _chk(inp, torch.float32, "in"); _chk(out, torch.uint8, "out")
    _ck(load().tvs_resample2d_u8(inp.data_ptr(), B, Hi, Wi, Ho, Wo, tab["iy"].data_ptr(), tab["wy"].data_ptr(), tab["ix"].data_ptr(),
                                 tab["wx"].data_ptr(), tab["ntaps"], out.data_ptr(), _stream()), "tvs_resample2d_u8")


def resample2d_bwd(dout, B, Hi, Wi, Ho, Wo, tab, tile, din):
    require_device()
    if dout.dtype not in (torch.float32, torch.bfloat16) or not dout.is_contiguous():
        raise TvsError("resample2d_bwd: dout must be contiguous f32 or bf16")
    _chk(din, torch.float32, "din")
    _ck(load().tvs_resample2d_bwd(dout.data_ptr(), int(dout.dtype == torch.bfloat16), B, Hi, Wi, Ho, Wo, tab["ty"].data_ptr(),
                                  tab["twy"].data_ptr(), tab["cy"].data_ptr(), tab["tx"].data_ptr(), tab["twx"].data_ptr(),
                                  tab["cx"].data_ptr(), tab["max_taps"], tile, din.data_ptr(), _stream()), "tvs_resample2d_bwd")


for _n in ("round_tf32", "pad_nhwc", "im2col_nhwc", "col2im_nhwc", "relu_mask", "avgpool2_nhwc", "upsample2x_fwd", "upsample2x_bwd", "cross_attn_fwd",
           "cross_attn_bwd", "dynconv_fwd", "dynconv_bwd", "resample2d_fwd", "resample2d_bwd", "resample2d_u8"):
    globals()[_n] = _wrap(globals()[_n], _n)


# ---- eval input transforms (preprocess.cu) ---------------------------------------------------------------------------------
def preproc_image_u8(img_u8, xofs, xcoef, yofs, ycoef, mean255, inv_std255, out_chw=None, out_u8=None):
    """img_u8: CUDA uint8 [Hi, Wi, 3] (row stride allowed); tables: CUDA int32; mean255 / inv_std255: sequences of 3 floats.
    out_chw: CUDA f32 [3, Ho, Wo]; out_u8: CUDA uint8 [Ho, Wo, 3] (the resized image before Normalize)."""
    require_device()
    if not (img_u8.is_cuda and img_u8.dtype == torch.uint8 and img_u8.dim() == 3 and img_u8.shape[2] == 3 and img_u8.stride(2) == 1
            and img_u8.stride(1) == 3):
        raise TvsError(f"preproc_image_u8: image must be a CUDA uint8 HWC tensor with packed pixels, got {tuple(img_u8.shape)} {img_u8.dtype}")
    for t in (xofs, xcoef, yofs, ycoef):
        _chk(t, torch.int32, "tap table")
    Ho, Wo = yofs.numel(), xofs.numel()
    if xcoef.numel() != 4 * Wo or ycoef.numel() != 4 * Ho:
        raise TvsError("preproc_image_u8: coefficient tables must be [n_out, 4]")
    if out_chw is not None:
        _chk(out_chw, torch.float32, "out_chw")
        if tuple(out_chw.shape) != (3, Ho, Wo):
            raise TvsError(f"preproc_image_u8: out_chw must be {(3, Ho, Wo)}, got {tuple(out_chw.shape)}")
    if out_u8 is not None:
        _chk(out_u8, torch.uint8, "out_u8")
        if tuple(out_u8.shape) != (Ho, Wo, 3):
            raise TvsError(f"preproc_image_u8: out_u8 must be {(Ho, Wo, 3)}, got {tuple(out_u8.shape)}")
    m = (c_float * 3)(*[float(v) for v in mean255])
    d = (c_float * 3)(*[float(v) for v in inv_std255])
    Hi, Wi, _ = img_u8.shape
    _ck(load().tvs_preproc_image_u8(img_u8.data_ptr(), Hi, Wi, img_u8.stride(0), xofs.data_ptr(), xcoef.data_ptr(), yofs.data_ptr(),
                                    ycoef.data_ptr(), ctypes.cast(m, c_void_p), ctypes.cast(d, c_void_p), Ho, Wo, _p(out_chw), _p(out_u8),
                                    _stream()), "tvs_preproc_image_u8")


def resize_nearest_f32(x, xofs, yofs, out):
    """x: CUDA f32 [Hi, Wi] (row stride allowed) -> out CUDA f32 [Ho, Wo] = x[yofs][:, xofs]."""
    require_device()
    if not (x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 and x.stride(1) == 1):
        raise TvsError("resize_nearest_f32: x must be a CUDA f32 [H, W] tensor with unit inner stride")
    _chk(xofs, torch.int32, "xofs"); _chk(yofs, torch.int32, "yofs"); _chk(out, torch.float32, "out")
    Ho, Wo = yofs.numel(), xofs.numel()
    if tuple(out.shape) != (Ho, Wo):
        raise TvsError(f"resize_nearest_f32: out must be {(Ho, Wo)}")
    _ck(load().tvs_resize_nearest_f32(x.data_ptr(), x.shape[0], x.shape[1], x.stride(0), xofs.data_ptr(), yofs.data_ptr(), Ho, Wo,
                                      out.data_ptr(), _stream()), "tvs_resize_nearest_f32")


# ---- train-time augmentations (preprocess.cu) --------------------------------------------------------------------------------
def _chk_hwc_u8(img_u8, what):
    if not (img_u8.is_cuda and img_u8.dtype == torch.uint8 and img_u8.dim() == 3 and img_u8.shape[2] == 3 and img_u8.stride(2) == 1
            and img_u8.stride(1) == 3):
        raise TvsError(f"{what}: image must be a CUDA uint8 HWC tensor with packed pixels, got {tuple(img_u8.shape)} {img_u8.dtype}")


def _chk_walk(walk, what):
    adelta, bdelta, x0, y0 = walk
    for t in walk:
        _chk(t, torch.int32, "walk table")
    if adelta.numel() != bdelta.numel() or x0.numel() != y0.numel():
        raise TvsError(f"{what}: walk tables must be (adelta [Wo], bdelta [Wo], x0 [Ho], y0 [Ho])")
    return y0.numel(), adelta.numel()


def warp_affine_u8(img_u8, walk, tab, mean255=None, inv_std255=None, lut=None, out_chw=None, out_u8=None):
    """cv2.warpAffine(INTER_CUBIC, BORDER_REPLICATE) of a CUDA uint8 [Hi, Wi, 3] image.  walk = (adelta, bdelta, x0, y0) CUDA int32
    tables of cv::warpAffine's fixed-point walk (rounding term 16 included); tab: CUDA int16 [32, 32, 4, 4] weight table;
    lut: optional CUDA uint8 [256] applied to the warped bytes; out_chw f32 [3, Ho, Wo] (normalised) and / or out_u8 [Ho, Wo, 3]."""
    require_device()
    _chk_hwc_u8(img_u8, "warp_affine_u8")
    Ho, Wo = _chk_walk(walk, "warp_affine_u8")
    if not (tab.is_cuda and tab.dtype == torch.int16 and tab.numel() == 32 * 32 * 16 and tab.is_contiguous()):
        raise TvsError("warp_affine_u8: tab must be a contiguous CUDA int16 [32, 32, 4, 4] tensor")
    if lut is not None and not (lut.is_cuda and lut.dtype == torch.uint8 and lut.numel() == 256 and lut.is_contiguous()):
        raise TvsError("warp_affine_u8: lut must be a contiguous CUDA uint8 [256] tensor")
    if out_chw is None and out_u8 is None:
        raise TvsError("warp_affine_u8: no output")
    if out_chw is not None:
        _chk(out_chw, torch.float32, "out_chw")
        if tuple(out_chw.shape) != (3, Ho, Wo) or mean255 is None or inv_std255 is None:
            raise TvsError(f"warp_affine_u8: out_chw must be {(3, Ho, Wo)} and needs mean255 / inv_std255")
    if out_u8 is not None:
        _chk(out_u8, torch.uint8, "out_u8")
        if tuple(out_u8.shape) != (Ho, Wo, 3):
            raise TvsError(f"warp_affine_u8: out_u8 must be {(Ho, Wo, 3)}, got {tuple(out_u8.shape)}")
    m = (c_float * 3)(*[float(v) for v in (mean255 if mean255 is not None else (0, 0, 0))])
    d = (c_float * 3)(*[float(v) for v in (inv_std255 if inv_std255 is not None else (1, 1, 1))])
    Hi, Wi, _ = img_u8.shape
    _ck(load().tvs_warp_affine_u8(img_u8.data_ptr(), Hi, Wi, img_u8.stride(0), walk[0].data_ptr(), walk[1].data_ptr(), walk[2].data_ptr(),
                                  walk[3].data_ptr(), tab.data_ptr(), _p(lut), ctypes.cast(m, c_void_p), ctypes.cast(d, c_void_p), Ho, Wo,
                                  _p(out_chw), _p(out_u8), _stream()), "tvs_warp_affine_u8")


def warp_affine_nearest_f32(x, walk, out):
    """cv2.warpAffine(INTER_NEAREST, BORDER_REPLICATE) of a CUDA f32 [Hi, Wi] mask; walk tables with the rounding term 512."""
    require_device()
    if not (x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 and x.stride(1) == 1):
        raise TvsError("warp_affine_nearest_f32: x must be a CUDA f32 [H, W] tensor with unit inner stride")
    Ho, Wo = _chk_walk(walk, "warp_affine_nearest_f32")
    _chk(out, torch.float32, "out")
    if tuple(out.shape) != (Ho, Wo):
        raise TvsError(f"warp_affine_nearest_f32: out must be {(Ho, Wo)}")
    _ck(load().tvs_warp_affine_nearest_f32(x.data_ptr(), x.shape[0], x.shape[1], x.stride(0), walk[0].data_ptr(), walk[1].data_ptr(),
                                           walk[2].data_ptr(), walk[3].data_ptr(), Ho, Wo, out.data_ptr(), _stream()), "tvs_warp_affine_nearest_f32")


def lut_normalize_u8(img_u8, lut, mean255, inv_std255, out_chw):
    """(lut[img] - mean255) * inv_std255 -> f32 [3, H, W]; img CUDA uint8 [H, W, 3], lut CUDA uint8 [256] or None."""
    require_device()
    _chk_hwc_u8(img_u8, "lut_normalize_u8")
    if lut is not None and not (lut.is_cuda and lut.dtype == torch.uint8 and lut.numel() == 256 and lut.is_contiguous()):
        raise TvsError("lut_normalize_u8: lut must be a contiguous CUDA uint8 [256] tensor")
    H, W, _ = img_u8.shape
    _chk(out_chw, torch.float32, "out_chw")
    if tuple(out_chw.shape) != (3, H, W):
        raise TvsError(f"lut_normalize_u8: out_chw must be {(3, H, W)}")
    m = (c_float * 3)(*[float(v) for v in mean255])
    d = (c_float * 3)(*[float(v) for v in inv_std255])
    _ck(load().tvs_lut_normalize_u8(img_u8.data_ptr(), H, W, img_u8.stride(0), _p(lut), ctypes.cast(m, c_void_p), ctypes.cast(d, c_void_p),
                                    out_chw.data_ptr(), _stream()), "tvs_lut_normalize_u8")
