"""CRIS weights container with the reference's constructor and ``state_dict`` layout
(src/models/components/cris_model/__init__.py:20-132, clip.py, layers.py).

The modules below only HOLD the frozen parameters under the reference's key names (``backbone.visual.layer1.0.conv1
.weight``, ``neck.f2_cat.0.weight``, ``decoder.layers.0.multihead_attn.in_proj_weight``, ``proj.txt.weight`` ...), so
``cris_best_single.pth`` / Lightning checkpoints load with ``strict=True``.  Nothing here computes: the B200 engine
(``tunevlseg_b200/engine_cris.py``) packs the tensors once (BatchNorm folded into GEMM operands) and runs its own
kernels; the containers' ``forward`` is never called.
"""
from __future__ import annotations

from collections import OrderedDict
from collections.abc import Mapping

import torch
from torch import nn

from .... import engine_cris


def _conv_layer(cin, cout, k=1, padding=0):
    return nn.Sequential(nn.Conv2d(cin, cout, k, 1, padding, bias=False), nn.BatchNorm2d(cout), nn.ReLU())


class _Holder(nn.Module):
    """A module tree that exists for its parameters only."""

    def forward(self, *a, **k):
        raise NotImplementedError("parameter container: the sm_100a engine runs this model (engine_cris.py)")


class _Bottleneck(_Holder):
    def __init__(self, inplanes, planes, stride):
        super().__init__()
        self.conv1, self.bn1 = nn.Conv2d(inplanes, planes, 1, bias=False), nn.BatchNorm2d(planes)
        self.conv2, self.bn2 = nn.Conv2d(planes, planes, 3, padding=1, bias=False), nn.BatchNorm2d(planes)
        self.conv3, self.bn3 = nn.Conv2d(planes, planes * 4, 1, bias=False), nn.BatchNorm2d(planes * 4)
        self.downsample = None
        if stride > 1 or inplanes != planes * 4:
            self.downsample = nn.Sequential(OrderedDict((("-1", nn.AvgPool2d(stride)), ("0", nn.Conv2d(inplanes, planes * 4, 1, bias=False)),
                                                         ("1", nn.BatchNorm2d(planes * 4)))))


class _AttentionPool2d(_Holder):
    def __init__(self, spacial_dim, embed_dim, num_heads, output_dim):
        super().__init__()
        self.spacial_dim, self.num_heads = spacial_dim, num_heads
        self.positional_embedding = nn.Parameter(torch.randn(spacial_dim ** 2 + 1, embed_dim) / embed_dim ** 0.5)
        self.k_proj, self.q_proj, self.v_proj = nn.Linear(embed_dim, embed_dim), nn.Linear(embed_dim, embed_dim), nn.Linear(embed_dim, embed_dim)
        self.c_proj = nn.Linear(embed_dim, output_dim or embed_dim)
        self.connect = nn.Sequential(nn.Conv2d(embed_dim, output_dim, 1, stride=1, bias=False), nn.BatchNorm2d(output_dim))


class _ModifiedResNet(_Holder):
    def __init__(self, layers, output_dim, heads, input_resolution, width):
        super().__init__()
        self.output_dim, self.input_resolution, self.layers = output_dim, input_resolution, tuple(layers)
        self.conv1, self.bn1 = nn.Conv2d(3, width // 2, 3, stride=2, padding=1, bias=False), nn.BatchNorm2d(width // 2)
        self.conv2, self.bn2 = nn.Conv2d(width // 2, width // 2, 3, padding=1, bias=False), nn.BatchNorm2d(width // 2)
        self.conv3, self.bn3 = nn.Conv2d(width // 2, width, 3, padding=1, bias=False), nn.BatchNorm2d(width)
        inpl = width
        for li, n in enumerate(layers, start=1):
            planes = width * 2 ** (li - 1)
            blocks = []
            for bi in range(n):
                blocks.append(_Bottleneck(inpl, planes, 2 if (bi == 0 and li > 1) else 1))
                inpl = planes * 4
            setattr(self, f"layer{li}", nn.Sequential(*blocks))
        self.attnpool = _AttentionPool2d(input_resolution // 32, width * 32, heads, output_dim)


class _ResidualAttentionBlock(_Holder):
    def __init__(self, d_model, n_head):
        super().__init__()
        self.attn = nn.MultiheadAttention(d_model, n_head)
        self.ln_1 = nn.LayerNorm(d_model)
        self.mlp = nn.Sequential(OrderedDict([("c_fc", nn.Linear(d_model, d_model * 4)), ("gelu", nn.Identity()),
                                              ("c_proj", nn.Linear(d_model * 4, d_model))]))
        self.ln_2 = nn.LayerNorm(d_model)


class _Transformer(_Holder):
    def __init__(self, width, layers, heads):
        super().__init__()
        self.width, self.layers = width, layers
        self.resblocks = nn.ModuleList(_ResidualAttentionBlock(width, heads) for _ in range(layers))


class CLIP(_Holder):
    """clip.py:389-467 (ModifiedResNet variant only: CRIS uses CLIP-RN50)."""

    def __init__(self, embed_dim, image_resolution, vision_layers, vision_width, context_length, vocab_size, transformer_width,
                 transformer_heads, transformer_layers):
        super().__init__()
        self.context_length, self.vocab_size = context_length, vocab_size
        self.visual = _ModifiedResNet(vision_layers, embed_dim, vision_width * 32 // 64, image_resolution, vision_width)
        self.transformer = _Transformer(transformer_width, transformer_layers, transformer_heads)
        self.token_embedding = nn.Embedding(vocab_size, transformer_width)
        self.positional_embedding = nn.Parameter(torch.empty(context_length, transformer_width).normal_(std=0.01))
        self.ln_final = nn.LayerNorm(transformer_width)
        self.text_projection = nn.Parameter(torch.empty(transformer_width, embed_dim).normal_(std=transformer_width ** -0.5))
        self.logit_scale = nn.Parameter(-torch.log(torch.tensor(0.07)))


_FP16_ROUND = (nn.Conv1d, nn.Conv2d, nn.Linear)


def build_model(state_dict: Mapping[str, torch.Tensor]) -> CLIP:
    """clip.py:578-647: infer every dimension from an OpenAI CLIP-RN50 state dict, load it non-strictly (the CRIS-only
    ``attnpool.connect`` keeps its fresh initialisation) and reproduce the fp16 round trip of ``convert_weights``
    (conv / linear / attention tensors and ``text_projection`` go through fp16 once; the caller casts back to fp32)."""
    if "visual.proj" in state_dict:
        raise NotImplementedError("CRIS is built on CLIP-RN50; the ViT image encoder of clip.py:346-386 is not on the path")
    sd = {k: v for k, v in state_dict.items() if k not in ("input_resolution", "context_length", "vocab_size")}
    layers = tuple(len({k.split(".")[2] for k in sd if k.startswith(f"visual.layer{b}")}) for b in range(1, 5))
    width = sd["visual.layer1.0.conv1.weight"].shape[0]
    out_w = round((sd["visual.attnpool.positional_embedding"].shape[0] - 1) ** 0.5)
    if out_w ** 2 + 1 != sd["visual.attnpool.positional_embedding"].shape[0]:
        raise ValueError("Wrong vision grid size or output width!")
    t_width = sd["ln_final.weight"].shape[0]
    model = CLIP(sd["text_projection"].shape[1], out_w * 32, layers, width, sd["positional_embedding"].shape[0],
                 sd["token_embedding.weight"].shape[0], t_width, t_width // 64,
                 len({k.split(".")[2] for k in sd if k.startswith("transformer.resblocks")}))
    model.load_state_dict(sd, strict=False)
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, _FP16_ROUND):
                m.weight.copy_(m.weight.half().float())
                if m.bias is not None:
                    m.bias.copy_(m.bias.half().float())
            if isinstance(m, nn.MultiheadAttention):
                m.in_proj_weight.copy_(m.in_proj_weight.half().float())
                m.in_proj_bias.copy_(m.in_proj_bias.half().float())
        model.text_projection.copy_(model.text_projection.half().float())
    return model.eval()


class FPN(_Holder):
    def __init__(self, in_channels=(512, 1024, 1024), out_channels=(256, 512, 1024)):
        if len(in_channels) < 3:
            raise ValueError("`FPN` module requires `in_channels` of length 3")
        if len(out_channels) < 3:
            raise ValueError("`FPN` module requires `out_channels` of length 3")
        super().__init__()
        i, o = in_channels, out_channels
        self.txt_proj = nn.Sequential(nn.Linear(i[2], o[2], False), nn.BatchNorm1d(o[2]), nn.ReLU())
        self.f1_v_proj = _conv_layer(i[2], o[2], 1, 0)
        self.norm_layer = nn.Sequential(nn.BatchNorm2d(o[2]), nn.ReLU())
        self.f2_v_proj, self.f2_cat = _conv_layer(i[1], o[1], 3, 1), _conv_layer(o[2] + o[1], o[1], 1, 0)
        self.f3_v_proj, self.f3_cat = _conv_layer(i[0], o[0], 3, 1), _conv_layer(o[0] + o[1], o[1], 1, 0)
        self.f4_proj5, self.f4_proj4, self.f4_proj3 = _conv_layer(o[2], o[1], 3, 1), _conv_layer(o[1], o[1], 3, 1), _conv_layer(o[1], o[1], 3, 1)
        self.aggr = _conv_layer(3 * o[1], o[1], 1, 0)
        coord = _Holder()
        coord.conv1 = _conv_layer(o[1] + 2, o[1], 3, 1)
        self.coordconv = nn.Sequential(coord, _conv_layer(o[1], o[1], 3, 1))


class _DecoderLayer(_Holder):
    def __init__(self, d_model, nhead, dim_feedforward, dropout):
        super().__init__()
        self.self_attn_norm, self.cross_attn_norm = nn.LayerNorm(d_model), nn.LayerNorm(d_model)
        self.self_attn = nn.MultiheadAttention(d_model, nhead, dropout=dropout)
        self.multihead_attn = nn.MultiheadAttention(d_model, nhead, dropout=dropout, kdim=d_model, vdim=d_model)
        self.ffn = nn.Sequential(nn.Linear(d_model, dim_feedforward), nn.ReLU(), nn.Dropout(dropout), nn.LayerNorm(dim_feedforward),
                                 nn.Linear(dim_feedforward, d_model))
        self.norm1, self.norm2, self.norm3 = nn.LayerNorm(d_model), nn.LayerNorm(d_model), nn.LayerNorm(d_model)
        self.dropout1, self.dropout2, self.dropout3 = nn.Dropout(dropout), nn.Dropout(dropout), nn.Dropout(dropout)


class TransformerDecoder(_Holder):
    def __init__(self, num_layers, d_model, nhead, dim_ffn, dropout, return_intermediate=False):
        super().__init__()
        if return_intermediate:
            raise NotImplementedError("return_intermediate=True (a list of per-layer outputs) is not used by any reference config")
        self.layers = nn.ModuleList(_DecoderLayer(d_model, nhead, dim_ffn, dropout) for _ in range(num_layers))
        self.num_layers, self.nhead = num_layers, nhead
        self.norm = nn.LayerNorm(d_model)


class Projector(_Holder):
    def __init__(self, word_dim=1024, in_dim=256, kernel_size=3):
        super().__init__()
        if kernel_size != 3:
            raise NotImplementedError("the dynamic-convolution kernel is written for the reference's 3x3 projector")
        self.in_dim, self.kernel_size = in_dim, kernel_size
        self.vis = nn.Sequential(nn.Upsample(scale_factor=2, mode="bilinear"), _conv_layer(in_dim * 2, in_dim * 2, 3, 1),
                                 nn.Upsample(scale_factor=2, mode="bilinear"), _conv_layer(in_dim * 2, in_dim, 3, 1),
                                 nn.Conv2d(in_dim, in_dim, 1))
        self.txt = nn.Linear(word_dim, in_dim * kernel_size * kernel_size + 1)


def load_checkpoint(path, allow_unsafe_pickle: bool | None = None):
    """``torch.load(path, map_location="cpu")`` as the reference calls it (cris_model/__init__.py:66).  The reference's
    conversion script writes with ``pickle_protocol=5`` (scripts/process_cris_checkpoint.py:22), whose FRAME opcode torch's
    weights-only unpickler rejects.  Such files are re-read with Python's own (all-protocol) unpickler **restricted to
    torch's weights-only allow-list of globals** - tensors, storages, dtypes, ``OrderedDict`` - so a checkpoint that names
    any other callable (the malicious case) still fails.  Full, arbitrary-code unpickling happens only on explicit
    opt-in: ``allow_unsafe_pickle=True`` or the environment variable ``TVS_ALLOW_UNSAFE_PICKLE=1``."""
    import os
    import pickle
    import types

    try:
        return torch.load(path, map_location="cpu")
    except pickle.UnpicklingError as first:
        from torch._weights_only_unpickler import _get_allowed_globals

        class _Restricted(pickle.Unpickler):
            def find_class(self, module, name):
                allowed = _get_allowed_globals()
                key = f"{module}.{name}"
                if key not in allowed:
                    raise pickle.UnpicklingError(f"{path}: global {key} is not on the weights-only allow-list")
                return allowed[key]

        mod = types.SimpleNamespace(Unpickler=_Restricted, load=lambda f, **kw: _Restricted(f, **kw).load(), __name__="tvs_restricted_pickle")
        try:
            return torch.load(path, map_location="cpu", weights_only=False, pickle_module=mod)
        except pickle.UnpicklingError:
            if allow_unsafe_pickle is None:
                allow_unsafe_pickle = os.environ.get("TVS_ALLOW_UNSAFE_PICKLE", "0") == "1"
            if not allow_unsafe_pickle:
                raise first
            return torch.load(path, map_location="cpu", weights_only=False)


class CRIS(nn.Module):
    """cris_model/__init__.py:20-132.  ``clip_pretrain`` is the TorchScript ``RN50.pt`` path as in the reference, or a
    CLIP ``state_dict`` mapping (tests / benchmarks hand over random weights: there is no checkpoint in this image)."""

    max_length = 77

    def __init__(self, clip_pretrain, fpn_in, fpn_out, vis_dim, word_dim, num_layers, num_head, dim_ffn, dropout,
                 return_intermediate, img_size=416, freeze_encoder=True, cris_pretrain=None, *args, **kwargs) -> None:
        nn.Module.__init__(self)
        self.img_size = img_size
        self.backbone = self.get_backbone(clip_pretrain)
        self.backbone.requires_grad_(not freeze_encoder)
        self.neck = FPN(in_channels=fpn_in, out_channels=fpn_out)
        self.decoder = TransformerDecoder(num_layers=num_layers, d_model=vis_dim, nhead=num_head, dim_ffn=dim_ffn, dropout=dropout,
                                          return_intermediate=return_intermediate)
        self.proj = Projector(word_dim, vis_dim // 2, 3)
        if cris_pretrain is not None:
            print("Loading CRIS pre-trained model from:", cris_pretrain)
            self.load_state_dict(load_checkpoint(cris_pretrain), strict=True)
        self.word_dim = word_dim

    @staticmethod
    def get_backbone(clip_pretrain) -> CLIP:
        if isinstance(clip_pretrain, Mapping):
            return build_model(clip_pretrain).float()
        clip_model = torch.jit.load(clip_pretrain, map_location="cpu")
        return build_model(clip_model.state_dict()).float()

    def get_pad_mask(self, input_ids, attention_mask):
        return ~(attention_mask.bool()) if attention_mask is not None else input_ids == 0

    # ---- engine plumbing ----------------------------------------------------------------------------------------
    @property
    def packed(self) -> engine_cris.PackedCris:
        dev = self.backbone.token_embedding.weight.device
        pk = getattr(self, "_tvs_packed", None)
        if pk is None or pk[0] != dev:
            from .... import abi
            abi.require_device()          # fails loudly without the CUDA library / an sm_100 device
            with torch.no_grad():
                sd = {k: v for k, v in self.state_dict().items() if k.startswith(("backbone.", "neck.", "decoder.", "proj."))}
                pk = (dev, engine_cris.PackedCris(sd, image_size=self.img_size, input_resolution=self.backbone.visual.input_resolution,
                                                  rn_layers=self.backbone.visual.layers, dec_layers=self.decoder.num_layers,
                                                  dec_heads=self.decoder.nhead))
            object.__setattr__(self, "_tvs_packed", pk)
        return pk[1]

    def repack(self) -> None:
        """Call after loading new frozen weights into an already-used model."""
        if hasattr(self, "_tvs_packed"):
            object.__delattr__(self, "_tvs_packed")

    def forward(self, text_input, image_input):
        raise NotImplementedError(
            "plain CRIS.forward is the reference's end-to-end fine-tuning / zero-shot path (configs/model/e2e_cris.yaml), "
            "outside the dgrad-only prompt-tuning hot path; use tunevlseg_b200.models.core_models.coop.COOPCRIS")
