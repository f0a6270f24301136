"""CLIPSeg + prompt learner nets: same classes, constructor arguments, call signature and error behaviour as
/root/reference/src/models/core_models/coop/{base_clipseg,coop_clipseg,vpt_clipseg,base_multimodal_clipseg,
maple_clipseg,shared_attn_learner_clipseg,shared_separate_learner_clipseg}.py - but the forward is three calls into
the B200 engine (vision tower, text tower, decoder) instead of ~400 ATen launches.

    net(text_input={"input_ids", "attention_mask"}, image_input=(B,3,H,W) f32) -> logits (B,1,H,W) f32
"""
from __future__ import annotations

from abc import ABC, abstractmethod

import os

import torch
from torch import nn

from .... import abi, engine
from ...components.hf_clipseg_wrapper import HFCLIPSegWrapper
from .context_learner import BaseVisualLearner


class BaseCLIPSeg(HFCLIPSegWrapper, ABC):
    """base_clipseg.py:24-199."""

    def __init__(self, model_cfg, freeze_all: bool = True, no_freeze_last_layer: bool = False,
                 use_new_last_layer: bool = False, new_last_layer_kernel_size=5, residual_ratio: float = 0.5) -> None:
        super().__init__(**model_cfg)
        self.assign_model_learnability(freeze_all, no_freeze_last_layer, use_new_last_layer, new_last_layer_kernel_size,
                                       residual_ratio)

    def assign_model_learnability(self, freeze_all, no_freeze_last_layer, use_new_last_layer, new_last_layer_kernel_size,
                                  residual_ratio):
        if not freeze_all:
            raise NotImplementedError("freeze_all=False trains the backbone (weight gradients): outside the dgrad-only "
                                      "prompt-tuning path (SURVEY.md section 8f, rank 4)")
        self.eval()
        self.requires_grad_(False)
        self.additive_decoder_layer = None
        if use_new_last_layer:
            ks = new_last_layer_kernel_size
            if not isinstance(ks, int):
                if ks[0] != ks[1]:
                    raise NotImplementedError("the fused head kernel supports square kernels only")
                ks = ks[0]
            if ks % 2 == 0 or ks > 7:
                raise NotImplementedError("the fused head kernel supports odd kernel sizes <= 7")
            # same module structure as the reference so that state_dict keys match (additive_decoder_layer.1.weight)
            self.additive_decoder_layer = nn.Sequential(
                nn.Upsample(scale_factor=float(self.model.config.vision_config.patch_size), mode="bilinear"),
                nn.Conv2d(self.model.config.reduce_dim, 1, kernel_size=ks, padding="same", padding_mode="replicate"),
            )
            self.residual_ratio = nn.Parameter(torch.tensor(residual_ratio))
        elif no_freeze_last_layer:
            # base_clipseg.py:73-80: the decoder's last layer (the transposed convolution) trains with the prompts.  Its weight
            # gradient is one small GEMM, featT [Dr, B*G*G] x dlogits-per-patchT [P*P, B*G*G] (engine.DecoderFn.backward).
            trans_conv = self.model.decoder.transposed_convolution
            if isinstance(trans_conv, nn.Sequential):
                raise NotImplementedError("use_complex_transposed_convolution=True (refined decoder) is outside the supported path")
            trans_conv.requires_grad_(True)

    # ---- shared pieces ------------------------------------------------------------------------------------------
    @property
    def packed(self) -> engine.PackedClipSeg:
        return engine.packed_for(self.model)

    def train(self, mode: bool = True):
        # frozen towers always run eval-mode semantics; only the learner / head follow `mode`
        super().train(mode)
        self.model.eval()
        return self

    def _check_inputs(self, input_ids, pixel_values):
        if pixel_values is None:
            raise ValueError("You have to specify pixel_values to use `CLIPSegForImageSegmentation`")
        if input_ids is None:
            raise ValueError("Invalid conditional, should be either provided as `input_ids` or `conditional_pixel_values`")
        if len(input_ids) != pixel_values.shape[0]:
            raise ValueError("Make sure to pass as many prompt texts as there are query images")
        abi.check_cuda_input(pixel_values)      # TvsError for host tensors: there is no CPU path

    def _text_condition(self, input_ids, attention_mask, learner, image_features=None, defer=None):
        """Conditional embedding (B, projection_dim).  ``learner`` None = stock HF text path (no prompts).
        ``defer``: see engine.TextTowerFn.forward (the tower's kernels are enqueued later by the closure appended to it)."""
        pk = self.packed
        tm = self.model.clip.text_model
        max_len = self.model.config.text_config.max_position_embeddings
        input_ids = input_ids.view(-1, input_ids.shape[-1])
        emb = tm.embeddings.token_embedding(input_ids)
        n = 0
        deep = None
        if learner is not None:
            n = learner.num_context
            emb = learner(input_embeddings=emb, max_length=max_len, image_features=image_features)
            if attention_mask is not None:
                attention_mask = learner.update_attention_mask_for_context(attention_mask, max_len)
            deep = learner.textual_deep_stack(len(pk.t_layers), image_features=image_features)
        S = emb.shape[1]
        emb = emb + tm.embeddings.position_embedding.weight[:S]
        ids = input_ids.to(torch.int)
        pre = ids if pk.eos_token_id == 2 else (ids == pk.eos_token_id).int()
        pool = pre.argmax(dim=-1) + n
        if learner is not None:
            pool = torch.clamp(pool, max=max_len - 1)
        key_mask = None if attention_mask is None else (attention_mask != 0).to(torch.uint8)
        return engine.TextTowerFn.apply(emb, deep, key_mask, pool, pk, n, defer)

    def _text_stream(self) -> torch.cuda.Stream:
        if os.environ.get("TVS_TEXT_STREAM", "1") == "0":      # A/B switch: text tower in line with the vision tower
            return torch.cuda.current_stream()
        st = getattr(self, "_tvs_text_stream", None)
        if st is None or st.device != torch.cuda.current_stream().device:
            # TVS_TEXT_PRIORITY=1 (experiment switch) gives the text tower's blocks precedence whenever both streams have
            # blocks pending; measured 10.24 vs 10.04 ms per step: the vision chain is the critical one, default priority stays
            st = torch.cuda.Stream(priority=-1 if os.environ.get("TVS_TEXT_PRIORITY", "0") == "1" else 0)
            object.__setattr__(self, "_tvs_text_stream", st)
        return st

    def _head_params(self):
        if self.additive_decoder_layer is None:
            return None, None, None
        conv = self.additive_decoder_layer[1]
        return conv.weight, conv.bias, self.residual_ratio

    def _tconv_params(self):
        """(weight, bias) of the transposed convolution when it trains (no_freeze_last_layer), else (None, None): the
        decoder then reads the live parameters instead of the copies packed once per device."""
        tc = self.model.decoder.transposed_convolution
        if isinstance(tc, nn.Sequential) or not tc.weight.requires_grad:
            return None, None
        return tc.weight, tc.bias

    @abstractmethod
    def model_forward(self, input_ids=None, pixel_values=None, attention_mask=None, **kwargs) -> torch.Tensor: ...

    def forward(self, text_input, image_input):
        B, _, H, W = image_input.shape
        logits = self.model_forward(**text_input, pixel_values=image_input)
        return logits.view(B, 1, H, W)


class COOPCLIPSeg(BaseCLIPSeg):
    """CoOp / CoCoOp: textual prompts only.  Stock vision tower (all layers, forward only - nothing upstream needs a
    gradient), pooled image feature feeds CoCoOp's meta-net, stock decoder: the additive layer and residual_ratio
    exist as parameters but are never used, as in the reference (coop_clipseg.py:418-484)."""

    def __init__(self, context_learner, *args, **kwargs) -> None:
        super().__init__(*args, **kwargs)
        cfg = self.model.config
        self.context_learner = context_learner(
            visual_dim=cfg.projection_dim,
            max_network_depth=min(cfg.text_config.num_hidden_layers, cfg.vision_config.num_hidden_layers),
            context_dim=cfg.text_config.hidden_size,
            embedding_layer=self.model.clip.text_model.embeddings.token_embedding,
        )

    def model_forward(self, input_ids=None, pixel_values=None, attention_mask=None, **kwargs):
        self._check_inputs(input_ids, pixel_values)
        pk = self.packed
        taps, feats = engine.vision_tower_stock(pk, pixel_values)
        cond = self._text_condition(input_ids, attention_mask, self.context_learner, image_features=feats)
        return engine.DecoderFn.apply(*taps, cond, None, None, None, pk, abi.BLEND_NONE, 0, *self._tconv_params())


class VPTCLIPSeg(BaseCLIPSeg):
    """VPT-deep: visual prompts only.  Stock (no-grad) text path; ``logits += additive(output)`` with no
    residual_ratio (vpt_clipseg.py:301-302, :321-395)."""

    def __init__(self, context_learner, *args, **kwargs) -> None:
        super().__init__(*args, **kwargs)
        cfg = self.model.config
        self.context_learner = context_learner(
            max_network_depth=min(cfg.text_config.num_hidden_layers, cfg.vision_config.num_hidden_layers),
            context_dim=cfg.vision_config.hidden_size,
        )

    def model_forward(self, input_ids=None, pixel_values=None, attention_mask=None, **kwargs):
        self._check_inputs(input_ids, pixel_values)
        pk = self.packed
        lr = self.context_learner
        with torch.no_grad():
            cond = self._text_condition(input_ids, attention_mask, None)
        n_run = max(pk.extract_layers) + 1
        taps = engine.VisionTowerFn.apply(lr.visual_stack(n_run), pixel_values, pk, lr.prompt_depth)
        w, b, _ = self._head_params()
        blend = abi.BLEND_NONE if w is None else abi.BLEND_ADD
        return engine.DecoderFn.apply(*taps, cond, w, b, None, pk, blend, lr.num_context, *self._tconv_params())


class BaseMultimodalCLIPSeg(BaseCLIPSeg):
    """MaPLe / shared-attn / shared-separate: vision first (so the shared learner's cache is filled by the visual
    branch, base_multimodal_clipseg.py:577-581), then text, then the decoder with the ratio blend."""

    def model_forward(self, input_ids=None, pixel_values=None, attention_mask=None, **kwargs):
        self._check_inputs(input_ids, pixel_values)
        pk = self.packed
        lr = self.context_learner
        for cache in ("_computed_textual_context_cache", "_computed_visual_context_cache"):
            if hasattr(lr, cache):
                getattr(lr, cache).clear()
        n_run = max(pk.extract_layers) + 1
        vis_ctx = lr.visual_stack(n_run)
        # The two towers are independent until the decoder.  The text tower is ~200 microsecond-scale kernels on
        # B*S ~ 400 rows: it runs on a side stream so that it hides behind the vision tower (autograd replays each
        # node's backward on its forward stream, so the backward overlaps too; under CUDA-graph capture the two
        # streams become parallel branches of the graph).
        main = torch.cuda.current_stream()
        side = self._text_stream()
        side.wait_stream(main)            # after the visual branch of a shared learner has filled its cache
        # Which branch is enqueued / captured first matters (same box, ms per step): text first (default) 9.18-9.19; vision forward
        # first with the text node still created first ("2", deferred text kernels) 9.24-9.26; vision first in forward AND hence the
        # text backward first ("1") 9.90-10.02 - the branch captured first is served first, and the text branch is the one whose
        # ~190 launch latencies must start early to stay hidden.  The switches stay for A/B only.
        mode = os.environ.get("TVS_VISION_FIRST", "0")
        if mode == "1":      # experiment: enqueue / capture the critical branch first (the text node then has the higher sequence
            taps = engine.VisionTowerFn.apply(vis_ctx, pixel_values, pk, lr.prompt_depth)      # number: ITS backward is enqueued first)
            with torch.cuda.stream(side):
                cond = self._text_condition(input_ids, attention_mask, lr)
        elif mode == "2":    # text node created first (so the vision backward is still enqueued first), its kernels enqueued after
            deferred: list = []                                                                    # the vision tower's
            with torch.cuda.stream(side):
                cond = self._text_condition(input_ids, attention_mask, lr, defer=deferred)
            taps = engine.VisionTowerFn.apply(vis_ctx, pixel_values, pk, lr.prompt_depth)
            with torch.cuda.stream(side):
                for run in deferred:
                    run()
        else:
            with torch.cuda.stream(side):
                cond = self._text_condition(input_ids, attention_mask, lr)
            taps = engine.VisionTowerFn.apply(vis_ctx, pixel_values, pk, lr.prompt_depth)
        main.wait_stream(side)
        cond.record_stream(main)
        w, b, r = self._head_params()
        blend = abi.BLEND_NONE if w is None else abi.BLEND_RATIO
        n_strip = lr.num_context if isinstance(lr, BaseVisualLearner) else 0
        return engine.DecoderFn.apply(*taps, cond, w, b, r, pk, blend, n_strip, *self._tconv_params())


class MapleCLIPSeg(BaseMultimodalCLIPSeg):
    def __init__(self, context_learner, *args, **kwargs) -> None:
        super().__init__(*args, **kwargs)
        cfg = self.model.config
        self.context_learner = context_learner(
            visual_dim=cfg.vision_config.hidden_size,
            max_network_depth=min(cfg.text_config.num_hidden_layers, cfg.vision_config.num_hidden_layers),
            context_dim=cfg.text_config.hidden_size,
            embedding_layer=self.model.clip.text_model.embeddings.token_embedding,
        )


class SharedSeparateCLIPSeg(BaseMultimodalCLIPSeg):
    def __init__(self, context_learner, *args, **kwargs) -> None:
        super().__init__(*args, **kwargs)
        cfg = self.model.config
        self.context_learner = context_learner(
            textual_dim=cfg.text_config.hidden_size,
            visual_dim=cfg.vision_config.hidden_size,
            max_network_depth=min(cfg.text_config.num_hidden_layers, cfg.vision_config.num_hidden_layers),
            context_dim=cfg.text_config.hidden_size,
            embedding_layer=self.model.clip.text_model.embeddings.token_embedding,
        )


class SharedAttnCLIPSeg(BaseMultimodalCLIPSeg):
    def __init__(self, context_learner, *args, **kwargs) -> None:
        super().__init__(*args, **kwargs)
        cfg = self.model.config
        self.context_learner = context_learner(
            textual_dim=cfg.text_config.hidden_size,
            visual_dim=cfg.vision_config.hidden_size,
            max_network_depth=min(cfg.text_config.num_hidden_layers, cfg.vision_config.num_hidden_layers),
            context_dim=cfg.text_config.hidden_size,
        )
