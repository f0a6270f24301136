"""CRIS + textual prompt learner (CoOp / CoCoOp): same class, constructor arguments, call signature and error
behaviour as /root/reference/src/models/core_models/coop/coop_cris.py:20-242 - the forward is the B200 engine
(``tunevlseg_b200/engine_cris.py``): forward-only bf16 CLIP-RN50, prompted text encoder, then neck / decoder /
projector / tail with dgrad back to the prompts.

    net(text_input={"input_ids"[, "attention_mask"]}, image_input=(B,3,H,W) f32) -> logits (B,1,H,W) f32
"""
from __future__ import annotations

import torch
from torch import nn

from .... import abi, engine_cris
from ...components.cris_model import CRIS
from .context_learner import CoCoOpContextLearner


class COOPCRIS(CRIS):
    def __init__(self, model_cfg, context_learner, freeze_all: bool = True, no_freeze_last_layer: bool = False,
                 use_new_last_layer: bool = False, new_last_layer_kernel_size=5, residual_ratio: float = 0.5) -> None:
        super().__init__(**model_cfg)
        self.assign_model_learnability(freeze_all, no_freeze_last_layer, use_new_last_layer, new_last_layer_kernel_size, residual_ratio)
        self.context_learner = context_learner(
            max_network_depth=self.backbone.transformer.layers,
            visual_dim=self.backbone.visual.output_dim,
            context_dim=self.word_dim,
            embedding_layer=self.backbone.token_embedding,
        )
        # CoCoOp conditions on the spatially pooled C5 map (coop_cris.py:49-56, :90-93)
        self.image_features_pooler_or_identity = (
            self._pool_4D_tensor if isinstance(self.context_learner, CoCoOpContextLearner) else nn.Identity())

    def assign_model_learnability(self, freeze_all, no_freeze_last_layer, use_new_last_layer, new_last_layer_kernel_size, residual_ratio):
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
            mid = 64
            self.additive_decoder_layer = nn.Sequential(
                nn.Conv2d(self.proj.in_dim * 2, mid, 1, bias=False),
                nn.Upsample(size=self.img_size, mode="bilinear"),
                nn.Conv2d(mid, 1, kernel_size=ks, padding="same", padding_mode="replicate"),
            )
            self.residual_ratio = nn.Parameter(torch.tensor(residual_ratio))
        elif no_freeze_last_layer:
            raise NotImplementedError("no_freeze_last_layer=True needs weight gradients of proj.txt / proj.vis[-1]: outside "
                                      "the dgrad-only path (SURVEY.md section 8f, rank 4)")

    def train(self, mode: bool = True):
        # frozen parts always run eval-mode semantics (BatchNorm running stats, no dropout)
        super().train(mode)
        for m in (self.backbone, self.neck, self.decoder, self.proj):
            m.eval()
        return self

    @staticmethod
    def _pool_4D_tensor(x: torch.Tensor) -> torch.Tensor:
        return x.mean((2, 3))

    def get_pad_mask(self, input_ids, attention_mask):
        pad_mask = super().get_pad_mask(input_ids, attention_mask)
        return self.context_learner.update_pad_mask_for_context(pad_mask=pad_mask, max_length=self.max_length)

    def forward(self, text_input, image_input):
        input_ids = text_input["input_ids"]
        attention_mask = text_input.get("attention_mask")
        abi.check_cuda_input(image_input)
        if len(input_ids) != image_input.shape[0]:
            raise ValueError("Make sure to pass as many prompt texts as there are query images")
        pk = self.packed
        learner = self.context_learner
        n = learner.num_context
        B = image_input.shape[0]
        pad_mask = self.get_pad_mask(input_ids, attention_mask)                       # (B, S) True = padding
        key_mask = (~pad_mask).to(torch.uint8).contiguous()

        vis = engine_cris.encode_image(pk, image_input.to(torch.float32))             # frozen, forward only
        feats = None
        if isinstance(learner, CoCoOpContextLearner):
            v5, h5, w5 = vis[-1]
            feats = v5.view(B, h5 * w5, -1).mean(1)                                    # = vis[-1].mean((2, 3))

        emb = self.backbone.token_embedding(input_ids)
        emb = learner(input_embeddings=emb, max_length=self.max_length, image_features=feats)
        S = emb.shape[1]
        emb = emb + self.backbone.positional_embedding[:S]
        depth = min(learner.prompt_depth, len(pk.t_layers))
        ctx_over = torch.stack([learner.get_textual_context(image_features=feats, index=i) for i in range(depth)])
        pool = torch.clamp(input_ids.argmax(dim=-1) + n, max=self.max_length - 1)
        words, state = engine_cris.CrisTextFn.apply(emb, ctx_over, key_mask, pool, pk, n)

        if self.additive_decoder_layer is None:
            w0 = w2 = b2 = ratio = None
        else:
            w0, w2, b2 = self.additive_decoder_layer[0].weight, self.additive_decoder_layer[2].weight, self.additive_decoder_layer[2].bias
            ratio = self.residual_ratio
        return engine_cris.head_forward(pk, vis, words, state, key_mask, w0, w2, b2, ratio)
