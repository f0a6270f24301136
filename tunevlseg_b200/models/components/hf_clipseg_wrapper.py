"""Weights container with the reference's constructor (src/models/components/hf_clipseg_wrapper.py:15-62).

The HuggingFace ``CLIPSegForImageSegmentation`` object is kept as the *container* of the frozen weights, so that
``from_pretrained`` paths, ``state_dict`` keys (``model.clip.*``, ``model.decoder.*``) and Lightning checkpoints stay
drop-in.  Its ``forward`` is never called on the prompt-tuning path: the B200 engine reads the parameters, packs
them once (``engine.packed_for``) and runs its own kernels.
"""
from __future__ import annotations

from torch import nn


class HFCLIPSegWrapper(nn.Module):
    def __init__(self, pretrained_model_name_or_path=None, freeze_encoder: bool = False, freeze_decoder: bool = False,
                 *args, **kwargs) -> None:
        super().__init__()
        model = self.get_pretrained_model(pretrained_model_name_or_path, *args, **kwargs)
        model.clip.requires_grad_(not freeze_encoder)
        model.decoder.requires_grad_(not freeze_decoder)
        self.model = model

    @staticmethod
    def get_pretrained_model(pretrained_model_name_or_path, *args, **kwargs):
        from transformers import CLIPSegForImageSegmentation

        if isinstance(pretrained_model_name_or_path, CLIPSegForImageSegmentation):
            return pretrained_model_name_or_path          # tests / benchmarks hand over a random-init model
        try:
            model = CLIPSegForImageSegmentation.from_pretrained(pretrained_model_name_or_path, *args, **kwargs)
        except TypeError:
            print("Unncessary arguments passed to `CLIPSegForImageSegmentation.from_pretrained`")
            model = CLIPSegForImageSegmentation.from_pretrained(pretrained_model_name_or_path)
        if not isinstance(model, CLIPSegForImageSegmentation):
            raise ValueError(f"Expected `CLIPSegForImageSegmentation` from {pretrained_model_name_or_path}, got {type(model)}")
        return model

    def forward(self, text_input, image_input):
        raise NotImplementedError(
            "HFCLIPSegWrapper.forward is the reference's end-to-end fine-tuning / zero-shot path (configs/model/"
            "e2e_clipseg.yaml), which needs weight gradients and is outside the dgrad-only prompt-tuning hot path "
            "this package implements (SURVEY.md section 8f, rank 4).")
