"""Batch collation of the reference's loaders ("tokenise + pad").

Mirror of ``src.data.components.data_collator.CustomDataCollatorWithPadding`` (data_collator.py:8-34; configured in
configs/experiment/coop/clipseg.yaml:133-143): the keys in ``padding_keys`` (``input_ids``, ``attention_mask`` - python lists of
different lengths, as ``ImageTextMaskDataset.__getitem__`` returns them from ``self.tokenizer(prompt)``,
image_text_mask_dataset.py:86-99) are padded the way ``transformers.DataCollatorWithPadding`` -> ``tokenizer.pad`` does it, every
other key goes through ``torch.utils.data.default_collate``.  Nothing here needs transformers at run time: the padding rules
(pad id / attention-mask 0, ``padding_side``, ``padding`` in {True, "longest", "max_length", False, "do_not_pad"}, ``max_length``,
``pad_to_multiple_of``) are restated, and pinned by fixtures generated with the reference's own class
(tests/golden/make_golden_collator.py).  Image / mask tensors that already live on the GPU (``GpuTrainTransforms``) are stacked
there; the small integer text tensors are built on the host and follow with ``.to(device)``.
"""
from __future__ import annotations

from collections.abc import Iterable
from typing import Any

import torch
from torch.utils.data import default_collate

_PAD_VALUE = {"attention_mask": 0, "token_type_ids": None, "special_tokens_mask": 1}      # token_type_ids: tokenizer.pad_token_type_id


class CustomDataCollatorWithPadding:
    def __init__(self, padding_keys: Iterable[str], tokenizer, padding: bool | str = True, max_length: int | None = None,
                 pad_to_multiple_of: int | None = None, return_tensors: str = "pt") -> None:
        self.padding_keys = set(padding_keys)
        if not self.padding_keys:
            msg = "`padding_keys` should not be empty."
            raise ValueError(msg)
        if return_tensors != "pt":
            raise ValueError(f"only return_tensors='pt' is supported, got {return_tensors!r}")
        self.tokenizer, self.padding, self.max_length, self.pad_to_multiple_of = tokenizer, padding, max_length, pad_to_multiple_of
        self.return_tensors = return_tensors

    # ---- tokenizer.pad, restated (transformers tokenization_utils_base.py: PreTrainedTokenizerBase.pad / _pad) ----
    def _target_length(self, lengths: list[int]) -> int | None:
        strategy = self.padding
        if strategy is True or strategy == "longest":
            n = max(lengths)
        elif strategy == "max_length":
            n = self.max_length if self.max_length is not None else getattr(self.tokenizer, "model_max_length", None)
            if n is None:
                raise ValueError("padding='max_length' needs max_length")
        elif strategy is False or strategy == "do_not_pad":
            return None
        else:
            raise ValueError(f"unknown padding strategy {strategy!r}")
        m = self.pad_to_multiple_of
        if m is not None and n % m != 0:
            n = (n // m + 1) * m
        return n

    def _pad(self, features: list[dict[str, Any]]) -> dict[str, torch.Tensor]:
        main = "input_ids" if "input_ids" in features[0] else next(iter(features[0]))
        rows = {k: [list(map(int, f[k].tolist() if torch.is_tensor(f[k]) else f[k])) for f in features] for k in features[0]}
        target = self._target_length([len(r) for r in rows[main]])
        side = getattr(self.tokenizer, "padding_side", "right")
        out = {}
        for key, seqs in rows.items():
            if key == main or key not in _PAD_VALUE:
                fill = self.tokenizer.pad_token_id if key == main else 0
            else:
                fill = _PAD_VALUE[key] if _PAD_VALUE[key] is not None else getattr(self.tokenizer, "pad_token_type_id", 0)
            if fill is None:
                raise ValueError("Asking to pad but the tokenizer does not have a padding token.")
            if target is not None:
                seqs = [(s + [fill] * (target - len(s))) if side == "right" else ([fill] * (target - len(s)) + s) if len(s) < target else s for s in seqs]
            if len({len(s) for s in seqs}) != 1:
                raise ValueError("Unable to create tensor, you should probably activate padding with 'padding=True' to have batched tensors "
                                 "with the same length.")
            out[key] = torch.tensor(seqs, dtype=torch.int64)
        return out

    def __call__(self, features: list[dict[str, Any]]) -> dict[str, Any]:
        features_to_pad = [{key: value for key, value in example.items() if key in self.padding_keys} for example in features]
        padded_features = self._pad(features_to_pad)
        # the default pytorch collate function for the leftover items (tensors are stacked on the device they live on)
        leftover_features = [{key: value for key, value in example.items() if key not in padded_features} for example in features]
        collate_features = default_collate(leftover_features)
        return {**collate_features, **padded_features}
