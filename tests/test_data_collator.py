"""The collator mirror (tunevlseg_b200/data/data_collator.py) against fixtures produced by the reference's own
CustomDataCollatorWithPadding + transformers' tokenizer.pad (tests/golden/make_golden_collator.py).  No GPU needed."""
import json
import os
from types import SimpleNamespace

import pytest
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "collator_reference.json")


def _features(g):
    out = []
    for i, (s, ids) in enumerate(zip(g["prompts"], g["token_lists"])):
        out.append({"image": torch.full((3, 2, 2), float(i)), "mask": torch.full((1, 2, 2), float(i % 2)), "mask_shape": torch.tensor([10 + i, 12]),
                    "mask_name": f"m{i}.png", "prompt": s, "input_ids": list(ids), "attention_mask": [1] * len(ids)})
    return out


def test_collator_matches_reference_fixtures():
    from tunevlseg_b200.data import CustomDataCollatorWithPadding

    g = json.load(open(GOLDEN))
    for case in g["cases"]:
        tok = SimpleNamespace(pad_token_id=g["vocab"]["<pad>"], padding_side=case["padding_side"], pad_token_type_id=0)
        kw = case["kwargs"]
        c = CustomDataCollatorWithPadding(padding_keys=["input_ids", "attention_mask"], tokenizer=tok, padding=kw["padding"],
                                          max_length=kw.get("max_length"), pad_to_multiple_of=kw.get("pad_to_multiple_of"), return_tensors="pt")
        out = c(_features(g))
        assert list(out.keys()) == case["keys"], case["kwargs"]
        assert out["input_ids"].dtype == torch.int64 and out["input_ids"].tolist() == case["input_ids"], (case["padding_side"], kw)
        assert out["attention_mask"].tolist() == case["attention_mask"], (case["padding_side"], kw)
        assert list(out["image"].shape) == case["image_shape"] and out["mask_shape"].tolist() == case["mask_shape"]
        assert out["mask_name"] == case["mask_name"] and out["prompt"] == case["prompt"]
        assert torch.equal(out["image"][2], torch.full((3, 2, 2), 2.0))


def test_collator_errors_and_tensor_inputs():
    from tunevlseg_b200.data import CustomDataCollatorWithPadding

    tok = SimpleNamespace(pad_token_id=0, padding_side="right")
    with pytest.raises(ValueError, match="padding_keys"):
        CustomDataCollatorWithPadding(padding_keys=[], tokenizer=tok)
    c = CustomDataCollatorWithPadding(padding_keys=["input_ids", "attention_mask"], tokenizer=tok, padding=False)
    with pytest.raises(ValueError, match="padding"):
        c([{"input_ids": [1, 2], "attention_mask": [1, 1]}, {"input_ids": [1], "attention_mask": [1]}])
    c = CustomDataCollatorWithPadding(padding_keys=["input_ids", "attention_mask"], tokenizer=tok)
    out = c([{"input_ids": torch.tensor([5, 6, 7]), "attention_mask": torch.tensor([1, 1, 1]), "x": torch.ones(2)},
             {"input_ids": torch.tensor([8]), "attention_mask": torch.tensor([1]), "x": torch.zeros(2)}])
    assert out["input_ids"].tolist() == [[5, 6, 7], [8, 0, 0]] and out["attention_mask"].tolist() == [[1, 1, 1], [1, 0, 0]] and out["x"].shape == (2, 2)
    with pytest.raises(ValueError, match="padding token"):
        CustomDataCollatorWithPadding(padding_keys=["input_ids"], tokenizer=SimpleNamespace(pad_token_id=None))([{"input_ids": [1]}, {"input_ids": [1, 2]}])
