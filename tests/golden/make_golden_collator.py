"""Golden vectors for batch collation, generated with the REFERENCE'S OWN src/data/components/data_collator.py (loaded from
/root/reference by file path: the package __init__ imports Lightning, which is absent) and transformers' DataCollatorWithPadding
over an in-memory word-level tokenizer (no tokenizer files exist offline).  python tests/golden/make_golden_collator.py"""
import importlib.util
import json
import os

import torch
from tokenizers import Tokenizer, models, pre_tokenizers
from transformers import PreTrainedTokenizerFast

spec = importlib.util.spec_from_file_location("ref_data_collator", "/root/reference/src/data/components/data_collator.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

VOCAB = {"<pad>": 0, "<bos>": 1, "<eos>": 2, "a": 3, "photo": 4, "of": 5, "cat": 6, "dog": 7, "the": 8, "left": 9, "kidney": 10, "[UNK]": 11, ".": 12}
PROMPTS = ["a photo of cat .", "the left kidney of the dog of the cat .", "dog", "a photo of the left kidney of a dog of a cat of the dog ."]


def tokenizer(padding_side="right"):
    tok = Tokenizer(models.WordLevel(VOCAB, unk_token="[UNK]"))
    tok.pre_tokenizer = pre_tokenizers.Whitespace()
    return PreTrainedTokenizerFast(tokenizer_object=tok, pad_token="<pad>", bos_token="<bos>", eos_token="<eos>", padding_side=padding_side)


def features(t):
    out = []
    for i, s in enumerate(PROMPTS):
        e = t(s)
        out.append({"image": torch.full((3, 2, 2), float(i)), "mask": torch.full((1, 2, 2), float(i % 2)), "mask_shape": torch.tensor([10 + i, 12]),
                    "mask_name": f"m{i}.png", "prompt": s, "input_ids": e["input_ids"], "attention_mask": e["attention_mask"]})
    return out


cases = []
for side in ("right", "left"):
    for kw in (dict(padding=True), dict(padding=True, pad_to_multiple_of=8), dict(padding="max_length", max_length=24), dict(padding="longest", pad_to_multiple_of=5)):
        t = tokenizer(side)
        c = ref.CustomDataCollatorWithPadding(padding_keys=["input_ids", "attention_mask"], tokenizer=t, max_length=kw.get("max_length"),
                                              padding=kw["padding"], pad_to_multiple_of=kw.get("pad_to_multiple_of"), return_tensors="pt")
        out = c(features(t))
        cases.append({"padding_side": side, "kwargs": kw, "keys": list(out.keys()), "input_ids": out["input_ids"].tolist(),
                      "attention_mask": out["attention_mask"].tolist(), "image_shape": list(out["image"].shape), "mask_shape": out["mask_shape"].tolist(),
                      "mask_name": out["mask_name"], "prompt": out["prompt"]})
t = tokenizer()
p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "collator_reference.json")
json.dump({"vocab": VOCAB, "prompts": PROMPTS, "token_lists": [t(s)["input_ids"] for s in PROMPTS], "cases": cases}, open(p, "w"), indent=1)
print(p, len(cases), "cases")
