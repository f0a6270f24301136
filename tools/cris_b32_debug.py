"""GPU debug: CRIS CoCoOp full geometry at B=32 - run-to-run determinism with a poisoned allocator (uninitialised reads?)"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cris as OCR
from tests.helpers import CRIS_FULL, build_cris_net, cris_oracle_head, cris_oracle_state, make_cris_batch

spec, B, L = CRIS_FULL, int(os.environ.get("DBG_B", "32")), 8
weights = OCR.init_weights(spec, seed=5)
net = build_cris_net("cocoop", spec, weights, seed=2)
img, ids, am, mask = make_cris_batch(spec, B, L, 9)
net = net.cuda()
ti = {"input_ids": ids.cuda(), "attention_mask": am.cuda()}
im = img.cuda()


def poison(val):
    blocks = []
    try:
        for mb in (2048, 1024, 512, 256, 128, 64, 32, 16, 8, 4, 2, 1):
            for _ in range(12):
                blocks.append(torch.full((mb * 1024 * 256,), val, device="cuda"))
    except torch.OutOfMemoryError:
        pass
    n = sum(b.numel() for b in blocks) * 4 / 2**30
    del blocks
    return n


if os.environ.get("DBG_ONCE") == "1":
    with torch.no_grad():
        out = net(text_input=ti, image_input=im)
    torch.cuda.synchronize()
    print("done", float(out.abs().max()))
    sys.exit(0)
with torch.no_grad():
    a = net(text_input=ti, image_input=im).clone()
    torch.cuda.synchronize()
    print("poisoned GiB:", poison(float("nan")))
    b = net(text_input=ti, image_input=im).clone()
    torch.cuda.synchronize()
    print("poisoned GiB:", poison(1e4))
    c = net(text_input=ti, image_input=im).clone()
    torch.cuda.synchronize()
print("run1 vs run2 (NaN-poisoned):", (a - b).abs().max().item(), "nan in run2:", bool(torch.isnan(b).any()))
print("run1 vs run3 (1e4-poisoned):", (a - c).abs().max().item())
per = [(round((a[i] - c[i]).abs().max().item(), 5)) for i in range(B)]
print("per-sample diff run1-run3:", per)
