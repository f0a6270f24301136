"""GPU debug: find the first kernel whose result is not reproducible run-to-run.  Every abi op is wrapped; after each call all
of its tensor arguments are check-summed (host sync - intra-kernel races still show).  The forward is repeated R times on
identical inputs and the per-op checksum sequences are compared with run 0.
    python tools/race_bisect.py [cris|maple|vpt] [B] [R]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tunevlseg_b200 import abi

which = sys.argv[1] if len(sys.argv) > 1 else "cris"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
R = int(sys.argv[3]) if len(sys.argv) > 3 else 6
os.environ["TVS_TEXT_STREAM"] = "0"

log = []
NAMES = [n for n in dir(abi) if callable(getattr(abi, n)) and not n.startswith("_") and n not in (
    "load", "lib_path", "require_device", "check_cuda_input", "launch_count", "set_profiler", "gemm_last_variant", "dicebce_scratch_bytes",
    "TvsError", "GemmArgs", "Structure", "POINTER", "byref", "c_char_p", "c_float", "c_int32", "c_int64", "c_void_p")]


def wrap(name, fn):
    def op(*a, **k):
        out = fn(*a, **k)
        sums = []
        for i, t in enumerate(list(a) + list(k.values())):
            if torch.is_tensor(t) and t.is_cuda and t.numel() > 0:
                x = t.detach()
                x = x.float() if x.is_floating_point() else x.double()
                sums.append((i, tuple(t.shape), str(t.dtype), float(torch.nan_to_num(x).sum()), float(torch.nan_to_num(x).abs().sum())))
        extra = abi.gemm_last_variant() if name == "gemm" else ""
        log.append((name, extra, sums))
        return out
    return op


for n in NAMES:
    f = getattr(abi, n)
    if callable(f) and getattr(f, "__module__", "") == abi.__name__ or n in ("gemm", "layernorm_fwd", "attn_fwd"):
        setattr(abi, n, wrap(n, f))

if which == "cris":
    from oracle import cris as OCR
    from tests.helpers import CRIS_FULL, build_cris_net, make_cris_batch
    net = build_cris_net("cocoop", CRIS_FULL, OCR.init_weights(CRIS_FULL, seed=5), seed=2).cuda()
    img, ids, am, mask = make_cris_batch(CRIS_FULL, B, 8, 9)
else:
    from oracle import clipseg as OC
    from tests.helpers import FULL, build_net, make_batch
    net = build_net(which, FULL, OC.init_weights(FULL, seed=7), seed=3).cuda()
    img, ids, am, mask = make_batch(FULL, B, 8, 4)
ti = {"input_ids": ids.cuda(), "attention_mask": am.cuda()}
im = img.cuda()
runs = []
outs = []
with torch.no_grad():
    for r in range(R):
        log.clear()
        o = net(text_input=ti, image_input=im)
        torch.cuda.synchronize()
        runs.append(list(log))
        outs.append(o.clone())
print(f"{which} B={B}: {len(runs[0])} ops per forward")
for r in range(1, R):
    d = (outs[r] - outs[0]).abs().max().item()
    first = None
    for i, (a, b) in enumerate(zip(runs[0], runs[r])):
        if a != b:
            first = i
            break
    print(f"run {r}: logits max diff vs run 0 = {d:.5f}; first differing op: {first}")
    if first is not None:
        for j in range(max(0, first - 1), min(len(runs[0]), first + 2)):
            print("   op", j, runs[0][j][0], runs[0][j][1])
            for sa, sb in zip(runs[0][j][2], runs[r][j][2]):
                flag = "  <-- differs" if sa != sb else ""
                print("      arg", sa[0], sa[1], sa[2], f"sum {sa[3]:.6g} / {sb[3]:.6g}", flag)
