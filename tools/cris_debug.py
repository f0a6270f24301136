"""Stage-by-stage comparison of the CRIS CUDA path with the CPU oracle (development aid; run on a GPU box)."""
import sys

import torch

sys.path.insert(0, ".")
from oracle import cris as OCR  # noqa: E402
from tests.helpers import CRIS_FULL, CRIS_SMALL, build_cris_net, cris_oracle_head, cris_oracle_state, make_cris_batch  # noqa: E402
from tunevlseg_b200 import engine_cris as E  # noqa: E402


def rel(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()


def main(full=False, case="cocoop"):
    spec = CRIS_FULL if full else CRIS_SMALL
    w = OCR.init_weights(spec, seed=7)
    net = build_cris_net(case, spec, w, seed=31)
    st, head = cris_oracle_state(case, net), cris_oracle_head(net)
    B = 2
    img, ids, am, mask = make_cris_batch(spec, B, 8, 32)
    with torch.no_grad():
        ref_logits, parts = OCR.net_forward(w, spec, st, head, ids, am, img, return_parts=True)
    net = net.cuda()
    pk = net.packed
    with torch.no_grad():
        vis = E.encode_image(pk, img.cuda())
        for (v, h, wd), r, name in zip(vis, parts["vis"], ("v3", "v4", "v5")):
            got = v.view(B, h, wd, -1).permute(0, 3, 1, 2).cpu()
            print(f"{name}: rel err {rel(got, r):.5f}  |ref|max {r.abs().max():.3f}", flush=True)
        # downstream stages fed with the ORACLE's inputs, so each error is local
        vis_ref = [(t.permute(0, 2, 3, 1).reshape(-1, t.shape[1]).contiguous().cuda(), t.shape[2], t.shape[3]) for t in parts["vis"]]
        state, words = parts["state"].cuda(), parts["words"].cuda()
        pad = OCR.pad_mask_with_context(st, ids, am, spec.max_length)
        km = (~pad).to(torch.uint8).cuda().contiguous()
        fq, H, W = E.fpn(pk, vis_ref, state, B)
        fq_ref = OCR.fpn(w, parts["vis"], parts["state"])
        print(f"fpn: rel err {rel(fq.view(B, H, W, -1).permute(0, 3, 1, 2).cpu(), fq_ref):.5f}", flush=True)
        fq_in = fq_ref.permute(0, 2, 3, 1).reshape(-1, fq_ref.shape[1]).contiguous().cuda()
        dec = E.transformer_decoder(pk, fq_in, words, km, B, H, W)
        print(f"decoder: rel err {rel(dec.view(B, H, W, -1).permute(0, 3, 1, 2).cpu(), parts['fq']):.5f}", flush=True)
        dec_in = parts["fq"].permute(0, 2, 3, 1).reshape(-1, parts["fq"].shape[1]).contiguous().cuda()
        pred = E.projector(pk, dec_in, state, B, H, W)
        print(f"projector: rel err {rel(pred.cpu(), parts['pred']):.5f}  |ref|max {parts['pred'].abs().max():.3f}", flush=True)
        logits = net(text_input={"input_ids": ids.cuda(), "attention_mask": am.cuda()}, image_input=img.cuda())
        print(f"logits: max-abs err {(logits.cpu() - ref_logits).abs().max():.5f}  |ref|max {ref_logits.abs().max():.3f}", flush=True)


if __name__ == "__main__":
    main(full="--full" in sys.argv, case=next((a for a in sys.argv[1:] if not a.startswith("--")), "cocoop"))
