"""Where does the bf16 error of the logits come from?  Compares product vs oracle at the taps and isolates the decoder."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import clipseg as OC, learners as OL
from tests.helpers import SMALL, FULL, build_net, make_batch, oracle_head, oracle_state
from tunevlseg_b200 import abi, engine

for spec_name, spec, B in (("SMALL", SMALL, 3), ("FULL", FULL, 2)):
    for case in ("maple", "vpt"):
        w = OC.init_weights(spec, seed=7)
        net = build_net(case, spec, w, seed=11)
        st, head = oracle_state(case, net, spec), oracle_head(net)
        img, ids, am, mask = make_batch(spec, B, 9, 12)
        with torch.no_grad():
            taps_ref = OC.vision_tower_prompted(w, spec, st, img)
            cond_ref = OC.text_tower(w, spec, st if st.is_textual else None, ids, am)
            blend = "ratio" if case == "maple" else "add"
            ref = OC.decoder(w, spec, taps_ref, cond_ref, st.num_context, head, blend)
            net = net.cuda()
            pk = net.packed
            lr = net.context_learner
            taps = engine.VisionTowerFn.apply(lr.visual_stack(10), img.cuda(), pk, lr.prompt_depth)
            for i, (a, b) in enumerate(zip(taps, taps_ref)):
                e = (a.cpu() - b).abs()
                print(f"{spec_name} {case} tap{i}: max-abs {e.max():.4f}  rel-to-max {e.max()/b.abs().max():.5f}  rms-rel {e.pow(2).mean().sqrt()/b.pow(2).mean().sqrt():.5f}  |tap|max {b.abs().max():.2f}")
            cond = net._text_condition(ids.cuda(), am.cuda(), lr if st.is_textual else None)
            e = (cond.cpu() - cond_ref).abs()
            print(f"{spec_name} {case} cond: max-abs {e.max():.5f} rel {e.max()/cond_ref.abs().max():.5f}")
            wa, ba, r = net._head_params()
            bl = abi.BLEND_RATIO if case == "maple" else abi.BLEND_ADD
            full = engine.DecoderFn.apply(*taps, cond, wa, ba, r if case == "maple" else None, pk, bl, lr.num_context)
            iso = engine.DecoderFn.apply(*[t.cuda() for t in taps_ref], cond_ref.cuda(), wa, ba, r if case == "maple" else None, pk, bl, lr.num_context)
            print(f"{spec_name} {case} logits: full-chain err {(full.cpu()-ref).abs().max():.4f}  decoder-only (exact taps/cond) err {(iso.cpu()-ref).abs().max():.4f}  |logit|max {ref.abs().max():.2f}")
