"""TEST / BENCH infrastructure (never imported by the product): run the UNMODIFIED reference classes on this image.

``install_shim()`` - the reference targets the transformers 4.3x API and this image ships 5.5: the two
4-D mask helpers are re-exported into ``modeling_clipseg`` and the encoder / decoder layer ``forward`` accepts the 4.x
positional ``(hidden, attention_mask, causal_attention_mask, output_attentions=)`` call and returns a 1-tuple.  Nothing of
the reference is edited; only the installed ``transformers`` classes it calls are adapted.

``reference_path()`` - where the reference package ``src`` can be imported from: ``/root/reference`` in the build container,
else ``oracle/_ref`` (byte-compiled there by ``oracle/build_ref.py``; git-ignored, travels to the GPU box), else None.

Users: tests/golden/make_golden.py (fixture generation) and bench.py's ``--impl reference`` arm.
"""
from __future__ import annotations

import os

HERE = os.path.dirname(os.path.abspath(__file__))


def reference_path() -> str | None:
    if os.environ.get("TVS_REF_FORCE_BUILT") != "1" and os.path.isdir("/root/reference/src/models/core_models/coop"):
        return "/root/reference"
    built = os.path.join(HERE, "_ref")
    if os.path.exists(os.path.join(built, "src", "models", "core_models", "coop", "__init__.pyc")):
        return built
    archive = os.path.join(built, "reference_src.zip")          # the same byte-compiled modules as one archive (zipimport)
    if os.path.exists(archive):
        return archive
    return None


def install_shim():
    import transformers.modeling_attn_mask_utils as mu
    import transformers.models.clipseg.modeling_clipseg as mc

    mc._create_4d_causal_attention_mask = mu._create_4d_causal_attention_mask
    mc._prepare_4d_attention_mask = mu._prepare_4d_attention_mask

    # the reference always passes output_attentions= by keyword (possibly None) -> detect by signature binding
    def wrap_kw(cls):
        orig = cls.forward

        def forward(self, hidden_states, *args, **kw):
            # reference-style call = BOTH causal_attention_mask and output_attentions supplied
            # (coop_clipseg.py:149-154, base_clipseg.py:117-122); stock 5.5 callers never do both
            ref_style = "output_attentions" in kw and (len(args) >= 2 or "causal_attention_mask" in kw)
            attention_mask = args[0] if len(args) > 0 else kw.pop("attention_mask", None)
            causal = args[1] if len(args) > 1 else kw.pop("causal_attention_mask", None)
            kw.pop("output_attentions", None)
            mask = attention_mask
            if causal is not None:
                mask = causal if mask is None else causal + mask
            out = orig(self, hidden_states, mask, **kw)
            return (out,) if ref_style else out

        cls.forward = forward

    wrap_kw(mc.CLIPSegEncoderLayer)
    wrap_kw(mc.CLIPSegDecoderLayer)

    # COOPCLIPSeg calls the stock vision model with output_hidden_states=True and reads .hidden_states
    # (coop_clipseg.py:353-358).  Give CLIPSegVisionTransformer the 4.x behaviour for that call.
    vt = mc.CLIPSegVisionTransformer
    orig_vt = vt.forward

    def vt_forward(self, pixel_values=None, output_attentions=None, output_hidden_states=None, return_dict=None, **kw):
        if not output_hidden_states:
            return orig_vt(self, pixel_values=pixel_values, **kw)
        h = self.pre_layrnorm(self.embeddings(pixel_values))
        states = [h]
        for layer in self.encoder.layers:
            h = layer(h, None)
            states.append(h)
        pooled = self.post_layernorm(h[:, 0, :])
        return mc.BaseModelOutputWithPooling(last_hidden_state=h, pooler_output=pooled, hidden_states=tuple(states))

    vt.forward = vt_forward

    # ... and the stock decoder with 4.x keyword arguments (coop_clipseg.py:462)
    dec = mc.CLIPSegDecoder
    orig_dec = dec.forward

    def dec_forward(self, hidden_states, conditional_embeddings, output_attentions=None, output_hidden_states=None,
                    return_dict=True, **kw):
        return orig_dec(self, hidden_states, conditional_embeddings)

    dec.forward = dec_forward
