"""Prompt-tuned CLIPSeg nets with the reference's names (src/models/core_models/coop/__init__.py:1-8).
COOPCRIS (CRIS / CLIP-RN50) is not built yet - see DESIGN.md "out of scope this round"."""
from .clipseg import (  # noqa: F401
    BaseCLIPSeg,
    BaseMultimodalCLIPSeg,
    COOPCLIPSeg,
    MapleCLIPSeg,
    SharedAttnCLIPSeg,
    SharedSeparateCLIPSeg,
    VPTCLIPSeg,
)
