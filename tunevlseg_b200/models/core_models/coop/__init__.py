"""Prompt-tuned CLIPSeg / CRIS nets with the reference's names (src/models/core_models/coop/__init__.py:1-8)."""
from .clipseg import (  # noqa: F401
    BaseCLIPSeg,
    BaseMultimodalCLIPSeg,
    COOPCLIPSeg,
    MapleCLIPSeg,
    SharedAttnCLIPSeg,
    SharedSeparateCLIPSeg,
    VPTCLIPSeg,
)
from .coop_cris import COOPCRIS  # noqa: F401,E402
