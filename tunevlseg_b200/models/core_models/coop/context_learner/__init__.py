"""Prompt learners with the reference's names and constructor keywords
(/root/reference/src/models/core_models/coop/context_learner/__init__.py:1-9)."""
from .learners import (  # noqa: F401
    BaseProjectorLearner,
    BaseSharedLearner,
    BaseUnimodalLearner,
    BaseVisualLearner,
    CoCoOpContextLearner,
    CoOpContextLearner,
    MapleContextLearner,
    SharedAttnLearner,
    SharedSeparateLearner,
    VPTContextLearner,
)
