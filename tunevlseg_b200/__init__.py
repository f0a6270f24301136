"""tunevlseg_b200 - a B200-native (sm_100a) implementation of TuneVLSeg's prompt-tuning train/eval step.

Package layout (only what the hot path needs):
    csrc/      hand-written CUDA kernels + the C ABI (include/tvs_b200.h) -> lib/libtvs_b200.so
    abi.py     ctypes binding (raw device pointers + stream); no fallback
    engine.py  weight packing, per-layer schedules, the four autograd nodes
    models/    host-side mirror of the reference's ``src.models`` interface (same class names / signatures)
    losses.py, metrics.py, optim.py   fused DiceCE loss, Dice/IoU metric objects, flat AdamW + NCCL all-reduce
"""
from __future__ import annotations

import importlib
import sys

__version__ = "0.1.0"

try:  # the text tower runs on a side stream by design: its AccumulateGrad nodes legitimately see two streams
    import torch as _torch

    _torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)
except Exception:  # noqa: BLE001 - older torch
    pass


def install_as_src() -> None:
    """Expose this package under the reference's module paths, so its Hydra configs resolve unchanged:
    ``_target_: src.models.core_models.coop.MapleCLIPSeg`` -> ``tunevlseg_b200.models.core_models.coop.MapleCLIPSeg``;
    ``monai.losses.DiceCELoss`` -> the fused loss when monai is absent."""
    names = ["models", "models.image_text_mask_module", "models.components", "models.components.hf_clipseg_wrapper", "models.components.cris_model",
             "models.core_models", "models.core_models.coop", "models.core_models.coop.context_learner"]
    if "src" not in sys.modules:
        import types
        sys.modules["src"] = types.ModuleType("src")
    for n in names:
        mod = importlib.import_module(f"{__name__}.{n}")
        sys.modules[f"src.{n}"] = mod
    setattr(sys.modules["src"], "models", sys.modules["src.models"])
    # collate_fn: _target_: src.data.components.data_collator.CustomDataCollatorWithPadding (configs/experiment/coop/clipseg.yaml:133-143)
    import types
    for n in ("data", "data.components"):
        sys.modules.setdefault(f"src.{n}", types.ModuleType(f"src.{n}"))
    sys.modules["src.data.components.data_collator"] = importlib.import_module(f"{__name__}.data.data_collator")
    setattr(sys.modules["src"], "data", sys.modules["src.data"])
    setattr(sys.modules["src.data"], "components", sys.modules["src.data.components"])
    setattr(sys.modules["src.data.components"], "data_collator", sys.modules["src.data.components.data_collator"])
    _alias_monai_loss()


def _alias_monai_loss() -> None:
    """``_target_: monai.losses.DiceCELoss`` (configs/model/maple_clipseg.yaml:29-33) -> ``tunevlseg_b200.losses.DiceCELoss``.
    When monai is importable its module object is kept and only the ``DiceCELoss`` attribute is re-pointed (the fused
    class raises ``NotImplementedError`` for any configuration it does not cover, so nothing is silently different);
    when it is absent, stand-in ``monai`` / ``monai.losses`` modules holding just that class are registered."""
    import types

    from .losses import DiceCELoss

    try:
        losses = importlib.import_module("monai.losses")
    except Exception:  # noqa: BLE001 - monai absent (or broken): register the stand-in
        monai = sys.modules.get("monai") or types.ModuleType("monai")
        losses = types.ModuleType("monai.losses")
        monai.losses = losses
        monai.__dict__.setdefault("__tvs_stand_in__", True)
        sys.modules["monai"], sys.modules["monai.losses"] = monai, losses
    losses.DiceCELoss = DiceCELoss
