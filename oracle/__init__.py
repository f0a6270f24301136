"""CPU oracle for the TuneVLSeg prompt-tuning train/eval step.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and there only as the checker.  The
product path (``tunevlseg_b200``) never imports this package and fails loudly
when its CUDA extension is missing.

What it restates (plain torch fp32 on the CPU, no transformers / monai /
torchmetrics / lightning imports):

* ``oracle.clipseg``   - HF CLIPSeg arithmetic (transformers 5.5.0
  ``models/clipseg/modeling_clipseg.py``) plus the reference wrappers' prompt
  plumbing (``src/models/core_models/coop/*.py``).
* ``oracle.learners``  - the six prompt learners as pure functions over a
  state dict (``src/models/core_models/coop/context_learner/*.py``).
* ``oracle.loss_metrics`` (+ ``loss_metrics.c``) - MONAI ``DiceCELoss`` and the
  torchmetrics ``Dice`` / ``JaccardIndex`` integer counters.

Parity pin: ``tests/golden/make_golden.py`` runs the *real* reference classes
from ``/root/reference`` (through a test-side transformers-5.x shim) and freezes
their outputs as fixtures; ``tests/test_oracle_golden.py`` holds the oracle to
them at <=1e-5.  MONAI and torchmetrics are not installed in this image, so the
loss/metric formulas are pinned only by hand-derived known-answer cases:
"parity unpinned" for those two third-party boundaries.
"""
