"""``FusedAdamW``: torch.optim.AdamW semantics (configs/model/maple_clipseg.yaml:36-39) over ONE flat fp32 buffer per
parameter group, updated by a single sm_100a kernel, with the data-parallel gradient all-reduce done on the same flat
buffer (one NCCL call per step).

Parameters are re-pointed at views of the flat buffer, their ``.grad`` at views of a flat gradient buffer that stays
allocated (``zero_grad`` is one memset), so parameters the step never uses (the reference's dead
``additive_decoder_layer`` under CoOp, ``residual_ratio`` under VPT) simply contribute zeros - no
``find_unused_parameters`` machinery.

Scheduler / checkpoint compatibility (the reference drives ``ReduceLROnPlateau`` and Lightning checkpoints,
maple_clipseg.yaml:50-55):
  * the kernel reads the learning rate from a device scalar (so a captured CUDA graph sees scheduler changes); the
    scalar is refreshed from ``param_groups[i]["lr"]`` by every eager ``step()`` and by ``GraphedTrainStep`` before each
    replay (``refresh_lr``) - a scheduler only ever has to touch ``param_groups``, as with ``torch.optim.AdamW``;
  * ``optimizer.state[p]`` holds ``step`` / ``exp_avg`` / ``exp_avg_sq`` in the torch.optim.AdamW layout (the moment
    tensors are views of the flat moment buffers), so ``state_dict()`` / ``load_state_dict()`` round-trip moments and
    the step count, and a torch.optim.AdamW state dict of the same parameters loads.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import abi


def _capturing() -> bool:
    return torch.cuda.is_available() and torch.cuda.is_current_stream_capturing()


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-2,
                 process_group=None) -> None:
        if lr < 0.0 or eps < 0.0 or not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0 or weight_decay < 0.0:
            raise ValueError("invalid AdamW hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.process_group = process_group
        self._flat = []
        for group in self.param_groups:
            ps = [p for p in group["params"] if p.requires_grad]
            if not ps:
                self._flat.append(None)
                continue
            dev = ps[0].device
            if dev.type != "cuda":
                raise abi.TvsError("FusedAdamW needs CUDA parameters (no CPU fallback)")
            n = sum(p.numel() for p in ps)
            n_pad = (n + 3) // 4 * 4
            flat = torch.zeros(n_pad, dtype=torch.float32, device=dev)
            gflat = torch.zeros(n_pad, dtype=torch.float32, device=dev)
            m, v = torch.zeros_like(flat), torch.zeros_like(flat)
            off = 0
            spans = []
            for p in ps:
                k = p.numel()
                flat[off:off + k].copy_(p.data.reshape(-1).to(torch.float32))
                p.data = flat[off:off + k].view(p.shape)
                p.grad = gflat[off:off + k].view(p.shape)
                spans.append((off, k))
                off += k
            self._flat.append(dict(param=flat, grad=gflat, m=m, v=v,
                                   step_dev=torch.zeros(1, dtype=torch.int32, device=dev),
                                   lr_dev=torch.full((1,), float(group["lr"]), dtype=torch.float32, device=dev),
                                   lr_host=float(group["lr"]), params=ps, spans=spans))
        self._bind_state()

    # ---- torch.optim.AdamW-shaped per-parameter state (views of the flat moment buffers) ---------------------------
    def _bind_state(self, steps: list[float] | None = None) -> None:
        for gi, f in enumerate(self._flat):
            if f is None:
                continue
            step = torch.tensor(float(steps[gi]) if steps is not None else 0.0, dtype=torch.float32)
            for p, (off, k) in zip(f["params"], f["spans"]):
                self.state[p] = {"step": step.clone(), "exp_avg": f["m"][off:off + k].view(p.shape),
                                 "exp_avg_sq": f["v"][off:off + k].view(p.shape)}

    def state_dict(self):
        """torch.optim.AdamW layout; moments are cloned out of the flat buffers, ``step`` is read from the device
        counter (one small D2H read - checkpoint time only)."""
        for f in self._flat:
            if f is None:
                continue
            step = float(f["step_dev"].item())
            for p in f["params"]:
                self.state[p]["step"] = torch.tensor(step, dtype=torch.float32)
        sd = super().state_dict()
        for st in sd["state"].values():
            for k, v in list(st.items()):
                if torch.is_tensor(v):
                    st[k] = v.detach().clone()
        return sd

    def load_state_dict(self, state_dict) -> None:
        """Restores moments, step count and learning rate into the flat buffers / device scalars (accepts the state dict
        of a ``torch.optim.AdamW`` over the same parameters)."""
        super().load_state_dict(state_dict)        # validates groups, casts tensors to the parameters' device
        steps = []
        with torch.no_grad():
            for f in self._flat:
                if f is None:
                    steps.append(0.0)
                    continue
                step = 0.0
                for p, (off, k) in zip(f["params"], f["spans"]):
                    st = self.state.get(p) or {}
                    if "exp_avg" in st:
                        f["m"][off:off + k].copy_(st["exp_avg"].reshape(-1).to(torch.float32))
                        f["v"][off:off + k].copy_(st["exp_avg_sq"].reshape(-1).to(torch.float32))
                        step = max(step, float(st.get("step", 0.0)))
                    else:
                        f["m"][off:off + k].zero_()
                        f["v"][off:off + k].zero_()
                f["step_dev"].fill_(int(step))
                steps.append(step)
        self._bind_state(steps)
        self.refresh_lr(force=True)

    def zero_grad(self, set_to_none: bool = False) -> None:   # grads stay allocated: they are views of the flat buffer
        for f in self._flat:
            if f is not None:
                f["grad"].zero_()

    @property
    def flat_grads(self):
        return [f["grad"] for f in self._flat if f is not None]

    def grad_bytes(self) -> int:
        return sum(g.numel() * 4 for g in self.flat_grads)

    def refresh_lr(self, force: bool = False) -> None:
        """Copy a changed group lr (scheduler, ``load_state_dict``) to the device scalar the kernel reads.  A host-side
        float compare per group; no device work unless the value changed.  Must not run inside a stream capture (the
        fill would be baked into the graph with the capture-time value) - the graph owner calls it before each replay."""
        for group, f in zip(self.param_groups, self._flat):
            if f is None:
                continue
            lr = float(group["lr"])
            if force or lr != f["lr_host"]:
                f["lr_dev"].fill_(lr)
                f["lr_host"] = lr

    sync_lr = refresh_lr       # round-1 name

    def _check_aliasing(self, f) -> None:
        g0, g1 = f["grad"].data_ptr(), f["grad"].data_ptr() + f["grad"].numel() * 4
        p0, p1 = f["param"].data_ptr(), f["param"].data_ptr() + f["param"].numel() * 4
        for p in f["params"]:      # autograd may have replaced .grad if it was None-d by foreign code
            if p.grad is None or not g0 <= p.grad.data_ptr() < g1:
                raise abi.TvsError("FusedAdamW: a parameter's .grad no longer aliases the flat buffer; call "
                                   "optimizer.zero_grad() (not set_to_none) between steps")
            if not p0 <= p.data_ptr() < p1:
                raise abi.TvsError("FusedAdamW: a parameter's storage no longer aliases the flat buffer (the module was moved or "
                                   "cast after the optimizer was built: model.to()/float()/load with assign=True); rebuild the optimizer")

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        world = 1
        if dist.is_available() and dist.is_initialized():
            world = dist.get_world_size(self.process_group)
        if not _capturing():
            self.refresh_lr()
        for group, f in zip(self.param_groups, self._flat):
            if f is None:
                continue
            self._check_aliasing(f)
            if world > 1:
                with abi.nvtx_range("ddp.all_reduce_flat_grad"):
                    dist.all_reduce(f["grad"], op=dist.ReduceOp.SUM, group=self.process_group)
            abi.counter_inc(f["step_dev"])
            b1, b2 = group["betas"]
            abi.adamw_flat(f["param"], f["grad"], f["m"], f["v"], float(group["lr"]), float(b1), float(b2), float(group["eps"]),
                           float(group["weight_decay"]), 0, 1.0 / world, step_dev=f["step_dev"], lr_dev=f["lr_dev"])
        return loss
