"""``FusedAdamW``: torch.optim.AdamW semantics (configs/model/maple_clipseg.yaml:36-39) over ONE flat fp32 buffer per
parameter group, updated by a single sm_100a kernel, with the data-parallel gradient all-reduce done on the same flat
buffer (one NCCL call per step).

Parameters are re-pointed at views of the flat buffer, their ``.grad`` at views of a flat gradient buffer that stays
allocated (``zero_grad`` is one memset), so parameters the step never uses (the reference's dead
``additive_decoder_layer`` under CoOp, ``residual_ratio`` under VPT) simply contribute zeros - no
``find_unused_parameters`` machinery.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import abi


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
            off = 0
            for p in ps:
                k = p.numel()
                flat[off:off + k].copy_(p.data.reshape(-1).to(torch.float32))
                p.data = flat[off:off + k].view(p.shape)
                p.grad = gflat[off:off + k].view(p.shape)
                off += k
            self._flat.append(dict(param=flat, grad=gflat, m=torch.zeros_like(flat), v=torch.zeros_like(flat),
                                   step_dev=torch.zeros(1, dtype=torch.int32, device=dev),
                                   lr_dev=torch.full((1,), float(group["lr"]), dtype=torch.float32, device=dev), params=ps))

    def zero_grad(self, set_to_none: bool = False) -> None:   # grads stay allocated: they are views of the flat buffer
        for f in self._flat:
            if f is not None:
                f["grad"].zero_()

    @property
    def flat_grads(self):
        return [f["grad"] for f in self._flat if f is not None]

    def grad_bytes(self) -> int:
        return sum(g.numel() * 4 for g in self.flat_grads)

    def sync_lr(self) -> None:
        """Copy the (possibly scheduler-modified) group lr to the device scalars the kernel reads."""
        for group, f in zip(self.param_groups, self._flat):
            if f is not None:
                f["lr_dev"].fill_(float(group["lr"]))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        world = 1
        if dist.is_available() and dist.is_initialized():
            world = dist.get_world_size(self.process_group)
        for group, f in zip(self.param_groups, self._flat):
            if f is None:
                continue
            for p in f["params"]:      # autograd may have replaced .grad if it was None-d by foreign code
                if p.grad is None or p.grad.data_ptr() < f["grad"].data_ptr() or p.grad.data_ptr() >= f["grad"].data_ptr() + f["grad"].numel() * 4:
                    raise abi.TvsError("FusedAdamW: a parameter's .grad no longer aliases the flat buffer; call "
                                       "optimizer.zero_grad() (not set_to_none) between steps")
            if world > 1:
                dist.all_reduce(f["grad"], op=dist.ReduceOp.SUM, group=self.process_group)
            abi.counter_inc(f["step_dev"])
            b1, b2 = group["betas"]
            abi.adamw_flat(f["param"], f["grad"], f["m"], f["v"], float(group["lr"]), float(b1), float(b2), float(group["eps"]),
                           float(group["weight_decay"]), 0, 1.0 / world, step_dev=f["step_dev"], lr_dev=f["lr_dev"])
        return loss
