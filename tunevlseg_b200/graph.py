"""Whole-step CUDA-graph capture: forward, fused loss/metrics, dgrad backward, gradient all-reduce and AdamW are
recorded once and replayed as one graph launch - the step is ~500 kernels, many of them microseconds long (the
12-layer text tower works on B*S ~ 400 rows), so launch latency, not arithmetic, bounds them when driven from Python.

The reference drives ~25 ATen launches per layer from Python with no graph.
"""
from __future__ import annotations

import os

import torch


class GraphedTrainStep:
    """Capture ``zero_grad -> module.training_step -> backward -> optimizer.step`` for a fixed batch shape.

    ``module`` is an ``ImageTextMaskModule`` (or anything with ``training_step(batch, idx) -> loss``); ``optimizer`` a
    ``FusedAdamW`` (device-side step counter, flat buffers -> capturable).  Call with a dict of (host-pinned or device)
    tensors of the captured shapes: they are copied into the static inputs, the graph is replayed and the static loss
    tensor is returned (read it with ``.item()`` after the replay if a host value is needed).
    """

    def __init__(self, module, optimizer, example_batch: dict, warmup: int = 3) -> None:
        self.module, self.optimizer = module, optimizer
        # a replayed graph cannot grow a Python list: every Dice metric of the module accumulates on the device instead
        # (same value as the cat-reduced per-sample lists; metrics.Dice.streaming)
        for name in getattr(module, "registered_metric_names", []):
            metric = getattr(module, name, None)
            if hasattr(metric, "streaming"):
                metric.streaming = True
        self.static = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in example_batch.items()}
        for v in self.static.values():
            if torch.is_tensor(v) and not v.is_cuda:
                raise ValueError("example_batch must live on the CUDA device")
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._eager_step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        # TVS_MAIN_PRIORITY=1 (experiment switch): capture the main chain on a high-priority stream, so that its thread
        # blocks are placed before those of the text-tower stream whenever both have blocks pending
        prio = os.environ.get("TVS_MAIN_PRIORITY", "0") == "1"
        with torch.cuda.graph(self.graph, stream=torch.cuda.Stream(priority=-1) if prio else None):
            self.loss = self._eager_step()

    def _eager_step(self) -> torch.Tensor:
        self.optimizer.zero_grad()
        loss = self.module.training_step(self.static, 0)
        loss.backward()
        self.optimizer.step()
        return loss.detach()

    def load(self, batch: dict) -> None:
        for k, v in batch.items():
            if torch.is_tensor(v):
                self.static[k].copy_(v, non_blocking=True)

    # ---- input prefetch: the host->device copy of step k+1 runs on a copy stream while step k computes -------------
    def prefetch(self, batch: dict) -> None:
        """Start copying ``batch`` (pinned host tensors) into a staging buffer on a side stream."""
        if not hasattr(self, "_copy_stream"):
            self._copy_stream = torch.cuda.Stream()
            self._staging = {k: torch.empty_like(v) for k, v in self.static.items() if torch.is_tensor(v)}
            self._staged = torch.cuda.Event()
            self._consumed = torch.cuda.Event()
            self._consumed.record()
        self._copy_stream.wait_event(self._consumed)          # the previous staging contents have been consumed
        with torch.cuda.stream(self._copy_stream):
            for k, v in batch.items():
                if torch.is_tensor(v):
                    self._staging[k].copy_(v, non_blocking=True)
            self._staged.record()

    def step_prefetched(self) -> torch.Tensor:
        """Consume the staged batch (device-to-device copy into the graph's static inputs) and replay."""
        cur = torch.cuda.current_stream()
        cur.wait_event(self._staged)
        for k, v in self._staging.items():
            self.static[k].copy_(v, non_blocking=True)
        self._consumed.record()
        self._replay()
        return self.loss

    def _replay(self) -> None:
        refresh = getattr(self.optimizer, "refresh_lr", None)
        if refresh is not None:        # a scheduler changed param_groups[i]["lr"]: update the device scalar the graph reads
            refresh()
        self.graph.replay()

    def __call__(self, batch: dict | None = None) -> torch.Tensor:
        if batch is not None:
            self.load(batch)
        self._replay()
        return self.loss
