"""``ImageTextMaskModule`` with the reference's constructor, step methods and optimizer grouping
(/root/reference/src/models/image_text_mask_module.py:23-395).  ``model_step`` makes ONE pass over logits and mask:
the fused kernel returns the Dice+BCE loss, the per-sample Dice counters and the batch confusion matrix, replacing
``loss_fn`` + ``sigmoid`` + ``mask.long()`` + two torchmetrics updates.

``pytorch_lightning`` is used when importable; otherwise a minimal stand-in base keeps the same method surface
(``log``, ``log_dict``, ``save_hyperparameters``, ``hparams``) so the step logic can run under a plain loop.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Any

import torch
from torch import nn

from ..losses import DiceCELoss
from ..metrics import Dice, JaccardIndex

try:  # pragma: no cover - lightning is not installed in the build image
    from pytorch_lightning import LightningModule as _Base
    _HAVE_LIGHTNING = True
except Exception:  # noqa: BLE001
    _HAVE_LIGHTNING = False

    class _Base(nn.Module):
        """Just enough of LightningModule for the step methods below."""

        def __init__(self) -> None:
            super().__init__()
            self.hparams = SimpleNamespace()
            self.logged: dict[str, Any] = {}
            self.logger = None
            self.global_step = 0

        def save_hyperparameters(self, ignore=(), frame_locals=None) -> None:
            for k, v in (frame_locals or {}).items():
                if k not in ignore and k not in ("self", "__class__", "args", "kwargs"):
                    setattr(self.hparams, k, v)

        def log(self, name, value, **kwargs) -> None:
            self.logged[name] = value

        def log_dict(self, values, **kwargs) -> None:
            self.logged.update(values)


class ImageTextMaskModule(_Base):
    plot_columns = ["Image", "Caption", "Label"]

    def __init__(self, net: nn.Module, loss_fn: nn.Module, optimizer, scheduler, compile: bool, task: str,
                 threshold: float = 0.5, weight_decay: float = 0.0, log_image_num: int = 8, lr_scheduler_config=None,
                 activation_fn=torch.sigmoid, cache_outputs: bool = False, *args, **kwargs) -> None:
        super().__init__()
        if _HAVE_LIGHTNING:
            self.save_hyperparameters(ignore=["net", "loss_fn", "optimizer", "scheduler"])
        else:
            self.save_hyperparameters(ignore=["net", "loss_fn", "optimizer", "scheduler"], frame_locals=dict(locals()))
        if compile:
            raise NotImplementedError("compile=True (torch.compile) is not used: the net already runs hand-written kernels")
        if activation_fn is not None and activation_fn is not torch.sigmoid:
            raise NotImplementedError("the fused loss/metric kernel implements the reference's sigmoid activation only")
        self.net, self.loss_fn, self.optimizer, self.scheduler = net, loss_fn, optimizer, scheduler
        self.activation_fn = nn.Identity() if activation_fn is None else activation_fn
        self.registered_metric_names: list[str] = []
        self._batch_conf: torch.Tensor | None = None
        self._batch_counts: torch.Tensor | None = None

    def forward(self, *args, **kwargs) -> torch.Tensor:
        return self.net(*args, **kwargs)

    def on_train_start(self) -> None:
        for name in self.registered_metric_names:
            if name.startswith("val"):
                reset = getattr(getattr(self, name, None), "reset", None)
                if reset is not None:
                    reset()

    # ---- the step ------------------------------------------------------------------------------------------------
    def model_step(self, batch, materialize: bool = False):
        """-> (loss, preds, targets).  With the fused ``DiceCELoss`` the metric counters of this batch are left in
        ``self._batch_counts`` / ``self._batch_conf`` and ``preds`` / ``targets`` are only materialised on request
        (validation image logging, predict) - nothing on the training path reads them."""
        logits = self.get_logits(batch)
        mask = batch["mask"]
        if isinstance(self.loss_fn, DiceCELoss):
            conf = torch.zeros(4, dtype=torch.int64, device=logits.device)
            loss, counts = self.loss_fn.forward_with_metrics(logits, mask, self.hparams.threshold, conf)
            self._batch_counts, self._batch_conf = counts, conf
            if not materialize:
                return loss, None, None
            return loss, self.activation_fn(logits.detach()), mask.long()
        loss = self.loss_fn(logits, mask)
        preds, targets = self.activation_fn(logits.detach()), mask.long()
        self._batch_counts = self._batch_conf = None
        return loss, preds, targets

    def _update_metrics(self, dice: Dice, iou: JaccardIndex, preds, targets):
        if self._batch_counts is not None:
            return dice.update_from_counts(self._batch_counts), iou.update_from_confmat(self._batch_conf)
        return dice(preds, targets), iou(preds, targets)

    def training_step(self, batch, batch_idx: int) -> torch.Tensor:
        loss, preds, targets = self.model_step(batch)
        d, i = self._update_metrics(self.train_dice, self.train_iou, preds, targets)
        n = len(batch["mask"])
        self.log_dict({"train_dice_step": d, "train_iou_step": i}, prog_bar=True, batch_size=n)
        self.log("train_loss", torch.nan_to_num(loss.detach(), nan=float("inf")), on_step=True, on_epoch=True, prog_bar=True,
                 logger=True, batch_size=n)
        return loss

    # ---- epoch-level metric values ----------------------------------------------------------------------------------
    # The reference logs the torchmetrics objects themselves (image_text_mask_module.py:118-125,144-148,162-170):
    # Lightning then reports ``metric.compute()`` over the state ACCUMULATED during the epoch and resets the metric
    # afterwards.  The metric objects here are filled by the fused kernel and are not ``torchmetrics.Metric``s, so
    # the same semantics are spelled out: per-step batch values for the training progress bar, and at every epoch end
    # the accumulated value is logged under the reference's key and the state is reset.
    def _log_epoch_metrics(self, stage: str, dice_key: str, iou_key: str) -> None:
        dice, iou = getattr(self, f"{stage}_dice", None), getattr(self, f"{stage}_iou", None)
        if dice is None or iou is None:
            return
        try:       # every rank takes part in the reductions inside compute(); "no update anywhere" is decided after them
            self.log_dict({dice_key: dice.compute(), iou_key: iou.compute()}, prog_bar=True)
        except RuntimeError as e:
            if "before any update" not in str(e):
                raise
        dice.reset()
        iou.reset()

    def on_train_epoch_end(self) -> None:
        self._log_epoch_metrics("train", "train_dice_epoch", "train_iou_epoch")

    def on_validation_epoch_end(self) -> None:
        self._log_epoch_metrics("val", "val_dice", "val_iou")

    def on_test_epoch_end(self) -> None:
        self._log_epoch_metrics("test", "test_dice", "test_iou")

    def validation_step(self, batch, batch_idx: int) -> None:
        want_images = self.logger is not None and (self.global_step == 0 or batch_idx == 0)
        loss, preds, targets = self.model_step(batch, materialize=want_images)
        d, i = self._update_metrics(self.val_dice, self.val_iou, preds, targets)
        # val_dice / val_iou are epoch values: logged from the accumulated state in on_validation_epoch_end
        self.log_dict({"val_loss": torch.nan_to_num(loss.detach(), nan=float("inf"))}, prog_bar=True, batch_size=len(batch["mask"]))
        if want_images and hasattr(self.logger, "log_image"):   # wandb tables/images: host-side glue, unchanged semantics
            k = self.hparams.log_image_num
            self.logger.log_image("val_pred", [p for p in preds[:k].float().cpu()])

    def test_step(self, batch, batch_idx: int = 0) -> None:
        loss, preds, targets = self.model_step(batch)
        d, i = self._update_metrics(self.test_dice, self.test_iou, preds, targets)
        self.log_dict({"test_loss": torch.nan_to_num(loss.detach(), nan=float("inf"))}, prog_bar=True, batch_size=len(batch["mask"]))

    def predict_step(self, batch, batch_idx: int = 0):
        logits = self.get_logits(batch)
        return {"preds": self.activation_fn(logits), "mask_name": batch["mask_name"], "mask_shape": batch["mask_shape"]}

    def get_logits(self, batch) -> torch.Tensor:
        text_input = {k: batch[k] for k in ("input_ids", "attention_mask")}
        if getattr(self.hparams, "cache_outputs", False):
            text_input["cache_name"] = batch["cache_name"]
        return self(image_input=batch["image"], text_input=text_input)

    # ---- metrics / optimisers --------------------------------------------------------------------------------------
    def store_and_register_metrics(self, metric_name: str, metric: Any) -> None:
        if isinstance(metric, nn.Module):          # setup() may run after the module was moved to the GPU
            p = next(self.parameters(), None)
            if p is not None:
                metric = metric.to(p.device)
        setattr(self, metric_name, metric)
        self.registered_metric_names.append(metric_name)

    def setup(self, stage: str) -> None:
        common = {"threshold": self.hparams.threshold, "zero_division": 1}
        dice_kw = {**common, "average": "samples"}
        iou_kw = {**common, "task": self.hparams.task}
        if stage == "fit":
            self.store_and_register_metrics("train_dice", Dice(**dice_kw))
            self.store_and_register_metrics("train_iou", JaccardIndex(**iou_kw))
        if stage in {"fit", "validate"}:
            self.store_and_register_metrics("val_dice", Dice(**dice_kw))
            self.store_and_register_metrics("val_iou", JaccardIndex(**iou_kw))
        if stage == "test":
            self.store_and_register_metrics("test_dice", Dice(**dice_kw))
            self.store_and_register_metrics("test_iou", JaccardIndex(**iou_kw))

    def get_optim_groups(self):
        """Decay / no-decay split by module type and parameter name (image_text_mask_module.py:304-361)."""
        if self.hparams.weight_decay <= 0:
            return self.parameters()
        decay, no_decay = set(), set()
        whitelist = (nn.Linear, nn.modules.conv._ConvNd)
        blacklist = (nn.Embedding, nn.GroupNorm, nn.LayerNorm, nn.modules.batchnorm._NormBase)
        for mn, m in self.named_modules():
            for pn, _ in m.named_parameters():
                fpn = f"{mn}.{pn}" if mn else pn
                if pn.endswith("proj_weight"):
                    decay.add(fpn)
                elif pn.endswith("weight"):
                    if isinstance(m, whitelist):
                        decay.add(fpn)
                    elif isinstance(m, blacklist):
                        no_decay.add(fpn)
                else:
                    no_decay.add(fpn)
        both = decay & no_decay
        if both:
            raise ValueError(f"parameters {both} made it into both decay/no_decay sets!")
        params = dict(self.named_parameters())
        extra = params.keys() - (decay | no_decay)
        if extra:
            raise ValueError(f"parameters {extra} were not separated into either decay/no_decay set!")
        return [{"params": [params[pn] for pn in sorted(decay)], "weight_decay": self.hparams.weight_decay},
                {"params": [params[pn] for pn in sorted(no_decay)], "weight_decay": 0.0}]

    def configure_optimizers(self):
        groups = self.get_optim_groups()
        optimizer = self.optimizer(groups)
        if self.scheduler is not None:
            scheduler = self.scheduler(optimizer=optimizer)
            return {"optimizer": optimizer,
                    "lr_scheduler": {"scheduler": scheduler, "monitor": "val_loss", "interval": "epoch", "frequency": 1,
                                     **(self.hparams.lr_scheduler_config or {})}}
        return {"optimizer": optimizer}
