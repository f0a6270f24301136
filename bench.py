#!/usr/bin/env python
"""Benchmark of the TuneVLSeg prompt-tuning TRAIN step (BASELINE.json: train img/s, CLIPSeg + MaPLe, 352x352).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

One "step" = one full training step over one synthetic batch: vision tower (10 layers, deep visual prompts), text
tower (12 layers, deep textual prompts), FiLM decoder + additive head, fused Dice+BCE loss and Dice/IoU counters,
dgrad-only backward, data-parallel gradient all-reduce (N > 1) and AdamW - nothing is skipped or cached.

Prints ONE JSON line (rank 0).  ``value`` is whole-job images/s with the batch already resident in HBM; ``e2e`` is the
same step driven through the public module API (``ImageTextMaskModule.training_step``) from PINNED HOST buffers, with
the host->device copies of image / mask / token ids and the device->host read of the loss inside the timed region.
``roofline`` describes the kernel with the largest share of the step (timed live with CUDA events on the launching
stream); ``cpu_baseline`` times the CPU oracle (a port of the reference's fp32 eager path) on the host cores on a
bounded sample.  ``--impl reference`` times that CPU path alone (the reference has no GPU-specific code; its own
Lightning/Hydra stack is not installable in this image - see DESIGN.md).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

WORKLOAD = "CLIPSeg ViT-B/16 + MaPLe multimodal prompts (depth 9, 4 ctx), 352x352, batch {B}/GPU, 1 binary class, text L=8"
GFLOP_PER_IMG = 166.6      # algorithmic fwd + dgrad-only bwd, counted on the reference modules
GFLOP_PER_IMG_CRIS = 211.5  # SURVEY.md section 8d cfg4 (FlopCounterMode on the reference COOPCRIS; RN50 forward 42.7)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


# ------------------------------------------------------------------------------------------------------------------
# synthetic workload: seed 12345, image ~ N(0,1), ~30 % foreground mask, [BOS, r.., EOS] ids
# ------------------------------------------------------------------------------------------------------------------
def synth_batch(B, seed=12345, L=8, size=352, vocab=49408, pinned=False):
    g = torch.Generator().manual_seed(seed)
    img = torch.randn(B, 3, size, size, generator=g)
    mask = (torch.rand(B, 1, size, size, generator=g) < 0.3).float()
    ids = torch.randint(1000, 40000, (B, L), generator=g)
    ids[:, 0], ids[:, -1] = 49406, 49407
    am = torch.ones(B, L, dtype=torch.long)
    batch = {"image": img, "mask": mask, "input_ids": ids, "attention_mask": am}
    if pinned:
        batch = {k: v.pin_memory() for k, v in batch.items()}
    return batch


def build_module(device, seed=0):
    """Random-init CLIPSeg (CIDAS/clipseg-rd64 geometry) + MaPLe learner inside the reference-shaped LightningModule."""
    from functools import partial

    from transformers import CLIPSegConfig, CLIPSegForImageSegmentation

    from tunevlseg_b200.losses import DiceCELoss
    from tunevlseg_b200.models.core_models.coop import MapleCLIPSeg
    from tunevlseg_b200.models.core_models.coop.context_learner import MapleContextLearner
    from tunevlseg_b200.models.image_text_mask_module import ImageTextMaskModule
    from tunevlseg_b200.optim import FusedAdamW

    torch.manual_seed(seed)
    cfg = CLIPSegConfig(vision_config=dict(image_size=352, patch_size=16), projection_dim=512, reduce_dim=64)
    hf = CLIPSegForImageSegmentation(cfg).eval()
    with torch.no_grad():      # HF inits LayerNorm affine to (1, 0) and biases to 0: make those terms non-trivial
        for name, p in hf.named_parameters():
            if name.endswith("bias"):
                p.normal_(0, 0.02)
            elif "layer_norm" in name or "layernorm" in name or "layrnorm" in name:
                p.add_(torch.randn_like(p) * 0.05)
    net = MapleCLIPSeg(
        model_cfg=dict(pretrained_model_name_or_path=hf, freeze_encoder=False, freeze_decoder=False),
        context_learner=partial(MapleContextLearner, prompt_depth=9, num_context=4, intermediate_dim=64, use_proj_norm=True,
                                use_unified_projection=False, use_lora_proj=False, context_initializer=None),
        freeze_all=True, no_freeze_last_layer=False, use_new_last_layer=True, new_last_layer_kernel_size=5, residual_ratio=0.5)
    module = ImageTextMaskModule(net=net, loss_fn=DiceCELoss(sigmoid=True, lambda_dice=1, lambda_ce=0.2),
                                 optimizer=partial(FusedAdamW, lr=2e-4, weight_decay=0.0), scheduler=None, compile=False,
                                 task="binary", threshold=0.5, weight_decay=0.0)
    module = module.to(device)
    module.setup("fit")
    opt = module.configure_optimizers()["optimizer"]
    return module, opt


CRIS_WORKLOAD = "CRIS CLIP-RN50 + CoCoOp (per-image meta-net context, depth 1, 4 ctx), 416x416, batch {B}/GPU, 1 binary class, text L=8"


def build_module_cris(device, seed=0):
    """Random-init CRIS (CLIP-RN50 geometry of RN50.pt, configs/model/cocoop/cris.yaml) + CoCoOp learner."""
    from functools import partial
    from types import SimpleNamespace

    from tunevlseg_b200.losses import DiceCELoss
    from tunevlseg_b200.models.components.cris_model import CLIP
    from tunevlseg_b200.models.core_models.coop import COOPCRIS
    from tunevlseg_b200.models.core_models.coop.context_learner import CoCoOpContextLearner
    from tunevlseg_b200.models.image_text_mask_module import ImageTextMaskModule
    from tunevlseg_b200.optim import FusedAdamW

    torch.manual_seed(seed)
    clip = CLIP(1024, 224, (3, 4, 6, 3), 64, 77, 49408, 512, 8, 12)
    with torch.no_grad():                      # default BN statistics are (0, 1): make every folded term non-trivial
        for name, buf in clip.named_buffers():
            if name.endswith("running_mean"):
                buf.normal_(0, 0.1)
            elif name.endswith("running_var"):
                buf.uniform_(0.8, 1.2)
    tok = lambda text, **kw: SimpleNamespace(input_ids=torch.tensor([[320, 1125, 539, 320]]))      # "a photo of a"  # noqa: E731
    net = COOPCRIS(
        model_cfg=dict(clip_pretrain=clip.state_dict(), fpn_in=[512, 1024, 1024], fpn_out=[256, 512, 1024], vis_dim=512, word_dim=1024,
                       num_layers=3, num_head=8, dim_ffn=2048, dropout=0.2, return_intermediate=False, img_size=416, freeze_encoder=True,
                       cris_pretrain=None),
        context_learner=partial(CoCoOpContextLearner, norm_image_features=False, prompt_depth=1, use_unified_projection=False,
                                intermediate_dim=64, use_proj_norm=True, use_lora_proj=False, num_context=4,
                                context_initializer="a photo of a", tokenizer=tok),
        freeze_all=True, no_freeze_last_layer=False, use_new_last_layer=True, new_last_layer_kernel_size=5, residual_ratio=0.5)
    module = ImageTextMaskModule(net=net, loss_fn=DiceCELoss(sigmoid=True, lambda_dice=1, lambda_ce=0.2),
                                 optimizer=partial(FusedAdamW, lr=2e-5, weight_decay=0.0), scheduler=None, compile=False,
                                 task="binary", threshold=0.5, weight_decay=0.0)
    module = module.to(device)
    module.setup("fit")
    return module, module.configure_optimizers()["optimizer"]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        mhz = [int(r[0]) for r in self.rows if r and r[0].isdigit()]
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": int(statistics.median(mhz)) if mhz else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(mhz)}


# ------------------------------------------------------------------------------------------------------------------
# CPU oracle step (cpu_baseline and --impl reference)
# ------------------------------------------------------------------------------------------------------------------
def oracle_step_factory_cris(B):
    from oracle import cris as OCR
    from oracle import learners as OL
    from oracle import loss_metrics as OLM

    spec = OCR.CrisSpec()
    w = OCR.init_weights(spec, seed=0)
    g = torch.Generator().manual_seed(1)
    params = {"context_vectors": (torch.randn(1, 4, 512, generator=g) * 0.02).requires_grad_(True),
              "projection_layers.0.0.weight": (torch.randn(64, 1024, generator=g) * 0.04).requires_grad_(True),
              "projection_layers.0.0.bias": torch.zeros(64, requires_grad=True),
              "projection_layers.0.2.weight": (torch.randn(512, 64, generator=g) * 0.1).requires_grad_(True),
              "projection_layers.0.3.weight": torch.ones(512, requires_grad=True), "projection_layers.0.3.bias": torch.zeros(512, requires_grad=True)}
    st = OL.LearnerState(kind="cocoop", prompt_depth=1, num_context=4, params=params, proj_style="mlp")
    head = {k: v.requires_grad_(True) for k, v in OCR.init_head(spec).items()}
    opt = torch.optim.AdamW(list(params.values()) + list(head.values()), lr=2e-5, weight_decay=0.0)
    batch = synth_batch(B, size=416)

    def step():
        opt.zero_grad(set_to_none=True)
        logits = OCR.net_forward(w, spec, st, head, batch["input_ids"], batch["attention_mask"], batch["image"])
        loss = OLM.dice_ce_loss(logits, batch["mask"])
        OLM.metric_counts(torch.sigmoid(logits.detach()), batch["mask"])
        loss.backward()
        opt.step()
        return float(loss.detach())

    return step


def oracle_step_factory(B):
    from oracle import clipseg as OC
    from oracle import learners as OL
    from oracle import loss_metrics as OLM

    spec = OC.ClipSegSpec()
    w = OC.init_weights(spec, seed=0)
    g = torch.Generator().manual_seed(1)
    params = {"context_vectors": (torch.randn(9, 4, 512, generator=g) * 0.02).requires_grad_(True)}
    for d in range(9):
        params[f"projection_layers.{d}.0.weight"] = (torch.randn(64, 512, generator=g) * 0.06).requires_grad_(True)
        params[f"projection_layers.{d}.0.bias"] = torch.zeros(64, requires_grad=True)
        params[f"projection_layers.{d}.2.weight"] = (torch.randn(768, 64, generator=g) * 0.1).requires_grad_(True)
        params[f"projection_layers.{d}.3.weight"] = torch.ones(768, requires_grad=True)
        params[f"projection_layers.{d}.3.bias"] = torch.zeros(768, requires_grad=True)
    st = OL.LearnerState(kind="maple", prompt_depth=9, num_context=4, params=params, proj_style="mlp")
    head = {k: v.requires_grad_(True) for k, v in OC.init_head(spec).items()}
    leaves = list(params.values()) + list(head.values())
    opt = torch.optim.AdamW(leaves, lr=2e-4, weight_decay=0.0)
    batch = synth_batch(B)

    def step():
        opt.zero_grad(set_to_none=True)
        logits = OC.net_forward(w, spec, st, head, batch["input_ids"], batch["attention_mask"], batch["image"])
        loss = OLM.dice_ce_loss(logits, batch["mask"])
        OLM.metric_counts(torch.sigmoid(logits.detach()), batch["mask"])
        loss.backward()
        opt.step()
        return float(loss.detach())

    return step


def time_oracle(B, steps, warmup, workload="maple"):
    torch.set_num_threads(os.cpu_count() or 1)
    step = oracle_step_factory_cris(B) if workload == "cris_cocoop" else oracle_step_factory(B)
    for _ in range(warmup):
        step()
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        ts.append(time.perf_counter() - t0)
    return B / statistics.median(ts), statistics.median(ts) * 1e3, torch.get_num_threads()


def run_reference(args, rank):
    """The reference arm: the reference's own CPU implementation of the path.  Its Lightning/Hydra/monai/torchmetrics
    stack cannot be installed here (no wheels, no network), so this is the oracle port (kind "port") on the host cores."""
    if rank != 0:
        return
    B = args.ref_batch
    cris = args.workload == "cris_cocoop"
    v, ms, cores = time_oracle(B, max(1, args.steps), min(args.warmup, 1), args.workload)
    print(json.dumps({
        "impl": "reference", "metric": "train img/s CRIS+CoCoOp 416x416" if cris else "train img/s CLIPSeg+MaPLe 352x352", "value": round(v, 3), "unit": "img/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": round(ms, 1), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic (seed 12345), random-init weights",
        "config": {"workload": (CRIS_WORKLOAD if cris else WORKLOAD).format(B=B), "note": "bounded sample: batch %d per step on the host cores" % B},
        "cpu_baseline": {"value": round(v, 3), "unit": "img/s", "cores": cores, "kind": "port",
                         "sample": f"{args.steps} full train steps (fwd+loss+metrics+bwd+AdamW) at batch {B}"},
        "e2e": {"value": round(v, 3), "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


# ------------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="images per GPU (BASELINE config: 32)")
    ap.add_argument("--ref-batch", type=int, default=4, help="batch of the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-kernels", action="store_true", help="also print the per-kernel share table to stderr")
    ap.add_argument("--no-graph", action="store_true", help="drive every kernel launch from Python instead of one CUDA graph")
    ap.add_argument("--workload", default="maple", choices=["maple", "cris_cocoop"],
                    help="maple = BASELINE.json's metric (default); cris_cocoop = configs[3], an extra line for the CRIS path")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if args.warmup < 3:
        args.warmup = 3
    from tunevlseg_b200 import abi

    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        from datetime import timedelta

        dist.init_process_group("nccl", device_id=device, timeout=timedelta(seconds=180))
    abi.require_device()
    B = args.batch
    cris = args.workload == "cris_cocoop"
    module, opt = build_module_cris(device, seed=0) if cris else build_module(device, seed=0)
    module.train()
    host = synth_batch(B, seed=12345 + rank, pinned=True, size=416 if cris else 352)
    resident = {k: v.to(device) for k, v in host.items()}
    h2d = sum(v.numel() * v.element_size() for v in host.values())

    def step_eager():
        opt.zero_grad()
        loss = module.training_step(resident, 0)
        loss.backward()
        opt.step()
        return loss

    use_graph = not args.no_graph
    if use_graph:
        from tunevlseg_b200.graph import GraphedTrainStep

        module.train_dice.streaming = True
        n_cap0 = abi.launch_count()
        graphed = GraphedTrainStep(module, opt, resident, warmup=3)
        launches_per_step = (abi.launch_count() - n_cap0) // 4        # 3 warm-up steps + 1 captured step
        step_resident = graphed

        graphed.prefetch(host)

        def step_e2e():
            # this step's inputs were copied host->device on the copy stream while the previous step computed; the copy
            # of the NEXT step's inputs is issued before this step's replay, so every step still pays one full H2D copy
            # inside the timed region - overlapped with compute instead of serialised in front of it
            loss = graphed.step_prefetched()
            graphed.prefetch(host)
            return loss.item()                       # device->host read of the loss (syncs)
    else:
        step_resident = step_eager

        def step_e2e():
            batch = {k: v.to(device, non_blocking=True) for k, v in host.items()}
            opt.zero_grad()
            loss = module.training_step(batch, 0)
            loss.backward()
            opt.step()
            return loss.item()          # device->host read of the step's result (syncs)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / steps

    for _ in range(args.warmup):
        step_resident()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    n0 = abi.launch_count()
    ms_step = timed(step_resident, args.steps)
    launches = abi.launch_count() - n0
    if use_graph:      # replays do not pass through the library's launch counter: kernels per captured step x steps
        launches = launches_per_step * args.steps
    clocks = sampler.stop() if rank == 0 else None

    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)

    # per-kernel shares + roofline of the dominant kernel: one profiled step with CUDA events on the launching stream.
    # EVERY rank runs the step (it contains the gradient all-reduce); only rank 0 records and reports.
    roof = None
    recs = [] if rank == 0 else None
    abi.set_profiler(recs)
    # Hold the stream back (~75 ms of device-side spinning) while the host enqueues the whole step: otherwise every kernel
    # that follows a short one is timed together with the host's ~25 us submission gap (measured: the QKV and fc1 GEMMs,
    # which follow a 17 us LayerNorm, looked 30-50 % slower than they run inside the captured graph).
    # The text tower is run in line for this one pass (TVS_TEXT_STREAM=0): on its own stream its ~250 small kernels share the
    # SMs with whatever is being timed (the same GEMM reads 103 us instead of 89 us), which says nothing about the kernel.
    # The ncu launch list (serialised by construction) agrees with the in-line numbers (profiles/r01b_ncu_launches_summary.md).
    torch.cuda._sleep(150_000_000)
    prev = os.environ.get("TVS_TEXT_STREAM")
    os.environ["TVS_TEXT_STREAM"] = "0"
    try:
        step_eager()
    finally:
        if prev is None:
            os.environ.pop("TVS_TEXT_STREAM", None)
        else:
            os.environ["TVS_TEXT_STREAM"] = prev
    torch.cuda.synchronize()
    abi.set_profiler(None)
    if rank == 0:
        agg = {}
        for name, key, a, b, fl, *_ in recs:
            k = f"{name}:{key}" if key else name
            t, f, c = agg.get(k, (0.0, 0.0, 0))
            agg[k] = (t + a.elapsed_time(b), f + fl, c + 1)
        total = sum(v[0] for v in agg.values())
        top = sorted(agg.items(), key=lambda kv: -kv[1][0])
        if args.profile_kernels:
            for k, (t, f, c) in top[:60]:
                sys.stderr.write(f"{k:48s} n={c:4d} {t:9.3f} ms {100 * t / total:5.1f}%  {f / max(t, 1e-9) / 1e9:8.1f} TFLOP/s\n")
        pk = peaks()
        k, (t, f, c) = next((kv for kv in top if kv[1][1] > 0), top[0])
        fam = k.split(":")[0]
        fam_t = sum(v[0] for kk, v in agg.items() if kk.split(":")[0] == fam)
        achieved = f / (t * 1e-3) / 1e12
        traffic = None      # dram__bytes_read + dram__bytes_write per launch from the committed ncu --set full capture, if it is this kernel
        tp = os.path.join(ROOT, "profiles", "r01_roofline_traffic.json")
        if os.path.exists(tp):
            tj = json.load(open(tp))
            if tj.get("kernel") == k:
                traffic = tj["dram_bytes_per_launch"]
        roof = {"bound": "tensor", "kernel": k, "achieved": round(achieved, 1), "peak": pk["tf_sustained"], "unit": "TFLOP/s",
                "frac": round(achieved / pk["tf_sustained"], 4), "traffic": traffic, "peak_source": pk["src"] + " (sustained cuBLAS bf16)",
                "launches_per_step": c, "avg_launch_us": round(1e3 * t / c, 1), "share_of_step": round(t / total, 3),
                "timing": "CUDA events around every launch of one eager step, text tower in line, host enqueued ahead of the device",
                "family_share_of_step": round(fam_t / total, 3)}

    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        v, ms, cores = time_oracle(args.ref_batch, 3, 1, args.workload)
        cpu = {"value": round(v, 3), "unit": "img/s", "cores": cores, "kind": "port",
               "sample": f"3 full train steps of the fp32 CPU oracle at batch {args.ref_batch} (median {ms:.0f} ms/step)"}

    if rank == 0:
        value = world * B / (ms_step * 1e-3)
        e2e = world * B / (ms_e2e * 1e-3)
        pk = peaks()
        out = {
            "metric": "train img/s CRIS+CoCoOp 416x416" if cris else "train img/s CLIPSeg+MaPLe 352x352", "value": round(value, 1),
            "unit": "img/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms_step, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "tf32" if cris else "bf16", "data": "synthetic (seed 12345), random-init weights",
            "config": {"workload": (CRIS_WORKLOAD if cris else WORKLOAD).format(B=B), "global_batch": world * B, "parallelism": f"dp{world}",
                       "cuda_graph": use_graph,
                       "l2": "per-step working set (>2 GB of activations) exceeds the 126 MB L2; no explicit flush",
                       "precision": ("fp32 activations, kind::tf32 tcgen05 GEMMs with round-to-nearest operands, fp32 text attention" if cris else
                                     "bf16 tcgen05 GEMMs + fp32 residual stream in the vision tower; tf32 GEMMs in text tower/decoder")},
            "model_tflops_per_gpu": round(value / world * (GFLOP_PER_IMG_CRIS if cris else GFLOP_PER_IMG) / 1e3, 1),
            "model_flops_frac_of_peak": round(value / world * (GFLOP_PER_IMG_CRIS if cris else GFLOP_PER_IMG) / 1e3 / pk["tf_sustained"], 4),
            "roofline": roof, "cpu_baseline": cpu,
            "e2e": {"value": round(e2e, 1), "unit": "img/s", "ms_per_step": round(ms_e2e, 3), "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
            "gpu_launches": int(launches), "clocks": clocks}
        print(json.dumps(out))
    if world > 1:
        # The captured graph holds NCCL work on the communicator; tearing the process group down under it can block
        # forever (observed).  All ranks are done and rank 0 has printed: synchronise, flush and leave without the
        # communicator teardown.
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
