#!/usr/bin/env python
"""Benchmark of the TuneVLSeg prompt-tuning TRAIN step (BASELINE.json: train img/s, CLIPSeg + MaPLe, 352x352).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME] [--no-secondary]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

One "step" = one full training step over one synthetic batch: vision tower (deep visual prompts), text tower (deep
textual prompts), FiLM decoder + additive head, fused Dice+BCE loss and Dice/IoU counters, dgrad-only backward,
data-parallel gradient all-reduce (N > 1) and AdamW - nothing is skipped or cached.

Prints ONE JSON line (rank 0).  The top-level keys describe BASELINE.json's metric (``configs[2]``: MaPLe depth 9, batch
32 per GPU); ``secondary`` holds the same measurements for the other BASELINE configurations (VPT-deep B=64, CoOp B=4,
CRIS+CoCoOp B=32 @416, and the batch-256 eval step with the fused Dice/IoU kernel), each through the same code path.

  value     whole-job images/s with the batch already resident in HBM (one CUDA-graph replay per step)
  e2e       the same step through the public module API (``ImageTextMaskModule.training_step``) from PINNED HOST buffers:
            host->device copies of image / mask / ids and the device->host read of the loss inside the timed region
  roofline  the tensor-bound kernel with the largest share of the step; ``roofline_hbm`` the HBM-bound one.  Durations come
            from CUDA-event NODES inside an instrumented capture of the same step (cudaEventRecordExternal around every
            kernel of our library), averaged over replays - i.e. measured inside the graph, next to the second stream, not
            from an eager pass
  cpu_baseline   the CPU oracle (a port of the reference's fp32 eager path) on the host cores, bounded sample, N=1 only
``--impl reference`` times the reference's own classes (byte-compiled from /root/reference into oracle/_ref by
oracle/build_ref.py, run through oracle/ref_shim.py) on the host cores; the oracle port is the fallback.
"""
from __future__ import annotations

import argparse
import gc
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

GFLOP_PER_IMG = {"maple": 166.6,         # algorithmic fwd + dgrad-only bwd, counted on the reference modules
                 "cris_cocoop": 211.5}   # SURVEY.md section 8d cfg4 (FlopCounterMode on the reference COOPCRIS; RN50 forward 42.7)

WORKLOADS = {
    # name: (BASELINE.json config, description, default batch per GPU, image size, metric name)
    "maple": ("configs[2]", "CLIPSeg ViT-B/16 + MaPLe multimodal prompts (depth 9, 4 ctx), 352x352, batch {B}/GPU, 1 binary class, text L=8", 32, 352,
              "train img/s CLIPSeg+MaPLe 352x352"),
    "vpt": ("configs[1]", "CLIPSeg ViT-B/16 + VPT-deep visual prompts (8 tokens, depth 12), 352x352, batch {B}/GPU, 1 binary class, text L=8", 64, 352,
            "train img/s CLIPSeg+VPT-deep 352x352"),
    "coop": ("configs[0]", "CLIPSeg ViT-B/16 + CoOp textual prompts (4 ctx tokens, depth 1), 352x352, batch {B}/GPU, 1 binary class, text L=8", 4, 352,
             "train img/s CLIPSeg+CoOp 352x352"),
    "cris_cocoop": ("configs[3]", "CRIS CLIP-RN50 + CoCoOp (per-image meta-net context, depth 1, 4 ctx), 416x416, batch {B}/GPU, 1 binary class, text L=8",
                    32, 416, "train img/s CRIS+CoCoOp 416x416"),
    "eval256": ("configs[4] (eval half)", "CRIS CLIP-RN50 + CoCoOp validation step (no grad) with the fused Dice/IoU counters, 416x416, batch {B}/GPU", 256, 416,
                "eval img/s CRIS 416x416 batch 256"),
}
SECONDARY = ["vpt", "coop", "cris_cocoop", "eval256"]
L2_NOTE = "per-step working set (>2 GB of activations) exceeds the 126 MB L2; no explicit flush"


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


def random_clipseg(seed=0):
    """Random-init CLIPSeg of the CIDAS/clipseg-rd64 geometry (no checkpoint exists in this image)."""
    from transformers import CLIPSegConfig, CLIPSegForImageSegmentation

    torch.manual_seed(seed)
    cfg = CLIPSegConfig(vision_config=dict(image_size=352, patch_size=16), projection_dim=512, reduce_dim=64)
    hf = CLIPSegForImageSegmentation(cfg).eval()
    with torch.no_grad():      # HF inits LayerNorm affine to (1, 0) and biases to 0: make those terms non-trivial
        for name, p in hf.named_parameters():
            if name.endswith("bias"):
                p.normal_(0, 0.02)
            elif "layer_norm" in name or "layernorm" in name or "layrnorm" in name:
                p.add_(torch.randn_like(p) * 0.05)
    return hf


def _wrap_module(net, device, lr):
    from functools import partial

    from tunevlseg_b200.losses import DiceCELoss
    from tunevlseg_b200.models.image_text_mask_module import ImageTextMaskModule
    from tunevlseg_b200.optim import FusedAdamW

    module = ImageTextMaskModule(net=net, loss_fn=DiceCELoss(sigmoid=True, lambda_dice=1, lambda_ce=0.2),
                                 optimizer=partial(FusedAdamW, lr=lr, weight_decay=0.0), scheduler=None, compile=False,
                                 task="binary", threshold=0.5, weight_decay=0.0)
    module = module.to(device)
    module.setup("fit")
    return module, module.configure_optimizers()["optimizer"]


def build_module(device, kind="maple", seed=0):
    """Random-init CLIPSeg + the named learner inside the reference-shaped LightningModule
    (configs/model/{maple_clipseg,vpt_clipseg,coop/clipseg}.yaml)."""
    from functools import partial

    from tunevlseg_b200.models.core_models.coop import COOPCLIPSeg, MapleCLIPSeg, VPTCLIPSeg
    from tunevlseg_b200.models.core_models.coop.context_learner import CoOpContextLearner, MapleContextLearner, VPTContextLearner

    hf = random_clipseg(seed)
    cls, learner = {
        "maple": (MapleCLIPSeg, partial(MapleContextLearner, prompt_depth=9, num_context=4, intermediate_dim=64, use_proj_norm=True,
                                        use_unified_projection=False, use_lora_proj=False, context_initializer=None)),
        "vpt": (VPTCLIPSeg, partial(VPTContextLearner, prompt_depth=12, num_context=8)),
        "coop": (COOPCLIPSeg, partial(CoOpContextLearner, prompt_depth=1, num_context=4, context_initializer=None)),
    }[kind]
    net = cls(model_cfg=dict(pretrained_model_name_or_path=hf, freeze_encoder=False, freeze_decoder=False), context_learner=learner,
              freeze_all=True, no_freeze_last_layer=False, use_new_last_layer=True, new_last_layer_kernel_size=5, residual_ratio=0.5)
    return _wrap_module(net, device, 2e-4)


def build_module_cris(device, seed=0):
    """Random-init CRIS (CLIP-RN50 geometry of RN50.pt, configs/model/cocoop/cris.yaml) + CoCoOp learner."""
    from functools import partial
    from types import SimpleNamespace

    from tunevlseg_b200.models.components.cris_model import CLIP
    from tunevlseg_b200.models.core_models.coop import COOPCRIS
    from tunevlseg_b200.models.core_models.coop.context_learner import CoCoOpContextLearner

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
    return _wrap_module(net, device, 2e-5)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
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
# CPU arms: the oracle port (cpu_baseline) and the reference's own classes (--impl reference)
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


def reference_step_factory(B):
    """The reference's OWN MapleCLIPSeg (src/models/core_models/coop/maple_clipseg.py, unmodified, imported from
    /root/reference or from its byte-compiled build oracle/_ref) on the CPU: forward, Dice+BCE, backward, AdamW.
    monai / torchmetrics / lightning have no wheel here, so the loss and the metric counters around the net are the
    oracle's restatement (kind "reference+shim").  Returns None when the reference cannot be imported."""
    from oracle import ref_shim

    path = ref_shim.reference_path()
    if path is None:
        return None
    import tempfile
    from functools import partial

    if path not in sys.path:
        sys.path.insert(0, path)
    ref_shim.install_shim()
    from src.models.core_models.coop import MapleCLIPSeg
    from src.models.core_models.coop.context_learner import MapleContextLearner

    from oracle import loss_metrics as OLM

    d = tempfile.mkdtemp(prefix="tvs_ref_hf_")
    random_clipseg(0).save_pretrained(d)
    net = MapleCLIPSeg(model_cfg=dict(pretrained_model_name_or_path=d, freeze_encoder=False, freeze_decoder=False),
                       context_learner=partial(MapleContextLearner, prompt_depth=9, num_context=4, intermediate_dim=64, use_proj_norm=True,
                                               use_unified_projection=False, use_lora_proj=False, context_initializer=None),
                       freeze_all=True, no_freeze_last_layer=False, use_new_last_layer=True, new_last_layer_kernel_size=5, residual_ratio=0.5)
    net.train()
    opt = torch.optim.AdamW([p for p in net.parameters() if p.requires_grad], lr=2e-4, weight_decay=0.0)
    batch = synth_batch(B)

    def step():
        opt.zero_grad(set_to_none=True)
        logits = net(text_input={"input_ids": batch["input_ids"], "attention_mask": batch["attention_mask"]}, image_input=batch["image"])
        loss = OLM.dice_ce_loss(logits, batch["mask"])
        OLM.metric_counts(torch.sigmoid(logits.detach()), batch["mask"])
        loss.backward()
        opt.step()
        return float(loss.detach())

    return step


def time_cpu_step(step, B, steps, warmup):
    for _ in range(warmup):
        step()
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        ts.append(time.perf_counter() - t0)
    return B / statistics.median(ts), statistics.median(ts) * 1e3


def host_threads():
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    torch.set_num_threads(n)          # torchrun exports OMP_NUM_THREADS=1: undo it for the host arm
    return torch.get_num_threads()


def workload_config(name, B, world):
    return {"workload": WORKLOADS[name][1].format(B=B), "baseline_config": WORKLOADS[name][0], "global_batch": world * B,
            "parallelism": f"dp{world}", "l2": L2_NOTE}


def run_reference(args, rank, world):
    """The reference arm: the reference's own CPU implementation of the path on the host cores (rank 0 only)."""
    if rank != 0:
        return
    name = args.workload
    if name not in ("maple", "cris_cocoop"):
        print(json.dumps({"impl": "reference", "unavailable": f"no CPU arm for workload {name}"}))
        return
    cores = host_threads()
    B_cfg = args.batch or WORKLOADS[name][2]
    B = args.ref_batch
    kind, step = "port", None
    if name == "maple" and os.environ.get("TVS_REF_ARM", "reference") != "port":
        try:
            step = reference_step_factory(B)
            kind = "reference+shim" if step is not None else "port"
        except Exception as e:  # noqa: BLE001 - fall back to the port, say why
            sys.stderr.write(f"[bench] reference classes unavailable ({type(e).__name__}: {e}); timing the oracle port\n")
            step = None
    if step is None:
        step = oracle_step_factory_cris(B) if name == "cris_cocoop" else oracle_step_factory(B)
    v, ms = time_cpu_step(step, B, max(1, args.steps), min(args.warmup, 1))
    what = ("the reference's own MapleCLIPSeg (oracle/_ref build of /root/reference, transformers-5 shim; loss / counters: oracle)"
            if kind == "reference+shim" else "the fp32 CPU oracle (port of the reference's eager path)")
    print(json.dumps({
        "impl": "reference", "metric": WORKLOADS[name][4], "value": round(v, 3), "unit": "img/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": round(ms, 1), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic (seed 12345), random-init weights",
        "config": workload_config(name, B_cfg, world),
        "cpu_baseline": {"value": round(v, 3), "unit": "img/s", "cores": cores, "kind": kind,
                         "sample": f"{args.steps} full train steps (fwd + loss + metrics + bwd + AdamW) of {what} at batch {B} per step "
                                   f"(bounded sample of the batch-{B_cfg} workload)"},
        "e2e": {"value": round(v, 3), "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


# ------------------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------------------
class Ctx:
    def __init__(self, args):
        self.args = args
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.device = torch.device("cuda", self.local_rank)

    def barrier(self):
        torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(self, fn, steps):
        """EXACTLY ``steps`` calls between two CUDA events, barrier + synchronize on both sides, MAX over ranks."""
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        self.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=self.device)
        if self.world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / steps


def summarise_kernels(recs):
    """recs: (name, key, e0, e1, flops, bytes) per op of ONE step (the events are re-stamped by every replay of an instrumented
    graph) -> {kernel: (total_ms, flops, bytes, launches)} of that step."""
    agg = {}
    for name, key, a, b, fl, by, *_ in recs:
        k = f"{name}:{key}" if key else name
        t, f, y, c = agg.get(k, (0.0, 0.0, 0.0, 0))
        agg[k] = (t + a.elapsed_time(b), f + fl, y + by, c + 1)
    return agg


def rooflines(agg_runs, timing):
    """agg_runs: per-replay dicts from summarise_kernels.  Picks the tensor-bound and the HBM-bound kernel with the largest
    share of the step and reports achieved rate / measured peak."""
    pk = peaks()
    keys = agg_runs[0].keys()
    mean = {k: (statistics.mean(r[k][0] for r in agg_runs), agg_runs[0][k][1], agg_runs[0][k][2], agg_runs[0][k][3]) for k in keys}
    total = sum(v[0] for v in mean.values())
    top = sorted(mean.items(), key=lambda kv: -kv[1][0])
    out = {"table": top, "total_ms": total}

    def entry(k, v, tensor):
        t, f, y, c = v
        fam = k.split(":")[0]
        fam_t = sum(vv[0] for kk, vv in mean.items() if kk.split(":")[0] == fam)
        if tensor:
            ach, peak, unit, src = f / (t * 1e-3) / 1e12, pk["tf_sustained"], "TFLOP/s", pk["src"] + " (sustained cuBLAS bf16)"
        else:
            ach, peak, unit, src = y / (t * 1e-3) / 1e9, pk["hbm"], "GB/s", pk["src"] + " (copy bandwidth)"
        e = {"bound": "tensor" if tensor else "hbm", "kernel": k, "achieved": round(ach, 1), "peak": peak, "unit": unit, "frac": round(ach / peak, 4),
             "traffic": None, "peak_source": src, "launches_per_step": c, "avg_launch_us": round(1e3 * t / c, 1),
             "share_of_kernel_time": round(t / total, 3), "family_share_of_kernel_time": round(fam_t / total, 3), "timing": timing}
        if tensor:
            e["algorithmic_gflop_per_launch"] = round(f / c / 1e9, 2)
        else:
            e["algorithmic_mb_per_launch"] = round(y / c / 1e6, 2)
        tp = os.path.join(ROOT, "profiles", "r02_roofline_traffic.json")
        if os.path.exists(tp):      # dram__bytes_read + dram__bytes_write per launch from the committed ncu --set full capture of THIS kernel
            tj = json.load(open(tp))
            if k in tj:
                e["traffic"] = tj[k]["dram_bytes_per_launch"]
                e["traffic_source"] = tj[k].get("source")
        return e

    # the dominant TENSOR kernel is a kernel INSTANCE (template instantiation): the GEMM launches of one instance are pooled
    # over their shapes and epilogues - summed algorithmic FLOPs over summed duration
    inst = {}
    for k, (t, f, y, c) in mean.items():
        if f <= 0:
            continue
        name, _, key = k.partition(":")
        ik, label = k, key
        if name == "gemm" and "|" in key:          # key = "<M>x<N>x<K>[_fmt]|<tile>x<stages> <kind> epi:<epilogue> cta_group::<n>"
            shape, variant = key.split("|")
            tile_kind, _, rest = variant.partition(" epi:")
            epi, _, cta = rest.partition(" ")
            ik, label = f"gemm_bf16_tcgen05_kernel<{tile_kind} {cta}>", f"{shape} {epi}"
        tt, ff, yy, cc, shapes = inst.get(ik, (0.0, 0.0, 0.0, 0, []))
        inst[ik] = (tt + t, ff + f, yy + y, cc + c, shapes + [(label, c, round(1e3 * t / c, 1), round(f / (t * 1e-3) / 1e12, 1))])
    # the dominant kernel: largest time among the instances that carry a real share (>= 10 %) of the step's algorithmic FLOPs
    # (the ~250 launch-bound text-tower GEMMs add up to a lot of overlapped time but hold < 1 % of the FLOPs)
    flops_total = sum(v[1] for v in inst.values())
    heavy = {k: v for k, v in inst.items() if v[1] >= 0.10 * flops_total} or inst
    t_k = max(heavy.items(), key=lambda kv: kv[1][0]) if inst else None
    # ... and the same rule for the HBM-bound pick: among the kernels that move a real share (>= 10 %) of the step's algorithmic
    # bytes, the one with the largest time (the launch-bound [B*L, 512] text-tower LayerNorms on the second stream wait for SMs:
    # a lot of overlapped time, < 1 % of the bytes)
    hbm = [kv for kv in top if kv[1][2] > 0 and kv[1][1] == 0]
    bytes_total = sum(kv[1][2] for kv in hbm)
    h_k = next((kv for kv in hbm if kv[1][2] >= 0.10 * bytes_total), hbm[0] if hbm else None)
    out["tensor"] = None
    if t_k:
        e = entry(t_k[0], t_k[1][:4], True)
        e["family_share_of_kernel_time"] = round(sum(v[0] for kk, v in inst.items() if kk.split("<")[0] == t_k[0].split("<")[0]) / total, 3)
        if t_k[0].startswith("gemm_bf16_tcgen05_kernel<"):
            e["launches"] = [{"shape_epilogue": sh, "n": n, "avg_us": us, "tflops": tf} for sh, n, us, tf in sorted(t_k[1][4], key=lambda r: -r[1] * r[2])]
        out["tensor"] = e
    out["hbm"] = entry(*h_k, False) if h_k else None
    return out


def run_train_workload(ctx: Ctx, name: str, steps: int, warmup: int, headline: bool):
    from tunevlseg_b200 import abi
    from tunevlseg_b200.graph import GraphedTrainStep

    args, device, rank, world = ctx.args, ctx.device, ctx.rank, ctx.world
    B = (args.batch if (headline and args.batch) else WORKLOADS[name][2])
    size = WORKLOADS[name][3]
    cris = name == "cris_cocoop"
    module, opt = build_module_cris(device, seed=0) if cris else build_module(device, name, seed=0)
    module.train()
    host = synth_batch(B, seed=12345 + rank, pinned=True, size=size)
    resident = {k: v.to(device) for k, v in host.items()}
    h2d = sum(v.numel() * v.element_size() for v in host.values())

    def step_eager():
        opt.zero_grad()
        loss = module.training_step(resident, 0)
        loss.backward()
        opt.step()
        return loss

    use_graph = not args.no_graph
    graphed = None
    if use_graph:
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

    for _ in range(warmup):
        step_resident()
    sampler = ClockSampler(ctx.local_rank)
    if rank == 0 and headline:
        sampler.start()
    n0 = abi.launch_count()
    ms_step = ctx.timed(step_resident, steps)
    launches = abi.launch_count() - n0
    if use_graph:      # replays do not pass through the library's launch counter: kernels per captured step x steps
        launches = launches_per_step * steps
    clocks = sampler.stop() if (rank == 0 and headline) else None

    for _ in range(2):
        step_e2e()
    ms_e2e = ctx.timed(step_e2e, steps)
    ctx.barrier()

    # ---- per-kernel durations INSIDE the graph: a second, instrumented capture of the same step whose kernels are bracketed
    # by event-record nodes (cudaEventRecordExternal); every rank captures and replays it (it contains the all-reduce)
    roof = None
    reps = 5
    timing = "CUDA-event record nodes around every library kernel inside an instrumented capture of the step, mean of %d replays" % reps
    recs: list = []
    runs = []
    try:
        if not use_graph or os.environ.get("TVS_BENCH_PROFILE", "graph") != "graph":
            raise RuntimeError("graph-level profiling switched off")
        abi.set_profiler(recs)
        try:
            probe = GraphedTrainStep(module, opt, resident, warmup=0)
        finally:
            abi.set_profiler(None)
        for _ in range(2):
            probe()
        torch.cuda.synchronize()
        t_probe = []
        for _ in range(reps):
            p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            p0.record()
            probe()
            p1.record()
            torch.cuda.synchronize()
            t_probe.append(p0.elapsed_time(p1))
            runs.append(summarise_kernels(recs))
        probe_ms = statistics.mean(t_probe)
        timing += "; the instrumented replay takes %.3f ms against %.3f ms for the plain graph" % (probe_ms, ms_step)
        if rank == 0 and headline and args.timeline:
            base = min(recs, key=lambda r: recs[0][2].elapsed_time(r[2]))[2]
            streams = {}
            with open(args.timeline, "w") as fh:
                fh.write("idx,stream,name,key,start_us,end_us\n")
                for i, (nm, key, a, b, *_rest) in enumerate(recs):
                    sid = streams.setdefault(_rest[2] if len(_rest) > 2 else 0, len(streams))
                    fh.write(f"{i},{sid},{nm},{key},{1e3 * base.elapsed_time(a):.1f},{1e3 * base.elapsed_time(b):.1f}\n")
        probe.graph.reset()
        del probe
    except Exception as e:  # noqa: BLE001 - fall back to one eager pass behind a device-side spin
        if rank == 0:
            sys.stderr.write(f"[bench] instrumented capture unavailable ({type(e).__name__}: {e}); timing one eager pass instead\n")
        abi.set_profiler(None)
        recs = []
        timing = "CUDA events around every launch of one eager step (host enqueued ahead of the device)"
        abi.set_profiler(recs)
        torch.cuda._sleep(150_000_000)
        try:
            step_eager()
        finally:
            abi.set_profiler(None)
        torch.cuda.synchronize()
        runs = [summarise_kernels(recs)]
    if rank == 0 and runs:
        roof = rooflines(runs, timing)
        if args.profile_kernels:
            sys.stderr.write(f"--- {name}: kernel time per step {roof['total_ms']:.3f} ms (graph step {ms_step:.3f} ms)\n")
            for k, (t, f, y, c) in roof["table"][:70]:
                rate = f"{f / max(t, 1e-9) / 1e9:8.1f} TFLOP/s" if f else (f"{y / max(t, 1e-9) / 1e6:8.1f} GB/s" if y else "")
                sys.stderr.write(f"{k:72s} n={c:4d} {t:9.3f} ms {100 * t / roof['total_ms']:5.1f}%  {rate}\n")

    out = None
    if rank == 0:
        value = world * B / (ms_step * 1e-3)
        e2e = world * B / (ms_e2e * 1e-3)
        pk = peaks()
        out = {"metric": WORKLOADS[name][4], "value": round(value, 1), "unit": "img/s", "n_gpus": world, "steps": steps, "warmup": warmup,
               "ms_per_step": round(ms_step, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": "tf32" if cris else "f16", "data": "synthetic (seed 12345), random-init weights",
               "config": workload_config(name, B, world),
               "execution": {"cuda_graph": use_graph,
                             "precision": ("fp32 activations, kind::tf32 tcgen05 GEMMs with round-to-nearest operands, fp32 text attention" if cris else
                                           "vision tower: fp16 forward operands / bf16 gradients on kind::f16 tcgen05 GEMMs, fp32 accumulate and fp32 "
                                           "residual stream; text tower and decoder: kind::tf32 GEMMs on fp32 activations")},
               "roofline": roof["tensor"] if roof else None, "roofline_hbm": roof["hbm"] if roof else None,
               "kernel_time_ms_per_step": round(roof["total_ms"], 3) if roof else None,
               "e2e": {"value": round(e2e, 1), "unit": "img/s", "ms_per_step": round(ms_e2e, 3), "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
               "gpu_launches": int(launches), "clocks": clocks}
        if name in GFLOP_PER_IMG:
            out["model_tflops_per_gpu"] = round(value / world * GFLOP_PER_IMG[name] / 1e3, 1)
            out["model_flops_frac_of_peak"] = round(value / world * GFLOP_PER_IMG[name] / 1e3 / pk["tf_sustained"], 4)
    # free everything of this workload (the captured graph first: it pins its memory pool and, at N > 1, NCCL work)
    if graphed is not None:
        graphed.graph.reset()
    del graphed, module, opt, resident, host, step_resident
    gc.collect()
    torch.cuda.empty_cache()
    return out


def run_eval_workload(ctx: Ctx, name: str, steps: int, warmup: int):
    """configs[4], the part that exists in the reference: a batch-256 validation step - CRIS forward
    under no_grad + ONE pass of the fused Dice+BCE / Dice / IoU kernel - through ``ImageTextMaskModule.validation_step``."""
    from tunevlseg_b200 import abi

    device, rank, world = ctx.device, ctx.rank, ctx.world
    B, size = WORKLOADS[name][2], WORKLOADS[name][3]
    module, _ = build_module_cris(device, seed=0)
    module.eval()
    host = synth_batch(B, seed=12345 + rank, pinned=True, size=size)
    resident = {k: v.to(device) for k, v in host.items()}
    h2d = sum(v.numel() * v.element_size() for v in host.values())

    @torch.no_grad()
    def step_resident():
        module.validation_step(resident, 1)

    @torch.no_grad()
    def step_e2e():
        batch = {k: v.to(device, non_blocking=True) for k, v in host.items()}
        module.validation_step(batch, 1)
        return float(module.logged["val_loss"])          # device->host read of the step's loss (syncs)

    for _ in range(warmup):
        step_resident()
    n0 = abi.launch_count()
    ms_step = ctx.timed(step_resident, steps)
    launches = abi.launch_count() - n0
    step_e2e()
    ms_e2e = ctx.timed(step_e2e, steps)
    recs: list = []
    runs = []
    for _ in range(3):
        recs.clear()
        abi.set_profiler(recs)
        torch.cuda._sleep(200_000_000)          # the host enqueues ahead of the device: no submission gaps inside the event pairs
        try:
            step_resident()
        finally:
            abi.set_profiler(None)
        torch.cuda.synchronize()
        runs.append(summarise_kernels(list(recs)))
    d, i = module.val_dice.compute(), module.val_iou.compute()          # metric state synchronisation across ranks (N > 1)
    out = None
    if rank == 0:
        roof = rooflines(runs, "CUDA events around every launch of an eager eval step (host enqueued ahead of the device), mean of 3 passes")
        value, e2e = world * B / (ms_step * 1e-3), world * B / (ms_e2e * 1e-3)
        # the fused loss / metric kernel on its own: 8 B per pixel (logits + mask read once)
        loss_k = next(((k, v) for k, v in roof["table"] if k.startswith("dicebce_metrics_fwd")), None)
        hbm = None
        if loss_k is not None:
            t, _, y, c = loss_k[1]
            pk = peaks()
            hbm = {"bound": "hbm", "kernel": "dicebce_metrics_fwd (dicebce_partial_kernel + finalize)", "achieved": round(y / (t * 1e-3) / 1e9, 1),
                   "peak": pk["hbm"], "unit": "GB/s", "frac": round(y / (t * 1e-3) / 1e9 / pk["hbm"], 4), "traffic": None,
                   "algorithmic_mb_per_launch": round(y / c / 1e6, 2), "avg_launch_us": round(1e3 * t / c, 1), "peak_source": pk["src"] + " (copy bandwidth)"}
        out = {"metric": WORKLOADS[name][4], "value": round(value, 1), "unit": "img/s", "n_gpus": world, "steps": steps, "warmup": warmup,
               "ms_per_step": round(ms_step, 3), "config": workload_config(name, B, world), "roofline": roof["tensor"], "roofline_hbm": hbm,
               "e2e": {"value": round(e2e, 1), "unit": "img/s", "ms_per_step": round(ms_e2e, 3), "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
               "gpu_launches": int(launches), "val_dice": round(float(d), 6), "val_iou": round(float(i), 6)}
    del module, resident, host
    gc.collect()
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=0, help="images per GPU of the headline workload (default: the BASELINE config's)")
    ap.add_argument("--ref-batch", type=int, default=4, help="batch of the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="measure only the headline workload")
    ap.add_argument("--profile-kernels", action="store_true", help="also print the per-kernel share table to stderr")
    ap.add_argument("--no-graph", action="store_true", help="drive every kernel launch from Python instead of one CUDA graph")
    ap.add_argument("--timeline", default="", help="write the per-kernel timeline of one instrumented replay of the headline step (CSV: stream, start, end)")
    ap.add_argument("--workload", default="maple", choices=list(WORKLOADS),
                    help="headline workload: maple = BASELINE.json's metric (default); the others are also reported under `secondary`")
    args = ap.parse_args()

    ctx = Ctx(args)
    rank, world = ctx.rank, ctx.world
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.warmup < 3:
        args.warmup = 3
    from tunevlseg_b200 import abi

    torch.cuda.set_device(ctx.local_rank)
    if world > 1:
        from datetime import timedelta

        dist.init_process_group("nccl", device_id=ctx.device, timeout=timedelta(seconds=300))
    abi.require_device()

    def run(name, steps, warmup, headline):
        if name == "eval256":
            return run_eval_workload(ctx, name, steps, warmup)
        return run_train_workload(ctx, name, steps, warmup, headline)

    out = run(args.workload, args.steps, args.warmup, True)
    secondary = []
    if not args.no_secondary and args.workload == "maple":
        sec_steps = max(3, min(args.steps, 10))
        for name in SECONDARY:
            try:
                res = run(name, sec_steps, 3, False)
            except Exception as e:  # noqa: BLE001 - a secondary workload must never take the headline line down with it
                res = {"metric": WORKLOADS[name][4], "error": f"{type(e).__name__}: {e}"[:300]} if rank == 0 else None
                torch.cuda.synchronize()
            if rank == 0:
                secondary.append(res)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and args.workload in ("maple", "cris_cocoop"):
        cores = host_threads()
        step = oracle_step_factory_cris(args.ref_batch) if args.workload == "cris_cocoop" else oracle_step_factory(args.ref_batch)
        v, ms = time_cpu_step(step, args.ref_batch, 3, 1)
        cpu = {"value": round(v, 3), "unit": "img/s", "cores": cores, "kind": "port",
               "sample": f"3 full train steps of the fp32 CPU oracle at batch {args.ref_batch} (median {ms:.0f} ms/step)"}

    if rank == 0:
        out["cpu_baseline"] = cpu
        if secondary:
            out["secondary"] = secondary
        print(json.dumps(out))
        sys.stdout.flush()
    if world > 1:
        # every captured graph was reset() above (they held NCCL work on the communicator); now the group can be torn down
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
