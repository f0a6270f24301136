"""World-size-2 checks on the CPU (gloo): the N > 1 host logic that does not need a GPU -
metric state synchronisation (confusion matrix: sum-reduce; per-sample Dice scores: sum + count, unequal per-rank sample counts) and the reference arm's
rank protocol under torchrun (rank 0 measures and prints one JSON line, other ranks exit 0 silently)."""
import json
import os
import subprocess
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _updates(rank):
    """UNEQUAL per-rank sample counts (the last validation batch is not dropped, image_text_mask_datamodule.py:40-47):
    rank 0 sees 3 batches of 4, rank 1 sees 2 batches of 4 and one of 1."""
    g = torch.Generator().manual_seed(100 + rank)
    for k in range(3):
        n = 1 if (rank == 1 and k == 2) else 4
        counts = torch.randint(0, 100, (n, 3), generator=g)
        counts[0] = 0                                    # an empty sample: zero_division -> 1
        yield counts, torch.randint(0, 1000, (4,), generator=g)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tunevlseg_b200.metrics import Dice, JaccardIndex

    dice, iou = Dice(threshold=0.5, zero_division=1, average="samples"), JaccardIndex(task="binary", threshold=0.5, zero_division=1)
    for counts, conf in _updates(rank):
        dice.update_from_counts(counts)
        iou.update_from_confmat(conf)
    d, i = dice.compute(), iou.compute()
    if rank == 0:
        torch.save({"dice": d, "iou": i}, out)
    dist.destroy_process_group()


def test_metric_sync_world2(tmp_path):
    from tunevlseg_b200.metrics import Dice, JaccardIndex

    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, 29531, out), nprocs=2, join=True)
    got = torch.load(out)
    # single-process reference over the union of both ranks' updates
    dice, iou = Dice(threshold=0.5, zero_division=1, average="samples"), JaccardIndex(task="binary", threshold=0.5, zero_division=1)
    for rank in range(2):
        for counts, conf in _updates(rank):
            dice.update_from_counts(counts)
            iou.update_from_confmat(conf)
    assert torch.allclose(got["dice"], dice.compute(), atol=1e-6)
    assert torch.allclose(got["iou"], iou.compute(), atol=1e-6)


def test_reference_arm_under_torchrun_world2():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29537", os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
           "--warmup", "0", "--ref-batch", "1"]
    res = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, res.stdout
    rec = json.loads(lines[0])
    assert rec["impl"] == "reference" and rec["n_gpus"] == 2 and rec["value"] > 0
    # "reference+shim": the reference's own MapleCLIPSeg (from /root/reference here, from its byte-compiled build oracle/_ref
    # on the GPU box); "port": the oracle, when neither is importable
    assert rec["cpu_baseline"]["kind"] in ("reference+shim", "port") and rec["e2e"]["h2d_bytes_per_step"] == 0
    assert rec["config"]["workload"].startswith("CLIPSeg ViT-B/16 + MaPLe") and "batch 32/GPU" in rec["config"]["workload"]


def test_reference_arm_runs_from_the_byte_compiled_build():
    """oracle/build_ref.py byte-compiles the reference into oracle/_ref (what travels to the GPU box, where /root/reference
    does not exist): the reference arm must import the reference's classes from there."""
    import pytest

    from oracle import build_ref

    if build_ref.build() is None and not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "src")):
        pytest.skip("no /root/reference and no prebuilt oracle/_ref")
    env = dict(os.environ, TVS_REF_FORCE_BUILT="1")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--ref-batch", "1"],
                         cwd=ROOT, capture_output=True, text=True, timeout=900, env=env)
    assert res.returncode == 0, res.stderr[-2000:]
    rec = json.loads([l for l in res.stdout.splitlines() if l.startswith("{")][0])
    assert rec["cpu_baseline"]["kind"] == "reference+shim", res.stderr[-1500:]
