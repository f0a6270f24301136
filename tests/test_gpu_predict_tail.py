"""GPU: the predict / eval tail - fused bicubic resize + PNG quantisation against
torchvision (what src/utils/save_utils.py calls), the save_predictions contract, and the offline metric script against
the monai formulas restated with numpy."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("src,dst", [((352, 352), (480, 640)), ((352, 352), (200, 133)), ((416, 416), (416, 416)), ((64, 64), (301, 7))])
def test_resize_quantise_matches_torchvision(src, dst):
    from torchvision.transforms import functional as TF

    from tunevlseg_b200.utils import resize_to_png_array

    g = torch.Generator(device="cuda").manual_seed(src[0] + dst[1])
    pred = torch.sigmoid(torch.randn(1, *src, device="cuda", generator=g) * 3)
    ours = resize_to_png_array(pred, dst).cpu()
    ref = TF.resize(pred.cpu().float(), size=list(dst), interpolation=TF.InterpolationMode.BICUBIC, antialias=False)
    ref_u8 = ref.mul(255).add_(0.5).clamp_(0, 255).to(torch.uint8)[0]            # torchvision.utils.save_image
    assert ours.shape == ref_u8.shape
    # byte work: bit-exact against the oracle (pinned to ATen's CPU kernel in tests/test_oracle_resize_u8.py) ...
    from oracle import resize_u8 as OR

    want = torch.from_numpy(OR.resize_to_png_array(pred.cpu().numpy(), dst))
    assert torch.equal(ours, want), f"{(ours != want).sum().item()} bytes differ from the oracle"
    # ... and against torchvision on the host, which is what the reference's save_predictions runs
    diff = (ours.int() - ref_u8.int()).abs()
    assert diff.max().item() == 0, f"{(diff != 0).sum().item()} bytes differ from torchvision"


def test_save_predictions_and_eval_script(tmp_path):
    import cv2

    from tunevlseg_b200.scripts.eval_metrics import evaluate
    from tunevlseg_b200.utils import save_predictions

    g = torch.Generator(device="cuda").manual_seed(3)
    shapes = [(120, 90), (120, 90), (77, 200)]
    preds = [torch.sigmoid(torch.randn(1, 64, 64, device="cuda", generator=g) * 4) for _ in shapes]
    preds.append(torch.zeros(1, 64, 64, device="cuda"))                        # empty prediction + empty ground truth -> 1
    shapes.append((50, 50))
    names = ["a/x0.png", "a/x1.png", "b/y0.png", "b/empty.png"]

    class Trainer:
        def predict(self, model, dataloaders, ckpt_path):
            yield {"preds": preds[:2], "mask_name": names[:2], "mask_shape": [torch.tensor(s) for s in shapes[:2]]}
            yield {"preds": preds[2:], "mask_name": names[2:], "mask_shape": shapes[2:]}

    class Log:
        def info(self, *a): pass
        def warning(self, *a): pass

    out = tmp_path / "masks"
    save_predictions({"output_masks_dir": str(out)}, Log(), Trainer(), None, None, None)
    assert sorted(str(p.relative_to(out)) for p in out.rglob("*.png")) == sorted(names)
    save_predictions({"output_masks_dir": str(out)}, Log(), Trainer(), None, None, None)      # exists, no overwrite flag: returns
    img = cv2.imread(str(out / names[0]), cv2.IMREAD_UNCHANGED)
    assert img.shape == (120, 90, 3) and (img[..., 0] == img[..., 1]).all()

    # ground truth = a shifted / noisy version; flat directory layout for the metric script
    seg, gt = tmp_path / "seg", tmp_path / "gt"
    seg.mkdir(); gt.mkdir()
    rng = np.random.default_rng(0)
    expect = {}
    for name, shape in zip(names, shapes):
        pred = cv2.imread(str(out / name), cv2.IMREAD_GRAYSCALE)
        flat = name.replace("/", "_")
        truth = np.zeros(shape, np.uint8) if "empty" in name else ((np.roll(pred, 3, axis=0) > 127) ^ (rng.random(shape) < 0.05)).astype(np.uint8) * 255
        cv2.imwrite(str(seg / flat), pred)
        cv2.imwrite(str(gt / flat), truth)
        p, t = pred > 127, truth > 127
        tp, fp, fn = int((p & t).sum()), int((p & ~t).sum()), int((~p & t).sum())
        dice = 100 * (2 * tp / (2 * tp + fp + fn) if (2 * tp + fp + fn) else 1.0)
        iou = 100 * (tp / (tp + fp + fn) if (tp + fp + fn) else 1.0)
        ones = 100 * (2 * t.sum() / (t.sum() + t.size) if t.size else 1.0)
        expect[flat] = (iou, dice, dice - ones)
    rows = evaluate(seg, gt, 127)
    assert [r["filename"] for r in rows] == sorted(expect)
    for r in rows:
        iou, dice, diff = expect[r["filename"]]
        assert abs(r["iou"] - iou) < 1e-9 and abs(r["dice"] - dice) < 1e-9 and abs(r["ones_dice_diff"] - diff) < 1e-9
    assert dict((r["filename"], r) for r in rows)["b_empty.png"]["dice"] == 100.0
