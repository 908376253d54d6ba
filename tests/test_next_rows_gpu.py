"""SURVEY.md §8(f) rows 3 and 4 on the GPU: the device-side scene generator (input pipeline) and checkpoint
save -> --resume (the reference only saves: train.py:85-90)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("shape", [(1, 128, 128), (3, 64, 64)])
def test_device_scene_generator_on_cuda(shape):
    """Schema of reference dataloader.py:23-33 / metric.py:21-22 (image in [0,1], bbox (x,y,w,h) px inside the canvas,
    count = number of live boxes) and determinism under a seeded device generator."""
    from spair_pytorch_b200.dataloader import scattered_sprites_gpu
    dev = torch.device("cuda")
    B, S = 16, 9

    def draw(seed):
        return scattered_sprites_gpu(B, shape, dev, torch.Generator(device=dev).manual_seed(seed), max_sprites=S)

    img, bbox, count = draw(5)
    assert img.is_cuda and img.shape == (B,) + shape and img.dtype == torch.float32
    assert bbox.shape == (B, S, 4) and count.shape == (B, 1)
    assert float(img.min()) >= 0.0 and float(img.max()) <= 1.0 and float(img.max()) > 0.5
    live = (bbox[..., 2] > 0).sum(1, keepdim=True).float()
    assert torch.equal(live, count) and int(count.min()) >= 1 and int(count.max()) <= S
    x0, y0, w, h = bbox.unbind(-1)
    assert bool(((x0 >= 0) & (y0 >= 0) & (x0 + w <= shape[2]) & (y0 + h <= shape[1])).all())
    # every live sprite leaves ink inside its box; nothing is drawn outside the union of the boxes
    mask = torch.zeros(B, shape[1], shape[2], device=dev, dtype=torch.bool)
    for b in range(B):
        for s in range(int(count[b])):
            xa, ya, ww, hh = (int(v) for v in bbox[b, s])
            assert float(img[b, :, ya:ya + hh, xa:xa + ww].max()) > 0.1
            mask[b, ya:ya + hh, xa:xa + ww] = True
    assert float((img.amax(1) * (~mask).float()).max()) == 0.0
    img2, bbox2, count2 = draw(5)
    assert torch.equal(img, img2) and torch.equal(bbox, bbox2) and torch.equal(count, count2)
    assert not torch.equal(img, draw(6)[0])


def _train(args, env):
    cmd = [sys.executable, os.path.join(ROOT, "train_dp.py")] + args
    res = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    return res.stdout


@pytest.mark.parametrize("mode", ["graph", "eager"])
def test_checkpoint_resume_equals_uninterrupted_run(tmp_path, mode):
    """6 steps in one go == 4 steps, checkpoint, `--resume`, 3 more steps — bitwise, with SPAIR_DETERMINISTIC=1 (model +
    Adam state + step travel in the checkpoint; noise and scenes are functions of (seed, rank, step))."""
    env = dict(os.environ, SPAIR_DETERMINISTIC="1", PYTHONPATH=ROOT)
    env.pop("RANK", None)
    common = ["--batch", "8", "--log-every", "1000", "--ckpt-every", "3"] + (["--eager"] if mode == "eager" else [])
    a, b = str(tmp_path / "a"), str(tmp_path / "b")
    os.makedirs(a), os.makedirs(b)
    _train(common + ["--start-step", "1000", "--steps", "6", "--ckpt-dir", a, "--save-final", os.path.join(a, "final.pt")], env)
    _train(common + ["--start-step", "1000", "--steps", "4", "--ckpt-dir", b], env)
    assert os.path.exists(os.path.join(b, "step_1002.pt"))
    out = _train(common + ["--resume", "--steps", "3", "--ckpt-dir", b, "--save-final", os.path.join(b, "final.pt")], env)
    assert "resumed from step 1002" in out
    sa, sb = torch.load(os.path.join(a, "final.pt")), torch.load(os.path.join(b, "final.pt"))
    assert list(sa) == list(sb)
    moved = 0
    ck = torch.load(os.path.join(b, "step_1002.pt"))["model"]
    for k in sa:
        assert torch.equal(sa[k], sb[k]), "parameter %s differs after resume" % k
        moved += int(not torch.equal(sa[k].cpu(), ck[k].cpu()))
    assert moved > 40          # the resumed run really trained on (every network moved after the checkpoint)
