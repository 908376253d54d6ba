"""Datasets under the reference's ``spair.dataloader`` names (dataloader.py:10-36).

``SimpleScatteredMNISTDataset`` reads the reference's HDF5 layout ``train/full/{image,bbox,digit_count}``
(needs h5py, imported lazily because the file is not shipped with the reference either).
``ScatteredSpritesDataset`` is the procedural stand-in used for benchmarking."""
import numpy as np
import torch
from torch.utils import data

from . import config as cfg


class SimpleScatteredMNISTDataset(data.Dataset):
    def __init__(self, in_file):
        super().__init__()
        import h5py
        self.dataset = h5py.File(in_file, 'r')['train/full']
        self.episode = None

    def __getitem__(self, index):
        obs = self.dataset['image'][index, ...]
        image = np.moveaxis(obs[..., None], -1, 0)          # (H, W) -> (1, H, W)
        return image, self.dataset['bbox'][index, ...], self.dataset['digit_count'][index, ...]

    def __len__(self):
        return self.dataset['image'].shape[0]


def scattered_sprites(batch, image_shape, seed=1234, max_sprites=9, sprite_px=(10, 20), return_boxes=False):
    """``k ~ U{1..max}`` soft glyph-like blobs per image, max-composited on black, values in [0,1].
    Boxes are (x, y, w, h) in pixels, as ``metric.mAP`` expects (metric.py:21-22)."""
    C, H, W = image_shape
    rng = np.random.RandomState(seed)
    imgs = np.zeros((batch, C, H, W), np.float32)
    boxes = np.zeros((batch, max_sprites, 4), np.float32)
    counts = np.zeros((batch, 1), np.float32)
    for b in range(batch):
        k = rng.randint(1, max_sprites + 1)
        counts[b, 0] = k
        for i in range(k):
            s = rng.randint(sprite_px[0], min(sprite_px[1], H, W) + 1)
            yy, xx = np.mgrid[0:s, 0:s].astype(np.float32) / max(s - 1, 1)
            sprite = np.zeros((s, s), np.float32)
            for _ in range(rng.randint(2, 5)):
                cy, cx, r = rng.uniform(0.2, 0.8), rng.uniform(0.2, 0.8), rng.uniform(0.08, 0.3)
                sprite = np.maximum(sprite, np.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / (2 * r * r)))
            sprite = np.clip(sprite * 1.2, 0, 1)
            y0, x0 = rng.randint(0, H - s + 1), rng.randint(0, W - s + 1)
            colour = np.ones(C, np.float32) if C == 1 else rng.uniform(0.3, 1.0, C).astype(np.float32)
            for c in range(C):
                imgs[b, c, y0:y0 + s, x0:x0 + s] = np.maximum(imgs[b, c, y0:y0 + s, x0:x0 + s], sprite * colour[c])
            boxes[b, i] = (x0, y0, s, s)
    if return_boxes:
        return torch.from_numpy(imgs), torch.from_numpy(boxes), torch.from_numpy(counts)
    return torch.from_numpy(imgs)


class ScatteredSpritesDataset(data.Dataset):
    """Procedural scattered-sprite scenes with the same item schema as the HDF5 dataset."""

    def __init__(self, length=4096, image_shape=None, seed=1234, max_sprites=9, sprite_px=(10, 20)):
        self.length, self.seed = length, seed
        self.image_shape = tuple(image_shape or cfg.INPUT_IMAGE_SHAPE)
        self.max_sprites, self.sprite_px = max_sprites, sprite_px

    def __getitem__(self, index):
        img, box, cnt = scattered_sprites(1, self.image_shape, self.seed + index, self.max_sprites, self.sprite_px, True)
        return img[0].numpy(), box[0].numpy(), cnt[0].numpy()

    def __len__(self):
        return self.length


def scattered_sprites_gpu(batch, image_shape, device, generator=None, max_sprites=9, sprite_px=(10, 20)):
    """Device-side scene generator (no host work, no H2D copy): the same family of scenes as
    ``scattered_sprites`` — ``k ~ U{1..max}`` glyph-like unions of 2-4 Gaussian strokes inside a random
    square, max-composited on black — built with batched tensor ops on ``device``.
    Returns (image [B,C,H,W] in [0,1], bbox [B,max_sprites,4] as (x,y,w,h) px, count [B,1])."""
    C, H, W = image_shape
    S, K = max_sprites, 4
    kw = dict(device=device, generator=generator)
    count = torch.randint(1, S + 1, (batch, 1), **kw)
    size = torch.randint(sprite_px[0], min(sprite_px[1], H, W) + 1, (batch, S), **kw).float()
    x0 = torch.floor(torch.rand(batch, S, **kw) * (W - size + 1))
    y0 = torch.floor(torch.rand(batch, S, **kw) * (H - size + 1))
    live = (torch.arange(S, device=device)[None, :] < count).float()                       # [B,S]
    n_strokes = torch.randint(2, 5, (batch, S, 1), **kw)
    stroke_on = (torch.arange(K, device=device)[None, None, :] < n_strokes).float()        # [B,S,K]
    cy, cx = (0.2 + 0.6 * torch.rand(batch, S, K, **kw) for _ in range(2))
    rad = 0.08 + 0.22 * torch.rand(batch, S, K, **kw)
    ys = torch.arange(H, device=device).float()[None, None, :, None]                       # [1,1,H,1]
    xs = torch.arange(W, device=device).float()[None, None, None, :]                       # [1,1,1,W]
    u = (xs - x0[..., None, None]) / (size[..., None, None] - 1).clamp(min=1)              # sprite-local coords [B,S,1,W]
    v = (ys - y0[..., None, None]) / (size[..., None, None] - 1).clamp(min=1)              # [B,S,H,1]
    inside = ((u >= 0) & (u <= 1)).float() * ((v >= 0) & (v <= 1)).float() * live[..., None, None]   # [B,S,H,W]
    sprite = torch.zeros(batch, S, H, W, device=device)
    for k in range(K):
        d2 = (u - cx[..., k, None, None]) ** 2 + (v - cy[..., k, None, None]) ** 2
        blob = torch.exp(-d2 / (2 * rad[..., k, None, None] ** 2)) * stroke_on[..., k, None, None]
        sprite = torch.maximum(sprite, blob)
    sprite = (sprite * 1.2).clamp(0, 1) * inside
    colour = torch.ones(batch, S, C, device=device) if C == 1 else 0.3 + 0.7 * torch.rand(batch, S, C, **kw)
    image = (sprite[:, :, None] * colour[..., None, None]).amax(dim=1)                     # max-composite -> [B,C,H,W]
    bbox = torch.stack([x0, y0, size, size], dim=-1) * live[..., None]
    return image, bbox, count.float()
