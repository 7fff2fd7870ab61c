"""Deterministic synthetic Tibetan-Plateau-shaped samples.

Mirrors the tensor layout of the reference's ``CustomDataset``
(``/root/reference/datasets.py:156-174``): per sample
``lr_grace_05 [1, 2h, 2w]``, ``lr_grace_025 [1, 4h, 4w]`` (the "real" field /
regression target) and ``hr_aux [C_aux, 4h, 4w]``, float32, standardised per
channel (the reference caches ``StandardScaler`` outputs,
``datasets.py:409-424``).  ``h x w`` is the generator-input / PAM grid.

Content: aux channels are unit-variance smooth Gaussian random
fields (three 5x5 box passes), the last aux channel is a static "DEM" (seed 99,
identical for every sample, cf. ``datasets.py:377-378``), two channels are
lat/lon ramps (``datasets.py:352-369``); the target is a fixed random linear mix
of the aux channels (seed 7) plus 0.3 x smooth noise, re-standardised;
``lr_grace_05`` is the 2x2 block mean of the target.

Everything is generated on the CPU with seeded ``torch.Generator`` objects so a
sample is identical on every machine; callers move batches to the GPU.
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.nn.functional as F

C_AUX_DEFAULT = 45


def _smooth(x: torch.Tensor, passes: int = 3) -> torch.Tensor:
    c = x.shape[0]
    k = torch.full((c, 1, 5, 5), 1.0 / 25.0, dtype=x.dtype)
    y = x[None]
    for _ in range(passes):
        y = F.conv2d(F.pad(y, (2, 2, 2, 2), mode="replicate"), k, groups=c)
    return y[0]


def _standardise(x: torch.Tensor) -> torch.Tensor:
    m = x.mean(dim=(-2, -1), keepdim=True)
    s = x.std(dim=(-2, -1), keepdim=True, unbiased=False).clamp_min(1e-6)
    return (x - m) / s


def make_sample(idx: int, h: int, w: int, c_aux: int = C_AUX_DEFAULT) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Return (lr_grace_05 [1,2h,2w], lr_grace_025 [1,4h,4w], hr_aux [c_aux,4h,4w])."""
    H, W = 4 * h, 4 * w
    g = torch.Generator().manual_seed(1234 + idx)
    aux = _standardise(_smooth(torch.randn(c_aux, H, W, generator=g)))
    if c_aux >= 3:
        aux[0] = _standardise(torch.linspace(-1, 1, H)[:, None].expand(H, W))
        aux[1] = _standardise(torch.linspace(-1, 1, W)[None, :].expand(H, W))
    dem = _standardise(_smooth(torch.randn(1, H, W, generator=torch.Generator().manual_seed(99))))
    aux[c_aux - 1] = dem[0]
    mix = torch.randn(c_aux, generator=torch.Generator().manual_seed(7)) / (c_aux ** 0.5)
    noise = _standardise(_smooth(torch.randn(1, H, W, generator=g)))
    target = _standardise((mix[:, None, None] * aux).sum(dim=0, keepdim=True) + 0.3 * noise)
    lr05 = F.avg_pool2d(target[None], 2)[0]
    return lr05.contiguous(), target.contiguous(), aux.contiguous()


def make_batch(first_idx: int, batch: int, h: int, w: int, c_aux: int = C_AUX_DEFAULT):
    """Stack ``batch`` consecutive samples: ([B,1,2h,2w], [B,1,4h,4w], [B,c_aux,4h,4w])."""
    parts = [make_sample(first_idx + i, h, w, c_aux) for i in range(batch)]
    return tuple(torch.stack([p[j] for p in parts]) for j in range(3))


def fast_batch(seed: int, batch: int, h: int, w: int, c_aux: int = C_AUX_DEFAULT, device="cpu"):
    """Cheap variant for throughput runs at large shapes: same layout and
    statistics class (smooth unit-variance fields), generated with pooled noise
    instead of per-sample box filtering.  Used by ``bench.py`` only."""
    g = torch.Generator().manual_seed(seed)
    H, W = 4 * h, 4 * w

    def field(c):
        z = torch.randn(batch, c, H // 4 + 2, W // 4 + 2, generator=g)
        z = F.interpolate(z, size=(H, W), mode="bilinear", align_corners=False)
        return (z - z.mean(dim=(2, 3), keepdim=True)) / z.std(dim=(2, 3), keepdim=True).clamp_min(1e-6)

    aux = field(c_aux)
    mix = torch.randn(c_aux, generator=torch.Generator().manual_seed(7)) / (c_aux ** 0.5)
    target = (mix[None, :, None, None] * aux).sum(dim=1, keepdim=True) + 0.3 * field(1)
    target = (target - target.mean(dim=(2, 3), keepdim=True)) / target.std(dim=(2, 3), keepdim=True).clamp_min(1e-6)
    lr05 = F.avg_pool2d(target, 2)
    return lr05.to(device), target.to(device), aux.to(device)
