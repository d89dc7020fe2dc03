"""Synthetic weights and inputs for benchmarks and smoke runs (no network: there are no
released checkpoints in the image).

`export_style_state_dict` draws weights with the statistics of a checkpoint written by the
reference's scripts/export_inference_model.py:26 (every conv filter L2-normalised per output
channel, fp32), see SURVEY.md section 8c.  `synthetic_input` builds the generator input the way
the callers do (scripts/demo.py:56-66): x = cat([mask - 0.5, img * mask], 1).
(tests/ check these reproduce the oracle's own seeded generators bit for bit.)
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

from . import arch


def _fir(gain: float) -> torch.Tensor:
    f = torch.tensor(arch.FIR_PROTOTYPE, dtype=torch.float32)
    f = torch.outer(f, f)
    f = f / f.sum()
    return f * gain


def export_style_state_dict(resolution: int, seed: int = 1) -> "OrderedDict[str, torch.Tensor]":
    g = torch.Generator().manual_seed(seed)
    sd = OrderedDict()
    for key, shape, _ in arch.state_entries(resolution):
        if key.endswith("filter.weight"):
            t = _fir(1.0 if "downsample" in key else 4.0).repeat(shape[0], 1, 1, 1).clone()
        elif key.endswith("filter_const"):
            t = torch.tensor([[1.0, 0.0], [0.0, 0.0]]).repeat(1, 1, shape[2] // 2, shape[3] // 2).clone()
        elif key.endswith("noise_const"):
            t = torch.randn(shape, generator=g)
        elif key.endswith("noise_strength") or key.endswith(".bias"):
            t = 0.1 * torch.randn(shape, generator=g)
        else:
            t = torch.randn(shape, generator=g)
            t = t * t.flatten(1).square().sum(1).add(1e-8).rsqrt().view(-1, 1, 1, 1)
        sd[key] = t.contiguous()
    return sd


def _stroke(mask, rng, s: int) -> None:
    """One free-form brush stroke: a random polyline of thick segments with round joints, rasterised with numpy
    (distance-to-segment test inside each segment's bounding box)."""
    n_vertex = int(rng.randint(4, 18))
    width = float(rng.uniform(12, 48)) * s / 512.0 + 1.0
    mean_angle, spread = 2 * math.pi / 5, 2 * math.pi / 15
    a_min, a_max = mean_angle - rng.uniform(0, spread), mean_angle + rng.uniform(0, spread)
    radius = math.hypot(s, s) / 8
    px, py = float(rng.randint(0, s)), float(rng.randint(0, s))
    half = width / 2
    for i in range(n_vertex):
        ang = rng.uniform(a_min, a_max)
        if i % 2 == 0:
            ang = 2 * math.pi - ang
        r = float(np.clip(rng.normal(radius, radius / 2), 0, 2 * radius))
        qx, qy = float(np.clip(px + r * math.cos(ang), 0, s)), float(np.clip(py + r * math.sin(ang), 0, s))
        x0, x1 = int(max(min(px, qx) - half, 0)), int(min(max(px, qx) + half + 1, s))
        y0, y1 = int(max(min(py, qy) - half, 0)), int(min(max(py, qy) + half + 1, s))
        if x1 > x0 and y1 > y0:
            yy, xx = np.mgrid[y0:y1, x0:x1].astype(np.float32)
            dx, dy = qx - px, qy - py
            t = np.clip(((xx - px) * dx + (yy - py) * dy) / max(dx * dx + dy * dy, 1e-6), 0.0, 1.0)
            d2 = (xx - (px + t * dx)) ** 2 + (yy - (py + t * dy)) ** 2
            mask[y0:y1, x0:x1][d2 <= half * half] = 0
        px, py = qx, qy


def free_form_mask(s: int, rng, hole_range=(0.0, 1.0)) -> "np.ndarray":
    """A free-form inpainting mask in the style of the evaluation protocol BASELINE.json names for configs[1]
    (scripts/evaluate_fid_lpips.py:44-121 draws random rectangles plus random brush strokes and keeps masks whose hole
    ratio lies in `hole_range`): 1 = known pixel, 0 = hole.  This is an independent numpy implementation of that
    recipe (no PIL), seeded through `rng` (np.random.RandomState); it does not reproduce the reference's random stream."""
    coef = min(hole_range[0] + hole_range[1], 1.0)
    while True:
        mask = np.ones((s, s), np.uint8)
        for max_tries, max_size in ((int(10 * coef), s // 2), (int(5 * coef), s)):
            for _ in range(int(rng.randint(max(max_tries, 1)))):
                w, h = int(rng.randint(max_size)), int(rng.randint(max_size))
                x = int(rng.randint(-(w // 2), s - w + w // 2))
                y = int(rng.randint(-(h // 2), s - h + h // 2))
                mask[max(y, 0):min(y + h, s), max(x, 0):min(x + w, s)] = 0
        for _ in range(int(rng.randint(max(int(20 * coef), 1)))):
            _stroke(mask, rng, s)
        if rng.random_sample() > 0.5:
            mask = mask[::-1]
        if rng.random_sample() > 0.5:
            mask = mask[:, ::-1]
        ratio = 1.0 - float(mask.mean())
        if hole_range[0] < ratio < hole_range[1]:
            return np.ascontiguousarray(mask)


def synthetic_input(resolution: int, n: int, seed: int = 1234, hole: float = 0.4, masks: str = "blocks") -> torch.Tensor:
    """img ~ U[-1,1]; x = [mask-0.5, img*mask] (scripts/demo.py:56-66).  masks = "blocks": blocky random holes (8x8 cells,
    1 = known); masks = "free_form": free-form masks (rectangles + brush strokes, hole ratio in (0, 1)), seed 0-style
    stream from np.random.RandomState(seed) -- the protocol BASELINE.json configs[1] names."""
    g = torch.Generator().manual_seed(seed)
    img = torch.rand(n, 3, resolution, resolution, generator=g) * 2 - 1
    if masks == "free_form":
        rng = np.random.RandomState(seed)
        mask = torch.from_numpy(np.stack([free_form_mask(resolution, rng) for _ in range(n)])).float()[:, None]
        return torch.cat([mask - 0.5, img * mask], dim=1).contiguous()
    cells = max(resolution // 8, 1)
    coarse = (torch.rand(n, 1, cells, cells, generator=g) > hole).float()
    mask = F.interpolate(coarse, size=(resolution, resolution), mode="nearest")
    return torch.cat([mask - 0.5, img * mask], dim=1).contiguous()
