"""Synthetic weights and inputs for benchmarks and smoke runs (no network: there are no
released checkpoints in the image).

`export_style_state_dict` draws weights with the statistics of a checkpoint written by the
reference's scripts/export_inference_model.py:26 (every conv filter L2-normalised per output
channel, fp32), see SURVEY.md section 8c.  `synthetic_input` builds the generator input the way
the callers do (scripts/demo.py:56-66): x = cat([mask - 0.5, img * mask], 1).
(tests/ check these reproduce the oracle's own seeded generators bit for bit.)
"""
from __future__ import annotations

from collections import OrderedDict

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


def synthetic_input(resolution: int, n: int, seed: int = 1234, hole: float = 0.4) -> torch.Tensor:
    """img ~ U[-1,1]; mask = blocky random holes (8x8 cells, 1 = known); x = [mask-0.5, img*mask]."""
    g = torch.Generator().manual_seed(seed)
    img = torch.rand(n, 3, resolution, resolution, generator=g) * 2 - 1
    cells = max(resolution // 8, 1)
    coarse = (torch.rand(n, 1, cells, cells, generator=g) > hole).float()
    mask = F.interpolate(coarse, size=(resolution, resolution), mode="nearest")
    return torch.cat([mask - 0.5, img * mask], dim=1).contiguous()
