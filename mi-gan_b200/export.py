"""Training snapshot -> inference weights on the GPU (scripts/export_inference_model.py:17-85).

`copy_weights(source, dest, resolution)` has the reference function's signature: `source` is the training generator
(lib/model_zoo/migan.py: `encoder.b{res}` / `synthesis.b{res}` blocks of separable convolutions whose `Conv2d`s either hold
`weight` or, re-parameterised, `w0 .. w{k-1}`), `dest` an inference `Generator` (this package's or the reference's).  Every
filter goes through `b200_reparam_filter` (csrc/reparam.cu): sum of the re-parameterisation tensors / sqrt(k), then unit
L2 norm per output filter -- on the device, in one launch per convolution.  Biases and noise tensors are taken over as they are.
`export_state_dict(source, resolution)` returns the state_dict `scripts/export_inference_model.py` saves.

The source is only read through attributes (`reparametrize`, `num_reparam_tensors`, `w{i}` / `weight`, `bias`,
`noise_const`, `noise_strength`): any module tree of that shape works; unpickling a training `.pkl` needs the reference's own
classes on the path, like the reference script."""
from __future__ import annotations

import ctypes
import math
from collections import OrderedDict
from typing import List

import torch
import torch.nn as nn

from . import _abi


def reparam_filter(tensors: List[torch.Tensor]) -> torch.Tensor:
    """[w0 .. w{k-1}] (CUDA float32, identical shapes [cout, cin/groups, kh, kw]) -> the inference filter (same shape)."""
    k = len(tensors)
    if k < 1 or k > 16:
        raise RuntimeError("reparam_filter takes 1 .. 16 tensors, got %d" % k)
    ts = []
    for t in tensors:
        if not (isinstance(t, torch.Tensor) and t.is_cuda):
            raise RuntimeError("reparam_filter expects CUDA tensors: migan_b200.export has no CPU path")
        ts.append(t.detach().to(torch.float32).contiguous())
    if any(t.shape != ts[0].shape or t.device != ts[0].device for t in ts) or ts[0].dim() != 4:
        raise RuntimeError("reparam_filter expects %d tensors of one 4-D shape on one device" % k)
    cout, fan = ts[0].shape[0], ts[0][0].numel()
    out = torch.empty_like(ts[0])
    ptrs = (ctypes.c_void_p * k)(*[t.data_ptr() for t in ts])
    with torch.cuda.device(out.device):
        _abi.check(_abi.load().b200_reparam_filter(ctypes.cast(ptrs, ctypes.c_void_p), k, cout, fan, out.data_ptr(),
                                                   torch.cuda.current_stream(out.device).cuda_stream))
    return out


def _source_filter(conv, device) -> nn.Parameter:
    if getattr(conv, "reparametrize", False):
        ws = [getattr(conv, "w%d" % i) for i in range(conv.num_reparam_tensors)]
    else:
        ws = [conv.weight]
    return nn.Parameter(reparam_filter([w.to(device) for w in ws]))


def _take_conv(src_conv, dst_conv, device) -> None:
    """One `Conv2d` of the training graph -> the matching conv of the inference graph: filter and bias."""
    dst_conv.weight = _source_filter(src_conv, device)
    if getattr(dst_conv, "bias", None) is not None:
        dst_conv.bias = nn.Parameter(src_conv.bias.detach().to(device, torch.float32))


def copy_weights(source, dest, resolution: int = 256, device=None) -> None:
    """Same call as scripts/export_inference_model.py:17.  `device`: where the filters are computed and left (default: the
    first CUDA device the source lives on, else cuda:0)."""
    if device is None:
        p = next(iter(source.parameters()), None)
        device = p.device if (p is not None and p.is_cuda) else torch.device("cuda", 0)
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("migan_b200.export computes the filters on a CUDA device; got %s" % device)
    levels = [2 ** i for i in range(2, int(math.log2(resolution)) + 1)]
    for side, head in (("encoder", "fromrgb"), ("synthesis", "torgb")):
        for res in levels:
            s_blk, d_blk = getattr(getattr(source, side), "b%d" % res), getattr(getattr(dest, side), "b%d" % res)
            if getattr(d_blk, head, None) is not None:                      # :33-36, :60-63
                _take_conv(getattr(s_blk, head), getattr(d_blk, head), device)
            for name in ("conv1", "conv2"):                                 # :38-52, :65-82
                s_sep, d_sep = getattr(s_blk, name), getattr(d_blk, name)
                _take_conv(s_sep.conv1, d_sep.conv1, device)
                _take_conv(s_sep.conv2, d_sep.conv2, device)
                if side == "synthesis" and getattr(d_sep, "use_noise", False):
                    d_sep.noise_const = s_sep.conv2.noise_const             # a buffer of the training conv's 1x1 stage
                    d_sep.noise_strength = s_sep.conv2.noise_strength


def export_state_dict(source, resolution: int = 256, device=None) -> "OrderedDict[str, torch.Tensor]":
    """The state_dict of the inference generator for a training generator (what export_inference_model.py:156 saves), on the CPU."""
    from .generator import Generator
    dest = Generator(resolution)
    copy_weights(source, dest, resolution, device)
    return OrderedDict((k, v.detach().cpu()) for k, v in dest.state_dict().items())
