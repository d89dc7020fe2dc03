"""CPU oracle for the arbitrary-resolution crop pipeline of the deployed (ONNX) form, SURVEY.md 8(f) row f4.
TEST INFRASTRUCTURE ONLY (same rules as oracle/migan_oracle.py: imported by tests/, bench.py's cpu legs and smoke()).

Restates ``MIGAN_Pipeline`` of scripts/create_onnx_pipeline.py:121-264 as plain functions:

* ``masked_bbox``   :133-227  bounding box of the hole (any mask value < 255), padded, squared and clipped -> crop window
* ``resize_nearest`` / ``resize_bilinear_aa``   what ``torchvision.transforms.functional.resize`` does to a *tensor*
  (torchvision 0.26 ``_functional_tensor.resize``: cast to float32, ``F.interpolate(..., antialias=True)`` for bilinear /
  ``mode='nearest'`` for nearest, round, cast back; identity when the size does not change)
* ``preprocess``    :229-236  crop -> model input x[1,4,res,res]
* ``postprocess``   :238-248  generator output -> resized to the crop, feathered composite (oracle/prepost_oracle.py)
* ``forward``       :250-264  the whole request; the image is updated in place like the reference

``aa_weights`` additionally restates ATen's anti-aliased bilinear weight computation in numpy float32 / float64 exactly
as the C++ evaluates it (weights in float, the filter argument through double, ``w / total``); the CUDA kernels build the
same table.  It was checked here against the weights torch applies (impulse responses), see tests/test_pipeline_emul.py.

Parity pin: tests/golden/make_golden_pipeline.py runs the reference's own ``MIGAN_Pipeline`` on seeded images / masks and
asserts bit-equality of every stage with the functions below; the vectors are committed as tests/golden/pipeline_*.npz.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import prepost_oracle as P


def masked_bbox(mask_u8: np.ndarray, res: int, padding: int = 128):
    """mask_u8 [H,W] (255 = known).  Returns (x_min, x_max, y_min, y_max) of the crop window, create_onnx_pipeline.py:133-227."""
    h, w = mask_u8.shape                                                     # :137-140 (the padded-sum trick just reads H, W)
    m = torch.from_numpy(np.ascontiguousarray(mask_u8)).to(torch.float32)
    cols = torch.nonzero(m.mean(dim=0) < 255.0).flatten().tolist()           # :146,148
    rows = torch.nonzero(m.mean(dim=1) < 255.0).flatten().tolist()           # :147,149
    x_min = min(list(cols) + [w]); x_max = max(list(cols) + [0])             # :151-152
    y_min = min(list(rows) + [h]); y_max = max(list(rows) + [0])             # :153-154
    x_min = min(x_min, x_max); x_max = max(x_min, x_max)                     # :156-164
    y_min = min(y_min, y_max); y_max = max(y_min, y_max)                     # :166-174
    cnt_x = (x_min + x_max) // 2; cnt_y = (y_min + y_max) // 2               # :176-177
    crop = max(x_max - x_min, y_max - y_min) + padding * 2                   # :179-182
    crop = max(crop, res)                                                    # :183-186
    off = crop // 2                                                          # :188
    x_min = max(cnt_x - off, 0); x_max = min(cnt_x + off, w)                 # :189-196
    y_min = max(cnt_y - off, 0); y_max = min(cnt_y + off, h)                 # :197-204
    x_ex = max(crop - (x_max - x_min), 0); y_ex = max(crop - (y_max - y_min), 0)   # :206-213
    x_min = max(x_min - x_ex, 0); x_max = min(x_max + x_ex, w)               # :215-222
    y_min = max(y_min - y_ex, 0); y_max = min(y_max + y_ex, h)               # :224-231
    return int(x_min), int(x_max), int(y_min), int(y_max)


def resize_nearest(t_u8: torch.Tensor, oh: int, ow: int) -> torch.Tensor:
    """tvF.resize(t, (oh, ow), interpolation=NEAREST) on a uint8 tensor [1,C,H,W]."""
    if (t_u8.shape[2], t_u8.shape[3]) == (oh, ow):
        return t_u8
    return torch.round(F.interpolate(t_u8.to(torch.float32), size=(oh, ow), mode="nearest")).to(torch.uint8)


def resize_bilinear_aa(t: torch.Tensor, oh: int, ow: int) -> torch.Tensor:
    """tvF.resize(t, (oh, ow), interpolation=BILINEAR) on a tensor [1,C,H,W]: anti-aliased; uint8 goes through float32 and round()."""
    if (t.shape[2], t.shape[3]) == (oh, ow):
        return t
    r = F.interpolate(t.to(torch.float32), size=(oh, ow), mode="bilinear", align_corners=False, antialias=True)
    return torch.round(r).to(torch.uint8) if t.dtype == torch.uint8 else r


def preprocess(image_u8: torch.Tensor, mask_u8: torch.Tensor, res: int) -> torch.Tensor:
    """:229-236.  image_u8 [1,3,h,w], mask_u8 [1,1,h,w] (the crop) -> x [1,4,res,res]."""
    image = resize_bilinear_aa(image_u8, res, res)
    mask = resize_nearest(mask_u8, res, res)
    image = image.to(torch.float32) * 2 / 255 - 1
    mask = mask.to(torch.float32) / 255
    return torch.cat([mask - 0.5, image * mask], dim=1)


def postprocess(image_u8: torch.Tensor, mask_u8: torch.Tensor, model_output: torch.Tensor) -> torch.Tensor:
    """:238-248.  The generator output is mapped to [0, 255], resized to the crop and blended with the feathered mask."""
    out = ((model_output * 0.5 + 0.5) * 255).clamp(0, 255)
    out = resize_bilinear_aa(out, image_u8.size(2), image_u8.size(3))
    image = image_u8.to(torch.float32)
    mask = mask_u8.to(torch.float32)
    mask = F.max_pool2d(mask, 3, stride=1, padding=1)
    mask = F.pad(mask, (2, 2, 2, 2), mode="reflect")
    mask = F.conv2d(mask, P.gaussian_kernel_5x5().view(1, 1, 5, 5), padding="valid")
    mask = mask / torch.tensor(255)
    composed = image * mask + out * (1 - mask)
    return composed.clamp(0, 255).to(torch.uint8)


def forward(generator, image_u8: torch.Tensor, mask_u8: torch.Tensor, res: int, padding: int = 128, taps: dict = None) -> torch.Tensor:
    """:250-264.  generator: x[1,4,res,res] -> y[1,3,res,res].  image_u8 [1,3,H,W] is modified in place and returned."""
    mask = resize_nearest(mask_u8, image_u8.size(2), image_u8.size(3))
    x0, x1, y0, y1 = masked_bbox(mask[0, 0].numpy(), res, padding)
    ci, cm = image_u8[:, :, y0:y1, x0:x1], mask[:, :, y0:y1, x0:x1]
    x = preprocess(ci, cm, res)
    y = generator(x)
    post = postprocess(ci, cm, y)
    if taps is not None:
        taps.update(box=(x0, x1, y0, y1), x=x, y=y, post=post)
    image_u8[:, :, y0:y1, x0:x1] = post
    return image_u8


# ---- the weight table of ATen's anti-aliased bilinear filter, evaluated like the C++ (UpSampleKernel.cpp, HelperInterpBase) ----
def aa_weights(in_size: int, out_size: int):
    """Returns (xmin[out], xsize[out], w[out][max_interp]) in numpy: float scale / support / centre, the filter argument
    through double ((j + xmin - center + 0.5) * invscale), weights normalised by division."""
    f32, f64 = np.float32, np.float64
    scale = f32(f32(in_size) / f32(out_size))
    if scale >= 1.0:
        support, invscale = f32(f64(1.0) * f64(scale)), f32(f64(1.0) / f64(scale))
    else:
        support, invscale = f32(1.0), f32(1.0)
    max_interp = int(math.ceil(support)) * 2 + 1
    xmin = np.zeros(out_size, np.int32); xsize = np.zeros(out_size, np.int32)
    w = np.zeros((out_size, max_interp), np.float32)
    for i in range(out_size):
        center = f32(f64(scale) * (i + 0.5))
        lo = max(int(f64(f32(center - support)) + 0.5), 0)
        n = min(int(f64(f32(center + support)) + 0.5), in_size) - lo
        n = min(max(n, 0), max_interp)
        total = f32(0)
        for j in range(n):
            a = abs(f32((f64(f32(f32(j + lo) - center)) + 0.5) * f64(invscale)))
            w[i, j] = f32(f64(1.0) - f64(a)) if a < 1.0 else f32(0)
            total = f32(total + w[i, j])
        if total != 0:
            for j in range(n):
                w[i, j] = f32(w[i, j] / total)
        xmin[i], xsize[i] = lo, n
    return xmin, xsize, w
