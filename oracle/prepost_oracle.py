"""CPU oracle for the callers' pre/post-processing around the generator.  TEST INFRASTRUCTURE ONLY
(same rules as oracle/migan_oracle.py).

Restates, for images already at the model resolution (PIL's resize to the same size is a copy, cv2.resize to the same
size likewise), scripts/demo.py:

* preprocess  :56-66   uint8 RGB image [H,W,3] + uint8 mask [H,W] (255 = known) -> x = cat([mask-0.5, img*mask]) [1,4,H,W]
* postprocess :135-142 result = (y*0.5+0.5).clamp(0,1)*255 -> uint8 (truncation) -> HWC; composed = img*mask + result*(1-mask)

Parity pin: tests/golden/make_golden_prepost.py runs the reference's own ``preprocess`` function on seeded PIL images
and asserts bit-equality with ``preprocess`` below; the post-processing lines of the reference are inline in ``main()``
and cannot be imported, so ``postprocess`` is a line-by-line restatement with the same torch / numpy calls.
"""
import numpy as np
import torch


def preprocess(img_u8: np.ndarray, mask_u8: np.ndarray) -> torch.Tensor:
    """img_u8 [N,H,W,3], mask_u8 [N,H,W] -> x [N,4,H,W] float32 (demo.py:59-66 per image)."""
    mask = torch.Tensor(mask_u8[..., np.newaxis] // 255).float()           # :60,62
    img = torch.Tensor(img_u8).float() * 2 / 255 - 1                       # :61
    img = img.permute(0, 3, 1, 2)                                          # :63
    mask = mask.permute(0, 3, 1, 2)                                        # :64
    return torch.cat([mask - 0.5, img * mask], dim=1)                      # :65


def postprocess(y: torch.Tensor, img_u8: np.ndarray, mask_u8: np.ndarray) -> np.ndarray:
    """y [N,3,H,W] float32 -> composed uint8 [N,H,W,3] (demo.py:135-141 per image)."""
    res = (y * 0.5 + 0.5).clamp(0, 1) * 255                                # :135
    res = res.to(torch.uint8).permute(0, 2, 3, 1).numpy()                  # :136
    m = mask_u8[..., np.newaxis] // 255                                    # :139
    return img_u8 * m + res * (1 - m)                                      # :140


def gaussian_kernel_5x5() -> torch.Tensor:
    """The 5x5 smoothing kernel of the deployed pipeline (scripts/create_onnx_pipeline.py:66-88, GaussianSmoothing with
    kernel_size=5, sigma=1.0 as constructed at :127-128).  Note the exponent is -((x - mean) / (2 sigma))^2, as written there."""
    import math
    size, std = 5, 1.0
    grids = torch.meshgrid([torch.arange(size, dtype=torch.float32) for _ in range(2)])
    kernel = 1
    mean = (size - 1) / 2
    for g in grids:
        kernel = kernel * (1 / (std * math.sqrt(2 * math.pi)) * torch.exp(-((g - mean) / (2 * std)) ** 2))
    return kernel / torch.sum(kernel)


def feather_composite(image_u8: torch.Tensor, mask_u8: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """The blend of the deployed (ONNX) pipeline for a crop already at the model resolution
    (scripts/create_onnx_pipeline.py:233-245; the bilinear resize to the crop size at :235 is the identity then):
    image_u8 [N,3,H,W] uint8, mask_u8 [N,1,H,W] uint8 (255 = known), y [N,3,H,W] generator output -> composed uint8 [N,3,H,W].
    The mask is dilated (3x3 max-pool), smoothed with the 5x5 kernel above on a reflect-padded border and used as the
    per-pixel blend weight."""
    import torch.nn.functional as F
    out = ((y * 0.5 + 0.5) * 255).clamp(0, 255)                                        # :234
    image = image_u8.to(torch.float32)                                                 # :236
    mask = mask_u8.to(torch.float32)                                                   # :237
    mask = F.max_pool2d(mask, 3, stride=1, padding=1)                                  # :238
    mask = F.pad(mask, (2, 2, 2, 2), mode="reflect")                                   # :117 (GaussianSmoothing.forward)
    mask = F.conv2d(mask, gaussian_kernel_5x5().view(1, 1, 5, 5), padding="valid")     # :118
    mask = mask / torch.tensor(255)                                                    # :240
    composed = image * mask + out * (1 - mask)                                         # :241
    return composed.clamp(0, 255).to(torch.uint8)                                      # :242
