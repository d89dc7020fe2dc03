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
