"""CPU oracle for the Co-Mod-GAN generator forward pass (BASELINE.json config 5).  TEST INFRASTRUCTURE ONLY.

Same rules as ``oracle/migan_oracle.py``: this file is the checker, never the product; only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import it, nothing under ``mi-gan_b200/`` does.

It restates, as plain functions over a ``state_dict``, the inference (eval-mode, fp32) algorithm of

* ``lib/model_zoo/comodgan.py``   Encoder :113-204, synthesis_block_first :207-258, synthesis_block :261-343,
                                    Synthesis :346-421, Generator :424-455
* ``lib/model_zoo/stylegan.py``   dense :62-99, modulated_conv2d :102-195, conv2d_layer :198-245,
                                    synthesis_layer :248-310, torgb_layer :313-344, Mapping :355-438
* ``torch_utils/ops/conv2d_resample.py``  _conv2d_wrapper :28-54, conv2d_resample :59-154
* ``lib/model_zoo/common/utils.py``       lrelu_agc :96-125 (the variant that takes a per-call ``gain``)

with the same torch CPU primitives the reference uses (``F.conv2d`` / ``F.conv_transpose2d`` / ``torch.addmm``).

Parity pin: ``tests/golden/make_golden_comodgan.py`` imports the real reference in the build container, loads the same
seeded ``state_dict`` into ``comodgan.Generator`` and asserts this oracle matches it (see the tolerance there) at
R=256 (the demo construction, ``scripts/demo.py:95-100``) and at small resolutions; the outputs are committed under
``tests/golden/``.  The reference ships no Co-Mod-GAN weights or golden tensors, so real-checkpoint parity is unpinned.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

from .migan_oracle import SQRT2, block_res, channels, check_resolution, encode_res, setup_filter, upfirdn2d_ref

Z_DIM = 512      # stylegan.py:357
W_DIM = 512      # stylegan.py:359, comodgan.py:348
W0_DIM = 1024    # comodgan.py:117 (encoder oc_n), :349
MAP_LAYERS = 8   # stylegan.py:361
MAP_LR = 0.01    # stylegan.py:365


def num_ws(resolution: int) -> int:
    """comodgan.py:371-374 gives 14 @256 and 16 @512 = one w per conv + the last torgb; 2*log2(R) - 2 in general."""
    return 2 * check_resolution(resolution) - 2


# --------------------------------------------------------------------------- #
# state_dict
# --------------------------------------------------------------------------- #
def state_dict_spec(resolution: int) -> "OrderedDict[str, Tuple[int, ...]]":
    """Keys and shapes in the reference's own ``state_dict()`` order (mapping, synthesis, encoder:
    stylegan.py:572-579 then comodgan.py:431-435)."""
    spec: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    spec["mapping.w_avg"] = (W_DIM,)
    for i in range(MAP_LAYERS):
        spec["mapping.fc%d.weight" % i] = (W_DIM, Z_DIM if i == 0 else W_DIM)
        spec["mapping.fc%d.bias" % i] = (W_DIM,)

    def synth_layer(p, cin, cout, res, k, *, filt, noise):
        spec[p + ".weight"] = (cout, cin, k, k)
        spec[p + ".bias"] = (cout,)
        if noise:
            spec[p + ".noise_strength"] = ()
        if filt:
            spec[p + ".resample_filter"] = (4, 4)
        if noise:
            spec[p + ".noise_const"] = (res, res)
        spec[p + ".affine.weight"] = (cin, W_DIM + W0_DIM)
        spec[p + ".affine.bias"] = (cin,)

    c4 = channels(4)
    spec["synthesis.b4.fc.weight"] = (c4 * 16, W0_DIM)
    spec["synthesis.b4.fc.bias"] = (c4 * 16,)
    # synthesis_block_first.conv is a synthesis_layer with the default resample_filter (comodgan.py:231)
    synth_layer("synthesis.b4.conv", c4, c4, 4, 3, filt=True, noise=True)
    synth_layer("synthesis.b4.torgb", c4, 3, 4, 1, filt=False, noise=False)
    for r in block_res(resolution)[1:]:
        ci, co = channels(r // 2), channels(r)
        spec["synthesis.b%d.resample_filter" % r] = (4, 4)
        synth_layer("synthesis.b%d.conv0" % r, ci, co, r, 3, filt=True, noise=True)
        synth_layer("synthesis.b%d.conv1" % r, co, co, r, 3, filt=False, noise=True)
        synth_layer("synthesis.b%d.torgb" % r, co, 3, r, 1, filt=False, noise=False)

    er = encode_res(resolution)
    for idx, r in enumerate(er[:-1]):
        ci, co = channels(r), channels(r // 2)
        p = "encoder.b%d" % r
        spec[p + ".resample_filter"] = (4, 4)
        if idx == 0:
            spec[p + ".fromrgb.weight"] = (ci, 4, 1, 1)
            spec[p + ".fromrgb.bias"] = (ci,)
        spec[p + ".conv0.weight"] = (ci, ci, 3, 3)
        spec[p + ".conv0.bias"] = (ci,)
        spec[p + ".conv1.weight"] = (co, ci, 3, 3)
        spec[p + ".conv1.bias"] = (co,)
        spec[p + ".conv1.resample_filter"] = (4, 4)
    spec["encoder.b4.conv.weight"] = (c4, c4, 3, 3)
    spec["encoder.b4.conv.bias"] = (c4,)
    spec["encoder.b4.fc.weight"] = (W0_DIM, c4 * 16)
    spec["encoder.b4.fc.bias"] = (W0_DIM,)
    return spec


def make_state_dict(resolution: int, seed: int = 1) -> "OrderedDict[str, torch.Tensor]":
    """Seeded weights.  Conv/dense weights are N(0,1) as in the constructors (the runtime ``weight_gain`` does the
    equalised-lr scaling, stylegan.py:77,217); unlike the constructors, biases, noise strengths and w_avg are non-zero
    so that every term of the forward is exercised."""
    g = torch.Generator().manual_seed(seed)
    fir = setup_filter([1, 3, 3, 1])
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for k, shape in state_dict_spec(resolution).items():
        if k.endswith("resample_filter"):
            v = fir.clone()
        elif k.endswith("noise_strength"):
            v = 0.1 * torch.randn((), generator=g)
        elif k.endswith("affine.bias"):
            v = 1.0 + 0.1 * torch.randn(shape, generator=g)           # bias_init=1 (stylegan.py:270)
        elif k.endswith("bias") or k == "mapping.w_avg":
            v = 0.1 * torch.randn(shape, generator=g)
        elif k.startswith("mapping.fc") and k.endswith("weight"):
            v = torch.randn(shape, generator=g) / MAP_LR              # stylegan.py:75
        else:
            v = torch.randn(shape, generator=g)
        sd[k] = v.to(torch.float32)
    return sd


def make_latent(n: int, seed: int = 4321) -> torch.Tensor:
    return torch.randn(n, Z_DIM, generator=torch.Generator().manual_seed(seed))


# --------------------------------------------------------------------------- #
# Ops
# --------------------------------------------------------------------------- #
def lrelu_agc(x, gain=1.0, alpha=0.2, act_gain=SQRT2, clamp=256.0):
    """common/utils.py:114-122: leaky_relu, * (act_gain*gain), clamp(+-clamp*gain)."""
    x = F.leaky_relu(x, negative_slope=alpha)
    g = act_gain * gain
    if g != 1:
        x = x * g
    if clamp is not None:
        x = x.clamp(-clamp * gain, clamp * gain)
    return x


def dense(sd, p, x, *, act: bool, lr_multi: float = 1.0):
    """stylegan.py:84-96."""
    w = sd[p + ".weight"]
    w = w * (lr_multi / np.sqrt(w.shape[1]))
    b = sd[p + ".bias"]
    if lr_multi != 1:
        b = b * lr_multi
    x = torch.addmm(b.unsqueeze(0), x, w.t())
    return lrelu_agc(x) if act else x


def _conv2d_wrapper(x, w, stride=1, padding=0, groups=1, transpose=False, flip_weight=True):
    """conv2d_resample.py:28-54 (without the cuDNN channels_last workaround, which does not change values)."""
    if not flip_weight:
        w = w.flip([2, 3])
    op = F.conv_transpose2d if transpose else F.conv2d
    return op(x, w, stride=stride, padding=padding, groups=groups)


def conv2d_resample_ref(x, w, f=None, up=1, down=1, padding=0, groups=1, flip_weight=True, flip_filter=False):
    """conv2d_resample.py:59-154, all six branches."""
    out_channels, in_channels_per_group, kh, kw = w.shape
    fw = 1 if f is None else int(f.shape[-1])          # _get_filter_size (conv2d_resample.py:... upfirdn2d.py:47-57)
    fh = 1 if f is None else int(f.shape[0])
    if isinstance(padding, int):
        padding = [padding] * 4
    elif len(padding) == 2:
        padding = [padding[0], padding[0], padding[1], padding[1]]
    px0, px1, py0, py1 = padding
    if up > 1:                                  # :94-98
        px0 += (fw + up - 1) // 2
        px1 += (fw - up) // 2
        py0 += (fh + up - 1) // 2
        py1 += (fh - up) // 2
    if down > 1:                                # :99-103
        px0 += (fw - down + 1) // 2
        px1 += (fw - down) // 2
        py0 += (fh - down + 1) // 2
        py1 += (fh - down) // 2
    if kw == 1 and kh == 1 and (down > 1 and up == 1):       # :106-109
        x = upfirdn2d_ref(x, f, down=down, padding=[px0, px1, py0, py1], flip_filter=flip_filter)
        return _conv2d_wrapper(x, w, groups=groups, flip_weight=flip_weight)
    if kw == 1 and kh == 1 and (up > 1 and down == 1):       # :112-115
        x = _conv2d_wrapper(x, w, groups=groups, flip_weight=flip_weight)
        return upfirdn2d_ref(x, f, up=up, padding=[px0, px1, py0, py1], gain=up ** 2, flip_filter=flip_filter)
    if down > 1 and up == 1:                                 # :118-121
        x = upfirdn2d_ref(x, f, padding=[px0, px1, py0, py1], flip_filter=flip_filter)
        return _conv2d_wrapper(x, w, stride=down, groups=groups, flip_weight=flip_weight)
    if up > 1:                                               # :124-142
        if groups == 1:
            w = w.transpose(0, 1)
        else:
            w = w.reshape(groups, out_channels // groups, in_channels_per_group, kh, kw)
            w = w.transpose(1, 2)
            w = w.reshape(groups * in_channels_per_group, out_channels // groups, kh, kw)
        px0 -= kw - 1
        px1 -= kw - up
        py0 -= kh - 1
        py1 -= kh - up
        pxt = max(min(-px0, -px1), 0)
        pyt = max(min(-py0, -py1), 0)
        x = _conv2d_wrapper(x, w, stride=up, padding=[pyt, pxt], groups=groups, transpose=True,
                            flip_weight=(not flip_weight))
        x = upfirdn2d_ref(x, f, padding=[px0 + pxt, px1 + pxt, py0 + pyt, py1 + pyt], gain=up ** 2,
                          flip_filter=flip_filter)
        if down > 1:
            x = upfirdn2d_ref(x, f, down=down, flip_filter=flip_filter)
        return x
    if up == 1 and down == 1:                                # :145-147
        if px0 == px1 and py0 == py1 and px0 >= 0 and py0 >= 0:
            return _conv2d_wrapper(x, w, padding=[py0, px0], groups=groups, flip_weight=flip_weight)
    x = upfirdn2d_ref(x, (f if up > 1 else None), up=up, padding=[px0, px1, py0, py1], gain=up ** 2,
                      flip_filter=flip_filter)               # :150-154
    x = _conv2d_wrapper(x, w, groups=groups, flip_weight=flip_weight)
    if down > 1:
        x = upfirdn2d_ref(x, f, down=down, flip_filter=flip_filter)
    return x


def conv2d_layer(sd, p, x, *, k: int, up=1, down=1, act=True, gain=1.0, taps=None):
    """stylegan.py:230-242."""
    w = sd[p + ".weight"]
    w = w * (1.0 / np.sqrt(w.shape[1] * k * k))
    f = sd.get(p + ".resample_filter")
    x = conv2d_resample_ref(x, w, f=f, up=up, down=down, padding=k // 2, flip_weight=(up == 1))
    b = sd.get(p + ".bias")
    if b is not None:
        x = x + b.view(1, -1, 1, 1)
    x = lrelu_agc(x, gain=gain) if act else x * gain
    if taps is not None:
        taps[p + ".out"] = x
    return x


def modulated_conv2d(x, weight, styles, noise=None, up=1, padding=0, resample_filter=None, demodulate=True,
                     flip_weight=True):
    """stylegan.py:128-195, fp32 eval path (fused_modconv=True: per-sample weights, grouped convolution)."""
    n = x.shape[0]
    cout, cin, kh, kw = weight.shape
    if demodulate:                                           # :144-146 ("Type StyleGan3")
        weight = weight * weight.square().mean([1, 2, 3], keepdim=True).rsqrt()
        styles = styles * styles.square().mean().rsqrt()     # mean over the WHOLE [N, I] tensor
    w = weight.unsqueeze(0) * styles.reshape(n, 1, -1, 1, 1)  # :149-150
    if demodulate:
        dcoefs = (w.square().sum(dim=[2, 3, 4]) + 1e-8).rsqrt()   # :154
        w = w * dcoefs.reshape(n, -1, 1, 1, 1)                    # :168
    x = x.reshape(1, -1, *x.shape[2:])                            # :188-193
    w = w.reshape(-1, cin, kh, kw)
    x = conv2d_resample_ref(x, w, f=resample_filter, up=up, padding=padding, groups=n, flip_weight=flip_weight)
    x = x.reshape(n, -1, *x.shape[2:])
    if noise is not None:
        x = x + noise
    return x


def _noise(sd, p, noise_mode: str, noise: Optional[Dict[str, torch.Tensor]]):
    """stylegan.py:284-289.  'random' needs caller-provided N(0,1) planes ``noise[p]`` of shape [N,1,r,r]."""
    if p + ".noise_strength" not in sd or noise_mode == "none":
        return None
    if noise_mode == "const":
        return sd[p + ".noise_const"] * sd[p + ".noise_strength"]
    return noise[p] * sd[p + ".noise_strength"]


def synthesis_layer(sd, p, x, w_long, *, up=1, gain=1.0, noise_mode="const", noise=None, taps=None):
    """stylegan.py:280-310."""
    styles = dense(sd, p + ".affine", w_long, act=False)
    x = modulated_conv2d(x, sd[p + ".weight"], styles, noise=_noise(sd, p, noise_mode, noise), up=up, padding=1,
                         resample_filter=sd.get(p + ".resample_filter"), flip_weight=(up == 1))
    x = x + sd[p + ".bias"].view(1, -1, 1, 1)
    x = lrelu_agc(x, gain=gain)
    if taps is not None:
        taps[p + ".out"] = x
    return x


def torgb_layer(sd, p, x, w_long, taps=None):
    """stylegan.py:330-344: styles scaled by weight_gain, no demodulation, bias, no activation."""
    cin = sd[p + ".weight"].shape[1]
    styles = dense(sd, p + ".affine", w_long, act=False) * (1.0 / np.sqrt(cin))
    x = modulated_conv2d(x, sd[p + ".weight"], styles, demodulate=False)
    x = x + sd[p + ".bias"].view(1, -1, 1, 1)
    if taps is not None:
        taps[p + ".out"] = x
    return x


# --------------------------------------------------------------------------- #
# Networks
# --------------------------------------------------------------------------- #
def mapping_forward(sd, z, num_ws_: int, truncation_psi=1.0, truncation_cutoff=None):
    """stylegan.py:401-438 (c_dim = 0, eval mode)."""
    x = z
    x = x * (x.square().mean(dim=1, keepdim=True) + 1e-8).rsqrt()
    for i in range(MAP_LAYERS):
        x = dense(sd, "mapping.fc%d" % i, x, act=True, lr_multi=MAP_LR)
    x = x.unsqueeze(1).repeat([1, num_ws_, 1])
    if truncation_psi != 1:
        w_avg = sd["mapping.w_avg"]
        if truncation_cutoff is None:
            x = w_avg.lerp(x, truncation_psi)
        else:
            x[:, :truncation_cutoff] = w_avg.lerp(x[:, :truncation_cutoff], truncation_psi)
    return x


def encoder_forward(sd, x_in, resolution: int, taps=None):
    """comodgan.py:187-204; encoder_block :33-60 (reslink=False), encoder_epilogue :93-110 (no mbstd, eval dropout)."""
    er = encode_res(resolution)
    feats = {}
    x = None
    for idx, r in enumerate(er[:-1]):
        p = "encoder.b%d" % r
        if idx == 0:
            x = conv2d_layer(sd, p + ".fromrgb", x_in, k=1, taps=taps)
        feat = conv2d_layer(sd, p + ".conv0", x, k=3, taps=taps)
        x = conv2d_layer(sd, p + ".conv1", feat, k=3, down=2, taps=taps)
        feats[r] = feat
    feat = conv2d_layer(sd, "encoder.b4.conv", x, k=3, taps=taps)
    x = dense(sd, "encoder.b4.fc", feat.flatten(1), act=True)
    feats[4] = feat
    if taps is not None:
        taps["encoder.b4.fc.out"] = x
    return x, feats


def synthesis_forward(sd, x_global, feats, ws, resolution: int, noise_mode="const", noise=None, taps=None):
    """comodgan.py:398-421, synthesis_block_first :235-258, synthesis_block :301-343 (res_link=False)."""
    br = block_res(resolution)
    w0 = x_global
    c4 = channels(4)
    wi = 0
    x = dense(sd, "synthesis.b4.fc", w0, act=True).view(-1, c4, 4, 4) + feats[4]
    x = synthesis_layer(sd, "synthesis.b4.conv", x, torch.cat([ws[:, wi], w0], 1), noise_mode=noise_mode,
                        noise=noise, taps=taps)
    img = torgb_layer(sd, "synthesis.b4.torgb", x, torch.cat([ws[:, wi + 1], w0], 1), taps=taps)
    wi += 1                                       # block.num_conv (the torgb w is shared with the next block's conv0)
    for r in br[1:]:
        p = "synthesis.b%d" % r
        x = synthesis_layer(sd, p + ".conv0", x, torch.cat([ws[:, wi], w0], 1), up=2, noise_mode=noise_mode,
                            noise=noise, taps=taps)
        x = x + feats[r]
        x = synthesis_layer(sd, p + ".conv1", x, torch.cat([ws[:, wi + 1], w0], 1), noise_mode=noise_mode,
                            noise=noise, taps=taps)
        img = upfirdn2d_ref(img, sd[p + ".resample_filter"], up=2, padding=[2, 1, 2, 1], gain=4)   # upsample2d
        img = img + torgb_layer(sd, p + ".torgb", x, torch.cat([ws[:, wi + 2], w0], 1), taps=taps)
        if taps is not None:
            taps[p + ".img"] = img
        wi += 2
    return img


def generator_forward(sd, x, z, resolution: int, *, truncation_psi=1.0, truncation_cutoff=None, noise_mode="const",
                      noise=None, taps=None, dtype=torch.float32):
    """comodgan.py:438-455."""
    if dtype != torch.float32:
        sd = {k: v.to(dtype) for k, v in sd.items()}
        x, z = x.to(dtype), z.to(dtype)
    with torch.no_grad():
        ws = mapping_forward(sd, z, num_ws(resolution), truncation_psi, truncation_cutoff)
        if taps is not None:
            taps["mapping.ws"] = ws
        xg, feats = encoder_forward(sd, x, resolution, taps)
        return synthesis_forward(sd, xg, feats, ws, resolution, noise_mode, noise, taps)


def noise_keys(resolution: int) -> List[Tuple[str, int]]:
    """(layer prefix, r) of every noise input, in forward order."""
    out = [("synthesis.b4.conv", 4)]
    for r in block_res(resolution)[1:]:
        out += [("synthesis.b%d.conv0" % r, r), ("synthesis.b%d.conv1" % r, r)]
    return out
