"""Shape bookkeeping of the MI-GAN inference generator (host-side mirror of the reference's
constructor arithmetic, lib/model_zoo/migan_inference.py:214-233 and :329-345).

`state_entries(resolution)` lists every state_dict entry (key, shape, kind) in the reference's
own order; the Generator module registers its parameters from it and the C library builds the
same list independently (tests compare the two with the reference's real state_dict fixture).
"""
from __future__ import annotations

import math
from typing import List, Tuple

CH_BASE = 32768
CH_MAX = 512
FIR_PROTOTYPE = (1.0, 3.0, 3.0, 1.0)  # setup_filter([1, 3, 3, 1]) at migan_inference.py:71, :95

PARAM, BUFFER = "param", "buffer"
Entry = Tuple[str, Tuple[int, ...], str]


def log2_resolution(resolution: int) -> int:
    """Raises ValueError for non powers of two, like the reference (:214-216, :329-331)."""
    if not isinstance(resolution, int) or resolution < 8:
        raise ValueError("resolution must be a power of two >= 8, got %r" % (resolution,))
    log2res = int(math.log2(resolution))
    if 2 ** log2res != resolution:
        raise ValueError("resolution must be a power of two, got %d" % resolution)
    return log2res


def nf(res: int) -> int:
    return min(CH_BASE // res, CH_MAX)


def encoder_resolutions(resolution: int) -> List[int]:
    return [2 ** i for i in range(log2_resolution(resolution), 1, -1)]


def synthesis_resolutions(resolution: int) -> List[int]:
    return [2 ** i for i in range(2, log2_resolution(resolution) + 1)]


def _sepconv(prefix: str, cin: int, cout: int, res_out=None, *, noise=False, down=False, up=False) -> List[Entry]:
    e: List[Entry] = []
    if noise:
        e.append((prefix + "noise_strength", (), PARAM))
        e.append((prefix + "noise_const", (res_out, res_out), BUFFER))
    e.append((prefix + "conv1.weight", (cin, 1, 3, 3), PARAM))
    e.append((prefix + "conv1.bias", (cin,), PARAM))
    e.append((prefix + "conv2.weight", (cout, cin, 1, 1), PARAM))
    if down:
        e.append((prefix + "downsample.filter.weight", (cin, 1, 4, 4), PARAM))
    if up:
        e.append((prefix + "upsample.filter_const", (1, 1, res_out, res_out), BUFFER))
        e.append((prefix + "upsample.filter.weight", (cout, 1, 4, 4), PARAM))
    return e


def state_entries(resolution: int) -> List[Entry]:
    out: List[Entry] = []
    sres = synthesis_resolutions(resolution)
    for i, r in enumerate(sres):
        p = "synthesis.b%d." % r
        c = nf(r)
        if i == 0:
            out += _sepconv(p + "conv1.", c, c)
            out += _sepconv(p + "conv2.", c, c)
        else:
            out += _sepconv(p + "conv1.", nf(sres[i - 1]), c, r, noise=True, up=True)
            out += _sepconv(p + "conv2.", c, c, r, noise=True)
        out += [(p + "torgb.weight", (3, c, 1, 1), PARAM), (p + "torgb.bias", (3,), PARAM)]
        if i > 0:
            out += [(p + "upsample.filter_const", (1, 1, r, r), BUFFER),
                    (p + "upsample.filter.weight", (3, 1, 4, 4), PARAM)]
    eres = encoder_resolutions(resolution)
    for i, r in enumerate(eres):
        p = "encoder.b%d." % r
        c = nf(r)
        last = i + 1 == len(eres)
        if i == 0:
            out += [(p + "fromrgb.weight", (c, 4, 1, 1), PARAM), (p + "fromrgb.bias", (c,), PARAM)]
        out += _sepconv(p + "conv1.", c, c)
        out += _sepconv(p + "conv2.", c, c if last else nf(eres[i + 1]), down=not last)
    return out
