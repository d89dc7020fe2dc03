"""Standalone op kernels (upfirdn2d, bias_act, conv2d_resample 1x1) against the CPU oracles
restated from torch_utils/ops (oracle/migan_oracle.py) and the committed golden vectors."""
import os

import numpy as np
import pytest
import torch

from migan_b200 import ops
from oracle import migan_oracle as O

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
F2 = O.setup_filter([1, 3, 3, 1])
F1 = torch.tensor([1., 2., 4., 2., 1., .5, .25, .125]) / 10.875
CASES = [(F2, 1, 1, (1, 2, 2, 1), False, 1.0), (F2, 2, 1, (2, 1, 2, 1), False, 4.0),
         (F2, 1, 2, (1, 1, 1, 1), False, 1.0), (F2, 2, 2, (3, 0, 1, 2), True, 2.0),
         (F1, 1, 1, (4, 3, 4, 3), False, 1.0), (F1, 2, 1, (5, 4, 5, 4), True, 4.0),
         (None, 1, 1, (0, 0, 0, 0), False, 1.0), (F2, 1, 1, (-1, 2, 3, -1), False, 1.0)]


def test_upfirdn2d_golden_vectors(cuda_device):
    z = np.load(os.path.join(GOLDEN, "ops.npz"))
    x = torch.from_numpy(z["x"]).to(cuda_device)
    for i, (f, up, down, pad, flip, gain) in enumerate(CASES):
        fd = None if f is None else f.to(cuda_device)
        got = ops.upfirdn2d(x, fd, up=up, down=down, padding=list(pad), flip_filter=flip, gain=gain)
        assert np.allclose(got.cpu().numpy().ravel(), z["up%d" % i], atol=2e-6), i


@pytest.mark.parametrize("shape", [(1, 3, 64, 64), (2, 5, 17, 33), (1, 64, 128, 128)])
@pytest.mark.parametrize("case", range(len(CASES)))
def test_upfirdn2d_matches_oracle(cuda_device, shape, case):
    f, up, down, pad, flip, gain = CASES[case]
    g = torch.Generator().manual_seed(case)
    x = torch.randn(*shape, generator=g)
    want = O.upfirdn2d_ref(x, f, up=up, down=down, padding=pad, flip_filter=flip, gain=gain)
    got = ops.upfirdn2d(x.to(cuda_device), None if f is None else f.to(cuda_device), up=up, down=down, padding=list(pad),
                        flip_filter=flip, gain=gain)
    assert got.shape == want.shape
    assert float((got.cpu() - want).abs().max()) < 1e-5


def test_up_down_sample_helpers_match_module_filters(cuda_device):
    """upsample2d / downsample2d with [1,3,3,1] == the generator's Upsample2d / Downsample2d (SURVEY 8c pin 2)."""
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 8, 16, 16, generator=g)
    f = ops.setup_filter([1, 3, 3, 1])
    assert torch.equal(f, F2)
    up = ops.upsample2d(x.to(cuda_device), f.to(cuda_device)).cpu()
    fc = torch.tensor([[1.0, 0.0], [0.0, 0.0]]).repeat(1, 1, 16, 16)
    want_up = O.upsample2d(x, O.setup_filter([1, 3, 3, 1], gain=4).repeat(8, 1, 1, 1), fc)
    assert float((up - want_up).abs().max()) < 1e-5
    dn = ops.downsample2d(x.to(cuda_device), f.to(cuda_device)).cpu()
    want_dn = O.downsample2d(x, O.setup_filter([1, 3, 3, 1]).repeat(8, 1, 1, 1))
    assert float((dn - want_dn).abs().max()) < 1e-5
    fl = ops.filter2d(x.to(cuda_device), f.to(cuda_device)).cpu()
    assert fl.shape == x.shape


@pytest.mark.parametrize("act", list(O._ACT.keys()))
@pytest.mark.parametrize("clamp", [None, 0.7])
def test_bias_act_matches_oracle(cuda_device, act, clamp):
    g = torch.Generator().manual_seed(3)
    for shape, dim in [((2, 5, 9, 12), 1), ((4, 7), 1), ((3, 6, 5), 0), ((1000,), 0)]:
        x = torch.randn(*shape, generator=g) * 3
        b = torch.randn(shape[dim], generator=g)
        want = O.bias_act_ref(x, b, dim=dim, act=act, clamp=clamp)
        got = ops.bias_act(x.to(cuda_device), b.to(cuda_device), dim=dim, act=act, clamp=clamp).cpu()
        assert float((got - want).abs().max()) < 2e-5, (act, shape)
    x = torch.randn(2, 3, 4, 4, generator=g)
    got = ops.bias_act(x.to(cuda_device), None, act=act, alpha=0.3, gain=0.5).cpu()
    assert float((got - O.bias_act_ref(x, None, act=act, alpha=0.3, gain=0.5)).abs().max()) < 2e-5


def test_bias_act_golden_vectors(cuda_device):
    z = np.load(os.path.join(GOLDEN, "ops.npz"))
    x = torch.from_numpy(z["x"]).to(cuda_device)
    b = torch.from_numpy(z["bvec"]).to(cuda_device)
    i = 0
    for act in O._ACT:
        for clamp in (None, 0.7):
            got = ops.bias_act(x, b, dim=1, act=act, clamp=clamp)
            assert np.allclose(got.cpu().numpy().ravel(), z["ba%d" % i], atol=2e-5), act
            i += 1


def test_conv2d_resample_1x1_branches(cuda_device):
    """conv2d_resample 1x1 + down / up / plain == what SeparableConv2d bakes in (conv2d_resample.py:106-116)."""
    g = torch.Generator().manual_seed(2)
    x = torch.randn(2, 64, 16, 16, generator=g)
    w = torch.randn(128, 64, 1, 1, generator=g) / 8
    f = O.setup_filter([1, 3, 3, 1])
    xd, wd, fd = x.to(cuda_device), w.to(cuda_device), f.to(cuda_device)
    conv = torch.nn.functional.conv2d
    want = conv(O.upfirdn2d_ref(x, f, down=2, padding=(1, 1, 1, 1)), w)
    assert float((ops.conv2d_resample(xd, wd, fd, down=2).cpu() - want).abs().max()) < 1e-4
    want = O.upfirdn2d_ref(conv(x, w), f, up=2, padding=(2, 1, 2, 1), gain=4)
    assert float((ops.conv2d_resample(xd, wd, fd, up=2).cpu() - want).abs().max()) < 1e-4
    assert float((ops.conv2d_resample(xd, wd).cpu() - conv(x, w)).abs().max()) < 1e-4


def test_ops_reject_cpu_tensors():
    with pytest.raises(RuntimeError, match="no CPU"):
        ops.upfirdn2d(torch.zeros(1, 1, 4, 4), None)
    with pytest.raises(RuntimeError, match="no CPU"):
        ops.bias_act(torch.zeros(4))
