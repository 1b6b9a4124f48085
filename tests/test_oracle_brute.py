"""Pin the oracle where the reference leaves the op unpinned:
compare oracle_ops.c / oracle python with independent float64 definitions
(oracle/brute.py) and check analytic-vs-numeric Jacobians at the reference's
own 1e-3 tolerance."""
import numpy as np
import pytest
import torch

from oracle import brute
from oracle import ops as oops
from oracle.image_warp import image_warp as oimage_warp
from test_oracle_reference_kats import analytic_jacobian, numeric_jacobian


def rnd(shape, seed, scale=1.0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed)) * scale


CORR_CASES = [
    dict(shape=(2, 5, 7, 9), kernel_size=1, max_displacement=4, pad=4, stride_1=1, stride_2=2),
    dict(shape=(1, 3, 6, 8), kernel_size=1, max_displacement=3, pad=3, stride_1=1, stride_2=1),
    dict(shape=(1, 4, 9, 10), kernel_size=3, max_displacement=2, pad=3, stride_1=1, stride_2=1),
    dict(shape=(2, 2, 9, 11), kernel_size=3, max_displacement=2, pad=3, stride_1=2, stride_2=2),
    dict(shape=(1, 33, 8, 8), kernel_size=1, max_displacement=2, pad=2, stride_1=1, stride_2=2),
    dict(shape=(1, 3, 8, 9), kernel_size=1, max_displacement=4, pad=2, stride_1=1, stride_2=2),
]


@pytest.mark.parametrize("case", CORR_CASES)
def test_correlation_vs_brute(case):
    case = dict(case)
    shape = case.pop("shape")
    a, b = rnd(shape, 1), rnd(shape, 2)
    got = oops.correlation(a, b, **case).numpy()
    want = brute.correlation(a.numpy(), b.numpy(), **case)
    assert got.shape == want.shape
    np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("case", CORR_CASES[:4])
def test_correlation_grad_vs_brute_fd(case):
    """oracle backward (restated CorrelateDataBackward0/1) vs finite differences of the
    float64 brute-force forward."""
    case = dict(case)
    shape = case.pop("shape")
    shape = (1, min(shape[1], 3), shape[2], shape[3])
    a, b = rnd(shape, 3), rnd(shape, 4)
    a.requires_grad_(True)
    b.requires_grad_(True)
    out = oops.correlation(a, b, **case)
    w = rnd(out.shape, 5)
    (out * w).sum().backward()
    wn = w.numpy().astype(np.float64)
    f = lambda x, y: float((brute.correlation(x, y, **case) * wn).sum())
    an, bn = a.detach().numpy().astype(np.float64), b.detach().numpy().astype(np.float64)
    eps = 1e-4
    rs = np.random.RandomState(0)
    for which, g in ((0, a.grad), (1, b.grad)):
        for _ in range(12):
            idx = tuple(rs.randint(0, s) for s in shape)
            d = np.zeros(shape)
            d[idx] = eps
            if which == 0:
                fd = (f(an + d, bn) - f(an - d, bn)) / (2 * eps)
            else:
                fd = (f(an, bn + d) - f(an, bn - d)) / (2 * eps)
            np.testing.assert_allclose(g[idx].item(), fd, rtol=1e-3, atol=1e-4)


def test_correlation_invalid_args():
    a = rnd((1, 2, 4, 4), 1)
    with pytest.raises(ValueError):
        oops.correlation(a, a, kernel_size=2)
    with pytest.raises(ValueError):
        oops.correlation(a, rnd((1, 2, 4, 5), 2))
    with pytest.raises(ValueError):
        oops.correlation(a, a, max_displacement=8, pad=0)


def test_backward_warp_vs_brute_zero_border():
    im = rnd((2, 6, 7, 3), 1)
    fl = rnd((2, 6, 7, 2), 2, 3.0)   # many taps leave the image
    np.testing.assert_allclose(oops.backward_warp(im, fl).numpy(),
                               brute.backward_warp_zero(im.numpy(), fl.numpy()), rtol=1e-5, atol=1e-5)


def test_image_warp_vs_brute_clamp_border():
    im = rnd((2, 6, 7, 3), 1)
    fl = rnd((2, 6, 7, 2), 2, 3.0)
    np.testing.assert_allclose(oimage_warp(im, fl).numpy(),
                               brute.backward_warp_clamp(im.numpy(), fl.numpy()), rtol=1e-5, atol=1e-5)


def test_zero_and_clamp_borders_differ():
    """R3: the op drops out-of-image taps, image_warp clamps them."""
    im = torch.ones(1, 4, 4, 1)
    fl = torch.full((1, 4, 4, 2), -1.5)
    z = oops.backward_warp(im, fl)
    c = oimage_warp(im, fl)
    assert torch.allclose(c, torch.ones_like(c))
    assert z[0, 0, 0, 0].item() == 0.0 and not torch.allclose(z, c)


def test_forward_warp_vs_brute():
    fl = rnd((2, 9, 11, 2), 7, 2.5)
    np.testing.assert_allclose(oops.forward_warp(fl).numpy(), brute.forward_warp(fl.numpy()),
                               rtol=1e-5, atol=1e-5)
    # zero flow: every pixel receives the full (clipped) Gaussian stencil
    z = oops.forward_warp(torch.zeros(1, 12, 12, 2))
    centre = sum(np.exp(-(x * x + y * y) / 2.0) for x in range(-4, 5) for y in range(-4, 5))
    np.testing.assert_allclose(z[0, 6, 6, 0].item(), centre, rtol=1e-5)


def test_downsample_vs_brute_and_errors():
    im = rnd((2, 8, 12, 3), 3)
    for s in (2, 4):
        np.testing.assert_allclose(oops.downsample(im, s).numpy(), brute.downsample(im.numpy(), s),
                                   rtol=1e-6, atol=1e-6)
    with pytest.raises(ValueError):
        oops.downsample(rnd((1, 6, 9, 1), 1), 2)


def test_backward_warp_jacobian_flow_only():
    im = rnd((1, 5, 5, 2), 1)
    fl = rnd((1, 5, 5, 2), 2, 1.2)
    fn = lambda f: oops.backward_warp(im, f)
    np.testing.assert_allclose(analytic_jacobian(fn, fl).numpy(), numeric_jacobian(fn, fl, 1e-3).numpy(),
                               rtol=1e-2, atol=2e-3)
    x = im.clone().requires_grad_(True)
    out = oops.backward_warp(x, fl)
    assert torch.autograd.grad(out.sum(), x, allow_unused=True)[0] is None  # ops.py:84
