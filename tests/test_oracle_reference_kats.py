"""Pin the CPU oracle with every live known-answer test the reference holds for
the hot path.  Vectors: tests/golden/reference_kats.json
(extracted from the reference test sources by tests/golden/make_reference_kats.py)."""
import numpy as np
import pytest
import torch

from oracle import ops as oops
from oracle import losses as olosses
from oracle.image_warp import image_warp as oimage_warp


def t(a):
    return torch.tensor(np.asarray(a), dtype=torch.float32)


def numeric_jacobian(fn, x, eps=1e-3):
    """Central-difference Jacobian like tf.test gradient_checker (delta=1e-3)."""
    x = x.clone()
    y0 = fn(x)
    J = torch.zeros(x.numel(), y0.numel())
    xf = x.view(-1)
    for i in range(xf.numel()):
        orig = xf[i].item()
        xf[i] = orig + eps
        yp = fn(x).reshape(-1).clone()
        xf[i] = orig - eps
        ym = fn(x).reshape(-1).clone()
        xf[i] = orig
        J[i] = (yp - ym) / (2 * eps)
    return J


def analytic_jacobian(fn, x):
    x = x.clone().requires_grad_(True)
    y = fn(x)
    J = torch.zeros(x.numel(), y.numel())
    for j in range(y.numel()):
        g = torch.zeros_like(y).reshape(-1)
        g[j] = 1.0
        (gx,) = torch.autograd.grad(y, x, g.view_as(y), retain_graph=True, allow_unused=True)
        J[:, j] = gx.reshape(-1)
    return J


@pytest.mark.parametrize("name", ["test_correlation_trivial", "test_correlation_batch"])
def test_correlation_kats(kats, arr, name):
    call = kats["correlation"][name]["calls"][0]
    in0, in1, expected = (arr(a) for a in call["args"])
    kw = call["kwargs"]
    out = oops.correlation(t(in0), t(in1), **kw)
    np.testing.assert_allclose(out.numpy(), expected, rtol=1e-6, atol=1e-6)
    # analytic vs numeric Jacobian at rtol=atol=1e-3 (test/ops/correlation.py:21-28)
    for which in (0, 1):
        def fn(x):
            return oops.correlation(x, t(in1), **kw) if which == 0 else oops.correlation(t(in0), x, **kw)
        x0 = t(in0) if which == 0 else t(in1)
        np.testing.assert_allclose(analytic_jacobian(fn, x0).numpy(), numeric_jacobian(fn, x0).numpy(),
                                   rtol=1e-3, atol=1e-3)


@pytest.mark.parametrize("name", ["test_move", "test_batches", "test_interpolate"])
def test_backward_warp_kats(kats, arr, name):
    call = kats["backward_warp"][name]["calls"][0]
    first, second, flow = (arr(a) for a in call["args"])
    pred = oops.backward_warp(t(second), t(flow))
    np.testing.assert_allclose(pred.numpy(), first, rtol=1e-6, atol=1e-6)
    fn = lambda f: oops.backward_warp(t(second), f)
    # test/ops/backward_warp.py:23-25: Jacobian w.r.t. flow at 1e-3.  The fixtures sit on
    # integer flows (bilinear kinks), where a central difference straddles two linear pieces;
    # the reference checker has the same property, so perturb off the kink first.
    f0 = t(flow) + 0.25
    np.testing.assert_allclose(analytic_jacobian(fn, f0).numpy(), numeric_jacobian(fn, f0).numpy(),
                               rtol=1e-3, atol=1e-3)


@pytest.mark.parametrize("name", ["test_move", "test_batches", "test_interpolate"])
def test_image_warp_kats(kats, arr, name):
    call = kats["image_warp"][name]["calls"][0]
    first, second, flow = (arr(a) for a in call["args"])
    pred = oimage_warp(t(second), t(flow))
    np.testing.assert_allclose(pred.numpy(), first, rtol=1e-6, atol=1e-6)


def test_downsample_kat(kats, arr):
    v = kats["downsample"]["test_downsample"]["vars"]
    first = arr(v["first"]).reshape(1, 4, 4, 1)
    second = arr(v["second"]).reshape(1, 2, 2, 1)
    np.testing.assert_allclose(oops.downsample(t(first), 2).numpy(), second, rtol=1e-6, atol=1e-6)


def test_smoothness_deltas_kat(kats, arr):
    e = kats["losses"]["test_smoothness_deltas"]
    flow = t(arr(e["vars"]["flow"]))
    du, dv, mask = olosses._smoothness_deltas(flow)
    du = du * mask
    dv = dv * mask
    got = {"mask": mask.numpy(), "delta_u": du.numpy(), "delta_v": dv.numpy()}
    assert len(e["calls"]) == 6
    for c in e["calls"]:
        expr = c["args"][0]["expr"]
        actual = eval(expr, {}, got)
        np.testing.assert_array_equal(actual, arr(c["args"][1]))


@pytest.mark.parametrize("name", ["test_create_outgoing_mask_all_directions",
                                  "test_create_outgoing_mask_large_movement"])
def test_outgoing_mask_kats(kats, arr, name):
    e = kats["losses"][name]
    flow = t(arr(e["vars"]["flow"]))
    mask = olosses.create_outgoing_mask(flow).numpy()
    c = e["calls"][0]
    np.testing.assert_array_equal(eval(c["args"][0]["expr"], {}, {"mask": mask}), arr(c["args"][1]))


def test_gradient_loss_kat(kats, arr):
    e = kats["losses"]["test_gradient_loss"]
    v = e["vars"]
    loss = olosses.gradient_loss(t(arr(v["im1"])), t(arr(v["im2"])), t(arr(v["mask"])))
    c = e["calls"][0]
    np.testing.assert_allclose(loss.item(), c["args"][1], atol=c["kwargs"]["atol"])


def test_forward_warp_jacobian():
    """test/ops/forward_warp.py:9-19: analytic vs numeric Jacobian at 1e-3 on a 1x10x10 flow."""
    g = torch.Generator().manual_seed(0)
    flow = torch.randn(1, 10, 10, 2, generator=g) * 1.5
    np.testing.assert_allclose(analytic_jacobian(oops.forward_warp, flow).numpy(),
                               numeric_jacobian(oops.forward_warp, flow).numpy(), rtol=1e-3, atol=1e-3)
