"""The reference's own CUDA kernels, compiled against stand-in TensorFlow headers into
oracle/_ref/libref_ops.so (oracle/ref_kernels.py), as ground truth for the four custom ops.

* CPU part (always on where the library exists): the correlation geometry comes from the reference's
  ``CorrelationState`` host code -- compared with the oracle and with the product's C ABI.
* GPU part (runs with every ``-m gpu`` pass): the oracle's C restatement and the product's kernels
  are compared with the reference kernels on the same inputs -- forward and gradients of all four
  ops, including displacement > 0 / C > 1 / K = 3 / strides, which the reference's own KATs never
  exercise (test/ops/correlation.py:30-89)."""
import ctypes
import itertools
import os

import numpy as np
import pytest
import torch

from oracle import ops as oops
from oracle import ref_kernels as RK

needs_lib = pytest.mark.skipif(not RK.available(), reason="oracle/_ref/libref_ops.so not built (needs the reference tree)")


@needs_lib
def test_correlation_geometry_from_the_reference_host_code():
    from unflow_b200 import _native
    lib = _native.lib()
    n = 0
    for H, W, ks, md, pad, s1, s2 in itertools.product((48, 37), (160, 64), (1, 3), (20, 4, 0), (20, 4, 0), (1, 2), (1, 2)):
        want = RK.correlation_out_shape(256, H, W, ks, md, pad, s1, s2)
        oc, oh, ow = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        rc = oops.lib().oracle_correlation_shape(H, W, ks, md, pad, s1, s2, ctypes.byref(oc), ctypes.byref(oh), ctypes.byref(ow))
        c2, h2, w2 = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        rc2 = lib.unflow_correlation_out_shape(H, W, ks, md, pad, s1, s2, ctypes.byref(c2), ctypes.byref(h2),
                                               ctypes.byref(w2))
        shape = (c2.value, h2.value, w2.value)
        if want[1] <= 0 or want[2] <= 0:            # the reference op rejects these (correlation_op.cc:60-61)
            assert rc != 0 and rc2 != 0
            continue
        n += 1
        assert rc == 0 and (oc.value, oh.value, ow.value) == want, (H, W, ks, md, pad, s1, s2)
        assert rc2 == 0 and tuple(shape) == want, (H, W, ks, md, pad, s1, s2)
    assert n > 100


def _close(a, b, tol=1e-5):
    a, b = a.detach().cpu(), b.detach().cpu()
    scale = max(float(b.abs().max()), 1e-12)
    assert float((a - b).abs().max()) <= tol * scale, float((a - b).abs().max()) / scale


@pytest.mark.gpu
@pytest.mark.parametrize("B,C,H,W,attrs", [(2, 32, 12, 20, dict(kernel_size=1, max_displacement=4, pad=4, stride_1=1, stride_2=2)),
                                           (1, 16, 10, 14, dict(kernel_size=3, max_displacement=3, pad=4, stride_1=2, stride_2=1)),
                                           (1, 256, 48, 160, dict(kernel_size=1, max_displacement=20, pad=20, stride_1=1, stride_2=2))])
def test_correlation_kernels(B, C, H, W, attrs):
    from unflow_b200.e2eflow import ops
    assert RK.available(), "oracle/_ref/libref_ops.so did not travel to the GPU box"
    g = torch.Generator().manual_seed(C + H)
    a, b = torch.randn(B, C, H, W, generator=g), torch.randn(B, C, H, W, generator=g)
    ref, p0, p1 = RK.correlation(a.cuda(), b.cuda(), **attrs)
    _close(oops.correlation(a, b, **attrs), ref)
    go = torch.randn(ref.shape, generator=g)
    r0, r1 = RK.correlation_grad(go.cuda(), p0, p1, (B, C, H, W), **attrs)
    ao, bo = a.clone().requires_grad_(True), b.clone().requires_grad_(True)
    oops.correlation(ao, bo, **attrs).backward(go)
    _close(ao.grad, r0)
    _close(bo.grad, r1)
    ac, bc = a.cuda().requires_grad_(True), b.cuda().requires_grad_(True)
    out = ops.correlation(ac, bc, **attrs)
    _close(out, ref)
    out.backward(go.cuda())
    _close(ac.grad, r0, 2e-5)
    _close(bc.grad, r1, 2e-5)


@pytest.mark.gpu
def test_warp_and_downsample_kernels():
    from unflow_b200.e2eflow import ops
    assert RK.available(), "oracle/_ref/libref_ops.so did not travel to the GPU box"
    g = torch.Generator().manual_seed(3)
    im = torch.rand(2, 18, 26, 3, generator=g)
    fl = torch.randn(2, 18, 26, 2, generator=g) * 4
    ref = RK.backward_warp(im.cuda(), fl.cuda())
    _close(oops.backward_warp(im, fl), ref)
    _close(ops.backward_warp(im.cuda(), fl.cuda()), ref)
    go = torch.randn(ref.shape, generator=g)
    rg = RK.backward_warp_grad(go.cuda(), im.cuda(), fl.cuda())
    fo = fl.clone().requires_grad_(True)
    oops.backward_warp(im, fo).backward(go)
    _close(fo.grad, rg)
    fc = fl.cuda().requires_grad_(True)
    ops.backward_warp(im.cuda(), fc).backward(go.cuda())
    _close(fc.grad, rg, 2e-5)

    ref = RK.forward_warp(fl.cuda())
    _close(oops.forward_warp(fl), ref, 1e-4)          # float atomics: order-dependent rounding
    _close(ops.forward_warp(fl.cuda()), ref, 1e-4)
    go = torch.randn(ref.shape, generator=g)
    rg = RK.forward_warp_grad(go.cuda(), fl.cuda())
    fo = fl.clone().requires_grad_(True)
    oops.forward_warp(fo).backward(go)
    _close(fo.grad, rg, 1e-4)
    fc = fl.cuda().requires_grad_(True)
    ops.forward_warp(fc).backward(go.cuda())
    _close(fc.grad, rg, 1e-4)

    x = torch.rand(2, 16, 24, 3, generator=g)
    for scale in (2, 4):
        ref = RK.downsample(x.cuda(), scale)
        _close(oops.downsample(x, scale), ref)
        _close(ops.downsample(x.cuda(), scale), ref)
