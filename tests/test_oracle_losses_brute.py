"""Pin the oracle's loss assembly (PARITY UNPINNED by the reference: its only ternary test is dead
code) with an independent float64 evaluation of the mathematical definitions (oracle/brute.py)."""
import numpy as np
import pytest
import torch

from oracle import brute
from oracle import losses as olosses
from unflow_b200 import synthetic


@pytest.mark.parametrize("mask_occlusion,use_border,dist", [('fb', True, 2), ('', False, 1), ('fb', False, 3)])
def test_compute_losses_vs_float64_definitions(mask_occlusion, use_border, dist):
    im1, im2, ffw, fbw = synthetic.level_inputs(1, 12, 16, seed=11, flow_mag=2.0)
    border = olosses.create_border_mask(im1, 0.1) if use_border else None
    got = olosses.compute_losses(im1, im2, ffw, fbw, border_mask=border, mask_occlusion=mask_occlusion,
                                 data_max_distance=dist)
    want, masks = brute.compute_losses(im1.numpy(), im2.numpy(), ffw.numpy(), fbw.numpy(),
                                       border_mask=None if border is None else border.numpy(),
                                       mask_occlusion=mask_occlusion, data_max_distance=dist)
    # the fb-occlusion decision is a hard threshold: identical unless a pixel sits within float32
    # rounding of it (none with this seed)
    assert np.array_equal(got['_aux']['mask_fw'].expand(1, 12, 16, 1).numpy(), masks[0])
    assert np.array_equal(got['_aux']['mask_bw'].expand(1, 12, 16, 1).numpy(), masks[1])
    for k, v in want.items():
        np.testing.assert_allclose(float(got[k]), v, rtol=5e-5, err_msg=k)


def test_tf_compat_known_answers():
    """TF 1.x conventions the oracle restates, on hand-computable cases."""
    from oracle import tf_compat as tfc
    # SAME padding of the FlowNet encoder at 384x1280
    assert tfc.same_pad(384, 7, 2) == (2, 3) and tfc.same_pad(192, 5, 2) == (1, 2)
    assert tfc.same_pad(48, 3, 2) == (0, 1) and tfc.same_pad(1280, 3, 1) == (1, 1)
    # legacy resize_bilinear (align_corners=False, no half-pixel centres): src = dst * in/out
    x = torch.tensor([0.0, 10.0]).view(1, 1, 2, 1)
    up = tfc.resize_bilinear_legacy(x, [1, 4]).view(-1).tolist()
    assert up == [0.0, 5.0, 10.0, 10.0]
    # resize_area by 2 == 2x2 box mean; rgb_to_grayscale weights
    y = torch.arange(16.0).view(1, 4, 4, 1)
    assert tfc.resize_area(y, [2, 2]).reshape(-1).tolist() == [2.5, 4.5, 10.5, 12.5]
    g = tfc.rgb_to_grayscale(torch.tensor([1.0, 1.0, 1.0]).view(1, 1, 1, 3))
    assert abs(float(g) - 0.9999) < 1e-6
    # conv2d_transpose(k=4, s=2, SAME) doubles the size and equals the adjoint of the SAME s2 conv
    xx = torch.randn(1, 3, 5, 6, dtype=torch.float64)
    w = torch.randn(3, 2, 4, 4, dtype=torch.float64)
    yy = tfc.conv2d_transpose_same(xx, w, None)
    assert yy.shape == (1, 2, 10, 12)
    z = torch.randn(1, 2, 10, 12, dtype=torch.float64)
    conv = torch.nn.functional.conv2d(torch.nn.functional.pad(z, (1, 1, 1, 1)), w, stride=2)
    assert abs(float((yy * z).sum() - (xx * conv).sum())) < 1e-9
