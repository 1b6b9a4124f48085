"""Row N3: flow metrics / visualisation / on-disk formats (CPU)."""
import numpy as np
import torch

from unflow_b200.e2eflow.core import flow_io, flow_util


def test_flo_roundtrip(tmp_path):
    rs = np.random.RandomState(0)
    flow = rs.randn(7, 9, 2).astype(np.float32) * 10
    flow[2, 3] = 1e10     # Middlebury "unknown" marker
    p = str(tmp_path / "a.flo")
    flow_io.write_flo(p, flow)
    got, mask = flow_io.read_flo(p)
    assert np.array_equal(got, flow) and mask.shape == (7, 9, 1)
    assert mask[2, 3, 0] == 0 and mask.sum() == 7 * 9 - 1
    raw = open(p, 'rb').read()
    assert raw[:4] == b'PIEH' and len(raw) == 12 + 7 * 9 * 8


def test_kitti_png16_roundtrip(tmp_path):
    rs = np.random.RandomState(1)
    flow = np.round(rs.randn(11, 13, 2) * 30 * 64) / 64.0      # representable at 1/64 px
    mask = (rs.rand(11, 13) > 0.3).astype(np.float32)
    p = str(tmp_path / "f.png")
    flow_io.write_kitti_flow(p, flow, mask)
    got, gmask = flow_io.read_kitti_flow(p)
    np.testing.assert_array_equal(got, flow.astype(np.float32))
    np.testing.assert_array_equal(gmask[:, :, 0], mask)
    # encoding rule of the reference reader: (v - 2^15) / 64
    arr = flow_io.read_png16(p)
    assert arr.dtype == np.uint16 and arr[0, 0, 0] == int(round(flow[0, 0, 0] * 64 + 2 ** 15))


def test_metrics():
    gt = torch.zeros(1, 4, 4, 2)
    fl = torch.zeros(1, 4, 4, 2)
    fl[0, 0, 0] = torch.tensor([3.0, 4.0])      # endpoint error 5
    fl[0, 1, 1] = torch.tensor([1.0, 0.0])      # error 1 (< 3 px: not an outlier)
    mask = torch.ones(1, 4, 4, 1)
    assert abs(float(flow_util.flow_error_avg(fl, gt, mask)) - 6.0 / 16) < 1e-6
    assert abs(float(flow_util.outlier_pct(gt, fl, mask)) - 100.0 / 16) < 1e-5
    mask[0, 0, 0, 0] = 0
    assert abs(float(flow_util.flow_error_avg(fl, gt, mask)) - 1.0 / 15) < 1e-6
    # relative threshold: 5 % of a 100 px ground-truth motion = 5 px > 3 px
    gt2 = torch.full((1, 2, 2, 2), 100.0 / 2 ** 0.5)
    assert float(flow_util.outlier_ratio(gt2, gt2 + 2.9, torch.ones(1, 2, 2, 1))) == 0.0


def test_flow_to_color_and_error_image():
    flow = torch.zeros(1, 2, 2, 2)
    flow[0, 0, 0] = torch.tensor([1.0, 0.0])     # hue 0 -> red
    flow[0, 0, 1] = torch.tensor([-1.0, 0.0])    # hue 0.5 -> cyan
    flow[0, 1, 0] = torch.tensor([0.0, 1.0])     # pi via the reference's atan2 special case
    flow[0, 1, 1] = torch.tensor([0.5, 0.5])
    im = flow_util.flow_to_color(flow, max_flow=8)
    assert im.shape == (1, 2, 2, 3)
    np.testing.assert_allclose(im[0, 0, 0].numpy(), [1.0, 0.0, 0.0], atol=1e-6)
    np.testing.assert_allclose(im[0, 0, 1].numpy(), [0.0, 1.0, 1.0], atol=1e-6)
    err = flow_util.flow_error_image(flow, torch.ones(1, 2, 2, 2) * 4, torch.ones(1, 2, 2, 1))
    assert err.shape == (1, 2, 2, 3) and float(err.min()) >= 0 and float(err.max()) <= 1
    err2 = flow_util.flow_error_image(flow, flow + 10.0, torch.ones(1, 2, 2, 1), log_colors=False)
    np.testing.assert_allclose(err2.numpy(), 1.0)


def test_resize_output_flow():
    f = torch.ones(1, 4, 6, 2)
    out = flow_io.resize_output_flow(f, 8, 18)
    assert out.shape == (1, 8, 18, 2)
    np.testing.assert_allclose(out[..., 0].numpy(), 3.0, rtol=1e-6)   # u scaled by 18/6
    np.testing.assert_allclose(out[..., 1].numpy(), 2.0, rtol=1e-6)   # v scaled by 8/4
