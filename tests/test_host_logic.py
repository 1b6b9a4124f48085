"""CPU tests of the host-side logic of the product (no CUDA compute): variable naming and
checkpoint layout, the conv/deconv wiring of FlowNetS against the oracle, TF-compat resizers,
mask helpers, and that the C-ABI library loads and exports every symbol of include/unflow.h."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from oracle import flownet as oflownet
from oracle import losses as olosses
from oracle import tf_compat as otf

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_abi_library_loads_and_exports_every_declared_symbol():
    from unflow_b200 import _native
    lib = _native.lib()
    header = open(os.path.join(ROOT, "include", "unflow.h")).read()
    declared = set(re.findall(r"\b(unflow_[a-z0-9_]+)\s*\(", header))
    assert declared, "no symbols parsed from include/unflow.h"
    for name in sorted(declared):
        assert hasattr(lib, name), "libunflow.so does not export %s" % name
    assert declared == set(_native.SIGNATURES), "ctypes table out of sync with include/unflow.h"
    assert lib.unflow_abi_version() >= 1


def test_abi_argument_checks_without_gpu():
    """The EINVAL paths mirror the reference's OP_REQUIRES checks and need no device."""
    from unflow_b200 import _native
    lib = _native.lib()
    oc, oh, ow = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    assert lib.unflow_correlation_out_shape(48, 160, 1, 20, 20, 1, 2, ctypes.byref(oc), ctypes.byref(oh),
                                            ctypes.byref(ow)) == 0
    assert (oc.value, oh.value, ow.value) == (441, 48, 160)
    assert lib.unflow_correlation_out_shape(48, 160, 2, 20, 20, 1, 2, None, None, None) == 1
    assert "kernel_size must be odd" in _native.last_error()
    assert lib.unflow_correlation_out_shape(4, 4, 1, 8, 0, 1, 2, None, None, None) == 1
    assert "Invalid correlation settings" in _native.last_error()
    assert lib.unflow_downsample(None, None, 1, 6, 9, 1, 2, None) == 1
    assert "divisible by scale" in _native.last_error()
    assert lib.unflow_backward_warp_fwd(None, None, None, 1, 4, 4, 3, 7, None) == 1
    # empty tensors are fine and launch nothing
    assert lib.unflow_backward_warp_fwd(None, None, None, 0, 4, 4, 3, 0, None) == 0
    assert lib.unflow_forward_warp_fwd(None, None, 0, 4, 4, None) == 0


def test_ops_refuse_cpu_tensors():
    from unflow_b200.e2eflow import ops
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.downsample(torch.zeros(1, 4, 4, 1), 2)
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.forward_warp(torch.zeros(1, 4, 4, 2))


@pytest.mark.parametrize("spec,full_res", [("S", False), ("C", False), ("CSS", False), ("s", False),
                                           ("cs", False), ("S", True), ("CS", True)])
def test_variable_inventory_matches_oracle(spec, full_res):
    from unflow_b200.e2eflow.core.flownet import FlowNetVariables
    v = FlowNetVariables(spec, full_res, seed=0)
    tf = v.to_tf_dict()
    want = oflownet.variable_shapes(spec, full_res)
    assert set(tf) == set(want)
    for k, shape in want.items():
        assert tuple(tf[k].shape) == tuple(shape), k


def test_flownet_c_parameter_count():
    from unflow_b200.e2eflow.core.flownet import FlowNetVariables
    v = FlowNetVariables("C", False, seed=0)
    assert sum(p.numel() for p in v.parameters()) == 39175298


def test_tf_dict_roundtrip():
    from unflow_b200.e2eflow.core.flownet import FlowNetVariables
    a = FlowNetVariables("s", False, seed=1)
    b = FlowNetVariables("s", False, seed=2)
    b.load_tf_dict(a.to_tf_dict())
    for (n1, p1), (n2, p2) in zip(a.named_parameters(), b.named_parameters()):
        assert n1 == n2 and torch.equal(p1, p2)


@pytest.mark.parametrize("spec,full_res,hw", [("s", False, (64, 128)), ("s", True, (64, 64)),
                                              ("S", False, (64, 64))])
def test_flownet_s_wiring_vs_oracle_cpu(spec, full_res, hw):
    """FlowNetS uses no custom op, so the product's layer wiring (TF SAME padding, deconv
    layout, concat order, bidirectional batching) can be checked on the CPU."""
    from unflow_b200.e2eflow.core.flownet import FlowNetVariables, flownet
    tfv = oflownet.init_variables(spec, full_res, seed=3)
    v = FlowNetVariables(spec, full_res, seed=0).load_tf_dict(tfv)
    g = torch.Generator().manual_seed(0)
    im1 = torch.rand(2, hw[0], hw[1], 3, generator=g) - 0.5
    im2 = torch.rand(2, hw[0], hw[1], 3, generator=g) - 0.5
    want_fw, want_bw = oflownet.flownet(tfv, im1, im2, spec, full_res, backward_flow=True)
    got_fw, got_bw = flownet(im1, im2, spec, full_res, backward_flow=True, variables=v)
    for w_list, g_list in ((want_fw[0], got_fw[0]), (want_bw[0], got_bw[0])):
        assert len(w_list) == len(g_list) == (7 if full_res else 5)
        for w, gg in zip(w_list, g_list):
            assert w.shape == gg.shape
            np.testing.assert_allclose(gg.detach().numpy(), w.numpy(), rtol=1e-4,
                                       atol=1e-5 * float(w.abs().max()))


def test_same_padding_rule():
    from unflow_b200.e2eflow.core.flownet import _same_pad
    assert _same_pad(384, 7, 2) == (2, 3)
    assert _same_pad(192, 5, 2) == (1, 2)
    assert _same_pad(48, 3, 2) == (0, 1)
    assert _same_pad(48, 3, 1) == (1, 1)
    assert _same_pad(7, 3, 2) == (1, 1)
    assert otf.same_pad(384, 7, 2) == (2, 3)


def test_resizers_vs_oracle():
    from unflow_b200.e2eflow.core import tf_image
    x = torch.randn(2, 6, 10, 2, generator=torch.Generator().manual_seed(0))
    np.testing.assert_allclose(tf_image.resize_bilinear(x, [24, 40]).numpy(),
                               otf.resize_bilinear_legacy(x, [24, 40]).numpy(), rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(tf_image.resize_area(x, [3, 5]).numpy(),
                               otf.resize_area(x, [3, 5]).numpy(), rtol=1e-6, atol=1e-6)
    # legacy bilinear: src = dst * in/out, no half-pixel shift -> exact copies on the grid
    up = tf_image.resize_bilinear(x, [12, 20])
    assert torch.equal(up[:, ::2, ::2], x)
    # area resize by an integer factor is the box mean
    y = torch.randn(1, 9, 15, 1, generator=torch.Generator().manual_seed(1))
    np.testing.assert_allclose(otf.resize_area(y, [3, 5]).numpy(),
                               y.reshape(1, 3, 3, 5, 3, 1).mean(dim=(2, 4)).numpy(), rtol=1e-5, atol=1e-6)


def test_loss_helpers_vs_oracle_cpu():
    """The stand-alone loss terms are plain torch and run on CPU tensors: check them against the
    oracle (which follows the reference's conv2d formulation)."""
    from unflow_b200.e2eflow.core import losses as L
    g = torch.Generator().manual_seed(0)
    flow = torch.randn(2, 7, 9, 2, generator=g) * 2
    im1 = torch.rand(2, 7, 9, 3, generator=g)
    im2 = torch.rand(2, 7, 9, 3, generator=g)
    mask = (torch.rand(2, 7, 9, 1, generator=g) > 0.3).float()
    for name, args in [("smoothness_loss", (flow,)), ("second_order_loss", (flow,)),
                       ("gradient_loss", (im1, im2, mask)), ("photometric_loss", (im1 - im2, mask)),
                       ("charbonnier_loss", (flow, mask)), ("divergence", (flow,)),
                       ("create_outgoing_mask", (flow,)), ("create_border_mask", (im1, 0.25))]:
        got, want = getattr(L, name)(*args), getattr(olosses, name)(*args)
        np.testing.assert_allclose(got.numpy(), want.numpy(), rtol=2e-5, atol=1e-6, err_msg=name)
    for d in (1, 2, 3):
        np.testing.assert_allclose(L.ternary_loss(im1, im2, mask, max_distance=d).item(),
                                   olosses.ternary_loss(im1, im2, mask, max_distance=d).item(), rtol=2e-5)
    for a, b in zip(L._smoothness_deltas(flow), olosses._smoothness_deltas(flow)):
        np.testing.assert_allclose(a.numpy(), b.numpy(), rtol=1e-6, atol=1e-6)
    for a, b in zip(L._second_order_deltas(flow), olosses._second_order_deltas(flow)):
        np.testing.assert_allclose(a.numpy(), b.numpy(), rtol=1e-6, atol=1e-6)
    assert torch.equal(L.create_mask(flow, [[1, 2], [0, 3]]), olosses.create_mask(flow, [[1, 2], [0, 3]]))


def test_install_as_e2eflow_alias():
    import unflow_b200
    unflow_b200.install_as_e2eflow()
    from e2eflow.core.losses import charbonnier_loss, DISOCC_THRESH  # noqa: F401
    from e2eflow.core.flownet import FLOW_SCALE
    from e2eflow.core.unsupervised import LOSSES
    assert FLOW_SCALE == 5.0 and DISOCC_THRESH == 0.8
    assert LOSSES == ['occ', 'sym', 'fb', 'grad', 'ternary', 'photo', 'smooth_1st', 'smooth_2nd']


def test_unsupervised_loss_needs_cuda_inputs():
    from unflow_b200.e2eflow.core.unsupervised import unsupervised_loss
    with pytest.raises(RuntimeError, match="CUDA"):
        unsupervised_loss((torch.zeros(1, 64, 64, 3),) * 2, {'flownet': 'S', 'ternary_weight': 1.0},
                          normalization=([0, 0, 0], 1.0), augment=False)


def test_every_product_module_imports_without_gpu():
    import importlib
    for m in ("ops", "core.image_warp", "core.losses", "core.fused_loss", "core.flownet",
              "core.unsupervised", "core.util", "core.tf_image"):
        importlib.import_module("unflow_b200.e2eflow." + m)


def test_run_config_semantics(tmp_path):
    """config.ini parsing as the reference does it (util.py:37-73, run.py:90-92)."""
    from unflow_b200 import run as R
    ini = tmp_path / "config.ini"
    ini.write_text("""
[dirs]
log = ../log
[run]
batch_size = 4
gpu_list = 0
dataset = kitti_ft
development = False
[train]
decay_interval = 100000
save_interval = 5000
flownet = C
pyramid_loss = True
ternary_weight = 1.0
smooth_2nd_weight = 3.0
[train_kitti_ft]
height = 320
width = 768
manual_decay_iters = 45000,20000
manual_decay_lrs = 0.5e-5,0.25e-5
""")
    cfg = R.config_dict(str(ini))
    assert cfg['run']['batch_size'] == 4 and cfg['run']['development'] is False
    assert cfg['train']['ternary_weight'] == 1.0 and cfg['train']['pyramid_loss'] is True
    assert cfg['train']['flownet'] == 'C' and isinstance(cfg['train']['decay_interval'], int)
    p = dict(cfg['train'])
    p.update(cfg['train_kitti_ft'])
    R.convert_input_strings(p)
    assert p['manual_decay_iters'] == [45000, 20000] and p['manual_decay_lrs'] == [0.5e-5, 0.25e-5]
    assert p['num_iters'] == 65000 and p['height'] == 320
    (tmp_path / "model.ckpt-5000.pt").write_bytes(b"")
    (tmp_path / "model.ckpt-15000.pt").write_bytes(b"")
    assert R.latest_checkpoint(str(tmp_path))[0] == 15000


def test_l2_mask_bytes_for_the_adam_kernel():
    """core/train.py: one bit per parameter marks the `weights` variables (slim.l2_regularizer, reference
    flownet.py:176) for the fused Adam kernel: bit k of byte i = element 4 * i + k of the flat buffer."""
    from unflow_b200.e2eflow.core.train import l2_mask_bytes
    # weights [0,5), biases [5,7), weights [7,10), padding [10,12)
    isw, packed = l2_mask_bytes(12, [0, 5, 7], [5, 2, 3], [True, False, True])
    assert isw.tolist() == [1, 1, 1, 1, 1, 0, 0, 1, 1, 1, 0, 0]
    assert packed.tolist() == [0b1111, 0b1001, 0b0011] and packed.dtype == torch.uint8
    for i in range(12):
        assert (int(packed[i // 4]) >> (i % 4)) & 1 == int(isw[i])


def test_nhwc_geometry_of_channel_and_batch_slices():
    """ops._nhwc_geometry: (batch stride, pixel pitch) of NCHW-shaped tensors with NHWC memory -- dense, a channel
    slice of a pitch-padded concat buffer, a batch slice; None for NCHW memory."""
    from unflow_b200.e2eflow.ops import _nhwc_geometry
    buf = torch.empty(4, 5, 6, 12)
    assert _nhwc_geometry(buf.permute(0, 3, 1, 2)) == (360, 12)
    assert _nhwc_geometry(buf[..., 2:9].permute(0, 3, 1, 2)) == (360, 12)
    assert _nhwc_geometry(buf[1:3, :, :, :8].permute(0, 3, 1, 2)) == (360, 12)
    assert _nhwc_geometry(torch.empty(2, 8, 5, 6)) is None
    assert _nhwc_geometry(torch.empty(2, 8, 5, 6).contiguous(memory_format=torch.channels_last)) == (240, 8)
