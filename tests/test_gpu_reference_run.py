"""The CUDA product against the golden vectors produced by the reference's own Python source
(tests/golden/reference_run.npz, see tests/golden/make_reference_run.py).  The tight comparisons
are product-vs-oracle (other test files) and oracle-vs-vectors (CPU); this file closes the triangle
directly, with tolerances one notch looser than those."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import flownet as oflownet
import synth

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_run.npz"))
WEIGHTS = dict(ternary=1.0, smooth_2nd=3.0, fb=0.2, occ=12.4, photo=0.5, grad=0.25, smooth_1st=0.75, sym=0.3)


def t(name):
    return torch.from_numpy(G[name]).clone().cuda()


def close(got, want, rtol, atol_rel, msg=""):
    want = np.asarray(want)
    atol = atol_rel * max(float(np.abs(want).max()), 1e-12)
    np.testing.assert_allclose(got.detach().cpu().numpy(), want, rtol=rtol, atol=atol, err_msg=msg)


def _variables(spec, seed, key):
    v = oflownet.init_variables(spec, False, seed=seed)
    s = sum(float(x.double().sum()) for x in v.values())
    a = sum(float(x.double().abs().sum()) for x in v.values())
    if not np.allclose([s, a], G[key], rtol=1e-12):
        pytest.skip("this torch build draws different random weights than the one the fixture was made with")
    return v


@pytest.mark.parametrize("tag,mode,use_border,dist", [('fb', 'fb', True, 3), ('none', '', False, 1), ('disocc', 'disocc', True, 2)])
def test_compute_losses_against_reference_run(tag, mode, use_border, dist):
    from unflow_b200.e2eflow.core import losses as L
    im1, im2 = t('L_im1'), t('L_im2')
    fw, bw = t('L_ffw').requires_grad_(True), t('L_fbw').requires_grad_(True)
    border = L.create_border_mask(im1, 0.1) if use_border else None
    res = L.compute_losses(im1, im2, fw, bw, border_mask=border, mask_occlusion=mode, data_max_distance=dist)
    total = 0.0
    for k in sorted(WEIGHTS):
        close(res[k], G['cl_%s_%s' % (tag, k)], rtol=1e-3, atol_rel=1e-5, msg=k)
        total = total + WEIGHTS[k] * res[k]
    total.backward()
    close(fw.grad, G['cl_%s_dfw' % tag], rtol=5e-3, atol_rel=5e-4, msg="dflow_fw")
    close(bw.grad, G['cl_%s_dbw' % tag], rtol=5e-3, atol_rel=5e-4, msg="dflow_bw")


def test_image_warp_against_reference_run():
    from unflow_b200.e2eflow.core.image_warp import image_warp
    close(image_warp(t('L_im1'), t('L_ffw')), G['warp_out'], rtol=1e-4, atol_rel=1e-5)


@pytest.mark.parametrize("tag,spec,seed", [('c', 'c', 21), ('s', 's', 22), ('cs', 'cs', 23)])
def test_flownet_against_reference_run(tag, spec, seed):
    from unflow_b200.e2eflow.core.flownet import FlowNetVariables, flownet
    v = FlowNetVariables(spec, False, seed=0).load_tf_dict(_variables(spec, seed, 'fn_%s_vars' % tag)).cuda()
    with torch.no_grad():
        fw, bw = flownet(t('fn_%s_im1' % tag), t('fn_%s_im2' % tag), spec, backward_flow=True, variables=v)
    for n in range(len(spec)):
        for lvl in range(5):
            close(fw[n][lvl], G['fn_%s_net%d_fw%d' % (tag, n, lvl)], rtol=2e-3, atol_rel=2e-4, msg="net %d fw %d" % (n, lvl))
            close(bw[n][lvl], G['fn_%s_net%d_bw%d' % (tag, n, lvl)], rtol=2e-3, atol_rel=2e-4, msg="net %d bw %d" % (n, lvl))


@pytest.mark.parametrize("tag,spec,seed,extra", [('c', 'c', 31, {}), ('s', 's', 32, {'pyramid_loss': False})])
def test_unsupervised_loss_against_reference_run(tag, spec, seed, extra):
    from unflow_b200.e2eflow.core.flownet import FlowNetVariables
    from unflow_b200.e2eflow.core.unsupervised import unsupervised_loss
    v = FlowNetVariables(spec, False, seed=0).load_tf_dict(_variables(spec, seed, 'ul_%s_vars' % tag)).cuda()
    params = dict(synth.KITTI_PARAMS, flownet=spec, **extra)
    with torch.no_grad():
        loss, ffw, fbw = unsupervised_loss((t('ul_%s_im1' % tag), t('ul_%s_im2' % tag)), params,
                                           synth.KITTI_NORMALIZATION, augment=False, return_flow=True, variables=v)
    close(loss, G['ul_%s_loss' % tag], rtol=1e-3, atol_rel=0.0)
    close(ffw, G['ul_%s_flow_fw' % tag], rtol=1e-3, atol_rel=5e-4)
    close(fbw, G['ul_%s_flow_bw' % tag], rtol=1e-3, atol_rel=5e-4)
