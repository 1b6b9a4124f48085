"""Host-side loss assembly on the CPU: the product's Python control flow (pyramid levels, weights,
masks, term table) with its CUDA entry points swapped for the oracle's CPU ops, compared with the
oracle's own assembly.  The GPU parity tests (test_gpu_model.py) cover the real kernels; these keep
the host logic checked where no GPU exists."""
import numpy as np
import pytest
import torch

from oracle import flownet as oflownet
from oracle import image_warp as oimage_warp
from oracle import losses as olosses
from oracle import ops as oops
from oracle import unsupervised as ounsup
from oracle import util as outil
import synth

TERMS = ['sym', 'occ', 'photo', 'grad', 'smooth_1st', 'smooth_2nd', 'fb', 'ternary']


@pytest.mark.parametrize("mask_occlusion,use_border,dist", [('fb', True, 3), ('', False, 1),
                                                            ('disocc', True, 2), ('disocc', False, 1)])
def test_unfused_term_table_matches_oracle(monkeypatch, mask_occlusion, use_border, dist):
    from unflow_b200.e2eflow.core import losses as L
    monkeypatch.setattr(L, 'image_warp', oimage_warp.image_warp)
    monkeypatch.setattr(L, 'forward_warp', oops.forward_warp)
    im1, im2, ffw, fbw = synth.level_inputs(2, 20, 28)
    border = olosses.create_border_mask(im1, 0.1) if use_border else None
    want = olosses.compute_losses(im1, im2, ffw, fbw, border_mask=border,
                                  mask_occlusion=mask_occlusion, data_max_distance=dist)
    got = L.compute_losses(im1, im2, ffw, fbw, border_mask=border, mask_occlusion=mask_occlusion,
                           data_max_distance=dist, _fused=False)
    assert list(got) == TERMS
    for k in TERMS:
        np.testing.assert_allclose(float(got[k]), float(want[k]), rtol=2e-5, atol=1e-7, err_msg=k)
    # a restricted term set returns exact zeros for the rest
    some = L.compute_losses(im1, im2, ffw, fbw, border_mask=border, mask_occlusion=mask_occlusion,
                            data_max_distance=dist, _fused=False, _terms=['ternary', 'occ'])
    for k in TERMS:
        if k in ('ternary', 'occ'):
            assert float(some[k]) == float(got[k])
        else:
            assert float(some[k]) == 0.0


@pytest.mark.parametrize("full_res,pyramid", [(False, True), (False, False), (True, True)])
def test_unsupervised_loss_host_flow_matches_oracle(monkeypatch, full_res, pyramid):
    from unflow_b200.e2eflow.core import unsupervised as U
    from unflow_b200.e2eflow.core.flownet import FlowNetVariables
    spec = 'S'
    params = dict(synth.KITTI_PARAMS, flownet=spec, full_res=full_res, pyramid_loss=pyramid)
    tfv = oflownet.init_variables(spec, full_res, seed=4)
    v = FlowNetVariables(spec, full_res, seed=0).load_tf_dict(tfv)

    def cpu_flownet(im1, im2, flownet_spec='S', full_resolution=False, train_all=False,
                    backward_flow=False, variables=None):
        assert variables is v
        return oflownet.flownet(variables.to_tf_dict(), im1, im2, flownet_spec=flownet_spec,
                                full_resolution=full_resolution, train_all=train_all,
                                backward_flow=backward_flow)

    def cpu_compute_losses(im1, im2, flow_fw, flow_bw, border_mask=None, mask_occlusion='',
                           data_max_distance=1, _terms=None):
        return olosses.compute_losses(im1, im2, flow_fw, flow_bw, border_mask=border_mask,
                                      mask_occlusion=mask_occlusion,
                                      data_max_distance=data_max_distance)

    monkeypatch.setattr(U, 'flownet', cpu_flownet)
    monkeypatch.setattr(U, 'compute_losses', cpu_compute_losses)
    monkeypatch.setattr(U, 'downsample', outil.downsample)

    im1, im2, _ = synth.image_pair(1, 128, 128, seed=6)
    want, wfw, wbw, wterms = ounsup.unsupervised_loss(tfv, (im1, im2), params, synth.KITTI_NORMALIZATION,
                                                      return_flow=True, return_terms=True)
    got, gfw, gbw = U.unsupervised_loss((im1, im2), params, synth.KITTI_NORMALIZATION, augment=False,
                                        return_flow=True, variables=v)
    np.testing.assert_allclose(float(got.detach()), float(want), rtol=1e-5)  # L2 term: fp64 vs fp32 summation
    np.testing.assert_allclose(gfw.numpy(), wfw.numpy(), rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(gbw.numpy(), wbw.numpy(), rtol=1e-6, atol=1e-6)
    assert float(U.tracked['loss/combined']) == float(got.detach())
    for t in U.LOSSES:
        np.testing.assert_allclose(float(U.tracked['loss/' + t]), float(wterms[t]), rtol=1e-6, err_msg=t)


def test_unsupervised_loss_vector_weighting_matches_oracle(monkeypatch):
    """The fused level-loss kernel hands its terms over as ONE vector (fused_loss.VECTOR_KEY) and
    unsupervised_loss weights them with one dot product per level; here the vector is supplied by a CPU stand-in,
    so the dot-product branch (weights, per-term tracking, pyramid weights) is held to the oracle and to the
    term-by-term branch without a GPU."""
    from unflow_b200.e2eflow.core import fused_loss
    from unflow_b200.e2eflow.core import unsupervised as U
    from unflow_b200.e2eflow.core.flownet import FlowNetVariables
    spec = 'S'
    params = dict(synth.KITTI_PARAMS, flownet=spec, full_res=False, pyramid_loss=True)
    tfv = oflownet.init_variables(spec, False, seed=4)
    v = FlowNetVariables(spec, False, seed=0).load_tf_dict(tfv)

    def cpu_flownet(im1, im2, flownet_spec='S', full_resolution=False, train_all=False,
                    backward_flow=False, variables=None):
        return oflownet.flownet(variables.to_tf_dict(), im1, im2, flownet_spec=flownet_spec,
                                full_resolution=full_resolution, train_all=train_all, backward_flow=backward_flow)

    calls = {'n': 0}

    def vector_compute_losses(im1, im2, flow_fw, flow_bw, border_mask=None, mask_occlusion='',
                              data_max_distance=1, _terms=None):
        d = olosses.compute_losses(im1, im2, flow_fw, flow_bw, border_mask=border_mask,
                                   mask_occlusion=mask_occlusion, data_max_distance=data_max_distance)
        wanted = set(_terms)
        vec = torch.stack([d[t] if t in wanted else torch.zeros(()) for t in fused_loss.TERM_ORDER])
        out = {t: vec[i] for i, t in enumerate(fused_loss.TERM_ORDER)}
        out[fused_loss.VECTOR_KEY] = vec
        calls['n'] += 1
        return out

    monkeypatch.setattr(U, 'flownet', cpu_flownet)
    monkeypatch.setattr(U, 'downsample', outil.downsample)
    im1, im2, _ = synth.image_pair(1, 128, 128, seed=6)
    res = {}
    for name, fn in (('scalar', lambda *a, **k: {t: x for t, x in vector_compute_losses(*a, **k).items()
                                                 if t != fused_loss.VECTOR_KEY}),
                     ('vector', vector_compute_losses)):
        monkeypatch.setattr(U, 'compute_losses', fn)
        loss = U.unsupervised_loss((im1, im2), params, synth.KITTI_NORMALIZATION, augment=False, variables=v)
        res[name] = (float(loss.detach()), {t: float(U.tracked['loss/' + t]) for t in U.LOSSES})
    want, _, _, wterms = ounsup.unsupervised_loss(tfv, (im1, im2), params, synth.KITTI_NORMALIZATION,
                                                  return_flow=True, return_terms=True)
    assert calls['n'] == 10                                  # 5 pyramid levels, two passes
    np.testing.assert_allclose(res['vector'][0], float(want), rtol=1e-5)
    np.testing.assert_allclose(res['vector'][0], res['scalar'][0], rtol=1e-6)
    for t in U.LOSSES:
        np.testing.assert_allclose(res['vector'][1][t], float(wterms[t]), rtol=1e-6, err_msg=t)
        np.testing.assert_allclose(res['vector'][1][t], res['scalar'][1][t], rtol=1e-6, atol=1e-12, err_msg=t)
