"""GPU parity of the model and the loss assembly against the CPU oracle."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import flownet as oflownet
from oracle import losses as olosses
from oracle import unsupervised as ounsup
import synth

TERMS = ['sym', 'occ', 'photo', 'grad', 'smooth_1st', 'smooth_2nd', 'fb', 'ternary']


def close(got, want, rtol=1e-4, atol_rel=1e-5, msg=""):
    want = want.detach().cpu()
    got = got.detach().cpu()
    atol = atol_rel * max(float(want.abs().max()), 1e-12)
    np.testing.assert_allclose(got.numpy(), want.numpy(), rtol=rtol, atol=atol, err_msg=msg)


@pytest.mark.parametrize("mask_occlusion,use_border,dist", [('fb', True, 3), ('', False, 1),
                                                            ('disocc', True, 2), ('fb', False, 2)])
def test_compute_losses_unfused_vs_oracle(mask_occlusion, use_border, dist):
    from unflow_b200.e2eflow.core import losses as L
    im1, im2, ffw, fbw = synth.level_inputs(2, 24, 40)
    border = olosses.create_border_mask(im1, 0.1) if use_border else None
    fo, bo = ffw.clone().requires_grad_(True), fbw.clone().requires_grad_(True)
    fg, bg = ffw.cuda().requires_grad_(True), fbw.cuda().requires_grad_(True)
    want = olosses.compute_losses(im1, im2, fo, bo, border_mask=border, mask_occlusion=mask_occlusion,
                                  data_max_distance=dist)
    got = L.compute_losses(im1.cuda(), im2.cuda(), fg, bg,
                           border_mask=border.cuda() if use_border else None,
                           mask_occlusion=mask_occlusion, data_max_distance=dist, _fused=False)
    assert set(got) == set(TERMS)
    for k in TERMS:
        close(got[k], want[k], rtol=2e-4, msg=k)
    wsum = lambda d: (d['ternary'] + 3.0 * d['smooth_2nd'] + 0.2 * d['fb'] + 12.4 * d['occ'] +
                      d['photo'] + d['grad'] + d['smooth_1st'] + d['sym'])
    wsum(want).backward()
    wsum(got).backward()
    close(fg.grad, fo.grad, rtol=1e-3, atol_rel=1e-4, msg="dflow_fw")
    close(bg.grad, bo.grad, rtol=1e-3, atol_rel=1e-4, msg="dflow_bw")


@pytest.mark.parametrize("spec,hw", [("C", (64, 128)), ("CS", (64, 128)), ("c", (64, 64))])
def test_flownet_vs_oracle(spec, hw):
    from unflow_b200.e2eflow.core.flownet import FlowNetVariables, flownet
    tfv = oflownet.init_variables(spec, False, seed=5)
    v = FlowNetVariables(spec, False, seed=0).load_tf_dict(tfv).cuda()
    im1, im2, _ = synth.image_pair(1, hw[0], hw[1], seed=3)
    im1, im2 = im1 / 255.0 - 0.4, im2 / 255.0 - 0.4
    want_fw, want_bw = oflownet.flownet(tfv, im1, im2, spec, backward_flow=True)
    with torch.no_grad():
        got_fw, got_bw = flownet(im1.cuda(), im2.cuda(), spec, backward_flow=True, variables=v)
    assert len(got_fw) == len(spec)
    for net in range(len(spec)):
        for w, g in zip(want_fw[net] + want_bw[net], got_fw[net] + got_bw[net]):
            close(g, w, rtol=1e-3, atol_rel=1e-4, msg="%s net %d" % (spec, net))
    # forward-only call returns the same forward flows (flownet.py:79-81)
    with torch.no_grad():
        only_fw = flownet(im1.cuda(), im2.cuda(), spec, variables=v)
    # (not bit-identical: the bidirectional pass runs 2B samples through cuDNN, which may pick other
    # algorithms than for B samples)
    for w, g in zip(got_fw[-1], only_fw[-1]):
        close(g, w, rtol=1e-3, atol_rel=1e-4)


@pytest.mark.parametrize("spec,hw", [("C", (128, 256)), ("S", (128, 192))])
def test_unsupervised_loss_vs_oracle(spec, hw):
    """configs[0]/[2] graph at reduced size: loss value, final flows (1e-4 relative) and the
    gradient of the loss w.r.t. a set of variables."""
    from unflow_b200.e2eflow.core.flownet import FlowNetVariables
    from unflow_b200.e2eflow.core.unsupervised import unsupervised_loss
    params = dict(synth.KITTI_PARAMS, flownet=spec)
    tfv = oflownet.init_variables(spec, False, seed=11)
    for k in tfv:  # leaves for the oracle's autograd
        tfv[k] = tfv[k].clone().requires_grad_(True)
    v = FlowNetVariables(spec, False, seed=0).load_tf_dict({k: t.detach() for k, t in tfv.items()}).cuda()
    im1, im2, _ = synth.image_pair(1, hw[0], hw[1], seed=21)
    want_loss, want_fw, want_bw = ounsup.unsupervised_loss(tfv, (im1, im2), params, synth.KITTI_NORMALIZATION,
                                                           augment=False, return_flow=True)
    got_loss, got_fw, got_bw = unsupervised_loss((im1.cuda(), im2.cuda()), params, synth.KITTI_NORMALIZATION,
                                                 augment=False, return_flow=True, variables=v)
    close(got_loss, want_loss, rtol=2e-4)
    # north_star tolerance: flow within 1e-4 relative (to the flow magnitude) of the reference
    close(got_fw, want_fw, rtol=1e-4, atol_rel=1e-4)
    close(got_bw, want_bw, rtol=1e-4, atol_rel=1e-4)
    want_loss.backward()
    got_loss.backward()
    # Gradients: the loss contains hard masks (fb occlusion `>`), so a 1e-7 difference in a flow
    # value can flip a mask pixel and move every gradient by O(1/pixels): compare in
    # the L2 norm per variable instead of element by element.
    for scope in v.kinds:
        w, b = v.weights(scope)
        for got, want, name in ((w.grad.cpu(), tfv[scope + '/weights'].grad.permute(3, 2, 0, 1), '/weights'),
                                (b.grad.cpu(), tfv[scope + '/biases'].grad, '/biases')):
            err = float((got - want).norm() / want.norm().clamp_min(1e-20))
            assert err < 5e-3, "%s%s: relative L2 gradient error %.3e" % (scope, name, err)


def test_adam_kernel_matches_tf_rule():
    """csrc/adam.cu against the TF AdamOptimizer update written out in float64."""
    from unflow_b200 import _native
    n = 4096 + 4
    g = torch.Generator().manual_seed(0)
    p = torch.randn(n, generator=g)
    grads = [torch.randn(n, generator=g) for _ in range(3)]
    pd, m, v = p.double(), torch.zeros(n, dtype=torch.float64), torch.zeros(n, dtype=torch.float64)
    pc, gc = p.cuda(), torch.zeros(n, device="cuda")
    mc, vc = torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    lr, b1, b2, eps = 1e-3, 0.9, 0.999, 1e-8
    for t, gr in enumerate(grads, 1):
        gc.copy_(gr * 2.0)                                   # "sum over 2 ranks", scale 0.5
        _native.check(_native.lib().unflow_adam_step(pc.data_ptr(), gc.data_ptr(), mc.data_ptr(), vc.data_ptr(),
                                                     n, lr, b1, b2, eps, t, 0.5, 1,
                                                     torch.cuda.current_stream().cuda_stream), "adam")
        assert float(gc.abs().max()) == 0.0                  # gradient cleared in the same pass
        gd = gr.double()
        m = b1 * m + (1 - b1) * gd
        v = b2 * v + (1 - b2) * gd * gd
        lr_t = lr * (1 - b2 ** t) ** 0.5 / (1 - b1 ** t)
        pd = pd - lr_t * m / (v.sqrt() + eps)
    np.testing.assert_allclose(pc.cpu().numpy(), pd.float().numpy(), rtol=1e-5, atol=1e-6)


def test_adam_kernel_l2_mask():
    """The regularisation gradient folded into the update: masked elements see grad * scale + l2 * p."""
    from unflow_b200 import _native
    n = 1024
    g = torch.Generator().manual_seed(1)
    p, gr = torch.randn(n, generator=g), torch.randn(n, generator=g)
    bits = (torch.rand(n, generator=g) < 0.5)
    q = bits.view(-1, 4).to(torch.uint8)
    mask = (q[:, 0] | (q[:, 1] << 1) | (q[:, 2] << 2) | (q[:, 3] << 3)).contiguous().cuda()
    l2, lr = 0.37, 1e-2
    st = torch.cuda.current_stream().cuda_stream
    out = []
    for fused in (True, False):
        pc, mc, vc = p.cuda(), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
        gc = (gr if fused else gr + l2 * p * bits).cuda()
        _native.check(_native.lib().unflow_adam_step_l2(pc.data_ptr(), gc.data_ptr(), mc.data_ptr(), vc.data_ptr(), n,
                                                        lr, 0.9, 0.999, 1e-8, 1, 1.0, 1,
                                                        mask.data_ptr() if fused else None, l2, st), "adam")
        out.append((pc.cpu(), mc.cpu(), vc.cpu()))
    for a, b in zip(*out):
        np.testing.assert_allclose(a.numpy(), b.numpy(), rtol=2e-6, atol=1e-7)


def test_trainer_step_applies_adam_to_the_flat_buffer():
    """One Trainer.step == Adam's first update on the gradient of the loss (the total loss itself
    need not decrease: the occlusion penalty is piecewise constant and grows as flows develop)."""
    from unflow_b200 import _native
    from unflow_b200.e2eflow.core.train import Trainer
    lr = 1e-3
    params = dict(synth.KITTI_PARAMS, learning_rate=lr)
    tr = Trainer(params, synth.KITTI_NORMALIZATION, "cuda", seed=3)
    im1, im2, _ = synth.image_pair(2, 128, 256, seed=8)
    im1, im2 = im1.cuda(), im2.cuda()
    tr.loss(im1, im2).backward()
    g = tr.flat_grad.clone() + tr.l2_gradient()           # the L2 term is added inside the Adam kernel
    assert tr.l2_mask is not None and float(tr.l2_gradient().abs().max()) > 0
    tr.flat_grad.zero_()
    p0 = tr.flat_param.clone()
    _native.reset_launch_count()
    loss = tr.step(im1, im2)
    assert _native.launch_count() >= 10 and np.isfinite(float(loss))
    assert float(tr.flat_grad.abs().max()) == 0.0          # cleared by the fused update
    want = -lr * g / (g.abs() + 1e-8 / (1 - 0.999) ** 0.5)   # first Adam step in closed form
    got = tr.flat_param - p0
    nz = g.abs() > 1e-3   # tiny gradients are summation noise (atomics) and differ run to run
    np.testing.assert_allclose(got[nz].cpu().numpy(), want[nz].cpu().numpy(), rtol=2e-3, atol=2e-6)
    losses = [float(tr.step(im1, im2)) for _ in range(3)]
    assert all(np.isfinite(losses)) and tr.iteration == 4


def test_cuda_graph_step_matches_eager_step():
    """Trainer.capture(): replaying the captured step must do what the eager step does."""
    from unflow_b200.e2eflow.core.train import Trainer
    params = dict(synth.KITTI_PARAMS, learning_rate=1e-4)
    im1, im2, _ = synth.image_pair(1, 128, 256, seed=4)
    im1, im2 = im1.cuda(), im2.cuda()
    jm1, jm2, _ = synth.image_pair(1, 128, 256, seed=5)
    jm1, jm2 = jm1.cuda(), jm2.cuda()
    a = Trainer(params, synth.KITTI_NORMALIZATION, "cuda", seed=9)
    b = Trainer(params, synth.KITTI_NORMALIZATION, "cuda", seed=9)
    p0 = b.flat_param.clone()
    b.capture(im1, im2)
    assert torch.equal(b.flat_param, p0) and b.iteration == 0      # capture leaves the state untouched
    assert float(b.adam_m.abs().max()) == 0.0
    la = [float(a.step(im1, im2)), float(a.step(jm1, jm2)), float(a.step(im1, im2))]
    lb = [float(b.step(im1, im2)), float(b.step(jm1, jm2)), float(b.step(im1, im2))]
    np.testing.assert_allclose(lb, la, rtol=2e-4)
    # Adam normalises every gradient to ~lr*sign(g): elements whose gradient is summation noise may
    # step the other way (|diff| up to 2*lr per step); everything else must agree closely
    diff = (a.flat_param - b.flat_param).abs()
    assert float((diff > 2e-5).float().mean()) < 0.02, float((diff > 2e-5).float().mean())
    assert b.graph_replays == 3 and b._graph_launches > 20
