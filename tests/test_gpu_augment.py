"""Row N4: GPU augmentation -- the deterministic cores against the oracle, the random wrappers for
their invariants, and unsupervised_loss(augment=True) end to end."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import augment as oaug
import synth


def test_transformer_vs_oracle():
    from unflow_b200.e2eflow.core import augment as A
    g = torch.Generator().manual_seed(0)
    U = torch.rand(3, 20, 28, 3, generator=g)
    tx, ty = torch.tensor([0.0, 0.1, -0.2]), torch.tensor([0.0, -0.05, 0.15])
    rot, scale = torch.tensor([0.0, 7.0, -12.0]), torch.tensor([1.0, 0.9, 1.1])
    flip = torch.tensor([1.0, -1.0, 1.0])
    theta = oaug.affine_matrices(tx, ty, rot, scale, flip)
    got_theta = A.affine_matrices(tx.cuda(), ty.cuda(), rot.cuda(), scale.cuda(), flip.cuda())
    np.testing.assert_allclose(got_theta.cpu().numpy(), theta.numpy(), rtol=1e-6, atol=1e-7)
    want = oaug.transformer(U, theta, (20, 28))
    got = A.transformer(U.cuda(), theta.cuda(), (20, 28))
    np.testing.assert_allclose(got.cpu().numpy(), want.numpy(), rtol=1e-5, atol=1e-5)
    # identity theta with the reference's (x+1)*W/2 sampling grid: interior pixels are blends of
    # neighbours, the output stays inside the input range, samples beyond W-1 are exactly zero
    ident = oaug.affine_matrices(torch.zeros(1), torch.zeros(1), torch.zeros(1), torch.ones(1))
    out = A.transformer(U[:1].cuda(), ident.cuda(), (20, 28)).cpu()
    assert float(out.max()) <= float(U[:1].max()) + 1e-5
    assert float(out[0, :, -1].abs().max()) < 1e-6     # weights from clamped indices cancel (to rounding)


def test_photometric_vs_oracle_and_random_wrappers():
    from unflow_b200.e2eflow.core import augment as A
    g = torch.Generator().manual_seed(1)
    im = torch.rand(2, 9, 11, 3, generator=g)
    contrast = torch.tensor([[0.2], [-0.3]]); gamma = torch.tensor([[0.8], [1.4]])
    colour = torch.tensor([[0.9, 1.0, 1.1], [1.05, 0.95, 1.0]])
    noise = torch.tensor([[0.01], [-0.02]]); bright = torch.tensor([[0.02], [-0.01]])
    want = oaug.photometric([im], contrast, gamma, colour, noise, bright)[0]
    got = A.photometric([im.cuda()], contrast.cuda(), gamma.cuda(), colour.cuda(), noise.cuda(), bright.cuda())[0]
    np.testing.assert_allclose(got.cpu().numpy(), want.numpy(), rtol=1e-5, atol=1e-6)
    # wrappers: same draw applied to every tensor of the list; reproducible under seed()
    A.seed(3)
    a1, a2 = A.random_affine([im.cuda(), im.cuda()], min_scale=0.9, max_scale=1.1, horizontal_flipping=True)
    assert torch.equal(a1, a2)
    A.seed(3)
    b1, _ = A.random_affine([im.cuda(), im.cuda()], min_scale=0.9, max_scale=1.1, horizontal_flipping=True)
    assert torch.equal(a1, b1)
    p1, p2 = A.random_photometric([im.cuda(), im.cuda()], noise_stddev=0.04, min_contrast=-0.3, max_contrast=0.3,
                                  brightness_stddev=0.02, min_colour=0.9, max_colour=1.1, min_gamma=0.7, max_gamma=1.5)
    assert torch.equal(p1, p2) and p1.shape == im.shape
    c1, c2 = A.random_crop([im, im], [2, 5, 6, 3], seed=4)
    assert c1.shape == (2, 5, 6, 3) and torch.equal(c1, c2)


def test_unsupervised_loss_with_augmentation_runs_and_backprops():
    from unflow_b200.e2eflow.core import augment as A
    from unflow_b200.e2eflow.core.flownet import FlowNetVariables
    from unflow_b200.e2eflow.core.unsupervised import unsupervised_loss
    v = FlowNetVariables('C', seed=1).cuda()
    im1, im2, _ = synth.image_pair(2, 128, 256, seed=2)
    A.seed(0)
    loss = unsupervised_loss((im1.cuda(), im2.cuda()), dict(synth.KITTI_PARAMS), synth.KITTI_NORMALIZATION,
                             augment=True, variables=v)
    loss.backward()
    w, _ = v.weights('flownet_c/conv3_1')
    assert np.isfinite(float(loss)) and torch.isfinite(w.grad).all() and float(w.grad.abs().max()) > 0
    A.seed(0)
    again = unsupervised_loss((im1.cuda(), im2.cuda()), dict(synth.KITTI_PARAMS), synth.KITTI_NORMALIZATION,
                              augment=True, variables=v)
    assert abs(float(again) - float(loss)) / abs(float(loss)) < 1e-4   # same draws -> same loss


def test_trainer_augment_changes_the_step_and_survives_graph_capture():
    """the reference trains with augment=True (train.py:160,169).  Trainer(augment=True)
    must (a) give a loss that depends on the augmentation seed, (b) reproduce with the same seed,
    (c) draw fresh parameters on every replay of the captured CUDA graph."""
    from unflow_b200.e2eflow.core import augment as A
    from unflow_b200.e2eflow.core.train import Trainer
    params = dict(synth.KITTI_PARAMS, learning_rate=0.0)       # lr 0: the variables stay put
    im1, im2, _ = synth.image_pair(1, 128, 256, seed=6)
    im1, im2 = im1.cuda(), im2.cuda()
    try:
        tr = Trainer(params, synth.KITTI_NORMALIZATION, "cuda", seed=2, augment=True)
        plain = Trainer(params, synth.KITTI_NORMALIZATION, "cuda", seed=2)
        A.seed(1); l1 = float(tr.step(im1, im2))
        A.seed(2); l2 = float(tr.step(im1, im2))
        A.seed(1); l3 = float(tr.step(im1, im2))
        l0 = float(plain.step(im1, im2))
        assert abs(l1 - l3) <= 1e-4 * abs(l1)
        assert abs(l1 - l2) > 1e-3 * abs(l1) and abs(l1 - l0) > 1e-3 * abs(l1)
        tr.capture(im1, im2)
        reps = [float(tr.step(im1, im2)) for _ in range(3)]
        assert all(np.isfinite(reps)) and len({round(r, 3) for r in reps}) == 3, reps
    finally:
        A.set_device_rng(False)
