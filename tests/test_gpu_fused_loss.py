"""Fused per-level loss kernels (csrc/level_loss.cu) against the CPU oracle and against the
unfused GPU path; binary masks must be bit-exact."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import losses as olosses
import synth

TERMS = ['sym', 'occ', 'photo', 'grad', 'smooth_1st', 'smooth_2nd', 'fb', 'ternary']
W = dict(ternary=1.0, smooth_2nd=3.0, fb=0.2, occ=12.4, photo=0.7, smooth_1st=1.3, sym=0.9, grad=0.4)


def wsum(d, terms=TERMS):
    return sum(W[k] * d[k] for k in terms)


def close(got, want, rtol=2e-4, atol_rel=1e-5, msg=""):
    want = want.detach().cpu()
    got = got.detach().cpu()
    atol = atol_rel * max(float(want.abs().max()), 1e-12)
    np.testing.assert_allclose(got.numpy(), want.numpy(), rtol=rtol, atol=atol, err_msg=msg)


CASES = [('fb', True, 3, (2, 24, 40)), ('', False, 1, (1, 17, 35)), ('disocc', True, 2, (2, 16, 33)),
         ('fb', False, 2, (1, 9, 70)), ('fb', True, 1, (3, 6, 20)), ('fb', True, 3, (1, 48, 64))]


@pytest.mark.parametrize("mask_occlusion,use_border,dist,shape", CASES)
def test_fused_vs_oracle(mask_occlusion, use_border, dist, shape):
    from unflow_b200.e2eflow.core import losses as L
    from unflow_b200.e2eflow.core import fused_loss
    B, h, w = shape
    im1, im2, ffw, fbw = synth.level_inputs(B, h, w, seed=h * 100 + w)
    border = olosses.create_border_mask(im1, 0.1) if use_border else None
    fo, bo = ffw.clone().requires_grad_(True), fbw.clone().requires_grad_(True)
    fg, bg = ffw.cuda().requires_grad_(True), fbw.cuda().requires_grad_(True)
    want = olosses.compute_losses(im1, im2, fo, bo, border_mask=border, mask_occlusion=mask_occlusion,
                                  data_max_distance=dist)
    bc = border.cuda() if use_border else None
    got, mfw, mbw = fused_loss.compute_losses_fused(im1.cuda(), im2.cuda(), fg, bg, bc, mask_occlusion, dist,
                                                    [t for t in TERMS if t != 'grad'], return_masks=True)
    # binary masks: bit exact against the oracle (identical input flows -> no tie band needed,
    # the kernel evaluates the mask predicates op by op in the reference's order); 'disocc'
    # depends on the atomically-summed splat map, allow its tie band only
    aux = want['_aux']
    for name, m_got, m_want in (("mask_fw", mfw, aux['mask_fw']), ("mask_bw", mbw, aux['mask_bw'])):
        m_want = m_want.expand_as(m_got.cpu())
        if mask_occlusion == 'disocc':
            assert (m_got.cpu() != m_want).float().mean() < 2e-3, name
        else:
            assert torch.equal(m_got.cpu(), m_want), name
    fused_terms = [t for t in TERMS if t != 'grad']
    for k in fused_terms:
        close(got[k], want[k], msg=k)
    wsum(want, fused_terms).backward()
    wsum(got, fused_terms).backward()
    close(fg.grad, fo.grad, rtol=2e-3, atol_rel=2e-4, msg="dflow_fw")
    close(bg.grad, bo.grad, rtol=2e-3, atol_rel=2e-4, msg="dflow_bw")
    # the public entry point (auto -> fused) also returns the unfused 'grad' term
    allt = L.compute_losses(im1.cuda(), im2.cuda(), ffw.cuda(), fbw.cuda(), border_mask=bc,
                            mask_occlusion=mask_occlusion, data_max_distance=dist)
    assert set(allt) == set(TERMS)
    for k in TERMS:
        close(allt[k], want[k], msg="public " + k)


@pytest.mark.parametrize("terms", [['ternary'], ['fb'], ['smooth_2nd'], ['smooth_1st'], ['photo'],
                                   ['occ', 'ternary', 'smooth_2nd', 'fb']])
def test_fused_term_subsets_and_grads_vs_unfused(terms):
    """Each term alone: value and gradient must match the unfused GPU path (torch autograd
    through the image_warp kernel), and unrequested terms are exact zeros."""
    from unflow_b200.e2eflow.core import losses as L
    im1, im2, ffw, fbw = synth.level_inputs(2, 30, 52, seed=5)
    border = olosses.create_border_mask(im1, 0.1).cuda()
    im1, im2 = im1.cuda(), im2.cuda()
    f1, b1 = ffw.cuda().requires_grad_(True), fbw.cuda().requires_grad_(True)
    f2, b2 = ffw.cuda().requires_grad_(True), fbw.cuda().requires_grad_(True)
    fused = L.compute_losses(im1, im2, f1, b1, border_mask=border, mask_occlusion='fb',
                             data_max_distance=2, _terms=terms, _fused=True)
    unf = L.compute_losses(im1, im2, f2, b2, border_mask=border, mask_occlusion='fb',
                           data_max_distance=2, _terms=terms, _fused=False)
    for k in TERMS:
        if k in terms:
            close(fused[k], unf[k], msg=k)
        else:
            assert float(fused[k]) == 0.0
    wsum(fused, terms).backward()
    wsum(unf, terms).backward()
    if terms != ['occ']:
        close(f1.grad, f2.grad, rtol=2e-3, atol_rel=2e-4)
        close(b1.grad, b2.grad, rtol=2e-3, atol_rel=2e-4)


def test_fused_full_resolution_properties():
    """Roofline geometry (B=4, 384x1280): properties instead of the oracle."""
    from unflow_b200.e2eflow.core import fused_loss
    B, h, w = 4, 384, 1280
    im1, im2, ffw, fbw = synth.level_inputs(1, h, w, seed=9)
    im1, im2 = im1.cuda().repeat(B, 1, 1, 1), im2.cuda().repeat(B, 1, 1, 1)
    ffw, fbw = ffw.cuda().repeat(B, 1, 1, 1).requires_grad_(True), fbw.cuda().repeat(B, 1, 1, 1).requires_grad_(True)
    terms = ['occ', 'fb', 'ternary', 'smooth_2nd']
    a = fused_loss.compute_losses_fused(im1, im2, ffw, fbw, None, 'fb', 3, terms)
    # batch of identical items == single item (means are normalised by B)
    b = fused_loss.compute_losses_fused(im1[:1], im2[:1], ffw[:1].detach(), fbw[:1].detach(), None, 'fb', 3, terms)
    for k in terms:
        close(a[k], b[k], rtol=1e-5, msg=k)
    # deterministic reduction: bitwise repeatable
    c = fused_loss.compute_losses_fused(im1, im2, ffw, fbw, None, 'fb', 3, terms)
    for k in terms:
        assert float(a[k]) == float(c[k])
    # swapping the roles of the two frames swaps nothing in the symmetric sum
    d = fused_loss.compute_losses_fused(im2, im1, fbw, ffw, None, 'fb', 3, terms)
    for k in terms:
        close(a[k], d[k], rtol=1e-5, msg="swap " + k)
    # perfect photometric + flow consistency: identical frames, zero flow -> data terms ~ 0
    z = torch.zeros_like(ffw)
    e = fused_loss.compute_losses_fused(im1, im1, z, z, None, 'fb', 3, ['ternary', 'fb', 'occ'])
    assert float(e['ternary']) < 1e-2 and float(e['occ']) < 1e-2
    (a['ternary'] + a['fb']).backward()
    assert torch.isfinite(ffw.grad).all() and float(ffw.grad.abs().max()) > 0


def test_fused_rejects_bad_arguments():
    from unflow_b200.e2eflow.core import fused_loss
    im = torch.rand(1, 8, 8, 3, device="cuda")
    fl = torch.zeros(1, 8, 8, 2, device="cuda")
    with pytest.raises(ValueError):
        fused_loss.compute_losses_fused(im, im, fl, fl, None, 'fb', 4, ['ternary'])
    with pytest.raises(ValueError):
        fused_loss.compute_losses_fused(im, im, fl, fl, None, 'fb', 1, ['grad'])
    with pytest.raises(ValueError):
        fused_loss.compute_losses_fused(im, im, fl, torch.zeros(1, 8, 9, 2, device="cuda"), None, 'fb', 1, ['fb'])
