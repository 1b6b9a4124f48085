"""Loss terms of the unsupervised objective -- same names, arguments, defaults and return
structures as the reference module (/root/reference/src/e2eflow/core/losses.py).

``compute_losses`` (the per-pyramid-level hot path, reference :16-87) runs on the fused CUDA
level-loss kernels (csrc/level_loss.cu) when they can serve the request; the individual
functions below are the stand-alone versions of each term (API parity), written with shifted
slices instead of the reference's one-hot / finite-difference ``tf.nn.conv2d`` filters -- the
values are the same, without materialising P*P-channel patch tensors through a convolution.

All tensors are float32 CUDA tensors in NHWC layout.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from ..ops import backward_warp, forward_warp  # noqa: F401  (reference imports both, :5)
from .image_warp import image_warp

DISOCC_THRESH = 0.8

_GRAY = (0.2989, 0.5870, 0.1140)  # tf.image.rgb_to_grayscale


def length_sq(x):
    return torch.sum(torch.square(x), 3, keepdim=True)


def compute_losses(im1, im2, flow_fw, flow_bw,
                   border_mask=None,
                   mask_occlusion='',
                   data_max_distance=1,
                   _terms=None, _fused=None):
    """Reference losses.py:16-87.  Returns the dict of the 8 named terms.

    ``_terms`` (private): iterable of term names actually needed.  TF graph mode only executes
    the terms whose weights are set (unsupervised.py:136-141); in eager mode the caller passes
    that set so unused terms are skipped (they are returned as exact 0-d zeros).
    ``_fused`` (private): force (True) / forbid (False) the fused CUDA path; None = auto.
    """
    from . import fused_loss
    terms = list(fused_loss.TERM_ORDER) if _terms is None else list(_terms)
    use_fused = fused_loss.available(im1, flow_fw, mask_occlusion, data_max_distance) \
        if _fused is None else _fused
    if not use_fused:
        return _compute_losses_unfused(im1, im2, flow_fw, flow_bw, border_mask, mask_occlusion,
                                       data_max_distance, _terms)
    fused_terms = [t for t in terms if t != 'grad']
    if 'grad' not in terms:
        return fused_loss.compute_losses_fused(im1, im2, flow_fw, flow_bw, border_mask,
                                               mask_occlusion, data_max_distance, fused_terms)
    # the Sobel 'grad' term ("NOT TESTED" in the reference config) is not fused: evaluate it
    # with the stand-alone functions on the masks the fused kernel produced
    losses, mask_fw, mask_bw = fused_loss.compute_losses_fused(
        im1, im2, flow_fw, flow_bw, border_mask, mask_occlusion, data_max_distance, fused_terms,
        return_masks=True)
    losses['grad'] = (gradient_loss(im1, image_warp(im2, flow_fw), mask_fw) +
                      gradient_loss(im2, image_warp(im1, flow_bw), mask_bw))
    losses.pop(fused_loss.VECTOR_KEY, None)       # the kernel's vector does not hold this term
    return losses


class _Direction:
    """Everything the eight terms need for one flow direction (a -> b): the second image and the
    opposite flow sampled along the flow, the forward-backward residual and its occlusion test
    (reference :21-22, :38-50)."""

    def __init__(self, im_a, im_b, flow_ab, flow_ba):
        self.image, self.flow = im_a, flow_ab
        self.other_warped = image_warp(im_b, flow_ab)
        self.image_diff = im_a - self.other_warped
        back = image_warp(flow_ba, flow_ab)
        self.flow_diff = flow_ab + back
        limit = 0.01 * (length_sq(flow_ab) + length_sq(back)) + 0.5
        self.fb_occluded = (length_sq(self.flow_diff) > limit).float()

    def disocclusion(self):
        return (forward_warp(self.flow) < DISOCC_THRESH).float()


def _compute_losses_unfused(im1, im2, flow_fw, flow_bw, border_mask, mask_occlusion,
                            data_max_distance, _terms=None):
    """The level loss term by term on stand-alone ops (used when the fused kernels cannot serve
    the request, and as their cross-check in the tests)."""
    from .fused_loss import TERM_ORDER
    wanted = set(TERM_ORDER if _terms is None else _terms)
    fw = _Direction(im1, im2, flow_fw, flow_bw)
    bw = _Direction(im2, im1, flow_bw, flow_fw)
    dirs = (fw, bw)

    if 'sym' in wanted or mask_occlusion == 'disocc':
        for d in dirs:
            d.disocc = d.disocclusion()

    # validity: the border mask if given, else "the flow stays inside the image"; then remove
    # what the chosen occlusion test marks (note the swap: a pixel of frame 1 is disoccluded when
    # the BACKWARD flow splats nothing onto it)
    for d, opposite in ((fw, bw), (bw, fw)):
        d.mask = create_outgoing_mask(d.flow) if border_mask is None else border_mask
        if mask_occlusion == 'fb':
            d.mask = d.mask * (1 - d.fb_occluded)
        elif mask_occlusion == 'disocc':
            d.mask = d.mask * (1 - opposite.disocc)

    def both(term):
        return term(fw, bw) + term(bw, fw)

    table = {
        'sym': lambda d, o: charbonnier_loss((1 - d.mask) - o.disocc),
        'occ': lambda d, o: charbonnier_loss(1 - d.mask),
        'photo': lambda d, o: photometric_loss(d.image_diff, d.mask),
        'grad': lambda d, o: gradient_loss(d.image, d.other_warped, d.mask),
        'smooth_1st': lambda d, o: smoothness_loss(d.flow),
        'smooth_2nd': lambda d, o: second_order_loss(d.flow),
        'fb': lambda d, o: charbonnier_loss(d.flow_diff, d.mask),
        'ternary': lambda d, o: ternary_loss(d.image, d.other_warped, d.mask,
                                             max_distance=data_max_distance),
    }
    zero = torch.zeros((), device=im1.device, dtype=torch.float32)
    return {name: (both(term) if name in wanted else zero) for name, term in table.items()}


def _rgb_to_gray(image):
    return (image[..., 0:1] * _GRAY[0] + image[..., 1:2] * _GRAY[1]) + image[..., 2:3] * _GRAY[2]


def _patches(x, radius):
    """[B,H,W,1] -> [B,H,W,P*P]: zero padded neighbourhood, channel k = (dy+r)*P + (dx+r)
    (what the reference's one-hot conv2d produces, losses.py:101-104)."""
    B, H, W, _ = x.shape
    P = 2 * radius + 1
    xp = F.pad(x[..., 0], (radius, radius, radius, radius))
    cols = [xp[:, dy:dy + H, dx:dx + W] for dy in range(P) for dx in range(P)]
    return torch.stack(cols, 3)


def ternary_loss(im1, im2_warped, mask, max_distance=1):
    def _ternary_transform(image):
        intensities = _rgb_to_gray(image) * 255
        patches = _patches(intensities, max_distance)
        transf = patches - intensities
        return transf / torch.sqrt(0.81 + torch.square(transf))

    def _hamming_distance(t1, t2):
        dist = torch.square(t1 - t2)
        dist_norm = dist / (0.1 + dist)
        return torch.sum(dist_norm, 3, keepdim=True)

    t1 = _ternary_transform(im1)
    t2 = _ternary_transform(im2_warped)
    dist = _hamming_distance(t1, t2)

    transform_mask = create_mask(mask, [[max_distance, max_distance],
                                        [max_distance, max_distance]])
    return charbonnier_loss(dist, mask * transform_mask)


def occlusion(flow_fw, flow_bw):
    mag_sq = length_sq(flow_fw) + length_sq(flow_bw)
    flow_bw_warped = image_warp(flow_bw, flow_fw)
    flow_fw_warped = image_warp(flow_fw, flow_bw)
    flow_diff_fw = flow_fw + flow_bw_warped
    flow_diff_bw = flow_bw + flow_fw_warped
    occ_thresh = 0.01 * mag_sq + 0.5
    occ_fw = (length_sq(flow_diff_fw) > occ_thresh).float()
    occ_bw = (length_sq(flow_diff_bw) > occ_thresh).float()
    return occ_fw, occ_bw


def _shift(x, dy, dx):
    """y[b,i,j,c] = x[b,i+dy,j+dx,c], zero outside (tf.nn.conv2d 'SAME' semantics)."""
    B, H, W, C = x.shape
    xp = F.pad(x, (0, 0, 1, 1, 1, 1))
    return xp[:, 1 + dy:1 + dy + H, 1 + dx:1 + dx + W, :]


def _sobel(x):
    """Per-channel Sobel x / y responses: [B,H,W,C] -> (gx, gy)."""
    gx = ((_shift(x, -1, 1) - _shift(x, -1, -1)) + 2 * (_shift(x, 0, 1) - _shift(x, 0, -1)) +
          (_shift(x, 1, 1) - _shift(x, 1, -1)))
    gy = ((_shift(x, 1, -1) - _shift(x, -1, -1)) + 2 * (_shift(x, 1, 0) - _shift(x, -1, 0)) +
          (_shift(x, 1, 1) - _shift(x, -1, 1)))
    return gx, gy


def divergence(flow):
    gx, _ = _sobel(flow[..., 0:1])
    _, gy = _sobel(flow[..., 1:2])
    return gx + gy


def norm(x, sigma):
    """exp(-x^2 / (2 sigma^2)): 1 at x = 0, falling off on the scale of sigma."""
    return torch.exp(-0.5 * torch.square(x / sigma))


def diffusion_loss(flow, im, occ):
    """8-neighbour flow differences, down-weighted where intensity or flow already differ and
    gated by the occlusion-label difference (reference losses.py:173-195; no shipped
    configuration uses it)."""
    def neighbor_diff(x):
        outs = []
        for c in range(x.shape[3]):
            xc = x[..., c:c + 1]
            for n in [0, 1, 2, 3, 5, 6, 7, 8]:
                outs.append(xc - _shift(xc, n // 3 - 1, n % 3 - 1))
        return torch.cat(outs, 3)

    occ_diff = neighbor_diff(occ)
    fd = neighbor_diff(flow)
    flow_diff_u, flow_diff_v = fd[..., :8], fd[..., 8:]
    flow_diff = torch.sqrt(torch.square(flow_diff_u) + torch.square(flow_diff_v))
    intensity_diff = torch.abs(neighbor_diff(_rgb_to_gray(im)))
    diff = norm(intensity_diff, 7.5 / 255) * norm(flow_diff, 0.5) * occ_diff * flow_diff
    return charbonnier_loss(diff)


def photometric_loss(im_diff, mask):
    return charbonnier_loss(im_diff, mask, beta=255)


def conv2d(x, weights):
    """tf.nn.conv2d(x, weights, [1,1,1,1], 'SAME') for NHWC x and HWIO weights."""
    k = weights.shape[0]
    w = weights.permute(3, 2, 0, 1).contiguous()
    y = F.conv2d(x.permute(0, 3, 1, 2), w, None, stride=1, padding=k // 2)
    return y.permute(0, 2, 3, 1)


def _smoothness_deltas(flow):
    mask_x = create_mask(flow, [[0, 0], [0, 1]])
    mask_y = create_mask(flow, [[0, 1], [0, 0]])
    mask = torch.cat([mask_x, mask_y], 3)

    def deltas(f):  # x: f(y,x) - f(y,x+1); y: f(y,x) - f(y+1,x)
        return torch.cat([f - _shift(f, 0, 1), f - _shift(f, 1, 0)], 3)

    return deltas(flow[..., 0:1]), deltas(flow[..., 1:2]), mask


def _gradient_delta(im1, im2_warped):
    def grads(im):
        gx, gy = _sobel(im)
        return torch.stack([gx, gy], 4).reshape(im.shape[0], im.shape[1], im.shape[2], -1)

    return grads(im1) - grads(im2_warped)


def gradient_loss(im1, im2_warped, mask):
    mask_x = create_mask(im1, [[0, 0], [1, 1]])
    mask_y = create_mask(im1, [[1, 1], [0, 0]])
    gradient_mask = torch.cat([mask_x, mask_y], 3).repeat(1, 1, 1, 3)
    diff = _gradient_delta(im1, im2_warped)
    return charbonnier_loss(diff, mask * gradient_mask)


def smoothness_loss(flow):
    delta_u, delta_v, mask = _smoothness_deltas(flow)
    loss_u = charbonnier_loss(delta_u, mask)
    loss_v = charbonnier_loss(delta_v, mask)
    return loss_u + loss_v


def _second_order_deltas(flow):
    mask_x = create_mask(flow, [[0, 0], [1, 1]])
    mask_y = create_mask(flow, [[1, 1], [0, 0]])
    mask_diag = create_mask(flow, [[1, 1], [1, 1]])
    mask = torch.cat([mask_x, mask_y, mask_diag, mask_diag], 3)

    def deltas(f):
        c2 = 2 * f
        return torch.cat([(_shift(f, 0, -1) + _shift(f, 0, 1)) - c2,
                          (_shift(f, -1, 0) + _shift(f, 1, 0)) - c2,
                          (_shift(f, -1, -1) + _shift(f, 1, 1)) - c2,
                          (_shift(f, -1, 1) + _shift(f, 1, -1)) - c2], 3)

    return deltas(flow[..., 0:1]), deltas(flow[..., 1:2]), mask


def second_order_loss(flow):
    delta_u, delta_v, mask = _second_order_deltas(flow)
    loss_u = charbonnier_loss(delta_u, mask)
    loss_v = charbonnier_loss(delta_v, mask)
    return loss_u + loss_v


def charbonnier_loss(x, mask=None, truncate=None, alpha=0.45, beta=1.0, epsilon=0.001):
    """Generalised Charbonnier penalty ((beta*x)^2 + epsilon^2)^alpha, averaged over ALL elements
    of ``x`` ([B,H,W,C]); elements where ``mask`` ([B,H,W,1] or [B,H,W,C], 0/1) is zero add
    nothing but still count in the denominator (reference :296-323)."""
    normalization = float(x.numel())
    error = torch.pow(torch.square(x * beta) + epsilon ** 2, alpha)
    if mask is not None:
        error = mask * error
    if truncate is not None:
        error = torch.clamp(error, max=truncate)
    return torch.sum(error) / normalization


def create_mask(tensor, paddings):
    B, H, W = tensor.shape[0], tensor.shape[1], tensor.shape[2]
    top, bottom = paddings[0]
    left, right = paddings[1]
    mask = torch.zeros(1, H, W, 1, device=tensor.device, dtype=torch.float32)
    if H - top - bottom > 0 and W - left - right > 0:
        mask[:, top:H - bottom, left:W - right, :] = 1.0
    return mask.expand(B, H, W, 1)


def create_border_mask(tensor, border_ratio=0.1):
    H, W = tensor.shape[1], tensor.shape[2]
    min_dim = np.float32(min(H, W))
    sz = int(math.ceil(np.float32(min_dim * np.float32(border_ratio))))
    return create_mask(tensor, [[sz, sz], [sz, sz]])


def create_outgoing_mask(flow):
    """[B,H,W,1] mask: 1 where pixel + flow lands inside [0,W-1] x [0,H-1], else 0."""
    B, H, W, _ = flow.shape
    grid_x = torch.arange(W, device=flow.device, dtype=torch.float32).view(1, 1, W)
    grid_y = torch.arange(H, device=flow.device, dtype=torch.float32).view(1, H, 1)
    pos_x = grid_x + flow[..., 0]
    pos_y = grid_y + flow[..., 1]
    inside = ((pos_x <= float(W - 1)) & (pos_x >= 0.0) &
              (pos_y <= float(H - 1)) & (pos_y >= 0.0))
    return inside.float().unsqueeze(3)
