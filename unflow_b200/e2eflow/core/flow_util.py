"""Flow metrics and visualisation -- reference src/e2eflow/core/flow_util.py (the step after the
hot path: evaluation).  Same function names and arguments; torch
tensors (any device) in NHWC layout."""
import math

import numpy as np
import torch


def atan2(y, x):
    """flow_util.py:5-18 (the reference builds atan2 from atan with explicit quadrant cases;
    x == y == 0 yields NaN there as well)."""
    angle = torch.where(x > 0.0, torch.atan(y / x), torch.zeros_like(x))
    angle = torch.where((x < 0.0) & (y >= 0.0), torch.atan(y / x) + math.pi, angle)
    angle = torch.where((x < 0.0) & (y < 0.0), torch.atan(y / x) - math.pi, angle)
    angle = torch.where((x == 0.0) & (y > 0.0), math.pi * torch.ones_like(x), angle)
    angle = torch.where((x == 0.0) & (y < 0.0), -math.pi * torch.ones_like(x), angle)
    angle = torch.where((x == 0.0) & (y == 0.0), float('nan') * torch.ones_like(x), angle)
    return angle


def _hsv_to_rgb(hsv):
    """tf.image.hsv_to_rgb."""
    h, s, v = hsv[..., 0], hsv[..., 1], hsv[..., 2]
    c = s * v
    m = v - c
    dh = h * 6.0
    fmodu = dh - 2.0 * torch.floor(dh / 2.0)
    x = c * (1.0 - torch.abs(fmodu - 1.0))
    hcat = torch.floor(dh).long()
    z = torch.zeros_like(c)
    rr = [c, x, z, z, x, c]
    gg = [x, c, c, x, z, z]
    bb = [z, z, x, c, c, x]
    r, g, b = z.clone(), z.clone(), z.clone()
    for k in range(6):
        sel = hcat == k
        r = torch.where(sel, rr[k], r)
        g = torch.where(sel, gg[k], g)
        b = torch.where(sel, bb[k], b)
    return torch.stack([r + m, g + m, b + m], -1)


def flow_to_color(flow, mask=None, max_flow=None):
    """Converts flow to 3-channel color image (flow_util.py:21-45).

    Args:
        flow: tensor of shape [num_batch, height, width, 2].
        mask: flow validity mask of shape [num_batch, height, width, 1].
    """
    n = 8
    B, H, W, _ = flow.shape
    mask = torch.ones(B, H, W, 1, device=flow.device) if mask is None else mask
    flow_u, flow_v = flow[..., 0], flow[..., 1]
    if max_flow is not None:
        max_flow = max(max_flow, 1)
    else:
        max_flow = torch.max(torch.abs(flow * mask))
    mag = torch.sqrt(torch.sum(torch.square(flow), 3))
    angle = atan2(flow_v, flow_u)
    im_h = torch.remainder(angle / (2 * math.pi) + 1.0, 1.0)
    im_s = torch.clamp(mag * n / max_flow, 0, 1)
    im_v = torch.clamp(n - im_s, 0, 1)
    im = _hsv_to_rgb(torch.stack([im_h, im_s, im_v], 3))
    return im * mask


_COLORMAP = np.asarray([
    [0, 0.0625, 49, 54, 149], [0.0625, 0.125, 69, 117, 180], [0.125, 0.25, 116, 173, 209],
    [0.25, 0.5, 171, 217, 233], [0.5, 1, 224, 243, 248], [1, 2, 254, 224, 144],
    [2, 4, 253, 174, 97], [4, 8, 244, 109, 67], [8, 16, 215, 48, 39],
    [16, 1000000000.0, 165, 0, 38]], dtype=np.float32)


def flow_error_image(flow_1, flow_2, mask_occ, mask_noc=None, log_colors=True):
    """Visualize the error between two flows as 3-channel color image (flow_util.py:48-95,
    adapted by the reference from the KITTI devkit)."""
    mask_noc = torch.ones_like(mask_occ) if mask_noc is None else mask_noc
    diff = torch.sqrt(torch.sum((flow_1 - flow_2) ** 2, 3, keepdim=True))
    if log_colors:
        cm = _COLORMAP.copy()
        cm[:, 2:5] = cm[:, 2:5] / 255
        mag = torch.sqrt(torch.sum(torch.square(flow_2), 3, keepdim=True))
        error = torch.minimum(diff / 3, 20 * diff / mag)
        im = torch.zeros(flow_1.shape[0], flow_1.shape[1], flow_1.shape[2], 3, device=flow_1.device)
        for i in range(cm.shape[0]):
            cond = (error >= float(cm[i, 0])) & (error < float(cm[i, 1]))
            col = torch.tensor(cm[i, 2:5], device=flow_1.device).view(1, 1, 1, 3)
            im = torch.where(cond.expand(-1, -1, -1, 3), col.expand_as(im), im)
        im = torch.where(mask_noc.bool().expand(-1, -1, -1, 3), im, im * 0.5)
        im = im * mask_occ
    else:
        error = (torch.clamp(diff, max=5) / 5) * mask_occ
        im = torch.cat([error, error * mask_noc, error * mask_noc], 3)
    return im


def euclidean(t):
    return torch.sqrt(torch.sum(t ** 2, 3, keepdim=True))


def flow_error_avg(flow_1, flow_2, mask):
    """Evaluates the average endpoint error between flow batches (flow_util.py:98-103)."""
    diff = euclidean(flow_1 - flow_2) * mask
    return torch.sum(diff) / torch.sum(mask)


def outlier_ratio(gt_flow, flow, mask, threshold=3.0, relative=0.05):
    diff = euclidean(gt_flow - flow) * mask
    if relative is not None:
        threshold = torch.clamp(euclidean(gt_flow) * relative, min=threshold)
    outliers = (diff >= threshold).float()
    return torch.sum(outliers) / torch.sum(mask)


def outlier_pct(gt_flow, flow, mask, threshold=3.0, relative=0.05):
    return outlier_ratio(gt_flow, flow, mask, threshold, relative) * 100
