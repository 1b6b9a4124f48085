"""CPU restatement of the reference's pure-TF backward warp (TEST INFRASTRUCTURE ONLY).

Follows /root/reference/src/e2eflow/core/image_warp.py:4-76: integer tap
positions come from ``pos + floor(flow)`` and are CLAMPED to the image,
bilinear weights come from ``flow - floor(flow)``; the four taps are combined
as wa*Ia + wb*Ib + wc*Ic + wd*Id (tf.add_n order).  Differentiable w.r.t. both
image and flow through torch autograd, as TF autodiff is in the reference.
"""
import torch


def image_warp(im, flow):
    B, H, W, C = im.shape
    im_flat = im.reshape(-1, C)
    flow_flat = flow.reshape(-1, 2)

    fl = torch.floor(flow_flat)
    fl_i = fl.to(torch.int64)
    wts = flow_flat - fl

    pos_x = torch.arange(W).repeat(H * B)
    pos_y = torch.arange(H).view(H, 1).expand(H, W).reshape(-1).repeat(B)

    xw, yw = wts[:, 0], wts[:, 1]
    wa = ((1 - xw) * (1 - yw)).unsqueeze(1)
    wb = ((1 - xw) * yw).unsqueeze(1)
    wc = (xw * (1 - yw)).unsqueeze(1)
    wd = (xw * yw).unsqueeze(1)

    x0 = pos_x + fl_i[:, 0]
    y0 = pos_y + fl_i[:, 1]
    x1 = (x0 + 1).clamp(0, W - 1)
    y1 = (y0 + 1).clamp(0, H - 1)
    x0 = x0.clamp(0, W - 1)
    y0 = y0.clamp(0, H - 1)

    base = (torch.arange(B) * (H * W)).view(B, 1).expand(B, H * W).reshape(-1)
    row0 = base + y0 * W
    row1 = base + y1 * W
    Ia = im_flat[row0 + x0]
    Ib = im_flat[row1 + x0]
    Ic = im_flat[row0 + x1]
    Id = im_flat[row1 + x1]

    out = ((wa * Ia + wb * Ib) + wc * Ic) + wd * Id
    return out.reshape(B, H, W, C)
