"""CPU restatement of the deterministic cores of the augmentation (TEST INFRASTRUCTURE ONLY).

Follows /root/reference/src/e2eflow/core/spatial_transformer.py:57-175 (``transformer`` with its
``_meshgrid`` / ``_interpolate``) and /root/reference/src/e2eflow/core/augment.py:30-48,91-106.
The random draws are inputs here.  Pinned against the reference files themselves, executed
unmodified under the TensorFlow-API stand-in of tests/golden/ with the random draws recorded
(transformer, random_affine incl. the flip branch, random_photometric:
tests/test_oracle_vs_reference_run.py); the reference has no test of its own for the augmentation."""
import math

import torch


def transformer(U, theta, out_size):
    B, H, W, C = U.shape
    oh, ow = out_size
    theta = theta.reshape(B, 2, 3).float()
    x_t = torch.ones(oh, 1) @ torch.linspace(-1.0, 1.0, ow).view(1, ow)
    y_t = torch.linspace(-1.0, 1.0, oh).view(oh, 1) @ torch.ones(1, ow)
    grid = torch.cat([x_t.reshape(1, -1), y_t.reshape(1, -1), torch.ones(1, oh * ow)], 0)
    T_g = theta @ grid.unsqueeze(0).expand(B, 3, oh * ow)
    x = T_g[:, 0].reshape(-1)
    y = T_g[:, 1].reshape(-1)
    # _interpolate
    x = (x + 1.0) * float(W) / 2.0
    y = (y + 1.0) * float(H) / 2.0
    x0 = torch.floor(x).long(); x1 = x0 + 1
    y0 = torch.floor(y).long(); y1 = y0 + 1
    x0 = x0.clamp(0, W - 1); x1 = x1.clamp(0, W - 1)
    y0 = y0.clamp(0, H - 1); y1 = y1.clamp(0, H - 1)
    base = (torch.arange(B) * (W * H)).view(B, 1).expand(B, oh * ow).reshape(-1)
    im_flat = U.reshape(-1, C).float()
    Ia = im_flat[base + y0 * W + x0]
    Ib = im_flat[base + y1 * W + x0]
    Ic = im_flat[base + y0 * W + x1]
    Id = im_flat[base + y1 * W + x1]
    x0f, x1f, y0f, y1f = x0.float(), x1.float(), y0.float(), y1.float()
    wa = ((x1f - x) * (y1f - y)).unsqueeze(1)
    wb = ((x1f - x) * (y - y0f)).unsqueeze(1)
    wc = ((x - x0f) * (y1f - y)).unsqueeze(1)
    wd = ((x - x0f) * (y - y0f)).unsqueeze(1)
    out = ((wa * Ia + wb * Ib) + wc * Ic) + wd * Id
    return out.reshape(B, oh, ow, C)


def affine_matrices(tx, ty, rot_deg, scale, flip=None):
    rad = rot_deg * math.pi / 180.0
    B = tx.shape[0]
    t = torch.zeros(B, 2, 3)
    for b in range(B):
        t1 = torch.tensor([[math.cos(rad[b]), -math.sin(rad[b]), float(tx[b])],
                           [math.sin(rad[b]), math.cos(rad[b]), float(ty[b])]])
        sx = float(scale[b]) * (float(flip[b]) if flip is not None else 1.0)
        t2 = torch.tensor([[sx, 0.0, 0.0], [0.0, float(scale[b]), 0.0], [0.0, 0.0, 1.0]])
        t[b] = t1 @ t2
    return t


def photometric(ims, contrast, gamma, colour, noise, brightness):
    out = []
    gamma_inv = 1.0 / gamma
    for im in ims:
        im_re = im.permute(1, 2, 0, 3)                       # [h, w, B, C]
        im_re = (im_re * (contrast + 1.0) + brightness) * colour
        im_re = torch.clamp(im_re, 0.0, 1.0)
        im_re = torch.pow(im_re, gamma_inv)
        im_re = im_re + noise
        out.append(im_re.permute(2, 0, 1, 3))
    return out
