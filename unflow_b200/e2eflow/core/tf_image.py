"""The two tf.image resizers the hot path calls, with TF 1.x semantics, on torch tensors.

* resize_bilinear: legacy ``align_corners=False`` kernel WITHOUT half-pixel centres
  (src = dst * in/out; call sites reference flownet.py:48, unsupervised.py:103-104).
  For the integer upscales on the path the interpolation matrix is tiny, so it is built
  once per shape on the host and applied as two small gathers + lerps on the GPU.
* resize_area: used only by util.downsample when H or W is odd (core/util.py:26).
"""
import math

import torch

_cache = {}


def _bilinear_axis(n_in, n_out, device):
    key = ("bil", n_in, n_out, str(device))
    if key not in _cache:
        scale = torch.tensor(n_in / n_out, dtype=torch.float32)
        src = torch.arange(n_out, dtype=torch.float32) * scale
        lo = src.floor().long()
        hi = torch.clamp(src.ceil().long(), max=n_in - 1)
        lerp = src - lo.float()
        _cache[key] = (lo.to(device), hi.to(device), lerp.to(device))
    return _cache[key]


def resize_bilinear(x, size):
    """x: [B,H,W,C] -> [B,size[0],size[1],C] (differentiable)."""
    B, H, W, C = x.shape
    oh, ow = int(size[0]), int(size[1])
    y0, y1, ly = _bilinear_axis(H, oh, x.device)
    x0, x1, lx = _bilinear_axis(W, ow, x.device)
    ly = ly.view(1, oh, 1, 1)
    lx = lx.view(1, 1, ow, 1)
    top_rows = x.index_select(1, y0)
    bot_rows = x.index_select(1, y1)
    tl, tr = top_rows.index_select(2, x0), top_rows.index_select(2, x1)
    bl, br = bot_rows.index_select(2, x0), bot_rows.index_select(2, x1)
    top = tl + (tr - tl) * lx
    bot = bl + (br - bl) * lx
    return top + (bot - top) * ly


def _area_axis(n_in, n_out, device):
    key = ("area", n_in, n_out, str(device))
    if key not in _cache:
        scale = n_in / n_out
        m = torch.zeros(n_out, n_in, dtype=torch.float64)
        for o in range(n_out):
            start, end = o * scale, (o + 1) * scale
            j = int(math.floor(start))
            while j < math.ceil(end):
                lo, hi = max(start, j), min(end, j + 1)
                if hi > lo:
                    m[o, min(max(j, 0), n_in - 1)] += hi - lo
                j += 1
        _cache[key] = (m / scale).float().to(device)
    return _cache[key]


def resize_area(x, size):
    B, H, W, C = x.shape
    wy = _area_axis(H, int(size[0]), x.device)
    wx = _area_axis(W, int(size[1]), x.device)
    out = torch.einsum("oh,bhwc->bowc", wy, x)
    return torch.einsum("pw,bowc->bopc", wx, out)
