"""conv / deconv layers of the FlowNet stacks with a selectable arithmetic mode.

'fp32'    plain cuDNN float32 (what the reference's slim.conv2d computes, flownet.py:174-233);
          no tensor cores.
'3xtf32'  the same contraction on the tensor cores at fp32-level accuracy: every operand is split
          x = hi + lo (hi = TF32-rounded, lo = exact residual, csrc/split.cu) and
          hi*hi' + hi*lo' + lo*hi' is evaluated by ONE cuDNN TF32 convolution whose contraction
          dimension carries the three products side by side (X' = [hi,hi,lo], W' = [hi',lo',hi']):
          fp32 accumulation inside the MMA, one output write, no extra adds.  Forward, dgrad and
          wgrad all use the same trick (wgrad concatenates along the batch).  Measured against the
          fp32 oracle it meets the same 1e-4 flow tolerance as the plain fp32 path.
'tf32'    single-pass TF32 (reduced precision; never used for parity or the headline number).
"""
import torch
import torch.nn.functional as F
from torch.nn import grad as nngrad

from ... import _native
from ..._native import check

_MODE = 'fp32'


def set_mode(mode):
    global _MODE
    if mode not in ('fp32', '3xtf32', 'tf32'):
        raise ValueError("conv precision must be 'fp32', '3xtf32' or 'tf32'")
    _MODE = mode
    # the flag is process-global because autograd runs the backward convolutions later
    torch.backends.cudnn.allow_tf32 = mode != 'fp32'
    torch.backends.cuda.matmul.allow_tf32 = False
    return mode


def get_mode():
    return _MODE


CL = torch.channels_last


def _cl(x):
    """Dense NHWC memory (channels_last) -- the layout cuDNN's tensor-core kernels compute in; with
    NCHW tensors cuDNN wraps every conv in nchwToNhwc / nhwcToNchw transform kernels (24 % of the
    step when measured)."""
    return x.contiguous(memory_format=CL)


def _split3(x, items, inner, order):
    """x: dense tensor viewed as [items][inner] -> 3 slabs per item (csrc/split.cu), flat output."""
    out = torch.empty((3 * items * inner,), device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        check(_native.lib().unflow_split3_tf32(x.data_ptr(), out.data_ptr(), items, inner, order,
                                               torch.cuda.current_stream().cuda_stream), "split3_tf32")
    return out


def _cat_channels(x, order):
    """[N,C,H,W] -> [N,3C,H,W] (channels_last): per pixel (hi,hi,lo) or (hi,lo,hi) along C."""
    x = _cl(x)
    N, C, H, W = x.shape
    return _split3(x, N * H * W, C, order).view(N, H, W, 3 * C).permute(0, 3, 1, 2)


def _cat_batch(x, order):
    """[N,C,H,W] -> [3N,C,H,W] (channels_last) = (hi;hi;lo) or (hi;lo;hi) along the batch."""
    x = _cl(x)
    N, C, H, W = x.shape
    return _split3(x, 1, x.numel(), order).view(3 * N, H, W, C).permute(0, 3, 1, 2)


class _Conv3x(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, stride, padding):
        ctx.save_for_backward(x, w)
        ctx.cfg = (stride, padding, b is not None)
        xs = _cat_channels(x, 0)                       # [N,3Ci,H,W]   hi,hi,lo
        ws = _cat_channels(w, 1)                       # [Co,3Ci,k,k]  hi,lo,hi
        return F.conv2d(xs, ws, b, stride=stride, padding=padding)

    @staticmethod
    def backward(ctx, g):
        x, w = ctx.saved_tensors
        stride, padding, has_b = ctx.cfg
        g = _cl(g)
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            gs = _cat_channels(g, 0)                   # [N,3Co,..]   hi,hi,lo
            wt = _cat_batch(w, 1)                      # [3Co,Ci,k,k] hi;lo;hi
            gx = nngrad.conv2d_input(x.shape, wt, gs, stride=stride, padding=padding)
        if ctx.needs_input_grad[1]:
            xb = _cat_batch(x, 0)                      # [3N,Ci,..]   hi;hi;lo
            gb3 = _cat_batch(g, 1)                     # [3N,Co,..]   hi;lo;hi
            gw = nngrad.conv2d_weight(xb, w.shape, gb3, stride=stride, padding=padding)
        if has_b and ctx.needs_input_grad[2]:
            gb = g.sum((0, 2, 3))
        return gx, gw, gb, None, None


class _Deconv3x(torch.autograd.Function):
    """conv_transpose2d(x, w[in,out,4,4], stride=2, padding=1)"""

    @staticmethod
    def forward(ctx, x, w, b):
        ctx.save_for_backward(x, w)
        ctx.has_b = b is not None
        xs = _cat_channels(x, 0)                       # [N,3Ci,h,w]
        ws = _cat_batch(w, 1)                          # [3Ci,Co,4,4]
        return F.conv_transpose2d(xs, ws, b, stride=2, padding=1)

    @staticmethod
    def backward(ctx, g):
        x, w = ctx.saved_tensors
        g = _cl(g)
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            gs = _cat_channels(g, 0)                   # [N,3Co,2h,2w]
            wc = _cat_channels(w, 1)                   # [Ci,3Co,4,4]
            gx = F.conv2d(gs, wc, None, stride=2, padding=1)
        if ctx.needs_input_grad[1]:
            # d/dw of conv_transpose == weight gradient of the conv whose input is g and output x
            gb3 = _cat_batch(g, 0)
            xb = _cat_batch(x, 1)
            gw = nngrad.conv2d_weight(gb3, w.shape, xb, stride=2, padding=1)
        if ctx.has_b and ctx.needs_input_grad[2]:
            gb = g.sum((0, 2, 3))
        return gx, gw, gb


def channels_last_active(x):
    """The tensor-core path computes in NHWC (cuDNN's TF32 kernels are NHWC; NCHW costs a layout
    transform around every conv); the exact-fp32 path keeps the reference's NCHW (cuDNN's fp32
    NHWC kernels measured 18 % slower than NCHW on this network)."""
    return _MODE != 'fp32' and x.is_cuda


def network_input(x_nhwc):
    """NHWC image batch -> the NCHW-shaped tensor the conv stack consumes.  In the tensor-core
    mode this is a free view (NHWC memory == channels_last); in fp32 mode an NCHW copy."""
    x = x_nhwc.permute(0, 3, 1, 2)
    return x if channels_last_active(x_nhwc) else x.contiguous()


def conv2d(x, w, b, stride, pads):
    """pads = (top, bottom, left, right) TF-SAME padding."""
    pt, pb, pl, pr = pads
    if pt == pb and pl == pr:
        padding = (pt, pl)
    else:
        x = F.pad(x, (pl, pr, pt, pb))
        padding = (0, 0)
    if _MODE == '3xtf32' and x.is_cuda:
        return _Conv3x.apply(x, w, b, stride, padding)
    return F.conv2d(x, w, b, stride=stride, padding=padding)


def conv_transpose2d(x, w, b):
    if _MODE == '3xtf32' and x.is_cuda:
        return _Deconv3x.apply(x, w, b)
    return F.conv_transpose2d(x, w, b, stride=2, padding=1)
