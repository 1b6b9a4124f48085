"""compute_losses on the fused CUDA level-loss kernels (csrc/level_loss.cu).

One forward launch (+ one backward launch) per pyramid level replaces the reference's graph of
several hundred TF ops (src/e2eflow/core/losses.py:16-87).  The 8 named terms come back as 0-d
tensors so the caller weights and sums them exactly as the reference does
(unsupervised.py:136-143); gradients w.r.t. the two flows are produced analytically by the
backward kernel (masks and occlusion maps are casts in the reference and carry none).
"""
import torch

from ... import _native
from ..._native import check
from ..ops import forward_warp, _prep, _stream, kernel_timer

TERM_ORDER = ['sym', 'occ', 'photo', 'grad', 'smooth_1st', 'smooth_2nd', 'fb', 'ternary']
_OCCL = {'': 0, 'fb': 1, 'disocc': 2}
_ALL_FUSED = [t for t in TERM_ORDER if t != 'grad']


def _bits(terms):
    m = 0
    for t in terms:
        m |= 1 << TERM_ORDER.index(t)
    return m


def available(im1, flow_fw, mask_occlusion, data_max_distance, terms=None):
    """Can the fused kernels serve this call?  (CUDA float32 inputs, census patch <= 7x7, no
    Sobel 'grad' term, no gradient requested for the images.)"""
    if not (torch.is_tensor(im1) and im1.is_cuda and im1.dtype == torch.float32):
        return False
    if im1.requires_grad and torch.is_grad_enabled():
        return False
    if not 1 <= int(data_max_distance) <= 3 or mask_occlusion not in _OCCL:
        return False
    if im1.shape[3] != 3 or flow_fw.shape[3] != 2:
        return False
    if terms is not None and 'grad' in terms:
        return False
    return True


class _LevelLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, flow_fw, flow_bw, im1, im2, border, fwarp_fw, fwarp_bw, occl, dist, bits,
                want_masks):
        B, h, w, _ = im1.shape
        dev = im1.device
        lib = _native.lib()
        losses = torch.empty(8, device=dev, dtype=torch.float32)
        need_saved = bool(bits >> 7 & 1)
        saved = torch.empty(4 * B * h * w, device=dev, dtype=torch.float32) if need_saved else None
        masks = torch.empty(2, B, h, w, 1, device=dev, dtype=torch.float32) if want_masks else None
        ws = torch.empty(int(lib.unflow_level_loss_workspace_bytes(B, h, w)), device=dev, dtype=torch.uint8)
        ptr = lambda t: t.data_ptr() if t is not None else None
        with torch.cuda.device(dev), kernel_timer.span("level_loss_fwd_%dx%d" % (h, w)):
            check(lib.unflow_level_loss_fwd(im1.data_ptr(), im2.data_ptr(), flow_fw.data_ptr(),
                                            flow_bw.data_ptr(), ptr(border), ptr(fwarp_fw), ptr(fwarp_bw),
                                            losses.data_ptr(), ptr(saved), ptr(masks), ws.data_ptr(),
                                            B, h, w, occl, dist, bits, _stream()), "level_loss")
        ctx.save_for_backward(flow_fw, flow_bw, im1, im2, border, fwarp_fw, fwarp_bw, saved)
        ctx.cfg = (occl, dist, bits)
        if want_masks:
            ctx.mark_non_differentiable(masks)
            return losses, masks
        return losses

    @staticmethod
    def backward(ctx, grad_losses, *unused):
        flow_fw, flow_bw, im1, im2, border, fwarp_fw, fwarp_bw, saved = ctx.saved_tensors
        occl, dist, bits = ctx.cfg
        B, h, w, _ = im1.shape
        grad_losses = grad_losses.contiguous().float()
        dfw = torch.empty_like(flow_fw)
        dbw = torch.empty_like(flow_bw)
        ptr = lambda t: t.data_ptr() if t is not None else None
        with torch.cuda.device(im1.device), kernel_timer.span("level_loss_bwd_%dx%d" % (h, w)):
            check(_native.lib().unflow_level_loss_bwd(
                grad_losses.data_ptr(), im1.data_ptr(), im2.data_ptr(), flow_fw.data_ptr(),
                flow_bw.data_ptr(), ptr(border), ptr(fwarp_fw), ptr(fwarp_bw), ptr(saved),
                dfw.data_ptr(), dbw.data_ptr(), B, h, w, occl, dist, bits, _stream()), "level_loss_grad")
        return (dfw, dbw) + (None,) * 9


VECTOR_KEY = '_vector'


def compute_losses_fused(im1, im2, flow_fw, flow_bw, border_mask=None, mask_occlusion='',
                         data_max_distance=1, terms=None, return_masks=False):
    terms = list(_ALL_FUSED if terms is None else terms)
    if 'grad' in terms:
        raise ValueError("the 'grad' term is not fused; use the unfused path")
    im1 = _prep(im1.detach(), "im1")
    im2 = _prep(im2.detach(), "im2")
    flow_fw = _prep(flow_fw, "flow_fw")
    flow_bw = _prep(flow_bw, "flow_bw")
    B, h, w, _ = im1.shape
    if tuple(flow_fw.shape) != (B, h, w, 2) or tuple(flow_bw.shape) != (B, h, w, 2) or im2.shape != im1.shape:
        raise ValueError("compute_losses: shape mismatch")
    border = None
    if border_mask is not None:
        border = _prep(border_mask.detach().expand(B, h, w, 1), "border_mask")
    occl = _OCCL[mask_occlusion]
    fwarp_fw = fwarp_bw = None
    if occl == 2 or 'sym' in terms:
        with torch.no_grad():
            fwarp_fw = forward_warp(flow_fw.detach())
            fwarp_bw = forward_warp(flow_bw.detach())
    out = _LevelLoss.apply(flow_fw, flow_bw, im1, im2, border, fwarp_fw, fwarp_bw, occl,
                           int(data_max_distance), _bits(terms), bool(return_masks))
    vec, masks = (out if return_masks else (out, None))
    zero = torch.zeros((), device=im1.device, dtype=torch.float32)
    parts = vec.unbind(0)
    losses = {name: (parts[i] if name in terms else zero) for i, name in enumerate(TERM_ORDER)}
    losses[VECTOR_KEY] = vec        # the terms as ONE tensor (TERM_ORDER): lets the caller weight them with one dot product
    if return_masks:
        return losses, masks[0], masks[1]
    return losses
