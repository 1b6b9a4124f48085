"""The four custom ops, same Python surface as the reference op module
(/root/reference/src/e2eflow/ops.py:69-107), backed by libunflow.so through the C ABI.

  correlation(first, second, **kwargs)   NCHW x NCHW -> NCHW cost volume (output 0 only)
  backward_warp(images, flows)           NHWC bilinear gather, zero outside
  forward_warp(flows)                    NHWC Gaussian splat count map [B,H,W,1]
  downsample(images, scale)              NHWC box mean

Gradient table (reference ops.py:80-107): BackwardWarp -> [None, dflow]; ForwardWarp ->
[dflow]; Correlation -> [g0, g1]; Downsample -> not differentiable.

Inputs must be float32 CUDA tensors; anything else raises -- there is no CPU fallback.
``python -m unflow_b200.e2eflow.ops`` rebuilds the library (the reference's compile entry,
ops.py:51-52).
"""
import torch

from .. import _native
from .._native import BORDER_CLAMP, BORDER_ZERO, check

# Register ops for compilation here (reference ops.py:11)
OP_NAMES = ['backward_warp', 'downsample', 'correlation', 'forward_warp']


def compile(op=None):
    """Reference ops.py:21-48 compiled one .so per op; here all ops live in libunflow.so."""
    from .. import build
    return build.build(force=True)


def _prep(t, name, ndim=4):
    if not torch.is_tensor(t):
        raise TypeError("%s must be a torch tensor" % name)
    if t.device.type != "cuda":
        raise RuntimeError("%s must be a CUDA tensor: the UnFlow ops have GPU kernels only "
                           "(as in the reference, which registers DEVICE_GPU kernels only)" % name)
    if t.dtype != torch.float32:
        raise TypeError("%s must be float32" % name)
    if t.dim() != ndim:
        raise ValueError("%s must have rank %d" % (name, ndim))
    return t.contiguous()


def _stream():
    return torch.cuda.current_stream().cuda_stream


class _KernelTimer:
    """Optional CUDA-event timing of individual launches on the launching stream (bench.py uses it
    to measure the per-launch duration of the hand-written kernels inside the timed region)."""

    def __init__(self):
        self.enabled = False
        self.records = []
        self.bytes = {}

    def enable(self):
        self.enabled = True
        self.records = []

    def collect(self):
        """-> {name: [ms, ...]}; call after a device synchronize.  ``self.bytes[name]`` holds the
        algorithmic bytes summed over the same launches (for spans that declared them)."""
        self.enabled = False
        out, self.bytes = {}, {}
        for name, s, e, nbytes in self.records:
            out.setdefault(name, []).append(s.elapsed_time(e))
            self.bytes[name] = self.bytes.get(name, 0) + nbytes
        self.records = []
        return out

    class _Span:
        def __init__(self, timer, name, nbytes=0):
            self.t, self.name, self.nbytes = timer, name, nbytes

        def __enter__(self):
            if self.t.enabled:
                self.s = torch.cuda.Event(enable_timing=True)
                self.e = torch.cuda.Event(enable_timing=True)
                self.s.record()
            return self

        def __exit__(self, *exc):
            if self.t.enabled:
                self.e.record()
                self.t.records.append((self.name, self.s, self.e, self.nbytes))
            return False

    def span(self, name, nbytes=0):
        return self._Span(self, name, nbytes)


kernel_timer = _KernelTimer()


_CORR_DEFAULTS = dict(kernel_size=1, max_displacement=20, pad=20, stride_1=1, stride_2=2)


def _corr_attrs(kwargs):
    a = dict(_CORR_DEFAULTS)
    for k, v in kwargs.items():
        if k not in a:
            raise TypeError("correlation() got an unexpected attribute %r" % k)
        a[k] = int(v)
    return (a['kernel_size'], a['max_displacement'], a['pad'], a['stride_1'], a['stride_2'])


class _Correlation(torch.autograd.Function):
    @staticmethod
    def forward(ctx, in0, in1, attrs):
        in0 = _prep(in0, "input_0")
        in1 = _prep(in1, "input_1")
        if in0.shape != in1.shape:
            raise ValueError("Input shapes have to be the same")
        B, C, H, W = in0.shape
        import ctypes
        oc, oh, ow = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        lib = _native.lib()
        check(lib.unflow_correlation_out_shape(H, W, *attrs, ctypes.byref(oc), ctypes.byref(oh),
                                               ctypes.byref(ow)), "correlation")
        out = torch.empty(B, oc.value, oh.value, ow.value, device=in0.device, dtype=torch.float32)
        with torch.cuda.device(in0.device), kernel_timer.span("correlation_fwd"):
            check(lib.unflow_correlation_fwd(in0.data_ptr(), in1.data_ptr(), out.data_ptr(),
                                             B, C, H, W, *attrs, _stream()), "correlation")
        ctx.save_for_backward(in0, in1)
        ctx.attrs = attrs
        return out

    @staticmethod
    def backward(ctx, gout):
        in0, in1 = ctx.saved_tensors
        B, C, H, W = in0.shape
        gout = gout.contiguous()
        g0 = torch.empty_like(in0)
        g1 = torch.empty_like(in1)
        with torch.cuda.device(in0.device), kernel_timer.span("correlation_bwd"):
            check(_native.lib().unflow_correlation_bwd(gout.data_ptr(), in0.data_ptr(), in1.data_ptr(),
                                                       g0.data_ptr(), g1.data_ptr(), B, C, H, W,
                                                       *ctx.attrs, _stream()), "correlation_grad")
        return g0, g1, None


def correlation(first, second, **kwargs):
    return _Correlation.apply(first, second, _corr_attrs(kwargs))


class _CorrelationBidir(torch.autograd.Function):
    """(correlation(a, b), correlation(b, a)) in one forward launch and one pair of backward launches:
    the reverse cost volume is a re-indexing of the forward one (include/unflow.h,
    unflow_correlation_fwd_bidir).  The reference builds the two volumes with two op instances
    (flownet.py:34-44); the values are bit-identical."""

    @staticmethod
    def forward(ctx, in0, in1, attrs):
        in0 = _prep(in0, "input_0")
        in1 = _prep(in1, "input_1")
        if in0.shape != in1.shape:
            raise ValueError("Input shapes have to be the same")
        B, C, H, W = in0.shape
        import ctypes
        oc, oh, ow = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        lib = _native.lib()
        check(lib.unflow_correlation_out_shape(H, W, *attrs, ctypes.byref(oc), ctypes.byref(oh),
                                               ctypes.byref(ow)), "correlation")
        out = torch.empty(B, oc.value, oh.value, ow.value, device=in0.device, dtype=torch.float32)
        rev = torch.empty_like(out)
        with torch.cuda.device(in0.device), kernel_timer.span("correlation_fwd_bidir"):
            check(lib.unflow_correlation_fwd_bidir(in0.data_ptr(), in1.data_ptr(), out.data_ptr(), rev.data_ptr(),
                                                   B, C, H, W, *attrs, _stream()), "correlation_bidir")
        ctx.save_for_backward(in0, in1)
        ctx.attrs = attrs
        return out, rev

    @staticmethod
    def backward(ctx, gout, grev):
        in0, in1 = ctx.saved_tensors
        B, C, H, W = in0.shape
        gout, grev = gout.contiguous(), grev.contiguous()
        geff = torch.empty_like(gout)
        g0 = torch.empty_like(in0)
        g1 = torch.empty_like(in1)
        lib = _native.lib()
        with torch.cuda.device(in0.device), kernel_timer.span("correlation_bwd_bidir"):
            check(lib.unflow_correlation_fold_grad(gout.data_ptr(), grev.data_ptr(), geff.data_ptr(), B, C, H, W,
                                                   *ctx.attrs, _stream()), "correlation_fold_grad")
            check(lib.unflow_correlation_bwd(geff.data_ptr(), in0.data_ptr(), in1.data_ptr(),
                                             g0.data_ptr(), g1.data_ptr(), B, C, H, W, *ctx.attrs, _stream()),
                  "correlation_grad")
        return g0, g1, None


def _nhwc_geometry(t):
    """(batch stride, pixel pitch) of an NCHW-shaped tensor whose memory is NHWC (channels contiguous, rows and
    pixels linear, any pitch / batch stride), or None."""
    N, C, H, W = t.shape
    sn, sc, sh, sw = t.stride()
    if (sc == 1 or C == 1) and sw >= C and sh == W * sw and (N == 1 or sn >= H * sh):
        return sn, sw
    return None


class _CorrelationBidirConcat(torch.autograd.Function):
    """The FlowNetC trunk input of the bidirectional pass, concat([conv_redir, corr], channels) for both
    directions at once (reference flownet.py:34-44 with the two directions batched), written into ONE
    pre-allocated NHWC buffer:

        buf[:B,  :, :, c0:c0+D] = correlation(feat[:B], feat[B:])        feat: [2B, C, H, W], NHWC memory
        buf[B:,  :, :, c0:c0+D] = correlation(feat[B:], feat[:B])

    The correlation kernels keep the reference op's NCHW layout; the tensors crossing into the NHWC conv stack
    go through the tiled transposes of csrc/relayout.cu instead of strided library copies, and the gradient
    of ``feat`` is produced as one NHWC buffer that the other consumer of ``feat`` (conv_redir's input-gradient
    kernel) accumulates into (conv_ops gradient slots) -- no zero fill, slice copy or add.
    ``members``: tensors already living in ``buf`` (conv_redir, written there by its conv's epilogue); their
    gradient is the matching channel slice of the incoming gradient."""

    @staticmethod
    def forward(ctx, buf, c0, feat, attrs, *members):
        from .core import conv_ops
        n2, C, H, W = feat.shape
        B = n2 // 2
        geo = _nhwc_geometry(feat)
        if geo is None:
            feat = feat.contiguous(memory_format=torch.channels_last)
            geo = _nhwc_geometry(feat)
        import ctypes
        oc, oh, ow = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        lib = _native.lib()
        check(lib.unflow_correlation_out_shape(H, W, *attrs, ctypes.byref(oc), ctypes.byref(oh),
                                               ctypes.byref(ow)), "correlation")
        D = oc.value
        assert (oh.value, ow.value) == (H, W) and tuple(buf.shape[:3]) == (n2, H, W) and buf.shape[3] >= c0 + D
        P = H * W
        planar = torch.empty(n2, C, H, W, device=feat.device, dtype=torch.float32)       # in0 | in1, NCHW
        vol = torch.empty(n2, D, H, W, device=feat.device, dtype=torch.float32)          # corr_ab | corr_ba
        pitch = buf.stride(2)
        with torch.cuda.device(feat.device):
            check(lib.unflow_interleaved_to_planar(feat.data_ptr(), geo[0], geo[1], planar.data_ptr(), C * P,
                                                   n2, C, P, _stream()), "interleaved_to_planar")
            with kernel_timer.span("correlation_fwd_bidir"):
                check(lib.unflow_correlation_fwd_bidir(planar[:B].data_ptr(), planar[B:].data_ptr(),
                                                       vol[:B].data_ptr(), vol[B:].data_ptr(),
                                                       B, C, H, W, *attrs, _stream()), "correlation_bidir")
            check(lib.unflow_planar_to_interleaved(vol.data_ptr(), D * P, buf.data_ptr() + 4 * c0, buf.stride(0),
                                                   pitch, n2, D, P, 0, _stream()), "planar_to_interleaved")
        ctx.save_for_backward(planar)
        ctx.attrs, ctx.c0, ctx.D = attrs, c0, D
        ctx.gen = conv_ops._generation
        ctx.feat_key = (feat.data_ptr(), tuple(feat.shape))
        ctx.member_spans = []
        off = 0
        for m in members:
            assert m.data_ptr() == buf.data_ptr() + 4 * off and m.stride(3) == pitch, "member not in its slot"
            ctx.member_spans.append((off, off + m.shape[1]))
            off += m.shape[1]
        assert off == c0
        return buf[..., :c0 + D].permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, g):
        from .core import conv_ops
        planar, = ctx.saved_tensors
        n2, C, H, W = planar.shape
        B, P, D, c0 = n2 // 2, H * W, ctx.D, ctx.c0
        geo = _nhwc_geometry(g)
        if geo is None:
            g = g.contiguous(memory_format=torch.channels_last)
            geo = _nhwc_geometry(g)
        lib = _native.lib()
        gvol = torch.empty(n2, D, H, W, device=g.device, dtype=torch.float32)
        geff = torch.empty(B, D, H, W, device=g.device, dtype=torch.float32)
        gplanar = torch.empty_like(planar)
        slot = conv_ops.grad_slot_get(ctx.gen, None, key=ctx.feat_key)
        if slot is not None and _nhwc_geometry(slot) is None:
            slot = None
        if slot is None:
            gbuf = torch.empty(n2, H, W, (C + 3) // 4 * 4, device=g.device, dtype=torch.float32)
            gfeat = gbuf[..., :C].permute(0, 3, 1, 2)
        else:
            gfeat = slot
        sgeo = _nhwc_geometry(gfeat)
        with torch.cuda.device(g.device):
            check(lib.unflow_interleaved_to_planar(g.data_ptr() + 4 * c0 * g.stride(1), geo[0], geo[1],
                                                   gvol.data_ptr(), D * P, n2, D, P, _stream()),
                  "interleaved_to_planar")
            with kernel_timer.span("correlation_bwd_bidir"):
                check(lib.unflow_correlation_fold_grad(gvol[:B].data_ptr(), gvol[B:].data_ptr(), geff.data_ptr(),
                                                       B, C, H, W, *ctx.attrs, _stream()), "correlation_fold_grad")
                check(lib.unflow_correlation_bwd(geff.data_ptr(), planar[:B].data_ptr(), planar[B:].data_ptr(),
                                                 gplanar[:B].data_ptr(), gplanar[B:].data_ptr(), B, C, H, W,
                                                 *ctx.attrs, _stream()), "correlation_grad")
            check(lib.unflow_planar_to_interleaved(gplanar.data_ptr(), C * P, gfeat.data_ptr(), sgeo[0], sgeo[1],
                                                   n2, C, P, 0 if slot is None else 1, _stream()),
                  "planar_to_interleaved")
        if slot is None:
            conv_ops.grad_slot_put(ctx.gen, ctx.feat_key, gfeat)
        return (None, None, gfeat if slot is None else None, None) + tuple(g[:, a:b] for a, b in ctx.member_spans)


def correlation_bidir_concat(buf, c0, feat, members, **kwargs):
    """See _CorrelationBidirConcat; None when the tiled kernel does not serve the attributes / shape."""
    attrs = _corr_attrs(kwargs)
    n2, C, H, W = feat.shape
    if not (feat.is_cuda and n2 % 2 == 0 and _native.lib().unflow_correlation_fwd_path(C, H, W, *attrs) == 1):
        return None
    return _CorrelationBidirConcat.apply(buf, c0, feat, attrs, *members)


def correlation_bidir(first, second, **kwargs):
    """(correlation(first, second), correlation(second, first)); one pass where the tiled kernel serves the
    attributes and shape, two ordinary calls otherwise."""
    attrs = _corr_attrs(kwargs)
    B, C, H, W = first.shape
    if first.is_cuda and first.dim() == 4 and _native.lib().unflow_correlation_fwd_path(C, H, W, *attrs) == 1:
        return _CorrelationBidir.apply(first, second, attrs)
    return _Correlation.apply(first, second, attrs), _Correlation.apply(second, first, attrs)


class _Warp(torch.autograd.Function):
    """Bilinear gather; ``mode`` selects the op semantics (zero) or image_warp (clamp)."""

    @staticmethod
    def forward(ctx, images, flows, mode, image_grad):
        images = _prep(images, "images")
        flows = _prep(flows, "flows")
        B, H, W, C = images.shape
        if tuple(flows.shape) != (B, H, W, 2):
            raise ValueError("flows must have shape [B,H,W,2] matching images")
        out = torch.empty_like(images)
        with torch.cuda.device(images.device):
            check(_native.lib().unflow_backward_warp_fwd(images.data_ptr(), flows.data_ptr(),
                                                         out.data_ptr(), B, H, W, C, mode, _stream()),
                  "backward_warp")
        ctx.save_for_backward(images, flows)
        ctx.mode = mode
        ctx.image_grad = image_grad
        return out

    @staticmethod
    def backward(ctx, grad):
        images, flows = ctx.saved_tensors
        B, H, W, C = images.shape
        grad = grad.contiguous()
        dflow = torch.empty_like(flows)
        want_dimg = ctx.image_grad and ctx.needs_input_grad[0]
        dimg = torch.zeros_like(images) if want_dimg else None
        with torch.cuda.device(images.device):
            check(_native.lib().unflow_backward_warp_bwd(
                grad.data_ptr(), images.data_ptr(), flows.data_ptr(), dflow.data_ptr(),
                dimg.data_ptr() if want_dimg else None, B, H, W, C, ctx.mode, _stream()),
                "backward_warp_grad")
        return dimg, dflow, None, None


def backward_warp(images, flows):
    # reference ops.py:80-84: gradient for the flow only
    return _Warp.apply(images, flows, BORDER_ZERO, False)


def _image_warp(im, flow):
    """core.image_warp.image_warp: clamped taps, gradients for image and flow."""
    return _Warp.apply(im, flow, BORDER_CLAMP, True)


class _ForwardWarp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, flows):
        flows = _prep(flows, "flows")
        B, H, W, two = flows.shape
        if two != 2:
            raise ValueError("flows must have shape [B,H,W,2]")
        out = torch.empty(B, H, W, 1, device=flows.device, dtype=torch.float32)
        with torch.cuda.device(flows.device):
            check(_native.lib().unflow_forward_warp_fwd(flows.data_ptr(), out.data_ptr(), B, H, W,
                                                        _stream()), "forward_warp")
        ctx.save_for_backward(flows)
        return out

    @staticmethod
    def backward(ctx, grad):
        (flows,) = ctx.saved_tensors
        B, H, W, _ = flows.shape
        grad = grad.contiguous()
        dflow = torch.empty_like(flows)
        with torch.cuda.device(flows.device):
            check(_native.lib().unflow_forward_warp_bwd(grad.data_ptr(), flows.data_ptr(),
                                                        dflow.data_ptr(), B, H, W, _stream()),
                  "forward_warp_grad")
        return dflow


def forward_warp(flows):
    return _ForwardWarp.apply(flows)


def downsample(images, scale=2):
    """Not differentiable (reference ops.py:107)."""
    images = _prep(images.detach() if torch.is_tensor(images) else images, "images")
    B, H, W, C = images.shape
    scale = int(scale)
    if scale < 1 or H % scale != 0 or W % scale != 0:
        raise ValueError("Input height and width must be divisible by scale")
    out = torch.empty(B, H // scale, W // scale, C, device=images.device, dtype=torch.float32)
    with torch.cuda.device(images.device):
        check(_native.lib().unflow_downsample(images.data_ptr(), out.data_ptr(), B, H, W, C, scale,
                                              _stream()), "downsample")
    return out


if __name__ == "__main__":
    print(compile())
