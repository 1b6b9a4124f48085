"""A minimal eager stand-in for the TensorFlow 1.x API surface the reference's hot-path Python uses
(TEST INFRASTRUCTURE; used only by tests/golden/make_reference_run.py).

TensorFlow cannot be installed here, so the reference's own source files
(/root/reference/src/e2eflow/core/{losses,image_warp,flownet,unsupervised,util}.py) cannot run as
they are.  This module lets them run UNMODIFIED: it is registered as ``tensorflow`` (plus
``tensorflow.contrib.{slim,layers,distributions}``) and implements, on torch CPU tensors, exactly the
calls those files make.  What the golden vectors generated this way pin is the reference's GRAPH --
which ops, in which order, with which constants, masks, scopes and variable names; what they do
NOT pin is the arithmetic inside the TensorFlow primitives, which is restated here from the TF 1.x
documentation (each primitive says so below).  The custom ops (``e2eflow.ops``) are served by the C
restatement in oracle/oracle_ops.c, which is pinned separately by the reference's own known-answer
tests.

Conventions: tensors are ``Tensor`` (a torch.Tensor subclass whose ``.shape`` has ``as_list()``);
``tf.shape`` returns a list of Python ints and scalar shape arithmetic stays in Python / numpy
float32, which is what a TF session would constant-fold to.
"""
import builtins
import contextlib
import math
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

float32, int32, bool = 'float32', 'int32', 'bool'          # noqa: A001  (tf.bool)
_TORCH_DTYPE = {'float32': torch.float32, 'int32': torch.int32, 'bool': torch.bool, 'uint16': torch.int32}


class Shape(tuple):
    def as_list(self):
        return list(self)


class Tensor(torch.Tensor):
    @property
    def shape(self):
        return Shape(super().shape)

    def get_shape(self):
        return self.shape

    # TensorFlow tensors are immutable: ``a *= b`` in the reference REBINDS ``a`` (losses.py:51-55 does
    # that with two names bound to one mask); torch would update the shared tensor in place.
    def __imul__(self, other):
        return self * other

    def __iadd__(self, other):
        return self + other

    def __isub__(self, other):
        return self - other

    def __itruediv__(self, other):
        return self / other


def _t(x, dtype=None):
    """anything -> Tensor (python scalars / lists / numpy included)."""
    if isinstance(x, torch.Tensor):
        out = x
    elif isinstance(x, (list, tuple)) and any(isinstance(e, (torch.Tensor, list, tuple)) for e in x):
        out = torch.stack([_t(e) for e in x])                  # nested lists holding tensors
    else:
        out = torch.as_tensor(np.asarray(x))
        if out.dtype == torch.float64:
            out = out.float()
        if out.dtype == torch.int64 and dtype is None:
            out = out.int()
    if dtype is not None:
        out = out.to(_TORCH_DTYPE[dtype] if isinstance(dtype, str) else dtype)
    return out.as_subclass(Tensor)


def _is_scalar(x):
    return isinstance(x, (int, float, np.integer, np.floating)) and not isinstance(x, builtins.bool)


# ---- graph bookkeeping -----------------------------------------------------------------------------
class GraphKeys:
    SUMMARIES = 'summaries'


class _State:
    def __init__(self):
        self.reset({})

    def reset(self, variables):
        self.variables = variables          # name -> torch leaf tensor (TF layout)
        self.scopes = []                    # current variable-scope path
        self.created = []                   # variable names in creation order
        self.reg_losses = []
        self.collections = {}
        self.arg_scope = []                 # stack of (functions, kwargs)
        self.rng = torch.Generator().manual_seed(0)
        self.draws = []                     # every random tensor handed out, in call order
        self.queues = []                    # file-name lists handed to string_input_producer, in call order
        self.decode_shape = (4, 6, 3)       # what the stand-in decode_png "reads"


STATE = _State()


class _Scope:
    def __init__(self, name):
        self.name = name

    def reuse_variables(self):
        pass                                # variables are looked up by name: reuse is implicit


@contextlib.contextmanager
def variable_scope(name_or_scope, default_name=None, values=None, reuse=None):
    name = name_or_scope.name if isinstance(name_or_scope, _Scope) else name_or_scope
    STATE.scopes.append(name)
    try:
        yield _Scope(name)
    finally:
        STATE.scopes.pop()


@contextlib.contextmanager
def name_scope(name, default_name=None, values=None):
    yield name or default_name


def add_to_collection(name, value):
    STATE.collections.setdefault(name, []).append(value)


def identity(x, name=None):
    return x


class _Summary:
    @staticmethod
    def scalar(*a, **k):
        return None


summary = _Summary()


class _Losses:
    @staticmethod
    def get_regularization_loss():
        """Sum of the regulariser terms recorded when variables were first created."""
        total = _t(0.0)
        for r in STATE.reg_losses:
            total = total + r
        return total


losses = _Losses()


# ---- element-wise / shape ops (documented TF semantics; broadcasting as numpy) ----------------------
def shape(x):
    return [int(s) for s in x.shape]


def constant(value, dtype=None, name=None):
    return _t(value, dtype)


def convert_to_tensor(value, dtype=None, name=None):
    return _t(value, dtype)


def cast(x, dtype, name=None):
    if _is_scalar(x):                       # shape arithmetic: numpy float32 / python int
        return np.float32(x) if dtype in ('float32', float32) else int(x)
    return _t(x).to(_TORCH_DTYPE[dtype]).as_subclass(Tensor)     # float -> int truncates toward zero


def to_int32(x):
    return cast(x, 'int32')


def zeros(shape=None, dtype='float32'):       # noqa: A002
    return _t(torch.zeros([int(s) for s in shape], dtype=_TORCH_DTYPE[dtype]))


def ones(shape=None, dtype='float32'):        # noqa: A002
    return _t(torch.ones([int(s) for s in shape], dtype=_TORCH_DTYPE[dtype]))


def ones_like(x):
    return _t(torch.ones_like(x))


def zeros_like(x):
    return _t(torch.zeros_like(x))


def reduce_max(x, axis=None, keepdims=False):
    x = _t(x)
    return _t(x.max() if axis is None else x.amax(dim=axis, keepdim=keepdims))


def range(*args):                      # noqa: A001  (tf.range)
    return _t(torch.arange(*[int(a) for a in args], dtype=torch.int32))


def reshape(x, shape_):
    return _t(x).reshape([int(s) for s in shape_]).as_subclass(Tensor)


def expand_dims(x, axis):
    return _t(x).unsqueeze(axis).as_subclass(Tensor)


def transpose(x, perm):
    return _t(x).permute(*perm).as_subclass(Tensor)


def matmul(a, b):
    a, b = _t(a), _t(b)
    if not a.is_floating_point():               # integer outer products (spatial_transformer._repeat)
        return _t((a.long() @ b.long()).to(a.dtype))
    return _t(a @ b)


def linspace(start, stop, num):
    return _t(torch.linspace(float(start), float(stop), int(num)))


def slice(x, begin, size):                  # noqa: A001  (tf.slice)
    x = _t(x)
    idx = tuple(builtins.slice(int(b), None if int(n) == -1 else int(b) + int(n)) for b, n in zip(begin, size))
    return x[idx].as_subclass(Tensor)


def random_uniform(shape=None, minval=0, maxval=1, dtype='float32', seed=None):       # noqa: A002
    u = torch.rand([int(s) for s in shape], generator=STATE.rng) * (float(maxval) - float(minval)) + float(minval)
    STATE.draws.append(u.clone())
    return _t(u)


def random_normal(shape=None, mean=0.0, stddev=1.0, dtype='float32', seed=None):     # noqa: A002
    g = torch.randn([int(s) for s in shape], generator=STATE.rng) * float(stddev) + float(mean)
    STATE.draws.append(g.clone())
    return _t(g)


def tile(x, multiples):
    return _t(x).repeat(*[int(m) for m in multiples]).as_subclass(Tensor)


def concat(values=None, axis=None, name=None, **kw):
    if isinstance(values, int):             # tf.concat(axis, values) of very old code
        values, axis = axis, values
    return _t(torch.cat([_t(v) for v in values], dim=axis))


def stack(values, axis=0):
    if all(_is_scalar(v) for v in values):      # shape vectors stay Python lists
        return [int(v) for v in values]
    return _t(torch.stack([_t(v) for v in values], dim=axis))


def unstack(value, num=None, axis=0):
    if isinstance(value, (list, tuple)):
        return list(value)
    return [t.as_subclass(Tensor) for t in torch.unbind(_t(value), dim=axis)]


def split(value=None, num_or_size_splits=None, axis=0, **kw):
    return [t.as_subclass(Tensor) for t in torch.chunk(_t(value), num_or_size_splits, dim=axis)]


def pad(x, paddings):
    flat = []
    for lo, hi in reversed([list(p) for p in paddings]):
        flat += [int(lo), int(hi)]
    return _t(F.pad(_t(x), flat))


def stop_gradient(x):
    return _t(x).detach().as_subclass(Tensor)


def gather(params, indices):
    return _t(params)[_t(indices).long()].as_subclass(Tensor)


def add_n(values):
    out = values[0]
    for v in values[1:]:
        out = out + v
    return out


def _ew(fn_t, fn_s):
    def op(*xs):
        if all(_is_scalar(x) for x in xs):
            return fn_s(*xs)
        return _t(fn_t(*[x if isinstance(x, torch.Tensor) else _t(x) for x in xs]))
    return op


square = _ew(torch.square, lambda x: x * x)
sqrt = _ew(torch.sqrt, lambda x: np.sqrt(np.float32(x)))
floor = _ew(torch.floor, lambda x: np.floor(x))
ceil = _ew(torch.ceil, lambda x: np.ceil(x))
abs = _ew(torch.abs, lambda x: np.abs(x))          # noqa: A001
minimum = _ew(torch.minimum, lambda a, b: min(a, b))
maximum = _ew(torch.maximum, lambda a, b: max(a, b))
multiply = _ew(torch.mul, lambda a, b: a * b)
logical_and = _ew(torch.logical_and, lambda a, b: builtins.bool(a) and builtins.bool(b))
greater = _ew(torch.gt, lambda a, b: a > b)
greater_equal = _ew(torch.ge, lambda a, b: a >= b)
less = _ew(torch.lt, lambda a, b: a < b)
equal = _ew(torch.eq, lambda a, b: a == b)
atan = _ew(torch.atan, lambda x: np.arctan(np.float32(x)))
mod = _ew(torch.remainder, lambda a, b: a % b)          # floored modulo, like tf.mod
sin = _ew(torch.sin, lambda x: np.sin(np.float32(x)))
cos = _ew(torch.cos, lambda x: np.cos(np.float32(x)))


def pow(x, y):                          # noqa: A001
    return _t(torch.pow(_t(x), y))


def clip_by_value(x, lo, hi):
    x = _t(x)
    return _t(torch.minimum(torch.maximum(x, torch.as_tensor(lo, dtype=x.dtype)), torch.as_tensor(hi, dtype=x.dtype)))


def reduce_sum(x, axis=None, keepdims=False, keep_dims=None):
    if keep_dims is not None:
        keepdims = keep_dims
    x = _t(x)
    return _t(x.sum() if axis is None else x.sum(dim=axis, keepdim=keepdims))


def where(cond, a, b):
    return _t(torch.where(cond, a, b))


def _unsupported(name):
    def fn(*a, **k):
        raise NotImplementedError("tf.%s is outside the augment=False hot path the shim serves" % name)
    return fn


for _n in ('round', 'placeholder', 'Variable', 'extract_image_patches'):
    globals()[_n] = _unsupported(_n)


# ---- TF primitives restated from the TF 1.x documentation ------------------------------------------
def _same_pad(size, k, stride):
    """SAME: out = ceil(in / stride); total padding split floor / ceil, the extra pixel at the end."""
    out = -(-size // stride)
    total = max((out - 1) * stride + k - size, 0)
    return total // 2, total - total // 2


def _conv2d_nchw(x, w_oihw, stride):
    kh, kw = w_oihw.shape[2], w_oihw.shape[3]
    pt, pb = _same_pad(x.shape[2], kh, stride)
    pl, pr = _same_pad(x.shape[3], kw, stride)
    return F.conv2d(F.pad(x, (pl, pr, pt, pb)), w_oihw, None, stride=stride)


class _NN:
    @staticmethod
    def conv2d(input, filter, strides, padding, name=None):     # noqa: A002
        """NHWC input, HWIO filter, SAME zero padding."""
        assert padding == 'SAME' and list(strides) == [1, 1, 1, 1]
        y = _conv2d_nchw(_t(input).permute(0, 3, 1, 2), _t(filter).permute(3, 2, 0, 1), 1)
        return _t(y.permute(0, 2, 3, 1))


nn = _NN()


class _Image:
    @staticmethod
    def rgb_to_grayscale(images):
        """Weighted sum with [0.2989, 0.5870, 0.1140] over the last axis (kept as size 1)."""
        w = torch.tensor([0.2989, 0.5870, 0.1140])
        return _t((_t(images) * w).sum(-1, keepdim=True))

    @staticmethod
    def resize_bilinear(images, size, align_corners=False):
        """TF 1.x kernel with align_corners=False: source coordinate = dst * (in / out) (no
        half-pixel offset), neighbours clamped at the border."""
        x = _t(images)
        B, H, W, C = x.shape
        oh, ow = int(size[0]), int(size[1])

        def axis(n_in, n_out):
            scale = np.float32(n_in) / np.float32(n_out)
            src = torch.arange(n_out, dtype=torch.float32) * float(scale)
            lo = src.floor().long().clamp(max=n_in - 1)
            hi = (lo + 1).clamp(max=n_in - 1)
            return lo, hi, (src - src.floor())

        y0, y1, fy = axis(H, oh)
        x0, x1, fx = axis(W, ow)
        top = x[:, y0][:, :, x0] + (x[:, y0][:, :, x1] - x[:, y0][:, :, x0]) * fx.view(1, 1, ow, 1)
        bot = x[:, y1][:, :, x0] + (x[:, y1][:, :, x1] - x[:, y1][:, :, x0]) * fx.view(1, 1, ow, 1)
        return _t(top + (bot - top) * fy.view(1, oh, 1, 1))

    @staticmethod
    def hsv_to_rgb(images):
        """TF's kernel (colorspace_op.h): h in [0,1) scaled to 6 sectors, piecewise-linear channels,
        clamped with cwiseMax / cwiseMin -- fmaxf / fminf on the GPU the reference runs on, which
        drop a NaN operand (flow_util.atan2 yields a NaN hue for a zero flow vector)."""
        x = _t(images)
        h, s_, v = x[..., 0], x[..., 1], x[..., 2]
        nh = h * 6
        zero, one = torch.zeros(()), torch.ones(())
        clamp01 = lambda z: torch.fmin(torch.fmax(z, zero), one)
        dr = clamp01((nh - 3).abs() - 1)
        dg = clamp01(2 - (nh - 2).abs())
        db = clamp01(2 - (nh - 4).abs())
        oms = 1 - s_
        return _t(torch.stack([(oms + s_ * dr) * v, (oms + s_ * dg) * v, (oms + s_ * db) * v], -1))

    @staticmethod
    def resize_image_with_crop_or_pad(image, target_height, target_width):
        """Centre crop ((in - target) // 2 from the start) and / or centre zero pad ((target - in) // 2
        before), per axis; [h,w,c] or [b,h,w,c]."""
        x = _t(image)
        hd = x.dim() - 3
        for d, target in ((hd, int(target_height)), (hd + 1, int(target_width))):
            n = x.shape[d]
            if n > target:
                x = x.narrow(d, (n - target) // 2, target)
            elif n < target:
                before = (target - n) // 2
                pad_spec = [0, 0] * (x.dim() - 1 - d) + [before, target - n - before]
                x = F.pad(x, pad_spec)
        return _t(x)

    @staticmethod
    def resize_area(images, size, align_corners=False):
        x = _t(images)
        oh, ow = int(size[0]), int(size[1])
        if x.shape[1] % oh or x.shape[2] % ow:
            raise NotImplementedError("resize_area: only integer factors are needed on the hot path")
        return _t(F.avg_pool2d(x.permute(0, 3, 1, 2), (x.shape[1] // oh, x.shape[2] // ow)).permute(0, 2, 3, 1))


image = _Image()


# ---- tensorflow.contrib.slim / layers ---------------------------------------------------------------
def _relu(x):
    return torch.relu(x)


def _scope_prefix(scope):
    parts = [s for s in STATE.scopes + [scope] if s]
    return '/'.join(parts)


def _variable(name, shape_, regularizer):
    if name not in STATE.variables:
        raise KeyError("the reference graph asks for variable '%s' %s which the supplied set does not have"
                       % (name, list(shape_)))
    v = STATE.variables[name]
    if list(v.shape) != list(shape_):
        raise ValueError("variable '%s': reference graph shape %s, supplied %s" % (name, list(shape_), list(v.shape)))
    if name not in STATE.created:
        STATE.created.append(name)
        if regularizer is not None:
            STATE.reg_losses.append(regularizer(v))
    return v


def _merge_arg_scope(fn, kwargs):
    merged = {}
    for fns, kw in STATE.arg_scope:
        if fn in fns:
            merged.update(kw)
    merged.update(kwargs)
    return merged


def slim_conv2d(inputs, num_outputs, kernel_size, **kwargs):
    """slim.conv2d: variables '<scope>/weights' [k,k,in,out] and '<scope>/biases' [out]; SAME padding;
    default activation relu unless overridden (arg_scope / activation_fn=None)."""
    kw = _merge_arg_scope(slim_conv2d, kwargs)
    assert kw.get('data_format') == 'NCHW' and kw.get('padding', 'SAME') == 'SAME'
    stride, k = kw.get('stride', 1), int(kernel_size)
    x = _t(inputs)
    prefix = _scope_prefix(kw['scope'])
    w = _variable(prefix + '/weights', [k, k, x.shape[1], num_outputs], kw.get('weights_regularizer'))
    b = _variable(prefix + '/biases', [num_outputs], None)
    y = _conv2d_nchw(x, w.permute(3, 2, 0, 1), stride) + b.view(1, -1, 1, 1)
    act = kw.get('activation_fn', _relu)
    return _t(act(_t(y)) if act is not None else y)


def slim_conv2d_transpose(inputs, num_outputs, kernel_size, **kwargs):
    """slim.conv2d_transpose: variables '<scope>/weights' [k,k,out,in], '<scope>/biases' [out]; by
    TF's definition the gradient of the SAME-padded stride-s conv2d (out -> in) w.r.t. its input, the
    output extent being in*stride."""
    kw = _merge_arg_scope(slim_conv2d_transpose, kwargs)
    assert kw.get('data_format') == 'NCHW' and kw.get('padding', 'SAME') == 'SAME'
    stride, k = kw.get('stride', 1), int(kernel_size)
    x = _t(inputs)
    prefix = _scope_prefix(kw['scope'])
    w = _variable(prefix + '/weights', [k, k, num_outputs, x.shape[1]], kw.get('weights_regularizer'))
    b = _variable(prefix + '/biases', [num_outputs], None)
    N, Cin, h, wd = x.shape
    H, W = h * stride, wd * stride
    pt, pb = _same_pad(H, k, stride)
    pl, pr = _same_pad(W, k, stride)
    # forward conv: [N,num_outputs,H,W] (padded by pt,pb,pl,pr) -> [N,Cin,h,w] with kernel [Cin,num_outputs,k,k]
    w_fwd = w.permute(3, 2, 0, 1)
    grad_padded = torch.nn.grad.conv2d_input((N, num_outputs, H + pt + pb, W + pl + pr), w_fwd, x, stride=stride, padding=0)
    y = grad_padded[:, :, pt:pt + H, pl:pl + W] + b.view(1, -1, 1, 1)
    act = kw.get('activation_fn', _relu)
    return _t(act(_t(y)) if act is not None else y)


@contextlib.contextmanager
def slim_arg_scope(fns, **kwargs):
    STATE.arg_scope.append((list(fns), kwargs))
    try:
        yield
    finally:
        STATE.arg_scope.pop()


def slim_l2_regularizer(scale):
    def reg(v):
        return _t(scale * (torch.sum(torch.square(v)) / 2))       # scale * tf.nn.l2_loss(v)
    return reg


# ---- input queues: only the FILE LISTS matter (core/input.py, kitti/input.py) --------------------
class _Train:
    @staticmethod
    def string_input_producer(string_tensor, num_epochs=None, shuffle=True, capacity=32, **kw):
        assert shuffle is False              # the reference never lets TF reorder the files
        STATE.queues.append([str(f) for f in string_tensor])
        return len(STATE.queues) - 1

    @staticmethod
    def batch(tensors, batch_size=1, num_threads=1, allow_smaller_final_batch=False, **kw):
        return tensors

    @staticmethod
    def get_checkpoint_state(path):
        return None


train = _Train()


class WholeFileReader:
    def read(self, queue):
        return None, queue


def _decode_png(contents, channels=None, dtype=None, name=None):
    h, w, c = STATE.decode_shape
    return _t(torch.zeros(h, w, c if channels is None else channels))


image.decode_png = staticmethod(_decode_png)
uint16 = 'uint16'


def install():
    """Register this module as ``tensorflow`` (+ contrib.slim / layers / distributions)."""
    me = sys.modules[__name__]
    contrib = types.ModuleType('tensorflow.contrib')
    slim = types.ModuleType('tensorflow.contrib.slim')
    slim.conv2d, slim.conv2d_transpose = slim_conv2d, slim_conv2d_transpose
    slim.arg_scope, slim.l2_regularizer = slim_arg_scope, slim_l2_regularizer
    layers = types.ModuleType('tensorflow.contrib.layers')
    layers.variance_scaling_initializer = lambda *a, **k: 'variance_scaling_initializer'
    dist = types.ModuleType('tensorflow.contrib.distributions')
    dist.Normal = object
    contrib.slim, contrib.layers, contrib.distributions = slim, layers, dist
    me.contrib = contrib
    sys.modules.update({'tensorflow': me, 'tensorflow.contrib': contrib, 'tensorflow.contrib.slim': slim,
                        'tensorflow.contrib.layers': layers, 'tensorflow.contrib.distributions': dist})
    return me
