"""CPU restatement of core/util.py helpers (TEST INFRASTRUCTURE ONLY).

/root/reference/src/e2eflow/core/util.py:12-26.
"""
from . import ops as _ops
from . import tf_compat as tfc


def resize_area(tensor, like):  # util.py:12-14
    return tfc.resize_area(tensor, like.shape[1:3]).detach()


def resize_bilinear(tensor, like):  # util.py:17-19
    return tfc.resize_bilinear_legacy(tensor, like.shape[1:3]).detach()


def downsample(tensor, num):  # util.py:21-26
    _, height, width, _ = tensor.shape
    if height % 2 == 0 and width % 2 == 0:
        return _ops.downsample(tensor, num)
    return tfc.resize_area(tensor, [int(height / num), int(width / num)])
