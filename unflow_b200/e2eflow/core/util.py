"""Resize helpers with the reference's names (src/e2eflow/core/util.py:12-26).

``resize_area`` / ``resize_bilinear`` resize to the spatial size of a second NHWC tensor and block
the gradient; ``downsample`` is the box-mean CUDA op for even sizes and TF's area resize otherwise.
(``summarized_placeholder`` belongs to the TensorBoard plumbing, which is out of scope.)
"""
from . import tf_image
from ..ops import downsample as _box_downsample


def _size_of(nhwc):
    return [int(nhwc.shape[1]), int(nhwc.shape[2])]


def resize_area(tensor, like):
    return tf_image.resize_area(tensor, _size_of(like)).detach()


def resize_bilinear(tensor, like):
    return tf_image.resize_bilinear(tensor, _size_of(like)).detach()


def downsample(tensor, num):
    rows, cols = _size_of(tensor)
    if rows % 2 or cols % 2:     # odd extent: the op's output grid is undefined, use area resize
        return tf_image.resize_area(tensor, [int(rows / num), int(cols / num)])
    return _box_downsample(tensor, num)
