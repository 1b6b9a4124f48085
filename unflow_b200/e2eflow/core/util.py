"""reference src/e2eflow/core/util.py:12-26 (resize helpers and the downsample wrapper)."""
import torch

from ..ops import downsample as downsample_ops
from . import tf_image


def resize_area(tensor, like):
    _, h, w, _ = like.shape
    return tf_image.resize_area(tensor, [h, w]).detach()


def resize_bilinear(tensor, like):
    _, h, w, _ = like.shape
    return tf_image.resize_bilinear(tensor, [h, w]).detach()


def downsample(tensor, num):
    _, height, width, _ = tensor.shape
    if height % 2 == 0 and width % 2 == 0:
        return downsample_ops(tensor, num)
    else:
        return tf_image.resize_area(tensor, [int(height / num), int(width / num)])
