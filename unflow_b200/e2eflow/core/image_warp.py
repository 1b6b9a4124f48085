"""Backward warp used on the training path (reference src/e2eflow/core/image_warp.py:4-76).

The reference builds it from ~40 TF ops (4 gathers + add_n); here it is one CUDA kernel
(csrc/warp.cu, border mode CLAMP) with the same semantics: integer taps pos + floor(flow)
clamped to the image, weights from flow - floor(flow), gradients w.r.t. both the image
(scatter-add) and the flow, as TF autodiff provides in the reference.
"""
from ..ops import _image_warp


def image_warp(im, flow):
    """Performs a backward warp of an image using the predicted flow.

    Args:
        im: Batch of images. [num_batch, height, width, channels]
        flow: Batch of flow vectors. [num_batch, height, width, 2]
    Returns:
        warped: transformed image of the same shape as the input image.
    """
    return _image_warp(im, flow)
