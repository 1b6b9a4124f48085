"""Small conveniences shared by the modules that call the C ABI directly."""
import torch

from .._native import BORDER_CLAMP, BORDER_ZERO, check  # noqa: F401

BORDER_STN = 2


def stream():
    return torch.cuda.current_stream().cuda_stream
