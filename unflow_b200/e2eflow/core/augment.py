"""Training-time augmentation -- reference src/e2eflow/core/augment.py:7-134 and the sampler of
src/e2eflow/core/spatial_transformer.py:18-175 (the step before the
hot path), on the GPU.

Same function names, keyword arguments and output structure.  The random draws come from torch's
generator instead of TF's, so individual samples differ from a TF run (augmentation cannot be
parity-pinned sample by sample); the deterministic cores ``transformer`` (given theta) and
``photometric`` (given the drawn parameters) are parity-tested against the oracle.  The bilinear
sampler is the hand-written gather kernel (csrc/warp.cu, border mode STN)."""
import math

import torch

from ... import _native
from .._native_shim import BORDER_STN, check, stream

_gen = None
_device_rng = False


def seed(s):
    """Seed the augmentation random stream (torch.Generator on the CPU; draws are tiny) and the
    device generator used in ``set_device_rng(True)`` mode."""
    global _gen
    _gen = torch.Generator().manual_seed(int(s))
    if torch.cuda.is_available():
        torch.cuda.manual_seed(int(s))


def set_device_rng(flag=True):
    """Draw the augmentation parameters on the device from torch's default CUDA generator instead of
    a CPU generator + host->device copies.  That generator is CUDA-graph safe (its Philox offset
    lives in device memory while capturing), so a training step captured by ``Trainer.capture``
    draws fresh parameters on every replay; the CPU path cannot be captured at all."""
    global _device_rng
    _device_rng = bool(flag)


def _uniform(n, lo, hi, device):
    if _device_rng and torch.device(device).type == "cuda":
        return torch.rand(n, device=device) * (hi - lo) + lo
    return (torch.rand(n, generator=_gen) * (hi - lo) + lo).to(device)


def _normal(n, std, device):
    if _device_rng and torch.device(device).type == "cuda":
        return torch.randn(n, device=device) * std
    return (torch.randn(n, generator=_gen) * std).to(device)


def transformer(U, theta, out_size):
    """spatial_transformer.transformer: U [B,H,W,C], theta [B,2,3] (or [B,6]) in normalised
    coordinates, out_size (height, width) -> [B,out_h,out_w,C]."""
    B, H, W, C = U.shape
    oh, ow = int(out_size[0]), int(out_size[1])
    theta = theta.reshape(B, 2, 3).to(U.device, torch.float32)
    x_t = torch.linspace(-1.0, 1.0, ow, device=U.device).view(1, ow).expand(oh, ow)
    y_t = torch.linspace(-1.0, 1.0, oh, device=U.device).view(oh, 1).expand(oh, ow)
    grid = torch.stack([x_t.reshape(-1), y_t.reshape(-1), torch.ones(oh * ow, device=U.device)], 0)
    T_g = torch.matmul(theta, grid)                     # [B,2,oh*ow]
    x = (T_g[:, 0] + 1.0) * float(W) / 2.0
    y = (T_g[:, 1] + 1.0) * float(H) / 2.0
    coords = torch.stack([x, y], 2).view(B, oh, ow, 2).contiguous()
    U = U.contiguous().float()
    if (oh, ow) != (H, W):
        raise NotImplementedError("transformer: out_size must equal the input size on this path")
    out = torch.empty_like(U)
    with torch.cuda.device(U.device):
        check(_native.lib().unflow_backward_warp_fwd(U.data_ptr(), coords.data_ptr(), out.data_ptr(),
                                                     B, H, W, C, BORDER_STN, stream()), "transformer")
    return out


def affine_matrices(tx, ty, rot_deg, scale, flip=None):
    """The 2x3 matrices random_affine builds from its draws (augment.py:30-48): t1 (rotation +
    translation) @ t2 (anisotropic scale with optional horizontal flip)."""
    rad = rot_deg * math.pi / 180.0
    B = tx.shape[0]
    zero, one = torch.zeros_like(tx), torch.ones_like(tx)
    t1 = torch.stack([torch.stack([torch.cos(rad), -torch.sin(rad), tx], 1),
                      torch.stack([torch.sin(rad), torch.cos(rad), ty], 1)], 1)      # [B,2,3]
    scale_x = scale if flip is None else scale * flip
    t2 = torch.stack([torch.stack([scale_x, zero, zero], 1),
                      torch.stack([zero, scale, zero], 1),
                      torch.stack([zero, zero, one], 1)], 1)                          # [B,3,3]
    return torch.matmul(t1, t2)


def random_affine(tensors, *,
                  max_translation_x=0.0, max_translation_y=0.0,
                  max_rotation=0.0, min_scale=1.0, max_scale=1.0,
                  horizontal_flipping=False):
    """Applies geometric augmentations to a list of tensors.

    Each element in the list is augmented in the same way.
    For all elements, num_batch must be equal while height, width and channels
    may differ."""
    dev = tensors[0].device
    B = tensors[0].shape[0]
    tx = _uniform(B, -max_translation_x, max_translation_x, dev)
    ty = _uniform(B, -max_translation_y, max_translation_y, dev)
    rot = _uniform(B, -max_rotation, max_rotation, dev)
    scale = _uniform(B, min_scale, max_scale, dev)
    flip = None
    if horizontal_flipping:
        f = _uniform(B, 0, 1, dev)
        flip = torch.where(f > 0.5, -torch.ones_like(f), torch.ones_like(f))
    t = affine_matrices(tx, ty, rot, scale, flip)
    return [transformer(x, t, (x.shape[1], x.shape[2])).detach() for x in tensors]


def photometric(ims, contrast, gamma, colour, noise, brightness):
    """The deterministic part of random_photometric (augment.py:91-106) for drawn parameters
    contrast/gamma/noise/brightness [B,1] and colour [B,3]."""
    gamma_inv = 1.0 / gamma
    out = []
    for im in ims:
        c = (contrast + 1.0).view(-1, 1, 1, 1)
        x = (im * c + brightness.view(-1, 1, 1, 1)) * colour.view(-1, 1, 1, 3)
        x = torch.clamp(x, 0.0, 1.0)
        x = torch.pow(x, gamma_inv.view(-1, 1, 1, 1))
        x = x + noise.view(-1, 1, 1, 1)
        out.append(x.detach())
    return out


def random_photometric(ims, *,
                       noise_stddev=0.0, min_contrast=0.0, max_contrast=0.0,
                       brightness_stddev=0.0, min_colour=1.0, max_colour=1.0,
                       min_gamma=1.0, max_gamma=1.0):
    """Applies photometric augmentations to a list of image batches (values in [0, 1]).

    Each image in the list is augmented in the same way."""
    dev = ims[0].device
    B = ims[0].shape[0]
    contrast = _uniform(B, min_contrast, max_contrast, dev).view(B, 1)
    gamma = _uniform(B, min_gamma, max_gamma, dev).view(B, 1)
    colour = _uniform(B * 3, min_colour, max_colour, dev).view(B, 3)
    noise = _normal(B, noise_stddev, dev).view(B, 1) if noise_stddev > 0.0 else torch.zeros(B, 1, device=dev)
    brightness = (_normal(B, brightness_stddev, dev).view(B, 1) if brightness_stddev > 0.0
                  else torch.zeros(B, 1, device=dev))
    return photometric(ims, contrast, gamma, colour, noise, brightness)


def random_crop(tensors, size, seed=None, name=None):
    """Randomly crops multiple tensors (of the same shape) to a given size.

    Each tensor is cropped in the same way (augment.py:111-134)."""
    shape = list(tensors[0].shape)
    if len(tensors) == 2:
        shape = [min(a, b) for a, b in zip(tensors[0].shape, tensors[1].shape)]
    g = _gen if seed is None else torch.Generator().manual_seed(int(seed))
    offset = [int(torch.randint(0, 2 ** 31 - 1, (1,), generator=g)) % (s - z + 1) for s, z in zip(shape, size)]
    return [t[tuple(slice(o, o + z) for o, z in zip(offset, size))] for t in tensors]
