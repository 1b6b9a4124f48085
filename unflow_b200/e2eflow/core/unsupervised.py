"""The unsupervised training objective -- reference
/root/reference/src/e2eflow/core/unsupervised.py:27-164 with the same signature:

    unsupervised_loss(batch, params, normalization=None, augment=True, return_flow=False)

``params`` holds the keys of config.ini [train] (+ the dataset section), exactly as the
reference reads them (flownet, full_res, train_all, pyramid_loss, mask_occlusion, border_mask,
<loss>_weight).  ``augment=True`` applies the reference's random affine / photometric augmentation
(core/augment.py) with torch's random stream in place of TF's.

Extra keyword ``variables`` (FlowNetVariables) replaces TF's implicit graph variable store; the
default is the module-level store of core.flownet.
"""
import torch

from .util import downsample
from .losses import compute_losses, create_border_mask
from .flownet import flownet, FLOW_SCALE, get_variables
from . import tf_image

# REGISTER ALL POSSIBLE LOSS TERMS
LOSSES = ['occ', 'sym', 'fb', 'grad', 'ternary', 'photo', 'smooth_1st', 'smooth_2nd']

# TF collects tracked scalars in graph collections (_track_loss, reference :18-19); the eager
# equivalent is a dict refreshed on every call.
tracked = {}

_const_cache = {}


def _device_constant(values, device):
    """Small host constants as cached device tensors (no host->device copy inside the step: the
    step must be capturable in a CUDA graph)."""
    key = (tuple(float(v) for v in values), str(device))
    if key not in _const_cache:
        _const_cache[key] = torch.tensor(list(values), device=device, dtype=torch.float32)
    return _const_cache[key]


def _track_loss(op, name):
    tracked[name] = op.detach() if torch.is_tensor(op) else op


def unsupervised_loss(batch, params, normalization=None, augment=True,
                      return_flow=False, variables=None):
    im1, im2 = batch
    channel_mean = _device_constant([v / 255.0 for v in normalization[0]], im1.device)
    im1 = im1 / 255.0
    im2 = im2 / 255.0
    im_shape = im1.shape[1:3]

    # -------------------------------------------------------------------------
    # Data & mask augmentation
    border_mask = create_border_mask(im1, 0.1)

    if augment:
        from .augment import random_affine, random_photometric
        im1_geo, im2_geo, border_mask_global = random_affine(
            [im1, im2, border_mask.contiguous()],
            horizontal_flipping=True,
            min_scale=0.9, max_scale=1.1
            )

        # augment locally
        im2_geo, border_mask_local = random_affine(
            [im2_geo, border_mask.contiguous()],
            min_scale=0.9, max_scale=1.1
            )
        border_mask = border_mask_local * border_mask_global

        im1_photo, im2_photo = random_photometric(
            [im1_geo, im2_geo],
            noise_stddev=0.04, min_contrast=-0.3, max_contrast=0.3,
            brightness_stddev=0.02, min_colour=0.9, max_colour=1.1,
            min_gamma=0.7, max_gamma=1.5)
    else:
        im1_geo, im2_geo = im1, im2
        im1_photo, im2_photo = im1, im2

    # Images for loss comparisons with values in [0, 1] (scale to original using * 255)
    im1_norm = im1_geo
    im2_norm = im2_geo
    # Images for neural network input with mean-zero values in [-1, 1]
    im1_photo = im1_photo - channel_mean
    im2_photo = im2_photo - channel_mean

    flownet_spec = params.get('flownet', 'S')
    full_resolution = params.get('full_res')
    train_all = params.get('train_all')

    if variables is None:
        variables = get_variables(flownet_spec, full_resolution, device=im1.device)

    flows_fw, flows_bw = flownet(im1_photo, im2_photo,
                                 flownet_spec=flownet_spec,
                                 full_resolution=full_resolution,
                                 backward_flow=True,
                                 train_all=train_all,
                                 variables=variables)

    flows_fw = flows_fw[-1]
    flows_bw = flows_bw[-1]

    # -------------------------------------------------------------------------
    # Losses
    layer_weights = [12.7, 4.35, 3.9, 3.4, 1.1]
    layer_patch_distances = [3, 2, 2, 1, 1]
    if full_resolution:
        layer_weights = [12.7, 5.5, 5.0, 4.35, 3.9, 3.4, 1.1]
        layer_patch_distances = [3, 3] + layer_patch_distances
        im1_s = im1_norm
        im2_s = im2_norm
        mask_s = border_mask
        final_flow_scale = FLOW_SCALE * 4
        final_flow_fw = flows_fw[0] * final_flow_scale
        final_flow_bw = flows_bw[0] * final_flow_scale
    else:
        im1_s = downsample(im1_norm, 4)
        im2_s = downsample(im2_norm, 4)
        mask_s = downsample(border_mask.contiguous(), 4)
        final_flow_scale = FLOW_SCALE
        final_flow_fw = tf_image.resize_bilinear(flows_fw[0], im_shape) * final_flow_scale * 4
        final_flow_bw = tf_image.resize_bilinear(flows_bw[0], im_shape) * final_flow_scale * 4

    combined_losses = dict()
    combined_loss = 0.0
    for loss in LOSSES:
        combined_losses[loss] = 0.0

    if params.get('pyramid_loss'):
        flow_enum = list(enumerate(zip(flows_fw, flows_bw)))
    else:
        flow_enum = [(0, (flows_fw[0], flows_bw[0]))]

    # graph mode evaluates only the terms that have a weight (reference :136-141)
    active = [loss for loss in LOSSES if params.get(loss + '_weight')]

    for i, flow_pair in flow_enum:
        flow_scale = final_flow_scale / (2 ** i)

        layer_weight = layer_weights[i]
        flow_fw_s, flow_bw_s = flow_pair

        mask_occlusion = params.get('mask_occlusion', '')
        assert mask_occlusion in ['fb', 'disocc', '']

        losses = compute_losses(im1_s, im2_s,
                                flow_fw_s * flow_scale, flow_bw_s * flow_scale,
                                border_mask=mask_s if params.get('border_mask') else None,
                                mask_occlusion=mask_occlusion,
                                data_max_distance=layer_patch_distances[i],
                                _terms=active)

        layer_loss = 0.0

        for loss in active:
            weight_name = loss + '_weight'
            _track_loss(losses[loss], loss)
            layer_loss = layer_loss + params[weight_name] * losses[loss]
            combined_losses[loss] = combined_losses[loss] + layer_weight * losses[loss]

        combined_loss = combined_loss + layer_weight * layer_loss

        if i + 1 < len(flow_enum):  # the reference builds (and TF prunes) one more level
            im1_s = downsample(im1_s, 2)
            im2_s = downsample(im2_s, 2)
            mask_s = downsample(mask_s, 2)

    regularization_loss = variables.regularization_loss()
    final_loss = combined_loss + regularization_loss

    _track_loss(final_loss, 'loss/combined')
    for loss in LOSSES:
        _track_loss(combined_losses[loss], 'loss/' + loss)

    if not return_flow:
        return final_loss

    return final_loss, final_flow_fw, final_flow_bw
