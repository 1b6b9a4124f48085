"""CPU restatement of the unsupervised loss assembly (TEST INFRASTRUCTURE ONLY).

Follows /root/reference/src/e2eflow/core/unsupervised.py:27-164 with
``augment=False`` (the random augmentation cannot be parity-pinned and is out
of scope).  PARITY UNPINNED (no reference test).
"""
import torch

from . import tf_compat as tfc
from .flownet import flownet, FLOW_SCALE
from .losses import compute_losses, create_border_mask
from .util import downsample

LOSSES = ['occ', 'sym', 'fb', 'grad', 'ternary', 'photo', 'smooth_1st', 'smooth_2nd']  # :15


def regularization_loss(variables, scale=0.0004):
    """slim.l2_regularizer(0.0004) on every conv / deconv ``weights`` variable
    (flownet.py:176,200,218): scale * sum(w^2) / 2, summed over variables."""
    total = torch.zeros(())
    for name, v in variables.items():
        if name.endswith('/weights'):
            total = total + scale * (torch.sum(torch.as_tensor(v).float() ** 2) / 2)
    return total


def unsupervised_loss(variables, batch, params, normalization=None, augment=False,
                      return_flow=False, return_terms=False):
    if augment:
        raise NotImplementedError("the oracle restates the augment=False graph only")
    channel_mean = torch.tensor(normalization[0], dtype=torch.float32) / 255.0
    im1, im2 = batch
    im1 = im1 / 255.0
    im2 = im2 / 255.0
    im_shape = im1.shape[1:3]

    border_mask = create_border_mask(im1, 0.1)
    im1_norm, im2_norm = im1, im2
    im1_photo = im1 - channel_mean
    im2_photo = im2 - channel_mean

    flownet_spec = params.get('flownet', 'S')
    full_resolution = params.get('full_res')
    train_all = params.get('train_all')

    flows_fw, flows_bw = flownet(variables, im1_photo, im2_photo, flownet_spec=flownet_spec,
                                 full_resolution=full_resolution, backward_flow=True,
                                 train_all=train_all)
    flows_fw = flows_fw[-1]
    flows_bw = flows_bw[-1]

    layer_weights = [12.7, 4.35, 3.9, 3.4, 1.1]
    layer_patch_distances = [3, 2, 2, 1, 1]
    if full_resolution:
        layer_weights = [12.7, 5.5, 5.0, 4.35, 3.9, 3.4, 1.1]
        layer_patch_distances = [3, 3] + layer_patch_distances
        im1_s, im2_s, mask_s = im1_norm, im2_norm, border_mask
        final_flow_scale = FLOW_SCALE * 4
        final_flow_fw = flows_fw[0] * final_flow_scale
        final_flow_bw = flows_bw[0] * final_flow_scale
    else:
        im1_s = downsample(im1_norm, 4)
        im2_s = downsample(im2_norm, 4)
        mask_s = downsample(border_mask, 4)
        final_flow_scale = FLOW_SCALE
        final_flow_fw = tfc.resize_bilinear_legacy(flows_fw[0], im_shape) * final_flow_scale * 4
        final_flow_bw = tfc.resize_bilinear_legacy(flows_bw[0], im_shape) * final_flow_scale * 4

    combined_losses = {loss: 0.0 for loss in LOSSES}
    combined_loss = 0.0

    if params.get('pyramid_loss'):
        flow_enum = list(enumerate(zip(flows_fw, flows_bw)))
    else:
        flow_enum = [(0, (flows_fw[0], flows_bw[0]))]

    for i, (flow_fw_s, flow_bw_s) in flow_enum:
        flow_scale = final_flow_scale / (2 ** i)
        layer_weight = layer_weights[i]
        mask_occlusion = params.get('mask_occlusion', '')
        assert mask_occlusion in ['fb', 'disocc', '']
        losses = compute_losses(im1_s, im2_s, flow_fw_s * flow_scale, flow_bw_s * flow_scale,
                                border_mask=mask_s if params.get('border_mask') else None,
                                mask_occlusion=mask_occlusion,
                                data_max_distance=layer_patch_distances[i])
        layer_loss = 0.0
        for loss in LOSSES:
            weight_name = loss + '_weight'
            if params.get(weight_name):
                layer_loss = layer_loss + params[weight_name] * losses[loss]
                combined_losses[loss] = combined_losses[loss] + layer_weight * losses[loss]
        combined_loss = combined_loss + layer_weight * layer_loss

        im1_s = downsample(im1_s, 2)
        im2_s = downsample(im2_s, 2)
        mask_s = downsample(mask_s, 2)

    final_loss = combined_loss + regularization_loss(variables)

    out = (final_loss,)
    if return_flow:
        out = out + (final_flow_fw, final_flow_bw)
    if return_terms:
        out = out + (combined_losses,)
    return out[0] if len(out) == 1 else out
