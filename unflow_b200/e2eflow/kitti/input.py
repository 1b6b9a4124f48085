"""KITTI evaluation inputs (reference src/e2eflow/kitti/input.py): image pairs with the occluded /
non-occluded ground-truth flow of the 2012 and 2015 training sets."""
import os
import random

import torch

from ..core.flow_io import read_kitti_flow
from ..core.input import Input, read_png_image, resize_image_with_crop_or_pad


class KITTIInput(Input):
    def _flow_files(self, flow_dir, hold_out_inv):
        """kitti/input.py:35-70: sorted flow_occ / flow_noc files; ``hold_out_inv`` keeps the first k
        of a ``random.seed(0)`` shuffle of each list (the same permutation, the lists being equally
        long)."""
        out = []
        for sub in ('flow_occ', 'flow_noc'):
            d = os.path.join(self.data.current_dir, flow_dir, sub)
            files = os.listdir(d)
            files.sort()
            if hold_out_inv is not None:
                random.seed(0)
                random.shuffle(files)
                files = files[:hold_out_inv]
            out.append([os.path.join(d, f) for f in files])
        assert len(out[0]) == len(out[1])
        return out

    def _input_train(self, image_dir, flow_dir, hold_out_inv=None):
        """One pass, batch 1: ``(im1, im2, input_shape, flow_occ, mask_occ, flow_noc, mask_noc)``,
        everything cropped / padded to ``dims`` like the reference's queues deliver it."""
        height, width = self.dims
        occ, noc = self._flow_files(flow_dir, hold_out_inv)
        for (fn1, fn2), f_occ, f_noc in zip(self.image_pairs(image_dir, hold_out_inv), occ, noc):
            raw1, raw2 = read_png_image(fn1), read_png_image(fn2)
            item = [self._preprocess_image(raw1).unsqueeze(0), self._preprocess_image(raw2).unsqueeze(0),
                    torch.tensor(raw1.shape).unsqueeze(0)]
            for path in (f_occ, f_noc):
                flow, mask = read_kitti_flow(path)
                flow, mask = torch.as_tensor(flow).float(), torch.as_tensor(mask).float()
                if mask.dim() == 2:
                    mask = mask.unsqueeze(-1)
                item += [resize_image_with_crop_or_pad(flow, height, width).unsqueeze(0),
                         resize_image_with_crop_or_pad(mask, height, width).unsqueeze(0)]
            yield tuple(item)

    def input_train_2015(self, hold_out_inv=None):
        return self._input_train('data_scene_flow/training/image_2', 'data_scene_flow/training', hold_out_inv)

    def input_test_2015(self, hold_out_inv=None):
        return self._input_test('data_scene_flow/testing/image_2', hold_out_inv)

    def input_train_2012(self, hold_out_inv=None):
        return self._input_train('data_stereo_flow/training/colored_0', 'data_stereo_flow/training', hold_out_inv)

    def input_test_2012(self, hold_out_inv=None):
        return self._input_test('data_stereo_flow/testing/colored_0', hold_out_inv)
