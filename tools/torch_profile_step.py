#!/usr/bin/env python
"""torch.profiler view of one training step: which aten ops own the GPU time (complements the ncu
launch list, which only has kernel names)."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from unflow_b200 import synthetic as synth  # noqa: E402
from unflow_b200.e2eflow.core import conv_ops  # noqa: E402
from unflow_b200.e2eflow.core.train import Trainer  # noqa: E402

conv_ops.set_mode(os.environ.get("UNFLOW_CONV_PRECISION", "3xtf32"))
dev = torch.device("cuda", 0)
tr = Trainer(dict(synth.KITTI_PARAMS, learning_rate=1e-5), synth.KITTI_NORMALIZATION, dev, seed=1)
im1, im2, _ = synth.image_pair(4, 384, 1280, seed=1)
im1, im2 = im1.to(dev), im2.to(dev)
for _ in range(3):
    tr.step(im1, im2)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
    tr.step(im1, im2)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="self_cuda_time_total", row_limit=45, max_name_column_width=60))
print(prof.key_averages(group_by_input_shape=True).table(sort_by="self_cuda_time_total", row_limit=40,
                                                         max_name_column_width=50, max_shapes_column_width=90))
