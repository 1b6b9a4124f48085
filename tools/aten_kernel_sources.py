#!/usr/bin/env python
"""Where do the library (ATen) kernels of one warm training step come from?  Groups the step's operators by
(name, input shapes, python stack inside this repository) and prints those with device time, largest first.
Used to hunt layout copies / gradient adds around the hand-written kernels."""
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
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True, with_stack=True) as prof:
    tr.step(im1, im2)
    torch.cuda.synchronize()

rows = []
for e in prof.key_averages(group_by_input_shape=True, group_by_stack_n=30):
    t = getattr(e, "self_device_time_total", 0.0)
    if t < 5.0 or not e.key.startswith("aten::"):
        continue
    stack = [f for f in (e.stack or []) if "unflow_b200" in f]
    where = " <- ".join(os.path.basename(f.split("(")[0].strip()) + ":" + f.split("(")[1].split(")")[0] + " " + f.split(":")[-1].strip()
                        if "(" in f else f for f in stack[:4])
    rows.append((t, e.count, e.key, str(e.input_shapes)[:90], where))
rows.sort(reverse=True)
print("| device us | calls | op | input shapes | python frames in unflow_b200 |\n|---:|---:|---|---|---|")
for t, n, k, sh, w in rows[:60]:
    print("| %.1f | %d | %s | %s | %s |" % (t, n, k, sh, w))
