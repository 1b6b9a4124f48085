"""Host-side input pipeline: the reference's file pairing / ordering rules
(src/e2eflow/core/input.py, kitti/input.py, kitti/data.py) on a miniature KITTI tree."""
import os
import random

import numpy as np
import pytest
import torch

cv2 = pytest.importorskip("cv2")

from unflow_b200.e2eflow.core import flow_io
from unflow_b200.e2eflow.core import input as inp
from unflow_b200.e2eflow.kitti.data import KITTIData
from unflow_b200.e2eflow.kitti.input import KITTIInput


def _frame(h, w, tag):
    """RGB image whose red/green channels encode the position and blue the frame tag."""
    yy, xx = np.meshgrid(np.arange(h), np.arange(w), indexing='ij')
    return np.stack([yy % 256, xx % 256, np.full((h, w), tag)], 2).astype(np.uint8)


def _write_rgb(path, rgb):
    os.makedirs(os.path.dirname(path), exist_ok=True)
    assert cv2.imwrite(path, np.ascontiguousarray(rgb[:, :, ::-1]))


@pytest.fixture
def tree(tmp_path):
    root = str(tmp_path)
    frames = {('2011_09_26', '2011_09_26_drive_0001_extract', 'image_02'): [0, 1, 2, 3, 5, 6],   # 4 missing
              ('2011_09_26', '2011_09_26_drive_0001_extract', 'image_03'): [0, 1, 2],
              ('2011_09_28', '2011_09_28_drive_0002_extract', 'image_02'): [10, 11],
              ('2011_09_28', '2011_09_28_drive_0002_extract', 'image_03'): [10, 11]}
    for (date, drive, view), nums in frames.items():
        for n in nums:
            _write_rgb(os.path.join(root, 'kitti_raw', date, drive, view, 'data', '%010d.png' % n), _frame(20, 30, n))
    for i in range(3):
        for j in (10, 11):
            _write_rgb(os.path.join(root, 'data_stereo_flow/training/colored_0', '%06d_%d.png' % (i, j)),
                       _frame(12, 18, 10 * i + j - 10))
        for sub, val in (('flow_occ', 1.0), ('flow_noc', 0.0)):
            flow = np.zeros((12, 18, 2), np.float32)
            flow[..., 0], flow[..., 1] = i + 0.5, -2.0 * i
            mask = np.full((12, 18), val)
            mask[0, 0] = 1.0
            os.makedirs(os.path.join(root, 'data_stereo_flow/training', sub), exist_ok=True)
            flow_io.write_kitti_flow(os.path.join(root, 'data_stereo_flow/training', sub, '%06d_10.png' % i), flow, mask)
    return root


def test_frame_numbers_and_crop_or_pad():
    assert inp.frame_name_to_num('0000000012.png') == 12 and inp.frame_name_to_num('0000000000.png') == 0
    t = torch.arange(5 * 6, dtype=torch.float32).reshape(5, 6, 1)
    c = inp.resize_image_with_crop_or_pad(t, 3, 2)                # centre crop: rows 1..3, cols 2..3
    assert torch.equal(c[..., 0], t[1:4, 2:4, 0])
    p = inp.resize_image_with_crop_or_pad(t, 8, 9)                # zero pad: 1 row above / 2 below, 1 col left / 2 right
    assert p.shape == (8, 9, 1) and torch.equal(p[1:6, 1:7], t) and float(p.sum()) == float(t.sum())
    m = inp.resize_image_with_crop_or_pad(t.unsqueeze(0), 7, 4)   # pad rows, crop cols, batched
    assert m.shape == (1, 7, 4, 1) and torch.equal(m[0, 1:6, :, 0], t[:, 1:5, 0])
    assert torch.equal(inp.resize_output_crop(t.unsqueeze(0), 3, 2, 1), c.unsqueeze(0))
    f = torch.ones(1, 4, 6, 2)
    r = inp.resize_output_flow(f, 8, 3)                           # u scales with the width, v with the height
    assert r.shape == (1, 8, 3, 2) and torch.allclose(r[..., 0], torch.full((1, 8, 3), 0.5)) \
        and torch.allclose(r[..., 1], torch.full((1, 8, 3), 2.0))
    assert inp.resize_output(torch.zeros(1, 4, 6, 3), 8, 12, 3).shape == (1, 8, 12, 3)


def test_kitti_raw_dirs_and_pairs(tree):
    data = KITTIData(tree)
    dirs = data.get_raw_dirs()
    assert len(dirs) == 4 and all(d.endswith('/data') for d in dirs)
    with pytest.raises(FileNotFoundError):
        KITTIData(os.path.join(tree, 'nope'))
    ki = KITTIInput(data, batch_size=2, dims=(16, 24), normalize=False, skipped_frames=True)
    pairs = ki.raw_pairs(swap_images=False, shift=0, seed=0)
    # consecutive frames only: 0-1,1-2,2-3,5-6 | 0-1,1-2 | 10-11 | 10-11  (3-5 is dropped)
    nums = sorted((os.path.basename(a), os.path.basename(b)) for a, b in pairs)
    assert len(pairs) == 8 and ('0000000003.png', '0000000005.png') not in nums
    assert all(inp.frame_name_to_num(os.path.basename(a)) + 1 == inp.frame_name_to_num(os.path.basename(b))
               and os.path.dirname(a) == os.path.dirname(b) for a, b in pairs)
    # the order is the reference's: python's random.seed(seed); random.shuffle on the collected list
    base = []
    for d in dirs:
        files = sorted(os.listdir(d))
        for i in range(len(files) - 1):
            if inp.frame_name_to_num(files[i]) + 1 == inp.frame_name_to_num(files[i + 1]):
                base.append((os.path.join(d, files[i]), os.path.join(d, files[i + 1])))
    random.seed(0)
    random.shuffle(base)
    assert pairs == base
    # without skipped_frames the 3-5 pair stays; swap doubles the list; shift rolls the FLATTENED names
    assert len(KITTIInput(data, 2, (16, 24), normalize=False).raw_pairs(swap_images=False)) == 9
    sw = ki.raw_pairs(swap_images=True)
    assert len(sw) == 16 and sw[0] == base[0] and sw[1] == base[0][::-1]
    flat = [x for p in base for x in p]
    rolled = ki.raw_pairs(swap_images=False, shift=3)
    assert [x for p in rolled for x in p] == flat[-3:] + flat[:-3]
    assert ki.raw_pairs(swap_images=False, shift=2) == base[-1:] + base[:-1]
    # skip=[0, 1]: the second pass strides by 2 up to len-2 but still pairs ADJACENT files
    # (core/input.py:148-163 takes files[i], files[i+1]); 6, 3, 2, 2 files -> 2 + 1 + 0 + 0 more
    k2 = KITTIInput(data, 2, (16, 24), normalize=False)
    assert len(k2.raw_pairs(swap_images=False, skip=[0, 1])) == 9 + 3


def test_raw_batches_crop_both_frames_alike_and_shard_by_rank(tree):
    data = KITTIData(tree)
    ki = KITTIInput(data, batch_size=2, dims=(16, 24), normalize=False, skipped_frames=True)
    pairs = ki.raw_pairs(swap_images=False)
    it = ki.input_raw(swap_images=False, center_crop=True, pin=False)
    batches = [next(it) for _ in range(5)]          # 8 pairs -> wraps around after 4 batches
    it.close()
    for bi, (a, b) in enumerate(batches):
        assert a.shape == (2, 16, 24, 3) and a.dtype == torch.float32
        for k in range(2):
            fn1, fn2 = pairs[(2 * bi + k) % len(pairs)]
            assert float(a[k, 0, 0, 2]) == inp.frame_name_to_num(os.path.basename(fn1))     # blue = frame tag
            assert float(b[k, 0, 0, 2]) == inp.frame_name_to_num(os.path.basename(fn2))
            assert torch.equal(a[k, ..., :2], b[k, ..., :2])          # same crop window in both frames
            oy, ox = int(a[k, 0, 0, 0]), int(a[k, 0, 0, 1])
            assert 0 <= oy <= 4 and 0 <= ox <= 6
            assert float(a[k, 15, 23, 0]) == oy + 15 and float(a[k, 15, 23, 1]) == ox + 23
    assert torch.equal(batches[4][0][..., 2], batches[0][0][..., 2])  # epoch wrap: same files again
    # rank r of 2 sees batches r, r+2, ...
    r1 = ki.input_raw(swap_images=False, rank=1, world_size=2, pin=False)
    x = next(r1)
    y = next(r1)
    r1.close()
    assert torch.equal(x[0], batches[1][0]) and torch.equal(y[0], batches[3][0])
    # normalisation (core/input.py:69-70)
    kn = KITTIInput(data, batch_size=1, dims=(16, 24), normalize=True, skipped_frames=True)
    itn = kn.input_raw(swap_images=False, pin=False)
    a, _ = next(itn)
    itn.close()
    want = (batches[0][0][0] - torch.tensor(kn.mean)) / kn.stddev
    assert torch.allclose(a[0], want, atol=1e-6)
    assert kn.get_normalization() == ([104.920005, 110.1753, 114.785955], 1 / 0.0039216)


def test_eval_inputs_2012(tree):
    data = KITTIData(tree, require=('data_stereo_flow',))
    ki = KITTIInput(data, batch_size=1, dims=(16, 24), normalize=False)
    items = list(ki.input_train_2012())
    assert len(items) == 3
    for i, (im1, im2, shape, focc, mocc, fnoc, mnoc) in enumerate(items):
        assert im1.shape == (1, 16, 24, 3) and shape.tolist() == [[12, 18, 3]]
        assert focc.shape == (1, 16, 24, 2) and mocc.shape == (1, 16, 24, 1)
        # padded by (2,2) rows and (3,3) cols; undoing it (resize_input's first step) restores the file
        inner = inp.resize_image_with_crop_or_pad(im1[0], 12, 18)
        assert float(inner[0, 0, 2]) == 10 * i and float(im2[0, 2, 3, 2]) == 10 * i + 1
        f = inp.resize_image_with_crop_or_pad(focc[0], 12, 18)
        assert torch.allclose(f[..., 0], torch.full((12, 18), i + 0.5)) and torch.allclose(f[..., 1], torch.full((12, 18), -2.0 * i))
        assert float(mocc.sum()) == 12 * 18 and float(mnoc.sum()) == 1.0
    held = list(ki.input_train_2012(hold_out_inv=2))
    names = sorted(os.listdir(os.path.join(tree, 'data_stereo_flow/training/flow_occ')))
    random.seed(0)
    random.shuffle(names)
    assert len(held) == 2
    # image pairs and flow files are shuffled with the same seed over equally long lists -> stay aligned
    for item, name in zip(held, names[:2]):
        i = int(name.split('_')[0])
        assert float(inp.resize_image_with_crop_or_pad(item[0][0], 12, 18)[0, 0, 2]) == 10 * i
        assert abs(float(inp.resize_image_with_crop_or_pad(item[3][0], 12, 18)[0, 0, 0]) - (i + 0.5)) < 1e-6
    r = inp.resize_input(items[1][0].reshape(-1), 12, 18, 16, 24)
    assert r.shape == (1, 16, 24, 3)
