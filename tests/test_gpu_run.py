"""python -m unflow_b200.run: config.ini semantics, checkpoint cadence and resume-by-iteration
(reference src/run.py, src/e2eflow/core/train.py:116-145,258-259) on a short synthetic run."""
import glob
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

CFG = """
[dirs]
log = {d}/log
checkpoints = {d}/log/checkpoints
data = {d}/nodata
[run]
batch_size = 2
gpu_list = 0
dataset = kitti
development = False
[train]
decay_interval = 100000
save_interval = 2
display_interval = 1
flownet = C
pyramid_loss = True
border_mask = True
ternary_weight = 1.0
smooth_2nd_weight = 3.0
[train_kitti]
height = 128
width = 256
num_iters = 6
learning_rate = 1.0e-5
decay_after = 100000
fb_weight = 0.2
mask_occlusion = fb
occ_weight = 12.4
"""


def test_run_trains_checkpoints_and_resumes(tmp_path, capsys):
    from unflow_b200 import run as R
    ini = tmp_path / "config.ini"
    ini.write_text(CFG.format(d=str(tmp_path)))
    R.main(["--ex", "t1", "--config", str(ini), "--synthetic", "--max-iters", "4"])
    out = capsys.readouterr().out
    assert "-- training from i = 1 to 4" in out and "-- train: i = 4, loss" in out
    ck = sorted(glob.glob(str(tmp_path / "log" / "checkpoints" / "t1" / "model.ckpt-*.pt")))
    assert [os.path.basename(c) for c in ck] == ["model.ckpt-2.pt", "model.ckpt-4.pt"]
    state = torch.load(ck[-1])
    assert "flownet_c/conv3_1/weights" in state["variables"]
    assert tuple(state["variables"]["flownet_c_features/conv1/weights"].shape) == (7, 7, 3, 64)  # HWIO
    # resume: parses the iteration from the file name and continues at 5
    R.main(["--ex", "t1", "--config", str(ini), "--synthetic"])
    out = capsys.readouterr().out
    assert "-- training from i = 5 to 6" in out
    assert os.path.exists(str(tmp_path / "log" / "checkpoints" / "t1" / "model.ckpt-6.pt"))
    # finished experiment: nothing left to do
    R.main(["--ex", "t1", "--config", str(ini), "--synthetic"])
    assert "max_iter reached" in capsys.readouterr().out
    # --ow starts over, --debug writes no checkpoints
    R.main(["--ex", "t1", "--config", str(ini), "--synthetic", "--ow", "--debug", "--max-iters", "2"])
    assert "-- training from i = 1 to 2" in capsys.readouterr().out
    assert not glob.glob(str(tmp_path / "log" / "checkpoints" / "t1" / "model.ckpt-*.pt"))
    # the reference's own checkpoint format: write, then resume from it
    R.main(["--ex", "t2", "--config", str(ini), "--synthetic", "--max-iters", "2", "--ckpt-format", "tf"])
    capsys.readouterr()
    d2 = tmp_path / "log" / "checkpoints" / "t2"
    assert sorted(os.listdir(d2)) == ["checkpoint", "model.ckpt-2.data-00000-of-00001", "model.ckpt-2.index"]
    R.main(["--ex", "t2", "--config", str(ini), "--synthetic", "--max-iters", "3", "--debug"])
    assert "-- training from i = 3 to 3" in capsys.readouterr().out


def test_evaluate_loop_on_synthetic_ground_truth():
    """Row N1: Trainer.eval counterpart -- AEE / outlier averages against a known flow."""
    from unflow_b200 import synthetic
    from unflow_b200.e2eflow.core.flownet import FlowNetVariables
    from unflow_b200.e2eflow.core.train import evaluate
    v = FlowNetVariables('C', seed=2).cuda()
    ex = []
    for s in (1, 2):
        im1, im2, flow = synthetic.image_pair(1, 96, 320, seed=s, max_flow=4.0)
        mask = torch.ones(1, 96, 320, 1)
        noc = (torch.rand(1, 96, 320, 1, generator=torch.Generator().manual_seed(s)) > 0.2).float()
        ex.append(tuple(t.cuda() for t in (im1, im2, flow, mask, flow, noc)))
    res, images = evaluate(v, dict(synthetic.KITTI_PARAMS), synthetic.KITTI_NORMALIZATION, ex, eval_size=(128, 384))
    assert res['num_examples'] == 2
    # untrained network: the error is of the order of the flow magnitudes involved, finite, > 0
    assert 0.0 < res['AEE/occluded'] < 100.0 and 0.0 < res['AEE/non-occluded'] < 100.0
    assert 0 <= res['outliers/non-occluded'] <= 100
    assert images['flow'].shape == (1, 96, 320, 3) and images['reverse disocc'].dtype == torch.bool


def test_run_trains_on_a_kitti_tree_and_evaluates(tmp_path, capsys):
    """Real-data path of run.py: KITTI raw PNG pairs through the reference's pairing rules, resume
    shift, evaluation on the 2012 training set after each save_interval chunk."""
    cv2 = pytest.importorskip("cv2")
    import numpy as np
    from unflow_b200 import run as R
    from unflow_b200.e2eflow.core import flow_io
    data = tmp_path / "data"
    rng = np.random.default_rng(0)
    base = cv2.GaussianBlur(rng.integers(0, 255, (150, 300, 3), dtype=np.uint8), (0, 0), 3)
    for view in ("image_02", "image_03"):
        d = data / "kitti_raw" / "2011_09_26" / "2011_09_26_drive_0001_extract" / view / "data"
        d.mkdir(parents=True)
        for n in range(6):
            assert cv2.imwrite(str(d / ("%010d.png" % n)), np.roll(base, 2 * n, axis=1))
    tr_dir = data / "data_stereo_flow" / "training"
    for sub in ("colored_0", "flow_occ", "flow_noc"):
        (tr_dir / sub).mkdir(parents=True)
    for i in range(2):
        im = np.roll(base, 5 * i, axis=0)[:120, :280]
        cv2.imwrite(str(tr_dir / "colored_0" / ("%06d_10.png" % i)), im)
        cv2.imwrite(str(tr_dir / "colored_0" / ("%06d_11.png" % i)), np.roll(im, 3, axis=1))
        flow = np.zeros((120, 280, 2), np.float32)
        flow[..., 0] = 3.0
        for sub in ("flow_occ", "flow_noc"):
            flow_io.write_kitti_flow(str(tr_dir / sub / ("%06d_10.png" % i)), flow, np.ones((120, 280)))
    ini = tmp_path / "config.ini"
    ini.write_text(CFG.format(d=str(tmp_path)).replace("data = %s/nodata" % tmp_path, "data = %s" % data))
    R.main(["--ex", "k1", "--config", str(ini), "--max-iters", "2"])
    out = capsys.readouterr().out
    assert "Training on 10 frame pairs." in out and "-- train: i = 2, loss" in out
    assert "-- eval: i = 2" in out and "AEE/occluded" in out and "num_examples = 2" in out
    assert os.path.exists(str(tmp_path / "log" / "checkpoints" / "k1" / "model.ckpt-2.pt"))
    R.main(["--ex", "k1", "--config", str(ini), "--max-iters", "4", "--debug"])
    assert "-- training from i = 3 to 4" in capsys.readouterr().out
