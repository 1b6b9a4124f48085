"""python -m unflow_b200.eval: per-example loop, metrics and benchmark files (reference
src/eval_gui.py) with a stub flow estimator on the CPU; the GPU path only swaps the estimator."""
import io
import os

import numpy as np
import pytest
import torch

cv2 = pytest.importorskip("cv2")

from unflow_b200 import eval as E
from unflow_b200.e2eflow.core import flow_io
from unflow_b200.e2eflow.kitti.data import KITTIData
from unflow_b200.e2eflow.kitti.input import KITTIInput


def test_flow_to_int16_clamps_and_truncates():
    f = np.zeros((1, 1, 4, 2), np.float32)
    f[0, 0, :, 0] = [0.0, 1.0 / 64 + 0.001, -600.0, 600.0]
    f[0, 0, :, 1] = [-0.01, 3.9 / 64, 0.5, -512.5]
    out = E.flow_to_int16(f)
    assert out.dtype == np.uint16 and out.shape == (1, 4, 3)
    assert out[0, :, 0].tolist() == [32768, 32769, 0, 65535]
    assert out[0, :, 1].tolist() == [32767, 32771, 32800, 0]       # truncation, not rounding
    assert out[0, :, 2].tolist() == [1, 1, 1, 1]


@pytest.fixture
def kitti2012(tmp_path):
    tr = tmp_path / "data_stereo_flow" / "training"
    for sub in ("colored_0", "flow_occ", "flow_noc"):
        (tr / sub).mkdir(parents=True)
    rng = np.random.default_rng(0)
    for i in range(3):
        im = rng.integers(0, 255, (20, 36, 3), dtype=np.uint8)
        cv2.imwrite(str(tr / "colored_0" / ("%06d_10.png" % i)), im)
        cv2.imwrite(str(tr / "colored_0" / ("%06d_11.png" % i)), im)
        flow = np.zeros((20, 36, 2), np.float32)
        flow[..., 0] = 2.0 + i
        noc = np.ones((20, 36))
        noc[:, :18] = 0
        flow_io.write_kitti_flow(str(tr / "flow_occ" / ("%06d_10.png" % i)), flow, np.ones((20, 36)))
        flow_io.write_kitti_flow(str(tr / "flow_noc" / ("%06d_10.png" % i)), flow, noc)
    return str(tmp_path)


def test_evaluate_examples_metrics_and_files(kitti2012, tmp_path):
    data = KITTIData(kitti2012, require=('data_stereo_flow',))
    ki = KITTIInput(data, batch_size=1, normalize=False, dims=(32, 48))
    seen = []

    def stub(im1, im2):          # a constant 3 px flow at the network size (32x48)
        assert im1.shape == (1, 32, 48, 3)
        seen.append(float(im1.mean()))
        fw = torch.zeros(1, 32, 48, 2)
        fw[..., 0] = 3.0 * 48 / 36               # resize_output_flow rescales u by w/old_w -> 3 px at file size
        return fw, -fw

    out_dir = str(tmp_path / "out")
    os.makedirs(out_dir)
    log = io.StringIO()
    avg = E.evaluate_examples("exA", ki.input_train_2012(), ki.dims, stub, torch.device("cpu"), num=-1,
                              out_dir=out_dir, output_benchmark=True, output_backward=True,
                              output_visual=True, log=log)
    # ground truth u = 2, 3, 4 -> endpoint errors 1, 0, 1; outliers need > 3 px AND > 5 %: none
    assert abs(avg['EPE_all'] - 2.0 / 3.0) < 1e-5 and abs(avg['EPE_noc'] - 2.0 / 3.0) < 1e-5
    assert avg['outliers_all'] == 0.0 and avg['outliers_noc'] == 0.0
    text = log.getvalue()
    assert "-- evaluating 'exA': 3/None" in text and "(exA) EPE_all = " in text
    files = sorted(os.listdir(out_dir))
    assert [f for f in files if f.endswith('_10.png')] == ['000000_10.png', '000001_10.png', '000002_10.png']
    assert '000001_01.png' in files and '000002_flow.png' in files and '000000_img.png' in files and '000000_err.png' in files
    flow, mask = flow_io.read_kitti_flow(os.path.join(out_dir, '000001_10.png'))
    assert flow.shape == (20, 36, 2) and np.allclose(flow[..., 0], 3.0, atol=1 / 64) and np.all(mask == 1)
    back, _ = flow_io.read_kitti_flow(os.path.join(out_dir, '000001_01.png'))
    assert np.allclose(back[..., 0], -3.0, atol=1 / 64)
    # --num limits the loop; .flo output
    out2 = str(tmp_path / "out2")
    os.makedirs(out2)
    avg2 = E.evaluate_examples("exA", ki.input_train_2012(), ki.dims, stub, torch.device("cpu"), num=2,
                               out_dir=out2, output_benchmark=True, output_png=False, log=io.StringIO())
    assert sorted(os.listdir(out2)) == ['000000_10.flo', '000001_10.flo'] and abs(avg2['EPE_all'] - 0.5) < 1e-5
    assert np.allclose(flow_io.read_flo(os.path.join(out2, '000000_10.flo'))[0][..., 0], 3.0, atol=1e-5)


def test_experiment_setup_prefers_log_ex_then_checkpoints(tmp_path):
    from unflow_b200.e2eflow.core import tf_checkpoint as ck
    from unflow_b200.e2eflow.core.flownet import FlowNetVariables
    log, ckd = tmp_path / "log", tmp_path / "ckpts"
    ini = tmp_path / "config.ini"
    ini.write_text("[dirs]\nlog = %s\ncheckpoints = %s\ndata = %s\n[run]\nbatch_size = 4\n[train]\nflownet = C\n"
                   "ternary_weight = 1.0\n[train_kitti]\nheight = 320\n" % (log, ckd, tmp_path))
    with pytest.raises(RuntimeError):
        E.experiment_setup("exA", str(ini), "kitti")
    v = FlowNetVariables("s", False, seed=1)
    ck.save_variables(v, str(ckd / "exA" / "model.ckpt-7"))
    params, ckpt, cfg = E.experiment_setup("exA", str(ini), "kitti")
    assert ckpt == (7, str(ckd / "exA" / "model.ckpt-7")) and cfg == str(ini)
    assert params['flownet'] == 'C' and params['height'] == 320
    # an experiment directory under log/ex with its own config.ini and checkpoint wins
    (log / "ex" / "exA").mkdir(parents=True)
    (log / "ex" / "exA" / "config.ini").write_text(ini.read_text().replace("flownet = C", "flownet = s"))
    ck.save_variables(v, str(log / "ex" / "exA" / "model.ckpt-9"))
    params, ckpt, cfg = E.experiment_setup("exA", str(ini), "kitti")
    assert ckpt[0] == 9 and params['flownet'] == 's' and cfg.endswith("log/ex/exA/config.ini")


def test_run_eval_glue_on_cpu(kitti2012, tmp_path, capsys):
    """Everything around the network (experiment lookup, checkpoint restore, KITTI input, output
    directory, flags) with a stub estimator."""
    import argparse
    from unflow_b200.e2eflow.core import tf_checkpoint as ck
    from unflow_b200.e2eflow.core.flownet import FlowNetVariables
    ini = tmp_path / "config.ini"
    ini.write_text("[dirs]\nlog = %s/log\ncheckpoints = %s/ckpts\ndata = %s\n[run]\nbatch_size = 4\n[train]\n"
                   "flownet = s\nternary_weight = 1.0\n" % (tmp_path, tmp_path, kitti2012))
    src = FlowNetVariables("s", False, seed=5)
    ck.save_variables(src, str(tmp_path / "ckpts" / "exB" / "model.ckpt-3"))
    got = {}

    def make(params, normalization, variables):
        got['params'], got['norm'] = params, normalization
        got['same'] = all(torch.equal(variables.to_tf_dict()[k], v) for k, v in src.to_tf_dict().items())

        def fn(im1, im2):
            assert im1.shape == (1, 384, 1280, 3)
            return torch.zeros(1, 384, 1280, 2), torch.zeros(1, 384, 1280, 2)
        return fn

    args = argparse.Namespace(dataset='kitti', variant='train_2012', ex='exB', num=2, gpu='0',
                              output_benchmark=True, output_visual=False, output_backward=False,
                              output_png=True, config=str(ini), out=str(tmp_path / "out"))
    res = E.run_eval(args, torch.device('cpu'), make_flow_fn=make)
    assert got['same'] and got['params']['flownet'] == 's' and got['norm'][0][0] == 104.920005
    assert abs(res['exB']['EPE_all'] - 2.5) < 1e-5            # zero flow against u = 2 and 3
    assert sorted(os.listdir(str(tmp_path / "out" / "exB"))) == ['000000_10.png', '000001_10.png', 'config.ini']
    assert "-- evaluating: on 2 pairs from kitti/train_2012" in capsys.readouterr().out
