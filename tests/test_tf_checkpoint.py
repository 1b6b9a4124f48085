"""TF checkpoint (tensor bundle) import/export, train.py:23-65 / SURVEY.md 8f N2.  No TensorFlow
here: the format restatement is pinned on published constants and hand-assembled bytes, plus
write -> read round trips."""
import os
import struct

import numpy as np
import pytest
import torch

from unflow_b200.e2eflow.core import tf_checkpoint as ck
from unflow_b200.e2eflow.core.flownet import FlowNetVariables


def test_crc32c_known_answers():
    assert ck.crc32c(b"123456789") == 0xe3069283               # the CRC-32C check value
    assert ck.crc32c(b"") == 0
    assert ck.crc32c(bytes(32)) == 0x8a9136aa                   # RFC 3720 B.4: 32 zero bytes
    assert ck.crc32c(bytes([0xff] * 32)) == 0x62a8ab43          # RFC 3720 B.4: 32 0xff bytes
    assert ck.crc32c(bytes(range(32))) == 0x46dd794e            # RFC 3720 B.4: ascending
    a = np.arange(1000, dtype=np.float32)
    assert ck.crc32c(a) == ck.crc32c(a.tobytes())
    assert ck.crc32c(b"456789", ck.crc32c(b"123")) == 0xe3069283   # incremental
    # LevelDB mask (crc32c.h): rotate + delta, invertible
    assert ck.mask_crc(0) == 0xa282ead8
    for c in (0, 1, 0xe3069283, 0xffffffff):
        assert ck.unmask_crc(ck.mask_crc(c)) == c and ck.mask_crc(c) != c


def test_varint_and_proto_round_trip():
    for n in (0, 1, 127, 128, 300, 2 ** 32, 2 ** 63 + 5):
        b = ck._put_varint(n)
        assert ck._get_varint(b, 0) == (n, len(b))
    assert ck._put_varint(300) == b"\xac\x02"                   # protobuf docs example
    e = ck._decode_entry(ck._encode_entry(1, (3, 3, 6, 64), 0, 4096, 13824, 0xdeadbeef))
    assert e == {'dtype': 1, 'shape': [3, 3, 6, 64], 'shard_id': 0, 'offset': 4096, 'size': 13824,
                 'crc32c': 0xdeadbeef, 'slices': 0}
    # hand-assembled BundleEntryProto: dtype=DT_FLOAT, shape {dim{size:2} dim{size:5}}, offset 7, size 40
    raw = bytes([0x08, 0x01, 0x12, 0x08, 0x12, 0x02, 0x08, 0x02, 0x12, 0x02, 0x08, 0x05, 0x20, 0x07,
                 0x28, 0x28, 0x35]) + struct.pack('<I', 123)
    e = ck._decode_entry(raw)
    assert (e['dtype'], e['shape'], e['offset'], e['size'], e['crc32c']) == (1, [2, 5], 7, 40, 123)


def test_snappy_uncompress_hand_built():
    # "abcabcabcabc": literal "abc" + copy(offset 3, length 9) with a 2-byte offset element
    comp = bytes([12, (3 - 1) << 2]) + b"abc" + bytes([((9 - 1) << 2) | 2, 3, 0])
    assert ck._snappy_uncompress(comp) == b"abcabcabcabc"
    # 1-byte-offset copy: length 4..11, offset < 2048
    comp = bytes([8, (4 - 1) << 2]) + b"wxyz" + bytes([((4 - 4) << 2) | 1, 4])
    assert ck._snappy_uncompress(comp) == b"wxyzwxyz"
    with pytest.raises(ck.CheckpointError):
        ck._snappy_uncompress(bytes([5, 0]) + b"a")


def test_table_layout_and_round_trip(tmp_path):
    path = str(tmp_path / "t.index")
    items = [(b"", b"hdr")] + [(("var%04d/weights" % i).encode(), os.urandom(40)) for i in range(5000)]
    ck._write_table(path, items)
    raw = open(path, "rb").read()
    assert struct.unpack("<Q", raw[-8:])[0] == 0xdb4775248b80fb57 and len(raw) > ck.BLOCK_SIZE
    # first entry of the first block: shared 0, key length 0, value length 3
    assert raw[:6] == b"\x00\x00\x03hdr"
    assert ck._read_table(path) == items
    corrupt = bytearray(raw)
    corrupt[100] ^= 0x40
    open(path, "wb").write(corrupt)
    with pytest.raises(ck.CheckpointError):
        ck._read_table(path)
    assert len(ck._read_table(path, verify=False)) == len(items)
    with pytest.raises(ck.CheckpointError):
        ck._write_table(path, [(b"b", b""), (b"a", b"")])


def test_bundle_round_trip_and_corruption(tmp_path):
    prefix = str(tmp_path / "model.ckpt-7")
    rng = np.random.default_rng(0)
    tensors = {"flownet_s/conv1/weights": rng.standard_normal((7, 7, 6, 64)).astype(np.float32),
               "flownet_s/conv1/biases": rng.standard_normal(64).astype(np.float32),
               "global_step": np.asarray(7, dtype=np.int64)}
    ck.write_bundle(prefix, tensors)
    assert sorted(os.listdir(tmp_path)) == ["model.ckpt-7.data-00000-of-00001", "model.ckpt-7.index"]
    r = ck.BundleReader(prefix)
    assert r.variables()["flownet_s/conv1/weights"] == (np.dtype("float32"), (7, 7, 6, 64))
    for k, v in tensors.items():
        got = r.tensor(k)
        assert got.dtype == v.dtype and got.shape == v.shape and np.array_equal(got, v)
    with pytest.raises(KeyError):
        r.tensor("nope")
    data = str(tmp_path / "model.ckpt-7.data-00000-of-00001")
    raw = bytearray(open(data, "rb").read())
    raw[-20] ^= 1
    open(data, "wb").write(raw)
    with pytest.raises(ck.CheckpointError):
        for k in tensors:
            ck.BundleReader(prefix).tensor(k)


def test_checkpoint_state_file(tmp_path):
    d = str(tmp_path)
    assert ck.get_checkpoint_state(d) is None
    ck.update_checkpoint_state(d, os.path.join(d, "model.ckpt-100"))
    ck.update_checkpoint_state(d, os.path.join(d, "model.ckpt-200"))
    text = open(os.path.join(d, "checkpoint")).read()
    assert text == ('model_checkpoint_path: "model.ckpt-200"\n'
                    'all_model_checkpoint_paths: "model.ckpt-100"\n'
                    'all_model_checkpoint_paths: "model.ckpt-200"\n')
    latest, every = ck.get_checkpoint_state(d)
    assert latest == os.path.join(d, "model.ckpt-200") and len(every) == 2
    assert ck.checkpoint_iteration(latest) == 200


def test_networks_through_a_checkpoint(tmp_path):
    """variables -> TF checkpoint -> fresh variables, including a stacked net, Adam slots, the
    partial (no full_res) restore of train.py:52-61 and per-network restores."""
    assert ck.net_names("CSS") == ["flownet_c", "stack_1_flownet", "stack_2_flownet"]
    src = FlowNetVariables("cs", False, seed=1)
    prefix = str(tmp_path / "ex" / "model.ckpt-42")
    slots = {n: (np.full(tuple(t.shape), 0.5, np.float32), np.full(tuple(t.shape), 0.25, np.float32))
             for n, t in src.to_tf_dict().items()}
    ck.save_variables(src, prefix, adam_slots=slots)
    assert ck.get_checkpoint_state(str(tmp_path / "ex"))[0] == prefix
    r = ck.BundleReader(prefix)
    assert "flownet_c_features/conv1/weights" in r and "stack_1_flownet/flownet_s/conv1/weights/Adam_1" in r
    assert r.variables()["flownet_c/deconv5/weights"][1][:2] == (4, 4)
    dst = FlowNetVariables("cs", False, seed=2)
    names = ck.restore_variables(dst, prefix)
    assert len(names) == len(src.variable_names())
    for k, v in src.to_tf_dict().items():
        assert torch.equal(dst.to_tf_dict()[k], v), k
    # only the first network
    dst2 = FlowNetVariables("cs", False, seed=3)
    before = dst2.to_tf_dict()
    ck.restore_variables(dst2, prefix, nets=[0])
    after = dst2.to_tf_dict()
    for k in after:
        if k.endswith('biases'):          # all-zero initial biases: nothing to tell apart
            continue
        same_as_src = torch.equal(after[k], src.to_tf_dict()[k])
        assert same_as_src == (not k.startswith("stack_")), k
        if k.startswith("stack_"):
            assert torch.equal(after[k], before[k])
    # a full_res model restored from a checkpoint without the full-resolution layers
    with pytest.raises(ValueError):
        FlowNetVariables("c", True)
    full = FlowNetVariables("s", True, seed=4)
    part_prefix = str(tmp_path / "part" / "model.ckpt-1")
    ck.save_variables(FlowNetVariables("s", False, seed=5), part_prefix)
    got = ck.restore_variables(full, part_prefix)
    assert not any("full_res" in n for n in got)
    with pytest.raises(KeyError):
        ck.restore_variables(full, part_prefix, allow_partial=False)
    with pytest.raises(KeyError):     # wrong architecture: missing non-full_res variables
        ck.restore_variables(FlowNetVariables("c", False, seed=6), part_prefix)


def test_run_checkpoint_helpers_both_formats(tmp_path):
    """run.py: save / newest-checkpoint lookup / resume incl. Adam moments / finetune resolution,
    for the .pt and the TF format (CPU Trainer: construction and state handling need no GPU)."""
    import sys
    sys.path.insert(0, os.path.dirname(__file__))
    import synth
    from unflow_b200 import run as R
    from unflow_b200.e2eflow.core.train import Trainer
    params = dict(synth.KITTI_PARAMS, flownet='cs')
    a = Trainer(params, synth.KITTI_NORMALIZATION, 'cpu', seed=1)
    assert all(n.startswith('stack_1_flownet/') for n in a.trainable_names)   # only the last net trains
    a.adam_m.copy_(torch.randn(a.adam_m.shape, generator=torch.Generator().manual_seed(0)))
    a.adam_v.copy_(torch.rand(a.adam_v.shape, generator=torch.Generator().manual_seed(1)))
    ex = tmp_path / "checkpoints" / "exA"
    ex.mkdir(parents=True)
    R.save_checkpoint(a, str(ex), 10, 'pt')
    assert R.latest_checkpoint(str(ex)) == (10, str(ex / "model.ckpt-10.pt"))
    R.save_checkpoint(a, str(ex), 20, 'tf')
    assert R.latest_checkpoint(str(ex)) == (20, str(ex / "model.ckpt-20"))
    for it, path in ((10, str(ex / "model.ckpt-10.pt")), (20, str(ex / "model.ckpt-20"))):
        b = Trainer(params, synth.KITTI_NORMALIZATION, 'cpu', seed=2)
        R.restore_checkpoint(b, path, with_optimizer=True)
        assert torch.equal(b.flat_param, a.flat_param), path
        assert torch.equal(b.adam_m[:b.num_params], a.adam_m[:a.num_params]), path
        assert torch.equal(b.adam_v[:b.num_params], a.adam_v[:a.num_params]), path
        for k, v in a.variables.to_tf_dict().items():
            assert torch.equal(b.variables.to_tf_dict()[k], v), k
    # finetune = exA: resolved through [dirs] checkpoints, then [dirs] log/ex (util.py:75-85)
    dirs = {'checkpoints': str(tmp_path / "checkpoints"), 'log': str(tmp_path / "log")}
    p = {'finetune': 'exA'}
    R.convert_input_strings(p, dirs)
    assert p['finetune'] == [(20, str(ex / "model.ckpt-20"))]
    (tmp_path / "log" / "ex" / "exB").mkdir(parents=True)
    R.save_checkpoint(a, str(tmp_path / "log" / "ex" / "exB"), 5, 'tf')
    p = {'finetune': 'exA,exB'}
    R.convert_input_strings(p, dirs)
    assert [c[0] for c in p['finetune']] == [20, 5]
    with pytest.raises(AssertionError):
        R.convert_input_strings({'finetune': 'nope'}, dirs)
    # network 0 only, from a single-network 'c' experiment (the usual C -> CS stacking workflow)
    c = Trainer(dict(params, flownet='c'), synth.KITTI_NORMALIZATION, 'cpu', seed=7)
    exC = tmp_path / "checkpoints" / "exC"
    exC.mkdir()
    R.save_checkpoint(c, str(exC), 3, 'tf')
    d = Trainer(params, synth.KITTI_NORMALIZATION, 'cpu', seed=8)
    before = d.variables.to_tf_dict()
    R.restore_checkpoint(d, R.latest_checkpoint(str(exC))[1], nets=[0])
    after, src = d.variables.to_tf_dict(), c.variables.to_tf_dict()
    for k in after:
        if k.endswith('weights'):
            assert torch.equal(after[k], src[k] if k in src else before[k]), k
