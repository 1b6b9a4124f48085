"""Host logic of csrc/tc_conv.cu on the CPU: the tap / output-class / tile plan the launcher builds
(unflow_tc_conv_plan) is executed here with plain tensor ops -- exactly the sum the kernel's
producer / MMA / epilogue roles implement, including the tile boxes and their masking -- and must
reproduce torch's conv2d / conv_transpose2d (the layers of reference flownet.py:166-233, :89-155 with
TF SAME padding, and their input gradients).  No GPU, no kernel launch."""
import ctypes

import pytest
import torch
import torch.nn.functional as F

from unflow_b200 import _native


def plan(N, Hin, Win, Cin, Hout, Wout, Cout, mode, stride, kh, kw, pt, pl):
    buf = (ctypes.c_int * 512)()
    n = _native.lib().unflow_tc_conv_plan(N, Hin, Win, Cin, Hout, Wout, Cout, mode, stride, kh, kw, pt, pl, buf, 512)
    assert n > 0, n
    v = list(buf[:n])
    keys = ["n_classes", "s_in", "s_out", "Hit", "Wit", "TW", "TH", "TN", "tiles_x", "tiles_y", "tiles_n",
            "n_blocks", "BN", "kblocks", "ntaps"]
    p = dict(zip(keys, v[:15]))
    p["class_start"] = v[15:20]
    p["class_pxy"] = [(v[20 + 2 * i], v[21 + 2 * i]) for i in range(4)]
    p["taps"] = [tuple(v[28 + 3 * i: 31 + 3 * i]) for i in range(p["ntaps"])]
    return p


def execute(p, x, w_taps, N, Hout, Wout, Cout):
    """x [N,Hin,Win,Cin] (NHWC), w_taps [taps][Cout][Cin]; tile by tile like the kernel."""
    _, Hin, Win, Cin = x.shape
    out = torch.full((N, Hout, Wout, Cout), float("nan"), dtype=x.dtype)
    written = torch.zeros((N, Hout, Wout), dtype=torch.int32)
    TW, TH, TN = p["TW"], p["TH"], p["TN"]
    assert TW * TH * TN <= 128
    for cls in range(p["n_classes"]):
        px, py = p["class_pxy"][cls]
        taps = p["taps"][p["class_start"][cls]:p["class_start"][cls + 1]]
        for tn in range(p["tiles_n"]):
            for ty in range(p["tiles_y"]):
                for tx in range(p["tiles_x"]):
                    n0, iy0, ix0 = tn * TN, ty * TH, tx * TW
                    acc = torch.zeros((TN, TH, TW, Cout), dtype=x.dtype)
                    for (dx, dy, widx) in taps:
                        # the TMA box: element stride s_in, zero fill outside the tensor
                        box = torch.zeros((TN, TH, TW, Cin), dtype=x.dtype)
                        for a in range(TN):
                            for b in range(TH):
                                for c in range(TW):
                                    n, yy, xx = n0 + a, p["s_in"] * (iy0 + b) + dy, p["s_in"] * (ix0 + c) + dx
                                    if n < N and 0 <= yy < Hin and 0 <= xx < Win:
                                        box[a, b, c] = x[n, yy, xx]
                        acc += box @ w_taps[widx].t()
                    for a in range(TN):
                        for b in range(TH):
                            for c in range(TW):
                                n, iy, ix = n0 + a, iy0 + b, ix0 + c
                                if n < N and iy < p["Hit"] and ix < p["Wit"]:
                                    oy, ox = p["s_out"] * iy + py, p["s_out"] * ix + px
                                    out[n, oy, ox] = acc[a, b, c]
                                    written[n, oy, ox] += 1
    assert int(written.min()) == 1 and int(written.max()) == 1     # every output exactly once
    return out


CONV = [(2, 5, 9, 11, 7, 3, 1, (1, 1, 1, 1)), (1, 4, 6, 10, 12, 1, 1, (0, 0, 0, 0)),
        (2, 3, 8, 12, 5, 3, 2, (0, 1, 0, 1)), (1, 3, 8, 12, 4, 5, 2, (1, 2, 1, 2)),
        (1, 2, 12, 16, 3, 7, 2, (2, 3, 2, 3)), (3, 2, 6, 20, 70, 3, 1, (1, 1, 1, 1))]


@pytest.mark.parametrize("N,Cin,H,W,Cout,k,stride,pads", CONV)
def test_conv_plan(N, Cin, H, W, Cout, k, stride, pads):
    pt, pb, pl, pr = pads
    g = torch.Generator().manual_seed(k * 7 + Cin)
    x = torch.randn(N, Cin, H, W, generator=g, dtype=torch.float64)
    w = torch.randn(Cout, Cin, k, k, generator=g, dtype=torch.float64)
    ref = F.conv2d(F.pad(x, (pl, pr, pt, pb)), w, stride=stride)
    Ho, Wo = ref.shape[2], ref.shape[3]
    p = plan(N, H, W, Cin, Ho, Wo, Cout, 0, stride, k, k, pt, pl)
    assert p["n_classes"] == 1 and p["ntaps"] == k * k and p["BN"] in (32, 64, 128)
    assert p["n_blocks"] * p["BN"] >= Cout > (p["n_blocks"] - 1) * p["BN"]
    w_taps = w.permute(2, 3, 0, 1).reshape(k * k, Cout, Cin)
    got = execute(p, x.permute(0, 2, 3, 1), w_taps, N, Ho, Wo, Cout)
    torch.testing.assert_close(got.permute(0, 3, 1, 2), ref, rtol=1e-12, atol=1e-12)


DECONV = [(2, 5, 3, 5, 7, 4, 2, 1, None), (1, 6, 4, 6, 3, 3, 2, 0, (8, 12)), (1, 4, 4, 6, 3, 5, 2, 1, (8, 12)),
          (1, 4, 6, 8, 3, 7, 2, 2, (12, 16)), (2, 4, 5, 7, 6, 3, 1, 1, None), (1, 3, 5, 7, 2, 1, 1, 0, None)]


@pytest.mark.parametrize("N,Cin,H,W,Cout,k,stride,pad,out_hw", DECONV)
def test_transposed_plan(N, Cin, H, W, Cout, k, stride, pad, out_hw):
    """deconvN forward (k4 s2 p1) and the input gradients of the stride-2 / stride-1 convolutions (the
    cropped transposed convolution: TF SAME's bottom / right padding rows are never produced)."""
    g = torch.Generator().manual_seed(k * 5 + Cout)
    x = torch.randn(N, Cin, H, W, generator=g, dtype=torch.float64)
    w = torch.randn(Cin, Cout, k, k, generator=g, dtype=torch.float64)
    Ho, Wo = (H - 1) * stride - 2 * pad + k, (W - 1) * stride - 2 * pad + k
    oph = opw = 0
    if out_hw:
        oph, opw = out_hw[0] - Ho, out_hw[1] - Wo
        Ho, Wo = out_hw
    ref = F.conv_transpose2d(x, w, stride=stride, padding=pad, output_padding=(max(oph, 0), max(opw, 0)))[:, :, :Ho, :Wo]
    if Ho % stride or Wo % stride:
        assert _native.lib().unflow_tc_conv_plan(N, H, W, Cin, Ho, Wo, Cout, 1, stride, k, k, pad, pad, None, 0) == -1
        return
    p = plan(N, H, W, Cin, Ho, Wo, Cout, 1, stride, k, k, pad, pad)
    assert p["n_classes"] == stride * stride and p["ntaps"] == k * k
    w_taps = w.permute(2, 3, 1, 0).reshape(k * k, Cout, Cin)
    got = execute(p, x.permute(0, 2, 3, 1), w_taps, N, Ho, Wo, Cout)
    torch.testing.assert_close(got.permute(0, 3, 1, 2), ref, rtol=1e-12, atol=1e-12)


def test_plan_tiles_of_the_flownet_shapes():
    """Tile boxes at the FlowNetC geometry (2B = 8 samples): utilisation of the 128 MMA rows."""
    for (H, W, want) in [(48, 160, 1.0), (24, 80, 1.0), (12, 40, 0.93), (6, 20, 0.93)]:
        p = plan(8, H, W, 64, H, W, 64, 0, 1, 3, 3, 1, 1)
        tiles = p["tiles_x"] * p["tiles_y"] * p["tiles_n"]
        util = 8 * H * W / (tiles * 128.0)
        assert util >= want - 1e-9, (H, W, p["TW"], p["TH"], p["TN"], util)


def test_plan_rejects_bad_arguments():
    lib = _native.lib()
    assert lib.unflow_tc_conv_plan(1, 8, 8, 4, 8, 8, 4, 0, 3, 3, 3, 1, 1, None, 0) == -1      # stride 3
    assert lib.unflow_tc_conv_plan(1, 8, 8, 4, 8, 8, 4, 2, 1, 3, 3, 1, 1, None, 0) == -1      # mode 2
    assert lib.unflow_tc_conv_plan(1, 8, 8, 4, 8, 8, 4, 0, 1, 9, 9, 4, 4, None, 0) == -1      # 81 taps
    assert lib.unflow_tc_conv_plan(1, 8, 8, 4, 8, 8, 4, 0, 1, 3, 3, 1, 1, None, 0) < -1       # needs a buffer
