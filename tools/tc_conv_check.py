#!/usr/bin/env python
"""GPU check of csrc/tc_conv.cu (tcgen05 3xTF32 implicit-GEMM conv): correctness against float64
convolutions on small cases of every mode, then accuracy and timing on the FlowNetC layer shapes next
to the library path it replaces (cuDNN 3xTF32 via core/conv_ops.py) and plain fp32 cuDNN.

    python tools/tc_conv_check.py [--quick]      (one JSON line per case)
"""
import json
import os
import sys
import time

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from unflow_b200.e2eflow.core import tc_conv as T  # noqa: E402
from unflow_b200.e2eflow.core import conv_ops  # noqa: E402

dev = torch.device("cuda")


def say(**kw):
    print(json.dumps(kw), flush=True)


def cl(t):
    return t.contiguous(memory_format=torch.channels_last)


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def make_x(N, C, H, W, pitch=None, seed=0, image_like=False):
    g = torch.Generator(device="cpu").manual_seed(seed)
    if image_like:
        x = torch.rand(N, C, H, W, generator=g) * 0.8 - 0.4
        x = F.avg_pool2d(F.pad(x, (1, 1, 1, 1), mode="replicate"), 3, 1)
    else:
        x = torch.randn(N, C, H, W, generator=g)
    pitch = pitch or T.round4(C)
    buf = torch.full((N, H, W, pitch), 7.25, device=dev)          # poison the slack channels
    view = buf[..., :C].permute(0, 3, 1, 2)
    view.copy_(x.to(dev))
    return view


def case_conv(name, N, Cin, Cout, H, W, k, stride, pads, bias=True, act=True, pitch=None, accumulate=False,
              image_like=False, time_it=False, compare_lib=False):
    pt, pb, pl, pr = pads
    x = make_x(N, Cin, H, W, pitch, seed=Cin + H, image_like=image_like)
    g = torch.Generator().manual_seed(Cout + k)
    w = cl((torch.randn(Cout, Cin, k, k, generator=g) * (2.0 / (Cin * k * k)) ** 0.5).to(dev))
    b = (torch.randn(Cout + 1, generator=g) * 0.1).to(dev)[1:] if bias else None     # 4-byte aligned only
    Ho, Wo = (H + pt + pb - k) // stride + 1, (W + pl + pr - k) // stride + 1
    ref = F.conv2d(F.pad(x.double(), (pl, pr, pt, pb)), w.double(), b.double() if bias else None, stride=stride)
    base = torch.randn(N, Cout, Ho, Wo, device=dev).contiguous(memory_format=torch.channels_last) if accumulate else None
    if act:
        ref = F.leaky_relu(ref, 0.1)
    if accumulate:
        ref = ref + base.double()
    planes = T.split_weights(w)
    out, obuf = out_buf(N, Cout, Ho, Wo)
    if accumulate:
        out.copy_(base)
    T.run(x, planes, out, mode=0, stride=stride, kh=k, kw=k, pad_t=pt, pad_l=pl, bias=b, act=act,
          accumulate=accumulate)
    torch.cuda.synchronize()
    rec = dict(case=name, mode="conv", N=N, Cin=Cin, Cout=Cout, H=H, W=W, k=k, stride=stride, err=rel(out, ref),
               finite=bool(torch.isfinite(out).all()), slack_untouched=bool((obuf[..., Cout:] == -3.5).all()))
    if compare_lib:
        conv_ops.set_mode("3xtf32")
        with torch.no_grad():
            y3 = conv_ops.conv2d(x if x.is_contiguous(memory_format=torch.channels_last) else cl(x), w, b, stride, pads, act=act)
        conv_ops.set_mode("fp32")
        with torch.no_grad():
            y32 = conv_ops.conv2d(x.contiguous(), w, b, stride, pads, act=act)
        rec["err_cudnn_3xtf32"] = rel(y3, ref)
        rec["err_cudnn_fp32"] = rel(y32, ref)
    if time_it:
        rec["us"] = bench(lambda: T.run(x, planes, out, mode=0, stride=stride, kh=k, kw=k, pad_t=pt, pad_l=pl,
                                        bias=b, act=act))
        flops = 2.0 * N * Ho * Wo * Cout * Cin * k * k
        rec["tflops_fp32_equiv"] = round(flops / rec["us"] / 1e6, 1)
        if compare_lib:
            conv_ops.set_mode("3xtf32")
            xc = x if x.is_contiguous(memory_format=torch.channels_last) else cl(x)
            with torch.no_grad():
                rec["us_cudnn_3xtf32_with_operand_passes"] = bench(lambda: conv_ops.conv2d(xc, w, b, stride, pads, act=act))
            conv_ops.set_mode("fp32")
    say(**rec)
    return rec


def case_deconv(name, N, Cin, Cout, H, W, k, stride, pad, bias=True, act=True, pitch=None, time_it=False,
                out_hw=None, compare_lib=False):
    """mode 1 against F.conv_transpose2d (output_padding chosen to reach out_hw)."""
    x = make_x(N, Cin, H, W, pitch, seed=Cin + W)
    g = torch.Generator().manual_seed(Cout + 3 * k)
    w = cl((torch.randn(Cin, Cout, k, k, generator=g) * (2.0 / (Cin * k * k / stride ** 2)) ** 0.5).to(dev))   # IOHW
    b = (torch.randn(Cout, generator=g) * 0.1).to(dev) if bias else None
    Ho = (H - 1) * stride - 2 * pad + k
    Wo = (W - 1) * stride - 2 * pad + k
    oph = opw = 0
    if out_hw is not None:
        oph, opw = out_hw[0] - Ho, out_hw[1] - Wo
        Ho, Wo = out_hw
    ref = F.conv_transpose2d(x.double(), w.double(), b.double() if bias else None, stride=stride, padding=pad,
                             output_padding=(max(oph, 0), max(opw, 0)))[:, :, :Ho, :Wo]   # crop = SAME's bottom/right pad
    if act:
        ref = F.leaky_relu(ref, 0.1)
    planes = T.split_weights(w, transpose=True)      # rows = Cout, contraction = Cin
    out, obuf = out_buf(N, Cout, Ho, Wo)
    T.run(x, planes, out, mode=1, stride=stride, kh=k, kw=k, pad_t=pad, pad_l=pad, bias=b, act=act)
    torch.cuda.synchronize()
    rec = dict(case=name, mode="transposed", N=N, Cin=Cin, Cout=Cout, H=H, W=W, k=k, stride=stride,
               err=rel(out, ref), finite=bool(torch.isfinite(out).all()),
               slack_untouched=bool((obuf[..., Cout:] == -3.5).all()))
    if time_it:
        rec["us"] = bench(lambda: T.run(x, planes, out, mode=1, stride=stride, kh=k, kw=k, pad_t=pad, pad_l=pad,
                                        bias=b, act=act))
        flops = 2.0 * N * H * W * Cout * Cin * k * k
        rec["tflops_fp32_equiv"] = round(flops / rec["us"] / 1e6, 1)
        if compare_lib and stride == 2 and k == 4:
            conv_ops.set_mode("3xtf32")
            xc = x if x.is_contiguous(memory_format=torch.channels_last) else cl(x)
            with torch.no_grad():
                rec["us_cudnn_3xtf32_with_operand_passes"] = bench(lambda: conv_ops.conv_transpose2d(xc, w, b, act=act))
            conv_ops.set_mode("fp32")
    say(**rec)
    return rec


def case_wgrad(name, N, Cin, Cout, H, W, k, stride, pads, pitch_x=None, time_it=False, deconv=False):
    """csrc/tc_wgrad.cu against torch's float64 weight gradient.  conv: y = conv(x, w); deconv: k4 s2 p1."""
    g = torch.Generator().manual_seed(Cin + 5 * k)
    if deconv:
        x = make_x(N, Cin, H, W, pitch_x, seed=Cin + 1)
        gy = make_x(N, Cout, 2 * H, 2 * W, None, seed=Cout + 2)
        xr = x.double().clone().requires_grad_(True)
        wr = torch.zeros(Cin, Cout, 4, 4, device=dev, dtype=torch.float64, requires_grad=True)
        F.conv_transpose2d(xr, wr, stride=2, padding=1).backward(gy.double())
        ref = wr.grad
        dw = torch.zeros(Cin, Cout, 4, 4, device=dev).contiguous(memory_format=torch.channels_last)
        run = lambda: T.wgrad(x, gy, dw, stride=2, kh=4, kw=4, pad_t=1, pad_l=1)
        flops = 2.0 * N * H * W * Cout * Cin * 16
    else:
        pt, pb, pl, pr = pads
        x = make_x(N, Cin, H, W, pitch_x, seed=Cin + 1)
        Ho, Wo = (H + pt + pb - k) // stride + 1, (W + pl + pr - k) // stride + 1
        gy = make_x(N, Cout, Ho, Wo, None, seed=Cout + 2)
        ref = torch.nn.grad.conv2d_weight(F.pad(x.double(), (pl, pr, pt, pb)), (Cout, Cin, k, k), gy.double(), stride=stride)
        dw = torch.zeros(Cout, Cin, k, k, device=dev).contiguous(memory_format=torch.channels_last)
        run = lambda: T.wgrad(gy, x, dw, stride=stride, kh=k, kw=k, pad_t=pt, pad_l=pl)
        flops = 2.0 * N * Ho * Wo * Cout * Cin * k * k
    run()
    torch.cuda.synchronize()
    rec = dict(case=name, mode="wgrad", N=N, Cin=Cin, Cout=Cout, H=H, W=W, k=k, stride=stride, err=rel(dw, ref),
               finite=bool(torch.isfinite(dw).all()), slack_untouched=True)
    if time_it:
        rec["us"] = bench(run)
        rec["tflops_fp32_equiv"] = round(flops / rec["us"] / 1e6, 1)
    say(**rec)
    return rec


def case_window(name, N, Ci, H, W, Co=64, time_it=True):
    """7x7 stride-2 first layer in the row-window form: forward + weight gradient, timing."""
    k = 7
    g = torch.Generator().manual_seed(Ci)
    x = (torch.rand(N, Ci, H, W, generator=g) - 0.4).to(dev)
    w = cl((torch.randn(Co, Ci, k, k, generator=g) * 0.05).to(dev))
    b = (torch.randn(Co, generator=g) * 0.1).to(dev)
    ref = F.leaky_relu(F.conv2d(F.pad(x.double(), (2, 3, 2, 3)), w.double(), b.double(), stride=2), 0.1)
    Ho, Wo = ref.shape[2:]
    cp = T.window_channels(Ci)
    xp = T.window_input(x, 2, 2, Wo)
    planes = T.split_weights(T.window_weights(w, cp))
    out = T.empty_nhwc(N, Co, Ho, Wo, dev)
    T.run_window(xp, planes, out, kh=k, stride=2, pad_t=2, bias=b, act=True)
    gy = cl(torch.randn(N, Co, Ho, Wo, generator=g).to(dev))
    dw = torch.zeros((Co, k, 1, 8 * cp), device=dev).permute(0, 3, 1, 2)
    T.wgrad_window(gy, xp, dw, kh=k, stride=2, pad_t=2)
    refw = torch.nn.grad.conv2d_weight(F.pad(x.double(), (2, 3, 2, 3)), (Co, Ci, k, k), gy.double(), stride=2)
    got = dw.permute(0, 2, 3, 1).reshape(Co, k, 8, cp)[:, :, :k, :Ci].permute(0, 3, 1, 2)
    rec = dict(case=name, mode="window", N=N, Cin=Ci, Cout=Co, H=H, W=W, k=k, stride=2, err=rel(out, ref),
               err_wgrad=rel(got, refw), finite=bool(torch.isfinite(out).all()), slack_untouched=True)
    if time_it:
        rec["us"] = bench(lambda: T.run_window(xp, planes, out, kh=k, stride=2, pad_t=2, bias=b, act=True))
        rec["us_wgrad"] = bench(lambda: T.wgrad_window(gy, xp, dw, kh=k, stride=2, pad_t=2))
        rec["us_window_input"] = bench(lambda: T.window_input(x, 2, 2, Wo))
    say(**rec)


def out_buf(N, C, H, W, fill=float("nan")):
    """NCHW-shaped view with NHWC memory and a channel pitch that is a multiple of 4 (poisoned slack)."""
    buf = torch.full((N, H, W, T.round4(C) + 4), -3.5, device=dev)
    v = buf[..., :C].permute(0, 3, 1, 2)
    v.fill_(fill)
    return v, buf


def bench(fn, iters=10, warmup=3):
    scratch = torch.empty(256 * 1024 * 1024 // 4, device=dev)
    for _ in range(warmup):
        fn()
    ts = []
    for _ in range(iters):
        scratch.zero_()                       # flush L2
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return round(ts[len(ts) // 2], 2)


def conv_layer_cases(time_it=True):
    B = 8
    case_conv("conv3_1", B, 473, 256, 48, 160, 3, 1, (1, 1, 1, 1), pitch=476, time_it=time_it)
    case_conv("conv4_1", B, 512, 512, 24, 80, 3, 1, (1, 1, 1, 1), time_it=time_it)
    case_conv("conv5_1", B, 512, 512, 12, 40, 3, 1, (1, 1, 1, 1), time_it=time_it)
    case_conv("conv6_1", B, 1024, 1024, 6, 20, 3, 1, (1, 1, 1, 1), time_it=time_it)
    case_conv("conv2 (5x5 s2)", B, 64, 128, 192, 640, 5, 2, (1, 2, 1, 2), time_it=time_it)
    case_conv("conv4 (s2)", B, 256, 512, 48, 160, 3, 2, (0, 1, 0, 1), time_it=time_it)
    case_deconv("deconv5", B, 1024, 512, 6, 20, 4, 2, 1, time_it=time_it)
    case_deconv("deconv3", B, 770, 128, 24, 80, 4, 2, 1, pitch=772, time_it=time_it)
    case_deconv("deconv2", B, 386, 64, 48, 160, 4, 2, 1, pitch=388, time_it=time_it)
    case_deconv("dgrad conv3_1", B, 256, 473, 48, 160, 3, 1, 1, bias=False, act=False, time_it=time_it)
    case_deconv("dgrad conv2 (5x5 s2)", B, 128, 64, 96, 320, 5, 2, 1, bias=False, act=False, out_hw=(192, 640),
                time_it=time_it)


def wgrad_layer_cases():
    B = 8
    case_wgrad("wgrad conv3_1", B, 473, 256, 48, 160, 3, 1, (1, 1, 1, 1), pitch_x=476, time_it=True)
    case_wgrad("wgrad conv4_1", B, 512, 512, 24, 80, 3, 1, (1, 1, 1, 1), time_it=True)
    case_wgrad("wgrad conv6_1", B, 1024, 1024, 6, 20, 3, 1, (1, 1, 1, 1), time_it=True)
    case_wgrad("wgrad conv4 (s2)", B, 256, 512, 48, 160, 3, 2, (0, 1, 0, 1), time_it=True)
    case_wgrad("wgrad conv3 (5x5 s2)", B, 128, 256, 96, 320, 5, 2, (1, 2, 1, 2), time_it=True)
    case_wgrad("wgrad deconv2", B, 386, 64, 48, 160, 4, 2, None, pitch_x=388, time_it=True, deconv=True)
    case_wgrad("wgrad deconv5", B, 1024, 512, 6, 20, 4, 2, None, time_it=True, deconv=True)


def set_pair(v):
    from unflow_b200 import _native
    assert _native.lib().unflow_set_int_option(b"tc_pair", v) == 0
    say(option="tc_pair", value=v)


def main():
    quick = "--quick" in sys.argv
    torch.backends.cudnn.benchmark = True
    t0 = time.time()
    if "--roles" in sys.argv:            # where does each warp role of tc_conv_kernel wait? (CTA 0's timers)
        from unflow_b200 import _native
        lib = _native.lib()
        buf = torch.zeros(16, dtype=torch.int64, device=dev)
        names = ["producer: wait free stage", "producer: total", "mma: wait free accumulator", "mma: wait operands",
                 "mma: total", "converter: wait TMA data", "converter: total", "epilogue: wait chunk", "epilogue: total"]

        def roles(label, fn):
            fn()                                   # warm
            torch.cuda.synchronize()
            assert lib.unflow_tc_conv_debug(buf.data_ptr()) == 0
            buf.zero_()
            fn()
            torch.cuda.synchronize()
            lib.unflow_tc_conv_debug(None)
            v = buf.tolist()
            say(case=label, **{n: int(x) for n, x in zip(names, v)})

        B = 8
        for pair in (0, 1):
            set_pair(pair)
            for (label, N, Cin, Cout, H, W, k, st, pads, pitch) in [
                    ("conv3_1", B, 473, 256, 48, 160, 3, 1, (1, 1, 1, 1), 476),
                    ("conv4_1", B, 512, 512, 24, 80, 3, 1, (1, 1, 1, 1), None),
                    ("conv6_1", B, 1024, 1024, 6, 20, 3, 1, (1, 1, 1, 1), None)]:
                x = make_x(N, Cin, H, W, pitch, seed=Cin + H)
                g = torch.Generator().manual_seed(Cout + k)
                w = cl((torch.randn(Cout, Cin, k, k, generator=g) * (2.0 / (Cin * k * k)) ** 0.5).to(dev))
                planes = T.split_weights(w)
                out, obuf = out_buf(N, Cout, H, W)
                roles("%s pair=%d" % (label, pair),
                      lambda: T.run(x, planes, out, mode=0, stride=st, kh=k, kw=k, pad_t=pads[0], pad_l=pads[2], bias=None, act=False))
            x = make_x(B, 386, 48, 160, 388, seed=3)
            g = torch.Generator().manual_seed(9)
            w = cl((torch.randn(386, 64, 4, 4, generator=g) * 0.05).to(dev))
            planes = T.split_weights(w, transpose=True)
            out, obuf = out_buf(B, 64, 96, 320)
            for ppx in (0, 1):
                assert lib.unflow_set_int_option(b"tc_pair_px", ppx) == 0
                fn = lambda: T.run(x, planes, out, mode=1, stride=2, kh=4, kw=4, pad_t=1, pad_l=1, bias=None, act=False)
                roles("deconv2 (64 channels) pair=%d pair_px=%d" % (pair, ppx), fn)
                say(case="deconv2 pair=%d pair_px=%d" % (pair, ppx), us=bench(fn))
            lib.unflow_set_int_option(b"tc_pair_px", 1)
        set_pair(1)
        # the weight-gradient kernel (same timers; [9] = converter blocked on a free operand slot)
        for pair in (0, 1):
            set_pair(pair)
            for (label, N, Cin, Cout, H, W) in [("wgrad conv3_1", B, 473, 256, 48, 160), ("wgrad conv4_1", B, 512, 512, 24, 80)]:
                x = make_x(N, Cin, H, W, T.round4(Cin), seed=Cin + 1)
                gy = make_x(N, Cout, H, W, None, seed=Cout + 2)
                dw = torch.zeros(Cout, Cin, 3, 3, device=dev).contiguous(memory_format=torch.channels_last)
                fn = lambda: T.wgrad(gy, x, dw, stride=1, kh=3, kw=3, pad_t=1, pad_l=1)
                for gs in (1, 3):
                    assert lib.unflow_set_int_option(b"tc_wgrad_gsplit", gs) == 0
                    fn(); torch.cuda.synchronize()
                    assert lib.unflow_tc_conv_debug(buf.data_ptr()) == 0
                    buf.zero_(); fn(); torch.cuda.synchronize(); lib.unflow_tc_conv_debug(None)
                    v = buf.tolist()
                    say(case="%s pair=%d gsplit=%d" % (label, pair, gs), us=bench(fn), slot_wait=int(v[9]),
                        **{n: int(xx) for n, xx in zip(names, v)})
                lib.unflow_set_int_option(b"tc_wgrad_gsplit", 2)
        set_pair(1)
        return
    if "--wgrad-trunc" in sys.argv:      # experiment: hi = raw fp32 (truncated by the tensor core), lo = x - trunc(x)
        from unflow_b200 import _native
        for tr in (0, 1):
            assert _native.lib().unflow_set_int_option(b"tc_wgrad_trunc", tr) == 0
            say(option="tc_wgrad_trunc", value=tr)
            B = 8
            case_wgrad("wgrad 3x3 s1 ragged", 2, 70, 50, 13, 21, 3, 1, (1, 1, 1, 1), pitch_x=80)
            case_wgrad("wgrad conv3_1", B, 473, 256, 48, 160, 3, 1, (1, 1, 1, 1), pitch_x=476, time_it=True)
            case_wgrad("wgrad conv4_1", B, 512, 512, 24, 80, 3, 1, (1, 1, 1, 1), time_it=True)
            case_wgrad("wgrad conv2 (5x5 s2)", B, 64, 128, 192, 640, 5, 2, (1, 2, 1, 2), time_it=True)
            case_wgrad("wgrad deconv2", B, 386, 64, 48, 160, 4, 2, None, pitch_x=388, time_it=True, deconv=True)
        _native.lib().unflow_set_int_option(b"tc_wgrad_trunc", 0)
        return
    if "--chunk-test" in sys.argv:       # K blocks per tensor-memory accumulation: accuracy and time
        from unflow_b200 import _native
        for ck in (4, 8, 16, 32):
            assert _native.lib().unflow_set_int_option(b"tc_chunk", ck) == 0
            say(option="tc_chunk", value=ck)
            B = 8
            case_conv("conv3_1", B, 473, 256, 48, 160, 3, 1, (1, 1, 1, 1), pitch=476, time_it=True)
            case_conv("conv4_1", B, 512, 512, 24, 80, 3, 1, (1, 1, 1, 1), time_it=True)
            case_conv("conv6_1", B, 1024, 1024, 6, 20, 3, 1, (1, 1, 1, 1), time_it=True)
            case_conv("conv2 (5x5 s2)", B, 64, 128, 192, 640, 5, 2, (1, 2, 1, 2), time_it=True)
            case_deconv("deconv2", B, 386, 64, 48, 160, 4, 2, 1, pitch=388, time_it=True)
            case_wgrad("wgrad conv3_1", B, 473, 256, 48, 160, 3, 1, (1, 1, 1, 1), pitch_x=476, time_it=True)
            case_wgrad("wgrad conv4_1", B, 512, 512, 24, 80, 3, 1, (1, 1, 1, 1), time_it=True)
        _native.lib().unflow_set_int_option(b"tc_chunk", 4)
        return
    if "--pair-test" in sys.argv:        # CTA pairs (cta_group::2) forced on: small cases of every mode, then the layers
        set_pair(2)
        case_conv("1x1 two tiles", 1, 32, 64, 16, 16, 1, 1, (0, 0, 0, 0), bias=False, act=False)
        case_conv("1x1 K=64 N=128", 1, 64, 128, 16, 16, 1, 1, (0, 0, 0, 0), bias=False, act=False)
        case_conv("3x3 s1", 2, 64, 128, 16, 24, 3, 1, (1, 1, 1, 1))
        case_conv("3x3 s1 ragged channels + pitch, odd tile count", 3, 70, 64, 13, 21, 3, 1, (1, 1, 1, 1), pitch=80)
        case_conv("3x3 s1 C_out tail 70", 1, 32, 70, 19, 21, 3, 1, (1, 1, 1, 1))
        case_conv("3x3 s1 accumulate, no act", 2, 64, 96, 12, 20, 3, 1, (1, 1, 1, 1), bias=False, act=False, accumulate=True)
        case_conv("3x3 s2 SAME(0,1)", 2, 64, 128, 16, 24, 3, 2, (0, 1, 0, 1))
        case_conv("5x5 s2 SAME(1,2)", 2, 64, 128, 16, 24, 5, 2, (1, 2, 1, 2))
        case_deconv("deconv k4 s2 p1", 2, 64, 128, 6, 10, 4, 2, 1)
        case_deconv("deconv k4 s2 p1 ragged", 2, 130, 64, 6, 20, 4, 2, 1, pitch=132)
        case_deconv("dgrad of 3x3 s2 SAME(0,1)", 2, 128, 64, 8, 12, 3, 2, 0, bias=False, act=False, out_hw=(16, 24))
        case_deconv("dgrad of 3x3 s1", 2, 128, 64, 9, 14, 3, 1, 1, bias=False, act=False)
        case_wgrad("wgrad 3x3 s1, 2 row blocks", 2, 64, 256, 16, 24, 3, 1, (1, 1, 1, 1))
        case_wgrad("wgrad 3x3 s1 ragged, 3 row blocks", 2, 70, 300, 13, 21, 3, 1, (1, 1, 1, 1), pitch_x=80)
        case_wgrad("wgrad 3x3 s2 SAME(0,1)", 2, 64, 256, 16, 24, 3, 2, (0, 1, 0, 1))
        case_wgrad("wgrad deconv k4 s2", 2, 260, 64, 6, 10, 4, 2, None, pitch_x=264, deconv=True)
        say(phase="small pair cases done", seconds=round(time.time() - t0, 1))
        for v in (2, 0, 1):
            set_pair(v)
            conv_layer_cases()
            if v != 2:
                wgrad_layer_cases()
        return
    if "--timing" in sys.argv:           # the FlowNetC layer shapes, all three kernels
        B = 8
        case_conv("3x3 s1 small", 2, 64, 128, 16, 24, 3, 1, (1, 1, 1, 1))
        case_conv("3x3 s1 ragged", 2, 70, 50, 13, 21, 3, 1, (1, 1, 1, 1), pitch=80)
        case_deconv("deconv small", 2, 130, 64, 6, 20, 4, 2, 1, pitch=132)
        case_conv("conv3_1", B, 473, 256, 48, 160, 3, 1, (1, 1, 1, 1), pitch=476, time_it=True)
        case_conv("conv4_1", B, 512, 512, 24, 80, 3, 1, (1, 1, 1, 1), time_it=True)
        case_conv("conv5_1", B, 512, 512, 12, 40, 3, 1, (1, 1, 1, 1), time_it=True)
        case_conv("conv6_1", B, 1024, 1024, 6, 20, 3, 1, (1, 1, 1, 1), time_it=True)
        case_conv("conv2 (5x5 s2)", B, 64, 128, 192, 640, 5, 2, (1, 2, 1, 2), time_it=True)
        case_conv("conv4 (s2)", B, 256, 512, 48, 160, 3, 2, (0, 1, 0, 1), time_it=True)
        case_deconv("deconv5", B, 1024, 512, 6, 20, 4, 2, 1, time_it=True)
        case_deconv("deconv3", B, 770, 128, 24, 80, 4, 2, 1, pitch=772, time_it=True)
        case_deconv("deconv2", B, 386, 64, 48, 160, 4, 2, 1, pitch=388, time_it=True)
        case_deconv("dgrad conv3_1", B, 256, 473, 48, 160, 3, 1, 1, bias=False, act=False, time_it=True)
        case_wgrad("wgrad 3x3 s1 ragged", 2, 70, 50, 13, 21, 3, 1, (1, 1, 1, 1), pitch_x=80)
        case_wgrad("wgrad deconv k4 s2", 2, 130, 64, 6, 10, 4, 2, None, pitch_x=132, deconv=True)
        case_wgrad("wgrad conv3_1", B, 473, 256, 48, 160, 3, 1, (1, 1, 1, 1), pitch_x=476, time_it=True)
        case_wgrad("wgrad conv4_1", B, 512, 512, 24, 80, 3, 1, (1, 1, 1, 1), time_it=True)
        case_wgrad("wgrad conv6_1", B, 1024, 1024, 6, 20, 3, 1, (1, 1, 1, 1), time_it=True)
        case_wgrad("wgrad conv2 (5x5 s2)", B, 64, 128, 192, 640, 5, 2, (1, 2, 1, 2), time_it=True)
        case_wgrad("wgrad deconv2", B, 386, 64, 48, 160, 4, 2, None, pitch_x=388, time_it=True, deconv=True)
        case_wgrad("wgrad deconv5", B, 1024, 512, 6, 20, 4, 2, None, time_it=True, deconv=True)
        case_window("window conv1 FlowNetC", B, 3, 384, 1280)
        return
    if "--profile2" in sys.argv:         # the half-width (BN = 64) layers and the pair kernels, for ncu
        B = 8
        case_deconv("deconv2", B, 386, 64, 48, 160, 4, 2, 1, pitch=388)
        case_deconv("dgrad conv2 (5x5 s2)", B, 128, 64, 96, 320, 5, 2, 1, bias=False, act=False, out_hw=(192, 640))
        case_wgrad("wgrad conv4_1", B, 512, 512, 24, 80, 3, 1, (1, 1, 1, 1))
        return
    if "--profile" in sys.argv:          # one launch of each big kernel, for ncu
        B = 8
        case_conv("conv3_1", B, 473, 256, 48, 160, 3, 1, (1, 1, 1, 1), pitch=476)
        case_deconv("deconv3", B, 770, 128, 24, 80, 4, 2, 1, pitch=772)
        case_wgrad("wgrad conv3_1", B, 473, 256, 48, 160, 3, 1, (1, 1, 1, 1), pitch_x=476)
        return
    if "--wgrad-only" in sys.argv:
        case_wgrad("wgrad 1x1", 1, 32, 128, 8, 16, 1, 1, (0, 0, 0, 0))
        case_wgrad("wgrad 3x3 s1", 2, 64, 128, 16, 24, 3, 1, (1, 1, 1, 1))
        case_wgrad("wgrad 3x3 s1 ragged", 2, 70, 50, 13, 21, 3, 1, (1, 1, 1, 1), pitch_x=80)
        case_wgrad("wgrad 3x3 s2 SAME(0,1)", 2, 64, 128, 16, 24, 3, 2, (0, 1, 0, 1))
        case_wgrad("wgrad 5x5 s2 SAME(1,2)", 2, 24, 64, 16, 24, 5, 2, (1, 2, 1, 2))
        case_wgrad("wgrad deconv k4 s2", 2, 130, 64, 6, 10, 4, 2, None, pitch_x=132, deconv=True)
        B = 8
        case_wgrad("wgrad conv3_1", B, 473, 256, 48, 160, 3, 1, (1, 1, 1, 1), pitch_x=476, time_it=True)
        case_wgrad("wgrad conv4_1", B, 512, 512, 24, 80, 3, 1, (1, 1, 1, 1), time_it=True)
        case_wgrad("wgrad conv6_1", B, 1024, 1024, 6, 20, 3, 1, (1, 1, 1, 1), time_it=True)
        case_wgrad("wgrad conv4 (s2)", B, 256, 512, 48, 160, 3, 2, (0, 1, 0, 1), time_it=True)
        case_wgrad("wgrad conv2 (5x5 s2)", B, 64, 128, 192, 640, 5, 2, (1, 2, 1, 2), time_it=True)
        case_wgrad("wgrad deconv2", B, 386, 64, 48, 160, 4, 2, None, pitch_x=388, time_it=True, deconv=True)
        case_wgrad("wgrad deconv5", B, 1024, 512, 6, 20, 4, 2, None, time_it=True, deconv=True)
        case_window("window conv1 FlowNetC", B, 3, 384, 1280)
        case_window("window conv1 FlowNetS", B, 6, 384, 1280)
        case_window("window conv1 stacked S", 4, 14, 384, 1280)
        return
    # --- smallest possible first: one tile, one K block, one tap -----------------------------------
    case_conv("1x1 one tile", 1, 32, 32, 8, 16, 1, 1, (0, 0, 0, 0), bias=False, act=False)
    case_conv("1x1 K=64 N=128", 1, 64, 128, 8, 16, 1, 1, (0, 0, 0, 0), bias=False, act=False)
    case_conv("3x3 s1", 2, 64, 128, 16, 24, 3, 1, (1, 1, 1, 1))
    case_conv("3x3 s1 ragged channels + pitch", 2, 70, 64, 13, 21, 3, 1, (1, 1, 1, 1), pitch=80)
    case_conv("3x3 s1 C_out tail 70", 1, 32, 70, 9, 11, 3, 1, (1, 1, 1, 1))
    case_conv("3x3 s1 C_out 30", 1, 40, 30, 9, 11, 3, 1, (1, 1, 1, 1))
    case_conv("3x3 s1 accumulate, no act", 2, 64, 96, 12, 20, 3, 1, (1, 1, 1, 1), bias=False, act=False, accumulate=True)
    case_conv("3x3 s2 SAME(0,1)", 2, 64, 128, 16, 24, 3, 2, (0, 1, 0, 1))
    case_conv("5x5 s2 SAME(1,2)", 2, 64, 128, 16, 24, 5, 2, (1, 2, 1, 2))
    case_conv("7x7 s2 SAME(2,3) 6ch image", 2, 6, 64, 32, 48, 7, 2, (2, 3, 2, 3), pitch=8, image_like=True)
    case_deconv("deconv k4 s2 p1", 2, 64, 128, 6, 10, 4, 2, 1)
    case_deconv("deconv k4 s2 p1 ragged", 2, 130, 64, 6, 20, 4, 2, 1, pitch=132)
    case_deconv("dgrad of 3x3 s2 SAME(0,1)", 2, 128, 64, 8, 12, 3, 2, 0, bias=False, act=False, out_hw=(16, 24))
    case_deconv("dgrad of 5x5 s2 SAME(1,2)", 2, 128, 64, 8, 12, 5, 2, 1, bias=False, act=False, out_hw=(16, 24))
    case_deconv("dgrad of 3x3 s1", 2, 128, 64, 9, 14, 3, 1, 1, bias=False, act=False)
    case_wgrad("wgrad 1x1", 1, 32, 128, 8, 16, 1, 1, (0, 0, 0, 0))
    case_wgrad("wgrad 3x3 s1", 2, 64, 128, 16, 24, 3, 1, (1, 1, 1, 1))
    case_wgrad("wgrad 3x3 s1 ragged", 2, 70, 50, 13, 21, 3, 1, (1, 1, 1, 1), pitch_x=80)
    case_wgrad("wgrad 3x3 s2 SAME(0,1)", 2, 64, 128, 16, 24, 3, 2, (0, 1, 0, 1))
    case_wgrad("wgrad 5x5 s2 SAME(1,2)", 2, 24, 64, 16, 24, 5, 2, (1, 2, 1, 2))
    case_wgrad("wgrad deconv k4 s2", 2, 130, 64, 6, 10, 4, 2, None, pitch_x=132, deconv=True)
    say(phase="small cases done", seconds=round(time.time() - t0, 1))
    if quick:
        return
    # --- FlowNetC layer shapes at 384x1280, 2B = 8 samples ---------------------------------------------
    B = 8
    case_conv("conv3_1", B, 473, 256, 48, 160, 3, 1, (1, 1, 1, 1), pitch=476, time_it=True, compare_lib=True)
    case_conv("conv4_1", B, 512, 512, 24, 80, 3, 1, (1, 1, 1, 1), time_it=True, compare_lib=True)
    case_conv("conv5_1", B, 512, 512, 12, 40, 3, 1, (1, 1, 1, 1), time_it=True, compare_lib=True)
    case_conv("conv6_1", B, 1024, 1024, 6, 20, 3, 1, (1, 1, 1, 1), time_it=True, compare_lib=True)
    case_conv("conv4 (s2)", B, 256, 512, 48, 160, 3, 2, (0, 1, 0, 1), time_it=True, compare_lib=True)
    case_conv("conv2 (5x5 s2)", B, 64, 128, 192, 640, 5, 2, (1, 2, 1, 2), time_it=True, compare_lib=True)
    case_conv("conv3 (5x5 s2)", B, 128, 256, 96, 320, 5, 2, (1, 2, 1, 2), time_it=True, compare_lib=True)
    case_conv("conv1 (7x7 s2, 3ch image)", B, 3, 64, 384, 1280, 7, 2, (2, 3, 2, 3), pitch=4, image_like=True,
              time_it=True, compare_lib=False)
    case_conv("conv_redir 1x1", B, 256, 32, 48, 160, 1, 1, (0, 0, 0, 0), time_it=True, compare_lib=True)
    case_deconv("deconv5", B, 1024, 512, 6, 20, 4, 2, 1, time_it=True, compare_lib=True)
    case_deconv("deconv4", B, 1026, 256, 12, 40, 4, 2, 1, pitch=1028, time_it=True, compare_lib=False)
    case_deconv("deconv3", B, 770, 128, 24, 80, 4, 2, 1, pitch=772, time_it=True, compare_lib=False)
    case_deconv("deconv2", B, 386, 64, 48, 160, 4, 2, 1, pitch=388, time_it=True, compare_lib=False)
    case_deconv("dgrad conv3_1", B, 256, 473, 48, 160, 3, 1, 1, bias=False, act=False, time_it=True)
    case_wgrad("wgrad conv3_1", B, 473, 256, 48, 160, 3, 1, (1, 1, 1, 1), pitch_x=476, time_it=True)
    case_wgrad("wgrad conv4_1", B, 512, 512, 24, 80, 3, 1, (1, 1, 1, 1), time_it=True)
    case_wgrad("wgrad conv6_1", B, 1024, 1024, 6, 20, 3, 1, (1, 1, 1, 1), time_it=True)
    case_wgrad("wgrad conv4 (s2)", B, 256, 512, 48, 160, 3, 2, (0, 1, 0, 1), time_it=True)
    case_wgrad("wgrad conv2 (5x5 s2)", B, 64, 128, 192, 640, 5, 2, (1, 2, 1, 2), time_it=True)
    case_wgrad("wgrad deconv2", B, 386, 64, 48, 160, 4, 2, None, pitch_x=388, time_it=True, deconv=True)
    case_wgrad("wgrad deconv5", B, 1024, 512, 6, 20, 4, 2, None, time_it=True, deconv=True)
    say(phase="done", seconds=round(time.time() - t0, 1))


if __name__ == "__main__":
    main()
