"""Index model of the staging / shared-memory layout of csrc/narrow_conv.cu, executed on the CPU.

The CUDA kernels cannot run here; what can be checked without a GPU is the integer arithmetic the
design rests on: the staging loader (row-wise cp.async, with a pixel pitch for channel slices)
covers each element of the (16+2)x(32+2)x16 window exactly once with the right source element, and
the strides chosen for the layouts (forward 706/38, wgrad 649/36, gradient tile pitch 66) give the
bank-conflict-free access patterns the comments claim.  The constants are parsed from the source so
the model cannot drift silently."""
import os
import re

import numpy as np

SRC = open(os.path.join(os.path.dirname(__file__), "..", "unflow_b200", "csrc", "narrow_conv.cu")).read()


def const(name):
    m = re.search(r"\b%s\s*=\s*([0-9]+)" % name, SRC)
    assert m, name
    return int(m.group(1))


TH, TW, KC, THREADS = const("TH"), const("TW"), const("KC"), const("THREADS")
SR, SC = TH + 2, TW + 2
PITCH, FWD_PITCH, FWD_CHS = const("PITCH"), const("FWD_PITCH"), const("FWD_CHS")
WG_CHS = SR * PITCH + 1
GPITCH = 2 * TW + 2


def window_reference(H, W, C, y0, x0, c0, XP=None):
    """ref[ch, pr, pc] = linear index into x[n] (pixel-major with pitch XP, channel-minor) or -1 for a zero."""
    XP = C if XP is None else XP
    ref = -np.ones((KC, SR, SC), dtype=np.int64)
    for pr in range(SR):
        for pc in range(SC):
            gy, gx = y0 - 1 + pr, x0 - 1 + pc
            if 0 <= gy < H and 0 <= gx < W:
                for ch in range(KC):
                    if c0 + ch < C:
                        ref[ch, pr, pc] = (gy * W + gx) * XP + c0 + ch
    return ref


def loader_rows(H, W, C, y0, x0, c0, CHS, P, XP=None):
    XP = C if XP is None else XP
    out = {}
    for tid in range(THREADS):
        ch, q = tid & 15, tid >> 4
        for pr in range(SR):
            gy = y0 - 1 + pr
            for j in range(5):
                pc = q + 8 * j
                if pc >= SC:
                    continue
                gx = x0 - 1 + pc
                ok = c0 + ch < C and 0 <= gy < H and 0 <= gx < W
                addr = ch * CHS + pr * P + q + 8 * j
                assert addr not in out
                out[addr] = (gy * W + gx) * XP + c0 + ch if ok else -1
    return out


def test_every_loader_fills_the_window_exactly():
    cases = [(37, 70, 194, 0, 0, 0, None), (37, 70, 194, 32, 64, 192, 196), (48, 160, 386, 16, 128, 368, 388),
             (5, 33, 16, 0, 32, 0, None), (16, 32, 18, 0, 0, 16, 20)]
    for H, W, C, y0, x0, c0, XP in cases:
        ref = window_reference(H, W, C, y0, x0, c0, XP)
        for CHS, P in ((FWD_CHS, FWD_PITCH), (WG_CHS, PITCH)):
            for loader in (loader_rows,):
                got = loader(H, W, C, y0, x0, c0, CHS, P, XP)
                assert len(got) == KC * SR * SC
                for ch in range(KC):
                    for pr in range(SR):
                        for pc in range(SC):
                            assert got[ch * CHS + pr * P + pc] == ref[ch, pr, pc], (loader.__name__, ch, pr, pc)


def banks(addresses_in_floats, width_floats):
    """Banks (4-byte words mod 32) touched by one access of ``width_floats`` per lane."""
    return [[(a + k) % 32 for k in range(width_floats)] for a in addresses_in_floats]


def conflict_free(per_lane_banks):
    flat = [b for lane in per_lane_banks for b in lane]
    return len(flat) == len(set(flat))


def test_layout_strides_are_bank_conflict_free():
    assert FWD_CHS >= SR * FWD_PITCH and FWD_CHS % 32 == 2 and FWD_PITCH % 4 == 2 and FWD_CHS % 2 == 0
    assert WG_CHS % 2 == 1 and (KC * WG_CHS) % 2 == 0 and GPITCH % 2 == 0
    # forward compute reads: thread (row = tid>>3, cg = tid&7) reads 8-byte pairs at 4*cg (+2, +4);
    # an LDS.64 is served per half-warp
    for ky in range(3):
        for off in (0, 2, 4):
            for half in range(2):
                lanes = range(16 * half, 16 * half + 16)
                addrs = [((t >> 3) + ky) * FWD_PITCH + 4 * (t & 7) + off for t in lanes]
                assert conflict_free(banks(addrs, 2))
    # staging stores (4-byte): async loaders, lanes = 16 channels x 2 adjacent pixels
    for CHS, P in ((FWD_CHS, FWD_PITCH),):
        for pix in range(0, SC - 1):
            addrs = [(t & 15) * CHS + pix + (t >> 4) for t in range(32)]
            assert conflict_free(banks(addrs, 1))
    # wgrad compute reads (4-byte): lanes = 16 channels x 2 row groups 4 rows apart, any row / column
    for r in range(0, 4):
        for c in range(0, SC):
            addrs = [(t & 15) * WG_CHS + (r + 4 * (t >> 4)) * PITCH + c for t in range(32)]
            assert conflict_free(banks(addrs, 1))
    # wgrad gradient tile: the two row groups of a warp read different 8-byte words without conflict
    for c in range(TW):
        addrs = sorted({(r + 4 * g) * GPITCH + 2 * c for g in range(2) for r in (0,)})
        assert conflict_free(banks(addrs, 2))


def test_wgrad_row_assignment_covers_the_tile_once():
    rows = []
    for tid in range(THREADS):
        h = tid >> 4
        base = (h >> 1) + 4 * (h & 1)
        if tid & 15 == 0:
            rows += [base, base + 8]
    assert sorted(rows) == list(range(TH))
    # the two row groups inside a warp are 4 rows apart (the bank argument above relies on it)
    for w in range(THREADS // 32):
        h0, h1 = 2 * w, 2 * w + 1
        assert ((h1 >> 1) + 4 * (h1 & 1)) - ((h0 >> 1) + 4 * (h0 & 1)) == 4
