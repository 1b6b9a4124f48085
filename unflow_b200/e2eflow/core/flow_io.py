"""On-disk flow formats either side of the path:

* KITTI 16-bit PNG flow: ``flow = (png[:, :, 0:2] - 2**15) / 64``, validity in channel 2
  (reference src/e2eflow/kitti/input.py:12-22; writer src/eval_gui.py:60-76);
* Middlebury ``.flo``: magic 202021.25 (b'PIEH'), int32 width, int32 height, float32 (u,v) rows
  (reference src/e2eflow/middlebury/input.py:10-28; writer src/eval_gui.py:79-93);
* ``resize_output_flow``: bilinear resize + rescale of (u,v) (src/e2eflow/core/input.py:28-34).
PNG coding uses a minimal self-contained 16-bit RGB PNG reader/writer (zlib) -- no image library
is required.
"""
import struct
import zlib

import numpy as np
import torch

from . import tf_image

FLO_MAGIC = 202021.25


def write_flo(path, flow):
    flow = np.asarray(flow, dtype=np.float32)
    h, w, _ = flow.shape
    with open(path, 'wb') as f:
        f.write(struct.pack('<f', FLO_MAGIC))
        f.write(struct.pack('<ii', w, h))
        f.write(flow.astype('<f4').tobytes())


def read_flo(path):
    """-> (flow [h,w,2] float32, mask [h,w,1] float32: both components < 1e9)."""
    with open(path, 'rb') as f:
        data = f.read()
    magic = struct.unpack('<f', data[:4])[0]
    if abs(magic - FLO_MAGIC) > 1e-3:
        raise ValueError("not a .flo file")
    w, h = struct.unpack('<ii', data[4:12])
    flow = np.frombuffer(data, dtype='<f4', count=2 * w * h, offset=12).reshape(h, w, 2).copy()
    mask = ((flow[:, :, 0] < 1e9) & (flow[:, :, 1] < 1e9)).astype(np.float32)[:, :, None]
    return flow, mask


def _png_chunk(tag, payload):
    return struct.pack('>I', len(payload)) + tag + payload + struct.pack('>I', zlib.crc32(tag + payload) & 0xffffffff)


def write_png16(path, arr):
    """arr: [h,w,3] uint16 -> 16-bit RGB PNG."""
    arr = np.asarray(arr, dtype=np.uint16)
    h, w, c = arr.shape
    assert c == 3
    raw = arr.astype('>u2').reshape(h, w * 3 * 2 // 2)
    rows = b''.join(b'\x00' + raw[y].tobytes() for y in range(h))
    png = b'\x89PNG\r\n\x1a\n' + _png_chunk(b'IHDR', struct.pack('>IIBBBBB', w, h, 16, 2, 0, 0, 0))
    png += _png_chunk(b'IDAT', zlib.compress(rows, 6)) + _png_chunk(b'IEND', b'')
    with open(path, 'wb') as f:
        f.write(png)


def read_png16(path):
    """16-bit RGB, non-interlaced PNG -> [h,w,3] uint16 (all five PNG filter types)."""
    with open(path, 'rb') as f:
        data = f.read()
    if data[:8] != b'\x89PNG\r\n\x1a\n':
        raise ValueError("not a PNG file")
    pos, idat, hdr = 8, [], None
    while pos < len(data):
        n, tag = struct.unpack('>I4s', data[pos:pos + 8])
        body = data[pos + 8:pos + 8 + n]
        if tag == b'IHDR':
            hdr = struct.unpack('>IIBBBBB', body)
        elif tag == b'IDAT':
            idat.append(body)
        pos += 12 + n
    w, h, depth, ctype, _, _, interlace = hdr
    if depth != 16 or ctype != 2 or interlace != 0:
        raise ValueError("expected a 16-bit RGB non-interlaced PNG")
    bpp, stride = 6, w * 6
    raw = np.frombuffer(zlib.decompress(b''.join(idat)), dtype=np.uint8)
    out = np.zeros((h, stride), dtype=np.uint8)
    prev = np.zeros(stride, dtype=np.int32)
    p = 0
    for y in range(h):
        ft = raw[p]
        line = raw[p + 1:p + 1 + stride].astype(np.int32)
        p += 1 + stride
        if ft == 0:
            cur = line
        elif ft == 2:
            cur = (line + prev) & 255
        else:
            cur = np.zeros(stride, dtype=np.int32)
            for i in range(stride):
                a = cur[i - bpp] if i >= bpp else 0
                b = prev[i]
                c = prev[i - bpp] if i >= bpp else 0
                if ft == 1:
                    pred = a
                elif ft == 3:
                    pred = (a + b) >> 1
                else:
                    pa, pb, pc = abs(b - c), abs(a - c), abs(a + b - 2 * c)
                    pred = a if (pa <= pb and pa <= pc) else (b if pb <= pc else c)
                cur[i] = (line[i] + pred) & 255
        out[y] = cur
        prev = cur
    return out.reshape(h, w, 3, 2).astype(np.uint16).dot(np.array([256, 1], dtype=np.uint16)).astype(np.uint16)


def write_kitti_flow(path, flow, mask=None):
    """eval_gui.py:60-76: u,v -> uint16 (64*f + 2^15), validity in channel 2."""
    flow = np.asarray(flow, dtype=np.float64)
    h, w, _ = flow.shape
    valid = np.ones((h, w)) if mask is None else np.asarray(mask).reshape(h, w)
    enc = np.clip(np.rint(flow * 64.0 + 2 ** 15), 0, 65535)
    arr = np.stack([enc[:, :, 0], enc[:, :, 1], valid], 2).astype(np.uint16)
    write_png16(path, arr)


def read_kitti_flow(path):
    """kitti/input.py:12-22 -> (flow [h,w,2] float32, mask [h,w,1] float32)."""
    gt = None
    try:                                   # OpenCV decodes a 375x1242 ground-truth file in milliseconds
        import cv2
        raw = cv2.imread(path, cv2.IMREAD_UNCHANGED)
        if raw is not None and raw.dtype == np.uint16 and raw.ndim == 3 and raw.shape[2] == 3:
            gt = raw[:, :, ::-1].astype(np.float32)       # BGR -> (u, v, valid)
    except ImportError:
        pass
    if gt is None:
        gt = read_png16(path).astype(np.float32)
    flow = (gt[:, :, 0:2] - 2 ** 15) / 64.0
    mask = gt[:, :, 2:3]
    return flow, mask


def resize_output_flow(t, height, width, channels=2):
    """core/input.py:28-34: bilinear (TF1 legacy) resize to [height,width] and rescale u, v."""
    _, old_height, old_width, _ = t.shape
    t = tf_image.resize_bilinear(t, [height, width])
    scale = torch.tensor([width / old_width, height / old_height], device=t.device, dtype=t.dtype)
    return t * scale
