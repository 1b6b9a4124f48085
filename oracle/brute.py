"""Independent float64 definitions used to pin the oracle where the reference
has no known-answer test (TEST INFRASTRUCTURE ONLY).

These are written from the mathematical definition of each op (SURVEY.md
section 2.2 formulas), NOT from the kernel loop structure, so that a shared
transcription error between oracle_ops.c and the CUDA product cannot hide.
"""
import numpy as np


def correlation(in0, in1, kernel_size=1, max_displacement=20, pad=20, stride_1=1, stride_2=2):
    """out[b,(p,o),y,x] = 1/(K^2 C) sum_{j,i,c} P0[b,c,y*s1+md+j, x*s1+md+i] *
    P1[b,c,y*s1+md+s2*p+j, x*s1+md+s2*o+i], P = zero padded by `pad`; j,i in [0,K)."""
    in0 = np.asarray(in0, np.float64)
    in1 = np.asarray(in1, np.float64)
    B, C, H, W = in0.shape
    K, md, s1, s2 = kernel_size, max_displacement, stride_1, stride_2
    kr = (K - 1) // 2
    border = md + kr
    r = md // s2
    D = 2 * r + 1
    ph, pw = H + 2 * pad, W + 2 * pad
    oh = int(np.ceil((ph - 2 * border) / s1))
    ow = int(np.ceil((pw - 2 * border) / s1))
    # extra margin so displaced windows never index out of the array
    m = max(0, md + K)  # generous
    P0 = np.zeros((B, C, ph + 2 * m, pw + 2 * m))
    P1 = np.zeros_like(P0)
    P0[:, :, m + pad:m + pad + H, m + pad:m + pad + W] = in0
    P1[:, :, m + pad:m + pad + H, m + pad:m + pad + W] = in1
    out = np.zeros((B, D * D, oh, ow))
    ys = np.arange(oh) * s1 + md + m
    xs = np.arange(ow) * s1 + md + m
    for pi, p in enumerate(range(-r, r + 1)):
        for oi, o in enumerate(range(-r, r + 1)):
            acc = np.zeros((B, oh, ow))
            for j in range(K):
                for i in range(K):
                    a = P0[:, :, (ys + j)[:, None], (xs + i)[None, :]]
                    b = P1[:, :, (ys + s2 * p + j)[:, None], (xs + s2 * o + i)[None, :]]
                    acc += np.sum(a * b, axis=1)
            out[:, pi * D + oi] = acc / (K * K * C)
    return out


def backward_warp_zero(images, flows):
    """Bilinear sample at (x+u, y+v); taps outside the image contribute 0."""
    images = np.asarray(images, np.float64)
    flows = np.asarray(flows, np.float64)
    B, H, W, C = images.shape
    out = np.zeros_like(images)
    for b in range(B):
        for y in range(H):
            for x in range(W):
                fx = x + flows[b, y, x, 0]
                fy = y + flows[b, y, x, 1]
                x0, y0 = int(np.floor(fx)), int(np.floor(fy))
                for (yy, xx, w) in ((y0, x0, (x0 + 1 - fx) * (y0 + 1 - fy)),
                                    (y0, x0 + 1, (fx - x0) * (y0 + 1 - fy)),
                                    (y0 + 1, x0, (x0 + 1 - fx) * (fy - y0)),
                                    (y0 + 1, x0 + 1, (fx - x0) * (fy - y0))):
                    if 0 <= xx < W and 0 <= yy < H:
                        out[b, y, x] += w * images[b, yy, xx]
    return out


def backward_warp_clamp(images, flows):
    """Bilinear sample with tap indices clamped to the image (image_warp semantics)."""
    images = np.asarray(images, np.float64)
    flows = np.asarray(flows, np.float64)
    B, H, W, C = images.shape
    out = np.zeros_like(images)
    for b in range(B):
        for y in range(H):
            for x in range(W):
                u, v = flows[b, y, x]
                fu, fv = np.floor(u), np.floor(v)
                xw, yw = u - fu, v - fv
                x0, y0 = x + int(fu), y + int(fv)
                cx = lambda t: min(max(t, 0), W - 1)
                cy = lambda t: min(max(t, 0), H - 1)
                out[b, y, x] = ((1 - xw) * (1 - yw) * images[b, cy(y0), cx(x0)] +
                                (1 - xw) * yw * images[b, cy(y0 + 1), cx(x0)] +
                                xw * (1 - yw) * images[b, cy(y0), cx(x0 + 1)] +
                                xw * yw * images[b, cy(y0 + 1), cx(x0 + 1)])
    return out


def forward_warp(flows):
    """out[b,ny,nx] = sum over source pixels of exp(-((nx-tx)^2+(ny-ty)^2)/2) for
    nx in [floor(tx-4), floor(tx+4)] (clipped), same in y; sigma=1, radius 4."""
    flows = np.asarray(flows, np.float64)
    B, H, W, _ = flows.shape
    out = np.zeros((B, H, W, 1))
    k = 4
    for b in range(B):
        for y in range(H):
            for x in range(W):
                tx = x + flows[b, y, x, 0]
                ty = y + flows[b, y, x, 1]
                if not (np.floor(tx - k) < W and np.floor(tx + k) >= 0 and
                        np.floor(ty - k) < H and np.floor(ty + k) >= 0):
                    continue
                x_lo = int(np.floor(tx - k)) if tx - k > 0 else 0
                y_lo = int(np.floor(ty - k)) if ty - k > 0 else 0
                x_hi = int(np.floor(tx + k)) if tx + k < W else W - 1
                y_hi = int(np.floor(ty + k)) if ty + k < H else H - 1
                for nx in range(x_lo, x_hi + 1):
                    for ny in range(y_lo, y_hi + 1):
                        out[b, ny, nx, 0] += np.exp(-((nx - tx) ** 2 + (ny - ty) ** 2) / 2.0)
    return out


def downsample(images, scale):
    images = np.asarray(images, np.float64)
    B, H, W, C = images.shape
    return images.reshape(B, H // scale, scale, W // scale, scale, C).mean(axis=(2, 4))
