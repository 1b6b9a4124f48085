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


# ---------------------------------------------------------------------------------------------
# Loss terms from their mathematical definition (float64, explicit loops).  Independent of the
# conv2d / gather formulation the oracle (and the reference) use.
# ---------------------------------------------------------------------------------------------
_GRAY = (0.2989, 0.5870, 0.1140)


def charbonnier(x, mask=None, beta=1.0, alpha=0.45, eps=0.001):
    x = np.asarray(x, np.float64)
    e = ((x * beta) ** 2 + eps ** 2) ** alpha
    if mask is not None:
        e = e * np.asarray(mask, np.float64)
    return e.sum() / x.size


def ternary_loss(im1, im2w, mask, max_distance):
    """sum_q mask(q) * interior(q) * charb( sum_k d_k/(0.1+d_k) ),  d_k = (t1_k - t2_k)^2,
    t_k = s_k / sqrt(0.81 + s_k^2),  s_k = I(q+k) - I(q) with zero intensity outside the image."""
    im1, im2w, mask = (np.asarray(a, np.float64) for a in (im1, im2w, mask))
    B, H, W, _ = im1.shape
    r = max_distance
    g1 = (im1[..., 0] * _GRAY[0] + im1[..., 1] * _GRAY[1] + im1[..., 2] * _GRAY[2]) * 255
    g2 = (im2w[..., 0] * _GRAY[0] + im2w[..., 1] * _GRAY[1] + im2w[..., 2] * _GRAY[2]) * 255
    total = 0.0
    for b in range(B):
        for y in range(H):
            for x in range(W):
                if not (r <= y < H - r and r <= x < W - r):
                    continue                      # transform mask (create_mask) removes the border
                dist = 0.0
                for dy in range(-r, r + 1):
                    for dx in range(-r, r + 1):
                        s1 = g1[b, y + dy, x + dx] - g1[b, y, x]
                        s2 = g2[b, y + dy, x + dx] - g2[b, y, x]
                        t1 = s1 / np.sqrt(0.81 + s1 * s1)
                        t2 = s2 / np.sqrt(0.81 + s2 * s2)
                        d = (t1 - t2) ** 2
                        dist += d / (0.1 + d)
                total += mask[b, y, x, 0] * (dist ** 2 + 1e-6) ** 0.45
    return total / (B * H * W)


def second_order_loss(flow):
    flow = np.asarray(flow, np.float64)
    B, H, W, _ = flow.shape
    total = 0.0
    for c in range(2):
        f = flow[..., c]
        for b in range(B):
            for y in range(H):
                for x in range(W):
                    ix, iy = 1 <= x < W - 1, 1 <= y < H - 1
                    if ix:
                        total += ((f[b, y, x - 1] + f[b, y, x + 1] - 2 * f[b, y, x]) ** 2 + 1e-6) ** 0.45
                    if iy:
                        total += ((f[b, y - 1, x] + f[b, y + 1, x] - 2 * f[b, y, x]) ** 2 + 1e-6) ** 0.45
                    if ix and iy:
                        total += ((f[b, y - 1, x - 1] + f[b, y + 1, x + 1] - 2 * f[b, y, x]) ** 2 + 1e-6) ** 0.45
                        total += ((f[b, y - 1, x + 1] + f[b, y + 1, x - 1] - 2 * f[b, y, x]) ** 2 + 1e-6) ** 0.45
    return total / (B * H * W * 4)


def first_order_loss(flow):
    flow = np.asarray(flow, np.float64)
    B, H, W, _ = flow.shape
    total = 0.0
    for c in range(2):
        f = flow[..., c]
        total += (((f[:, :, :-1] - f[:, :, 1:]) ** 2 + 1e-6) ** 0.45).sum()
        total += (((f[:, :-1, :] - f[:, 1:, :]) ** 2 + 1e-6) ** 0.45).sum()
    return total / (B * H * W * 2)


def compute_losses(im1, im2, flow_fw, flow_bw, border_mask=None, mask_occlusion='', data_max_distance=1):
    """The terms of losses.py:16-87 that do not need the splat map (occ, photo, smooth_1st,
    smooth_2nd, fb, ternary), from the definitions above."""
    im1, im2, flow_fw, flow_bw = (np.asarray(a, np.float64) for a in (im1, im2, flow_fw, flow_bw))
    B, H, W, _ = im1.shape
    out = {k: 0.0 for k in ('occ', 'photo', 'smooth_1st', 'smooth_2nd', 'fb', 'ternary')}
    masks = []
    for (A, Bi, f, g) in ((im1, im2, flow_fw, flow_bw), (im2, im1, flow_bw, flow_fw)):
        Bw = backward_warp_clamp(Bi, f)
        gw = backward_warp_clamp(g, f)
        fd = f + gw
        if border_mask is None:
            ys, xs = np.meshgrid(np.arange(H), np.arange(W), indexing='ij')
            px, py = xs[None] + f[..., 0], ys[None] + f[..., 1]
            m = ((px <= W - 1) & (px >= 0) & (py <= H - 1) & (py >= 0)).astype(np.float64)[..., None]
        else:
            m = np.asarray(border_mask, np.float64).copy()
        if mask_occlusion == 'fb':
            lsq = (fd ** 2).sum(-1, keepdims=True)
            mag = (f ** 2).sum(-1, keepdims=True) + (gw ** 2).sum(-1, keepdims=True)
            m = m * (1 - (lsq > 0.01 * mag + 0.5))
        masks.append(m)
        out['occ'] += charbonnier(1 - m)
        out['photo'] += charbonnier(A - Bw, m, beta=255)
        out['fb'] += charbonnier(fd, m)
        out['smooth_1st'] += first_order_loss(f)
        out['smooth_2nd'] += second_order_loss(f)
        out['ternary'] += ternary_loss(A, Bw, m, data_max_distance)
    return out, masks
