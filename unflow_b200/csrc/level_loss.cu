// level_loss.cu -- one pyramid level of the bidirectional unsupervised loss, fused.
//
// Replaces the reference's compute_losses graph (src/e2eflow/core/losses.py:16-87: hundreds of
// small TF kernels per level, every intermediate through HBM) with ONE pixel-parallel forward
// kernel and ONE backward kernel per level.  Per pixel and per direction (fw: A=im1,B=im2,
// f=flow_fw,g=flow_bw; bw: the mirror image) the forward kernel does
//   * image_warp(B, f) and image_warp(g, f)            (image_warp.py:4-76, clamped taps)
//   * forward/backward consistency occlusion + masks   (losses.py:31-59)
//   * occ / sym / photo / fb charbonnier terms          (losses.py:61-79, 298-322)
//   * 1st- and 2nd-order smoothness                     (losses.py:206-222, 250-295)
//   * the soft census ("ternary") data term over a PxP patch, P = 2*max_distance+1 <= 7
//                                                       (losses.py:90-122)
// and reduces the 8 named loss terms deterministically (per-CTA partials, last CTA sums them in
// a fixed order in double).  The `grad` (Sobel) term -- "NOT TESTED" in the reference's config and
// disabled everywhere -- is not fused; callers that request it use the unfused path.
//
// HBM-bound by design: algorithmic bytes fwd = 44*B*h*w read (im1 3, im2 3, 2 flows 4, mask 1)
// + 16*B*h*w written (two saved planes per direction for the backward pass);
// bwd = 60*B*h*w + 16*B*h*w.  The census stencil runs out of shared memory tiles
// (tile + halo of max_distance); gathers of the warps hit L1/L2.
//
// Arithmetic notes.  Mask decisions (fb occlusion `>`, outgoing mask `<=`/`>=`, disocclusion
// `<`) are evaluated with explicitly un-contracted float ops (__fmul_rn/__fadd_rn) in the
// reference's operation order, so the binary masks are bit-exact against an op-by-op float32
// evaluation of the reference graph (TF executes one kernel per op: no FMA contraction).
#include "common.cuh"

namespace unflow {
namespace ll {

constexpr int TX = 32, TY = 8, NT = TX * TY;
constexpr int RMAXL = 3;
constexpr int RW = TX + 2 * RMAXL, RH = TY + 2 * RMAXL;  // 38 x 14
enum { T_SYM = 0, T_OCC, T_PHOTO, T_GRAD, T_S1, T_S2, T_FB, T_TERN, NTERMS };
enum { OCC_NONE = 0, OCC_FB = 1, OCC_DISOCC = 2 };

constexpr float kEps2 = 1e-6f;   // epsilon^2, epsilon = 0.001 (losses.py:298)
constexpr float kAlpha = 0.45f;
constexpr float kGrayR = 0.2989f, kGrayG = 0.5870f, kGrayB = 0.1140f;
constexpr float kDisoccThresh = 0.8f;

struct Params {
  const float *im1, *im2, *ffw, *fbw, *border, *fwarp_fw, *fwarp_bw;
  float *partials;          // [nblocks][NTERMS]
  unsigned *counter;        // last-CTA-done ticket
  float *losses;            // [NTERMS]
  float *saved;             // [4][B*h*w]: g2w_fw, W_fw, g1w_bw, W_bw
  float *masks_out;         // optional [2][B*h*w] (mask_fw, mask_bw) for parity tests
  const float *gl;          // bwd: upstream dL/dloss_k [NTERMS]
  float *dffw, *dfbw;       // bwd outputs (accumulated with atomics, pre-zeroed)
  int B, h, w, occl, r;
  unsigned terms;           // bit k set = term k requested
};

struct Taps {
  int ia, ib, ic, id;  // pixel indices (y*w+x) of (y0,x0) (y1,x0) (y0,x1) (y1,x1)
  float wa, wb, wc, wd, xw, yw, omx, omy;
};

// image_warp.py:26-54, op-by-op float32 rounding
__device__ __forceinline__ Taps clamp_taps(float u, float v, int x, int y, int h, int w) {
  Taps t;
  const float fu = floorf(u), fv = floorf(v);
  t.xw = __fsub_rn(u, fu);
  t.yw = __fsub_rn(v, fv);
  const int x0 = x + (int)fu, y0 = y + (int)fv;
  const int cx0 = min(max(x0, 0), w - 1), cx1 = min(max(x0 + 1, 0), w - 1);
  const int cy0 = min(max(y0, 0), h - 1), cy1 = min(max(y0 + 1, 0), h - 1);
  t.ia = cy0 * w + cx0; t.ib = cy1 * w + cx0; t.ic = cy0 * w + cx1; t.id = cy1 * w + cx1;
  t.omx = __fsub_rn(1.0f, t.xw);
  t.omy = __fsub_rn(1.0f, t.yw);
  t.wa = __fmul_rn(t.omx, t.omy);
  t.wb = __fmul_rn(t.omx, t.yw);
  t.wc = __fmul_rn(t.xw, t.omy);
  t.wd = __fmul_rn(t.xw, t.yw);
  return t;
}
__device__ __forceinline__ float bil(const Taps &t, float Ia, float Ib, float Ic, float Id) {
  // tf.add_n([wa*Ia, wb*Ib, wc*Ic, wd*Id])
  return __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(t.wa, Ia), __fmul_rn(t.wb, Ib)), __fmul_rn(t.wc, Ic)),
                   __fmul_rn(t.wd, Id));
}
__device__ __forceinline__ float gray255(float r, float g, float b) {
  return __fmul_rn(__fadd_rn(__fadd_rn(__fmul_rn(r, kGrayR), __fmul_rn(g, kGrayG)), __fmul_rn(b, kGrayB)), 255.0f);
}
// generalized charbonnier of one element: ((x*beta)^2 + eps^2)^alpha
// The argument is >= eps^2 = 1e-6 > 0, so pow is exp2(alpha*log2(.)): two MUFU ops instead of the
// ~60-instruction IEEE powf (relative error ~1e-6, far inside the 1e-4 loss tolerance).
__device__ __forceinline__ float pow_pos(float x, float a) { return exp2f(a * __log2f(x)); }
__device__ __forceinline__ float charb(float x) { return pow_pos(__fadd_rn(__fmul_rn(x, x), kEps2), kAlpha); }
// d/dx ((x*beta)^2+eps^2)^alpha
__device__ __forceinline__ float charb_grad(float x, float beta) {
  const float xb = x * beta;
  return 2.0f * kAlpha * xb * beta * pow_pos(xb * xb + kEps2, kAlpha - 1.0f);
}
__device__ __forceinline__ float tern_t(float s) { return s * rsqrtf(0.81f + s * s); }
// d/ds2 of  d/(0.1+d), d = (t(s1)-t(s2))^2
__device__ __forceinline__ float tern_phi(float s1, float s2) {
  const float q2 = 0.81f + s2 * s2;
  const float rs2 = rsqrtf(q2);
  const float t1 = s1 * rsqrtf(0.81f + s1 * s1), t2 = s2 * rs2;
  const float dt = t1 - t2, d = dt * dt;
  const float den = 0.1f + d;
  return __fdividef(0.1f, den * den) * (-2.0f * dt) * (0.81f * rs2 * rs2 * rs2);
}

struct DirView {
  const float *A, *Bimg, *f, *g, *fwarp_other;  // per image-b base pointers
};

__device__ __forceinline__ bool inside_mask(float2 f, int x, int y, int h, int w) {
  // create_outgoing_mask (losses.py:347-366)
  const float px = __fadd_rn((float)x, f.x), py = __fadd_rn((float)y, f.y);
  return px <= (float)(w - 1) && px >= 0.0f && py <= (float)(h - 1) && py >= 0.0f;
}

// mask of one direction at pixel q (losses.py:31-59); also returns flow_diff and the taps
__device__ __forceinline__ float dir_mask(const Params &p, const DirView &d, int q, int x, int y,
                                          float2 f, const Taps &t, float2 &fd, float2 &gw) {
  const float2 *g2 = reinterpret_cast<const float2 *>(d.g);
  const float2 ga = __ldg(g2 + t.ia), gb = __ldg(g2 + t.ib), gc = __ldg(g2 + t.ic), gd = __ldg(g2 + t.id);
  gw.x = bil(t, ga.x, gb.x, gc.x, gd.x);
  gw.y = bil(t, ga.y, gb.y, gc.y, gd.y);
  fd.x = __fadd_rn(f.x, gw.x);
  fd.y = __fadd_rn(f.y, gw.y);
  float m = p.border ? __ldg(p.border + q) : (inside_mask(f, x, y, p.h, p.w) ? 1.0f : 0.0f);
  if (p.occl == OCC_FB) {
    const float mag = __fadd_rn(__fadd_rn(__fmul_rn(f.x, f.x), __fmul_rn(f.y, f.y)),
                                __fadd_rn(__fmul_rn(gw.x, gw.x), __fmul_rn(gw.y, gw.y)));
    const float thresh = __fadd_rn(__fmul_rn(0.01f, mag), 0.5f);
    const float lsq = __fadd_rn(__fmul_rn(fd.x, fd.x), __fmul_rn(fd.y, fd.y));
    m = __fmul_rn(m, lsq > thresh ? 0.0f : 1.0f);
  } else if (p.occl == OCC_DISOCC) {
    m = __fmul_rn(m, __ldg(d.fwarp_other + q) < kDisoccThresh ? 0.0f : 1.0f);
  }
  return m;
}

// ---------------------------------------------------------------------------------------------
// Forward
// ---------------------------------------------------------------------------------------------
template <int R>
__global__ void __launch_bounds__(NT)
level_loss_fwd_kernel(Params p) {
  __shared__ float sg[4][RH][RW];  // 0: gray(im1) 1: gray(im2) 2: gray(im2 warped by ffw) 3: gray(im1 warped by fbw)
  __shared__ double sred[NT / 32][NTERMS];
  __shared__ bool s_last;
  const int tid = threadIdx.y * TX + threadIdx.x;
  const int b = blockIdx.z;
  const int tx0 = blockIdx.x * TX, ty0 = blockIdx.y * TY;
  const int h = p.h, w = p.w;
  const long long npix = (long long)h * w, ioff = (long long)b * npix;
  const float *im1 = p.im1 + ioff * 3, *im2 = p.im2 + ioff * 3;
  const float *ffw = p.ffw + ioff * 2, *fbw = p.fbw + ioff * 2;
  const bool want_tern = (p.terms >> T_TERN) & 1u;

  // ---- phase 1: grey tiles (tile + halo R) of both images and both warped images ----
  if (want_tern) {
    constexpr int rw = TX + 2 * R, rh = TY + 2 * R;
    for (int i = tid; i < rw * rh; i += NT) {
      const int ly = i / rw, lx = i - ly * rw;
      const int y = ty0 - R + ly, x = tx0 - R + lx;
      float g1 = 0.f, g2 = 0.f, g2w = 0.f, g1w = 0.f;
      if (x >= 0 && x < w && y >= 0 && y < h) {
        const int q = y * w + x;
        g1 = gray255(__ldg(im1 + q * 3), __ldg(im1 + q * 3 + 1), __ldg(im1 + q * 3 + 2));
        g2 = gray255(__ldg(im2 + q * 3), __ldg(im2 + q * 3 + 1), __ldg(im2 + q * 3 + 2));
        {
          const float2 f = __ldg(reinterpret_cast<const float2 *>(ffw) + q);
          const Taps t = clamp_taps(f.x, f.y, x, y, h, w);
          float c[3];
#pragma unroll
          for (int k = 0; k < 3; ++k)
            c[k] = bil(t, __ldg(im2 + t.ia * 3 + k), __ldg(im2 + t.ib * 3 + k), __ldg(im2 + t.ic * 3 + k),
                       __ldg(im2 + t.id * 3 + k));
          g2w = gray255(c[0], c[1], c[2]);
        }
        {
          const float2 f = __ldg(reinterpret_cast<const float2 *>(fbw) + q);
          const Taps t = clamp_taps(f.x, f.y, x, y, h, w);
          float c[3];
#pragma unroll
          for (int k = 0; k < 3; ++k)
            c[k] = bil(t, __ldg(im1 + t.ia * 3 + k), __ldg(im1 + t.ib * 3 + k), __ldg(im1 + t.ic * 3 + k),
                       __ldg(im1 + t.id * 3 + k));
          g1w = gray255(c[0], c[1], c[2]);
        }
      }
      sg[0][ly][lx] = g1; sg[1][ly][lx] = g2; sg[2][ly][lx] = g2w; sg[3][ly][lx] = g1w;
    }
  }
  __syncthreads();

  // ---- phase 2: one pixel per thread, both directions ----
  float acc[NTERMS];
#pragma unroll
  for (int k = 0; k < NTERMS; ++k) acc[k] = 0.f;
  const int x = tx0 + threadIdx.x, y = ty0 + threadIdx.y;
  if (x < w && y < h) {
    const int q = y * w + x;
#pragma unroll
    for (int dir = 0; dir < 2; ++dir) {
      DirView d;
      d.A = dir == 0 ? im1 : im2;
      d.Bimg = dir == 0 ? im2 : im1;
      d.f = dir == 0 ? ffw : fbw;
      d.g = dir == 0 ? fbw : ffw;
      d.fwarp_other = dir == 0 ? p.fwarp_bw : p.fwarp_fw;
      if (d.fwarp_other) d.fwarp_other += ioff;
      const float2 f = __ldg(reinterpret_cast<const float2 *>(d.f) + q);
      const Taps t = clamp_taps(f.x, f.y, x, y, h, w);
      float2 fd, gw;
      Params pl = p;
      if (pl.border) pl.border += ioff;
      const float m = dir_mask(pl, d, q, x, y, f, t, fd, gw);
      if (p.masks_out) p.masks_out[(long long)dir * p.B * npix + ioff + q] = m;
      const float occ = __fsub_rn(1.0f, m);

      if ((p.terms >> T_OCC) & 1u) acc[T_OCC] += charb(occ);
      if ((p.terms >> T_SYM) & 1u) {
        const float dis = __ldg(d.fwarp_other + q) < kDisoccThresh ? 1.0f : 0.0f;
        acc[T_SYM] += charb(__fsub_rn(occ, dis));
      }
      if ((p.terms >> T_FB) & 1u) acc[T_FB] += m * (charb(fd.x) + charb(fd.y));
      if ((p.terms >> T_PHOTO) & 1u) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const float bw_ = bil(t, __ldg(d.Bimg + t.ia * 3 + k), __ldg(d.Bimg + t.ib * 3 + k),
                                __ldg(d.Bimg + t.ic * 3 + k), __ldg(d.Bimg + t.id * 3 + k));
          const float diff = __fmul_rn(__fsub_rn(__ldg(d.A + q * 3 + k), bw_), 255.0f);
          acc[T_PHOTO] += m * charb(diff);
        }
      }
      const float2 *f2 = reinterpret_cast<const float2 *>(d.f);
      if ((p.terms >> T_S1) & 1u) {
        // f(y,x) - f(y,x+1) masked on the last column; f(y,x) - f(y+1,x) masked on the last row
        if (x < w - 1) { const float2 n = __ldg(f2 + q + 1); acc[T_S1] += charb(f.x - n.x) + charb(f.y - n.y); }
        if (y < h - 1) { const float2 n = __ldg(f2 + q + w); acc[T_S1] += charb(f.x - n.x) + charb(f.y - n.y); }
      }
      if ((p.terms >> T_S2) & 1u) {
        const bool ix = x >= 1 && x < w - 1, iy = y >= 1 && y < h - 1;
        if (ix) { const float2 a = __ldg(f2 + q - 1), c = __ldg(f2 + q + 1);
                  acc[T_S2] += charb((a.x + c.x) - 2.f * f.x) + charb((a.y + c.y) - 2.f * f.y); }
        if (iy) { const float2 a = __ldg(f2 + q - w), c = __ldg(f2 + q + w);
                  acc[T_S2] += charb((a.x + c.x) - 2.f * f.x) + charb((a.y + c.y) - 2.f * f.y); }
        if (ix && iy) {
          float2 a = __ldg(f2 + q - w - 1), c = __ldg(f2 + q + w + 1);
          acc[T_S2] += charb((a.x + c.x) - 2.f * f.x) + charb((a.y + c.y) - 2.f * f.y);
          a = __ldg(f2 + q - w + 1); c = __ldg(f2 + q + w - 1);
          acc[T_S2] += charb((a.x + c.x) - 2.f * f.x) + charb((a.y + c.y) - 2.f * f.y);
        }
      }
      if (want_tern) {
        const int ly = threadIdx.y + R, lx = threadIdx.x + R;
        const float(*gA)[RW] = sg[dir == 0 ? 0 : 1];
        const float(*gB)[RW] = sg[dir == 0 ? 2 : 3];
        const bool tm = x >= R && x < w - R && y >= R && y < h - R;  // create_mask(mask, [[R,R],[R,R]])
        const float mt = tm ? m : 0.0f;
        float wq = 0.0f;
        if (mt != 0.0f) {
          const float cA = gA[ly][lx], cB = gB[ly][lx];
          float dist = 0.0f;
#pragma unroll
          for (int dy = -R; dy <= R; ++dy)
#pragma unroll
            for (int dx = -R; dx <= R; ++dx) {
              const float t1 = tern_t(gA[ly + dy][lx + dx] - cA);
              const float t2 = tern_t(gB[ly + dy][lx + dx] - cB);
              const float dd = (t1 - t2) * (t1 - t2);
              dist += __fdividef(dd, 0.1f + dd);
            }
          acc[T_TERN] += mt * charb(dist);
          wq = mt * charb_grad(dist, 1.0f);
        }
        if (p.saved) {
          float *sv = p.saved + (long long)(2 * dir) * p.B * npix + ioff;
          sv[q] = gB[ly][lx];
          sv[(long long)p.B * npix + q] = wq;
        }
      }
    }
  }

  // ---- deterministic reduction: warp shuffle -> CTA -> per-CTA partial -> last CTA sums all ----
  const int lane = tid & 31, wrp = tid >> 5;
#pragma unroll
  for (int k = 0; k < NTERMS; ++k) {
    double v = (double)acc[k];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
    if (lane == 0) sred[wrp][k] = v;
  }
  __syncthreads();
  const int nblocks = gridDim.x * gridDim.y * gridDim.z;
  const int bid = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
  if (tid < NTERMS) {
    double v = 0.0;
    for (int i = 0; i < NT / 32; ++i) v += sred[i][tid];
    p.partials[(long long)bid * NTERMS + tid] = (float)v;
    __threadfence();
  }
  __syncthreads();
  if (tid == 0) {
    const unsigned ticket = atomicAdd(p.counter, 1u);
    s_last = (ticket == (unsigned)nblocks - 1);
  }
  __syncthreads();
  if (s_last) {
    __threadfence();
    // warp k sums term k: lane-strided double partial sums, then a fixed shuffle tree
    if (wrp < NTERMS) {
      double v = 0.0;
      for (int i = lane; i < nblocks; i += 32) v += (double)__ldcg(p.partials + (long long)i * NTERMS + wrp);
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
      if (lane == 0) {
        const double n1 = (double)p.B * h * w;
        double norm = n1;                       // sym, occ, ternary: 1 channel
        if (wrp == T_PHOTO) norm = n1 * 3;      // 3 channels
        if (wrp == T_GRAD) norm = n1 * 6;
        if (wrp == T_S1 || wrp == T_FB) norm = n1 * 2;
        if (wrp == T_S2) norm = n1 * 4;
        p.losses[wrp] = (float)(v / norm);
      }
    }
    if (tid == 0) *p.counter = 0u;
  }
}

// ---------------------------------------------------------------------------------------------
// Backward: d(sum_k gl[k] * loss_k) / d(flow_fw, flow_bw)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float psi2(const float2 *f2, int q, int off, int comp, float center2) {
  // c'(delta) of a 2nd-order delta at pixel q along offset `off`
  const float2 a = __ldg(f2 + q - off), c = __ldg(f2 + q + off);
  const float delta = comp == 0 ? (a.x + c.x) - center2 : (a.y + c.y) - center2;
  return charb_grad(delta, 1.0f);
}

template <int R>
__global__ void __launch_bounds__(NT)
level_loss_bwd_kernel(Params p) {
  __shared__ float sg[6][RH][RW];  // 0: gray im1, 1: gray im2, 2: g2w (fw), 3: g1w (bw), 4: W_fw, 5: W_bw
  const int tid = threadIdx.y * TX + threadIdx.x;
  const int b = blockIdx.z;
  const int tx0 = blockIdx.x * TX, ty0 = blockIdx.y * TY;
  const int h = p.h, w = p.w;
  const long long npix = (long long)h * w, ioff = (long long)b * npix, plane = (long long)p.B * npix;
  const float *im1 = p.im1 + ioff * 3, *im2 = p.im2 + ioff * 3;
  const float *ffw = p.ffw + ioff * 2, *fbw = p.fbw + ioff * 2;
  const bool want_tern = ((p.terms >> T_TERN) & 1u) && p.saved;
  const double n1 = (double)p.B * h * w;
  const float u_photo = ((p.terms >> T_PHOTO) & 1u) ? (float)((double)__ldg(p.gl + T_PHOTO) / (n1 * 3)) : 0.f;
  const float u_fb = ((p.terms >> T_FB) & 1u) ? (float)((double)__ldg(p.gl + T_FB) / (n1 * 2)) : 0.f;
  const float u_s1 = ((p.terms >> T_S1) & 1u) ? (float)((double)__ldg(p.gl + T_S1) / (n1 * 2)) : 0.f;
  const float u_s2 = ((p.terms >> T_S2) & 1u) ? (float)((double)__ldg(p.gl + T_S2) / (n1 * 4)) : 0.f;
  const float u_tern = want_tern ? (float)((double)__ldg(p.gl + T_TERN) / n1) : 0.f;

  if (want_tern) {
    constexpr int rw = TX + 2 * R, rh = TY + 2 * R;
    for (int i = tid; i < rw * rh; i += NT) {
      const int ly = i / rw, lx = i - ly * rw;
      const int y = ty0 - R + ly, x = tx0 - R + lx;
      float v[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (x >= 0 && x < w && y >= 0 && y < h) {
        const int q = y * w + x;
        v[0] = gray255(__ldg(im1 + q * 3), __ldg(im1 + q * 3 + 1), __ldg(im1 + q * 3 + 2));
        v[1] = gray255(__ldg(im2 + q * 3), __ldg(im2 + q * 3 + 1), __ldg(im2 + q * 3 + 2));
        const float *sv = p.saved + ioff + q;
        v[2] = __ldg(sv); v[4] = __ldg(sv + plane); v[3] = __ldg(sv + 2 * plane); v[5] = __ldg(sv + 3 * plane);
      }
#pragma unroll
      for (int k = 0; k < 6; ++k) sg[k][ly][lx] = v[k];
    }
  }
  __syncthreads();

  const int x = tx0 + threadIdx.x, y = ty0 + threadIdx.y;
  if (x >= w || y >= h) return;
  const int q = y * w + x;
#pragma unroll
  for (int dir = 0; dir < 2; ++dir) {
    DirView d;
    d.A = dir == 0 ? im1 : im2;
    d.Bimg = dir == 0 ? im2 : im1;
    d.f = dir == 0 ? ffw : fbw;
    d.g = dir == 0 ? fbw : ffw;
    d.fwarp_other = dir == 0 ? p.fwarp_bw : p.fwarp_fw;
    if (d.fwarp_other) d.fwarp_other += ioff;
    float *df_out = (dir == 0 ? p.dffw : p.dfbw) + ioff * 2;   // gradient of this direction's flow
    float *dg_out = (dir == 0 ? p.dfbw : p.dffw) + ioff * 2;   // gradient of the other flow (scatter)
    const float2 *f2 = reinterpret_cast<const float2 *>(d.f);
    const float2 f = __ldg(f2 + q);
    const Taps t = clamp_taps(f.x, f.y, x, y, h, w);
    float2 fd, gw;
    Params pl = p;
    if (pl.border) pl.border += ioff;
    const float m = dir_mask(pl, d, q, x, y, f, t, fd, gw);
    float du = 0.f, dv = 0.f;

    // ---- data terms that go through the warped image: photo + ternary ----
    float dG = 0.f;  // dL/d(gray255 of warped B at q)
    if (want_tern) {
      const int ly = threadIdx.y + R, lx = threadIdx.x + R;
      const float(*gA)[RW] = sg[dir == 0 ? 0 : 1];
      const float(*gB)[RW] = sg[dir == 0 ? 2 : 3];
      const float(*Wq)[RW] = sg[dir == 0 ? 4 : 5];
      const float cA = gA[ly][lx], cB = gB[ly][lx], wc = Wq[ly][lx];
      // dL/dgB(q) = sum_k E_k(q-k) - sum_k E_k(q),  E_k(c) = W(c) * phi(gA(c+k)-gA(c), gB(c+k)-gB(c)).
      // phi is odd (phi(-s1,-s2) = -phi(s1,s2)), so pixel q as a neighbour of centre n and n as a
      // neighbour of centre q share one evaluation:  dL/dgB(q) = sum_n (W(n) + W(q)) * phi(cA-gA(n), cB-gB(n)).
      float ssum = 0.f;
#pragma unroll
      for (int dy = -R; dy <= R; ++dy)
#pragma unroll
        for (int dx = -R; dx <= R; ++dx) {
          const float wsum = Wq[ly + dy][lx + dx] + wc;
          if (wsum != 0.f) ssum += wsum * tern_phi(cA - gA[ly + dy][lx + dx], cB - gB[ly + dy][lx + dx]);
        }
      dG = u_tern * ssum;
    }
    if (dG != 0.f || (u_photo != 0.f && m != 0.f)) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const float Ia = __ldg(d.Bimg + t.ia * 3 + k), Ib = __ldg(d.Bimg + t.ib * 3 + k);
        const float Ic = __ldg(d.Bimg + t.ic * 3 + k), Id = __ldg(d.Bimg + t.id * 3 + k);
        const float wgt = k == 0 ? kGrayR : (k == 1 ? kGrayG : kGrayB);
        float dB = dG * 255.0f * wgt;
        if (u_photo != 0.f && m != 0.f) {
          const float diff = __ldg(d.A + q * 3 + k) - bil(t, Ia, Ib, Ic, Id);
          dB -= u_photo * m * charb_grad(diff, 255.0f);
        }
        du += dB * ((Ic - Ia) * t.omy + (Id - Ib) * t.yw);
        dv += dB * ((Ib - Ia) * t.omx + (Id - Ic) * t.xw);
      }
    }

    // ---- forward-backward consistency term ----
    if (u_fb != 0.f && m != 0.f) {
      const float2 *g2 = reinterpret_cast<const float2 *>(d.g);
      const float2 ga = __ldg(g2 + t.ia), gb = __ldg(g2 + t.ib), gc = __ldg(g2 + t.ic), gd = __ldg(g2 + t.id);
      const float ex = u_fb * m * charb_grad(fd.x, 1.0f), ey = u_fb * m * charb_grad(fd.y, 1.0f);
      du += ex + ex * ((gc.x - ga.x) * t.omy + (gd.x - gb.x) * t.yw) + ey * ((gc.y - ga.y) * t.omy + (gd.y - gb.y) * t.yw);
      dv += ey + ex * ((gb.x - ga.x) * t.omx + (gd.x - gc.x) * t.xw) + ey * ((gb.y - ga.y) * t.omx + (gd.y - gc.y) * t.xw);
      atomicAdd(dg_out + t.ia * 2, t.wa * ex); atomicAdd(dg_out + t.ia * 2 + 1, t.wa * ey);
      atomicAdd(dg_out + t.ib * 2, t.wb * ex); atomicAdd(dg_out + t.ib * 2 + 1, t.wb * ey);
      atomicAdd(dg_out + t.ic * 2, t.wc * ex); atomicAdd(dg_out + t.ic * 2 + 1, t.wc * ey);
      atomicAdd(dg_out + t.id * 2, t.wd * ex); atomicAdd(dg_out + t.id * 2 + 1, t.wd * ey);
    }

    // ---- smoothness (stencil transposes, recomputed from the flow) ----
    if (u_s1 != 0.f) {
      float gx = 0.f, gy = 0.f;
      if (x < w - 1) { const float2 n = __ldg(f2 + q + 1); gx += charb_grad(f.x - n.x, 1.f); gy += charb_grad(f.y - n.y, 1.f); }
      if (x >= 1)    { const float2 n = __ldg(f2 + q - 1); gx -= charb_grad(n.x - f.x, 1.f); gy -= charb_grad(n.y - f.y, 1.f); }
      if (y < h - 1) { const float2 n = __ldg(f2 + q + w); gx += charb_grad(f.x - n.x, 1.f); gy += charb_grad(f.y - n.y, 1.f); }
      if (y >= 1)    { const float2 n = __ldg(f2 + q - w); gx -= charb_grad(n.x - f.x, 1.f); gy -= charb_grad(n.y - f.y, 1.f); }
      du += u_s1 * gx; dv += u_s1 * gy;
    }
    if (u_s2 != 0.f) {
      // filters k: offsets e_k = +-1 (x), +-w (y), +-(w+1) (diag1), +-(w-1) (diag2);
      // psi_k(c) = mask_k(c) * c'(f(c-e)+f(c+e)-2f(c));  df(q) = sum_k psi_k(q-e)+psi_k(q+e)-2psi_k(q)
      float gx = 0.f, gy = 0.f;
      const int offs[4] = {1, w, w + 1, w - 1};
      const int dxs[4] = {1, 0, 1, -1}, dys[4] = {0, 1, 1, 1};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const bool need_x = k != 1, need_y = k != 0;
#pragma unroll
        for (int s = -1; s <= 1; ++s) {
          const int cx = x + s * dxs[k], cy = y + s * dys[k];
          if (cx < 0 || cx >= w || cy < 0 || cy >= h) continue;
          if (need_x && !(cx >= 1 && cx < w - 1)) continue;
          if (need_y && !(cy >= 1 && cy < h - 1)) continue;
          const int c = cy * w + cx;
          const float2 fc = __ldg(f2 + c);
          const float coef = s == 0 ? -2.f : 1.f;
          gx += coef * psi2(f2, c, offs[k], 0, 2.f * fc.x);
          gy += coef * psi2(f2, c, offs[k], 1, 2.f * fc.y);
        }
      }
      du += u_s2 * gx; dv += u_s2 * gy;
    }
    if (du != 0.f) atomicAdd(df_out + q * 2, du);
    if (dv != 0.f) atomicAdd(df_out + q * 2 + 1, dv);
  }
}

static int check_common(int B, int h, int w, int max_distance, int mask_occlusion, unsigned terms) {
  UNFLOW_REQUIRE(B >= 1 && h >= 1 && w >= 1, "level_loss: bad shape");
  UNFLOW_REQUIRE(B <= 65535, "level_loss: batch too large");
  UNFLOW_REQUIRE((long long)h * w * 3 < (1ll << 31), "level_loss: image too large");
  UNFLOW_REQUIRE(max_distance >= 1 && max_distance <= RMAXL, "level_loss: max_distance must be 1..3");
  UNFLOW_REQUIRE(mask_occlusion >= 0 && mask_occlusion <= 2, "level_loss: bad mask_occlusion");
  UNFLOW_REQUIRE(!((terms >> T_GRAD) & 1u), "level_loss: the 'grad' term is not fused");
  UNFLOW_REQUIRE(terms < (1u << NTERMS), "level_loss: bad term mask");
  return UNFLOW_OK;
}

}  // namespace ll
}  // namespace unflow

using namespace unflow;
using namespace unflow::ll;

extern "C" size_t unflow_level_loss_workspace_bytes(int B, int h, int w) {
  if (B < 1 || h < 1 || w < 1) return 0;
  const size_t nblocks = (size_t)ceil_div(w, TX) * ceil_div(h, TY) * B;
  return nblocks * NTERMS * sizeof(float) + 256;
}

extern "C" int unflow_level_loss_fwd(const float *im1, const float *im2, const float *flow_fw,
                                     const float *flow_bw, const float *border_mask,
                                     const float *fwarp_fw, const float *fwarp_bw, float *losses,
                                     float *saved, float *masks_out, void *workspace, int B, int h,
                                     int w, int mask_occlusion, int max_distance, unsigned terms,
                                     void *stream) {
  int rc = check_common(B, h, w, max_distance, mask_occlusion, terms);
  if (rc) return rc;
  UNFLOW_REQUIRE(im1 && im2 && flow_fw && flow_bw && losses && workspace, "level_loss: null pointer");
  const bool need_fwarp = mask_occlusion == OCC_DISOCC || ((terms >> T_SYM) & 1u);
  UNFLOW_REQUIRE(!need_fwarp || (fwarp_fw && fwarp_bw), "level_loss: forward_warp maps required");
  Params p{};
  p.im1 = im1; p.im2 = im2; p.ffw = flow_fw; p.fbw = flow_bw; p.border = border_mask;
  p.fwarp_fw = fwarp_fw; p.fwarp_bw = fwarp_bw; p.losses = losses; p.saved = saved;
  p.masks_out = masks_out; p.B = B; p.h = h; p.w = w; p.occl = mask_occlusion; p.r = max_distance;
  p.terms = terms;
  dim3 grid(ceil_div(w, TX), ceil_div(h, TY), B), block(TX, TY);
  const size_t nblocks = (size_t)grid.x * grid.y * grid.z;
  p.partials = (float *)workspace;
  p.counter = (unsigned *)((char *)workspace + nblocks * NTERMS * sizeof(float));
  cudaStream_t s = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(p.counter, 0, sizeof(unsigned), s);
  if (e != cudaSuccess) { set_error("level_loss memset: %s", cudaGetErrorString(e)); return UNFLOW_ECUDA; }
  switch (max_distance) {
    case 1: level_loss_fwd_kernel<1><<<grid, block, 0, s>>>(p); break;
    case 2: level_loss_fwd_kernel<2><<<grid, block, 0, s>>>(p); break;
    default: level_loss_fwd_kernel<3><<<grid, block, 0, s>>>(p); break;
  }
  count_launch();
  return check_launch("level_loss_fwd");
}

extern "C" int unflow_level_loss_bwd(const float *grad_losses, const float *im1, const float *im2,
                                     const float *flow_fw, const float *flow_bw,
                                     const float *border_mask, const float *fwarp_fw,
                                     const float *fwarp_bw, const float *saved, float *dflow_fw,
                                     float *dflow_bw, int B, int h, int w, int mask_occlusion,
                                     int max_distance, unsigned terms, void *stream) {
  int rc = check_common(B, h, w, max_distance, mask_occlusion, terms);
  if (rc) return rc;
  UNFLOW_REQUIRE(grad_losses && im1 && im2 && flow_fw && flow_bw && dflow_fw && dflow_bw,
                 "level_loss_grad: null pointer");
  UNFLOW_REQUIRE(!((terms >> T_TERN) & 1u) || saved, "level_loss_grad: saved planes required for ternary");
  UNFLOW_REQUIRE(mask_occlusion != OCC_DISOCC || (fwarp_fw && fwarp_bw), "level_loss_grad: forward_warp maps required");
  Params p{};
  p.im1 = im1; p.im2 = im2; p.ffw = flow_fw; p.fbw = flow_bw; p.border = border_mask;
  p.fwarp_fw = fwarp_fw; p.fwarp_bw = fwarp_bw; p.saved = const_cast<float *>(saved);
  p.gl = grad_losses; p.dffw = dflow_fw; p.dfbw = dflow_bw;
  p.B = B; p.h = h; p.w = w; p.occl = mask_occlusion; p.r = max_distance; p.terms = terms;
  cudaStream_t s = (cudaStream_t)stream;
  const size_t bytes = (size_t)B * h * w * 2 * sizeof(float);
  cudaError_t e = cudaMemsetAsync(dflow_fw, 0, bytes, s);
  if (e == cudaSuccess) e = cudaMemsetAsync(dflow_bw, 0, bytes, s);
  if (e != cudaSuccess) { set_error("level_loss_grad memset: %s", cudaGetErrorString(e)); return UNFLOW_ECUDA; }
  dim3 grid(ceil_div(w, TX), ceil_div(h, TY), B), block(TX, TY);
  switch (max_distance) {
    case 1: level_loss_bwd_kernel<1><<<grid, block, 0, s>>>(p); break;
    case 2: level_loss_bwd_kernel<2><<<grid, block, 0, s>>>(p); break;
    default: level_loss_bwd_kernel<3><<<grid, block, 0, s>>>(p); break;
  }
  count_launch();
  return check_launch("level_loss_bwd");
}
