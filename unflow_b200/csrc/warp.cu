// warp.cu -- bilinear backward warp (gather), NHWC, forward + backward.
//
// One kernel family, two border semantics:
//   UNFLOW_BORDER_ZERO  = the BackwardWarp op  (reference ops/backward_warp_op.cu.cc:14-138):
//       x = src_x + u; x0 = floorf(x); weights w_right = x - x0, w_left = x1 - x;
//       taps outside the image contribute nothing; gradient w.r.t. the flow only.
//   UNFLOW_BORDER_CLAMP = image_warp, the warp the training path really uses
//       (reference src/e2eflow/core/image_warp.py:4-76): integer taps = pos + floor(flow),
//       clamped to the image; weights from flow - floor(flow); TF autodiff gives a gradient
//       w.r.t. the flow (through the weights) and w.r.t. the image (gather -> scatter-add).
//
// HBM-bound.  Algorithmic bytes: fwd 4*B*H*W*(2C+2); bwd 4*B*H*W*(2C+4) (+C if d/dimage).
// One thread per output pixel: the flow is read as one float2, the C channels of each tap are
// contiguous (NHWC) so neighbouring lanes gather from the same / adjacent 128 B lines (L1 hits
// for smooth flows); C is a template parameter for the shapes on the path (1 mask, 2 flow,
// 3 image) so the channel loop is fully unrolled and the taps are issued back to back.
#include "common.cuh"

namespace unflow {

struct Taps {
  int x0, x1, y0, y1;      // tap coordinates (clamped in CLAMP mode)
  bool vx0, vx1, vy0, vy1; // tap validity (always true in CLAMP mode)
  float wa, wb, wc, wd;    // weights of (y0,x0) (y1,x0) (y0,x1) (y1,x1)
  float xw, yw;            // CLAMP: fractional parts; ZERO: w_right, w_bottom
  float w_left, w_top;     // ZERO only
};

template <int MODE>
__device__ __forceinline__ Taps make_taps(int sx, int sy, float u, float v, int H, int W) {
  Taps t;
  if (MODE == UNFLOW_BORDER_ZERO) {
    const float x = sx + u, y = sy + v;
    t.x0 = (int)floorf(x); t.x1 = t.x0 + 1;
    t.y0 = (int)floorf(y); t.y1 = t.y0 + 1;
    t.xw = x - t.x0;       // w_right
    t.w_left = t.x1 - x;
    t.yw = y - t.y0;       // w_bottom
    t.w_top = t.y1 - y;
    t.vx0 = t.x0 >= 0 && t.x0 < W; t.vx1 = t.x1 >= 0 && t.x1 < W;
    t.vy0 = t.y0 >= 0 && t.y0 < H; t.vy1 = t.y1 >= 0 && t.y1 < H;
    t.wa = t.w_left * t.w_top;  // (y0,x0)
    t.wc = t.xw * t.w_top;      // (y0,x1)
    t.wb = t.w_left * t.yw;     // (y1,x0)
    t.wd = t.xw * t.yw;         // (y1,x1)
  } else if (MODE == UNFLOW_BORDER_STN) {
    // spatial_transformer._interpolate (reference core/spatial_transformer.py:57-114): (u,v) are
    // ABSOLUTE sample coordinates; indices are clamped and the weights are formed from the CLAMPED
    // indices, so samples outside [0,W-1]x[0,H-1] cancel to exactly zero.
    const float x = u, y = v;
    const int x0 = (int)floorf(x), y0 = (int)floorf(y);
    t.x0 = min(max(x0, 0), W - 1); t.x1 = min(max(x0 + 1, 0), W - 1);
    t.y0 = min(max(y0, 0), H - 1); t.y1 = min(max(y0 + 1, 0), H - 1);
    t.vx0 = t.vx1 = t.vy0 = t.vy1 = true;
    t.w_left = (float)t.x1 - x; t.xw = x - (float)t.x0;
    t.w_top = (float)t.y1 - y;  t.yw = y - (float)t.y0;
    t.wa = t.w_left * t.w_top;
    t.wb = t.w_left * t.yw;
    t.wc = t.xw * t.w_top;
    t.wd = t.xw * t.yw;
  } else {
    const float fu = floorf(u), fv = floorf(v);
    t.xw = u - fu; t.yw = v - fv;
    int x0 = sx + (int)fu, y0 = sy + (int)fv;
    int x1 = x0 + 1, y1 = y0 + 1;
    t.x0 = min(max(x0, 0), W - 1); t.x1 = min(max(x1, 0), W - 1);
    t.y0 = min(max(y0, 0), H - 1); t.y1 = min(max(y1, 0), H - 1);
    t.vx0 = t.vx1 = t.vy0 = t.vy1 = true;
    t.w_left = 1.0f - t.xw; t.w_top = 1.0f - t.yw;
    t.wa = t.w_left * t.w_top;
    t.wb = t.w_left * t.yw;
    t.wc = t.xw * t.w_top;
    t.wd = t.xw * t.yw;
  }
  return t;
}

template <int MODE, int CT>
__global__ void __launch_bounds__(256)
backward_warp_fwd_kernel(const float *__restrict__ img, const float *__restrict__ flow,
                         float *__restrict__ out, int B, int H, int W, int Crt, long long npix) {
  const int C = CT > 0 ? CT : Crt;
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < npix;
       p += (long long)gridDim.x * blockDim.x) {
    const int sx = (int)(p % W);
    const int sy = (int)((p / W) % H);
    const long long b = p / ((long long)W * H);
    const float2 f = __ldg(reinterpret_cast<const float2 *>(flow) + p);
    const Taps t = make_taps<MODE>(sx, sy, f.x, f.y, H, W);
    const float *base = img + b * (long long)H * W * C;
    const float *pa = base + ((long long)t.y0 * W + t.x0) * C;
    const float *pb = base + ((long long)t.y1 * W + t.x0) * C;
    const float *pc = base + ((long long)t.y0 * W + t.x1) * C;
    const float *pd = base + ((long long)t.y1 * W + t.x1) * C;
    const bool va = t.vx0 && t.vy0, vb = t.vx0 && t.vy1, vc = t.vx1 && t.vy0, vd = t.vx1 && t.vy1;
    float *o = out + p * C;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      float s;
      if (MODE == UNFLOW_BORDER_ZERO) {
        // reference order: top-left, top-right, bottom-left, bottom-right
        s = 0.0f;
        if (va) s += t.wa * __ldg(pa + c);
        if (vc) s += t.wc * __ldg(pc + c);
        if (vb) s += t.wb * __ldg(pb + c);
        if (vd) s += t.wd * __ldg(pd + c);
      } else {
        // tf.add_n([wa*Ia, wb*Ib, wc*Ic, wd*Id])
        s = t.wa * __ldg(pa + c) + t.wb * __ldg(pb + c);
        s += t.wc * __ldg(pc + c);
        s += t.wd * __ldg(pd + c);
      }
      o[c] = s;
    }
  }
}

template <int MODE, int CT, bool DIMAGE>
__global__ void __launch_bounds__(256)
backward_warp_bwd_kernel(const float *__restrict__ grad, const float *__restrict__ img,
                         const float *__restrict__ flow, float *__restrict__ dflow,
                         float *__restrict__ dimg, int B, int H, int W, int Crt, long long npix) {
  const int C = CT > 0 ? CT : Crt;
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < npix;
       p += (long long)gridDim.x * blockDim.x) {
    const int sx = (int)(p % W);
    const int sy = (int)((p / W) % H);
    const long long b = p / ((long long)W * H);
    const float2 f = __ldg(reinterpret_cast<const float2 *>(flow) + p);
    const Taps t = make_taps<MODE>(sx, sy, f.x, f.y, H, W);
    const long long boff = b * (long long)H * W * C;
    const long long ia = boff + ((long long)t.y0 * W + t.x0) * C;
    const long long ib = boff + ((long long)t.y1 * W + t.x0) * C;
    const long long ic = boff + ((long long)t.y0 * W + t.x1) * C;
    const long long id = boff + ((long long)t.y1 * W + t.x1) * C;
    const bool va = t.vx0 && t.vy0, vb = t.vx0 && t.vy1, vc = t.vx1 && t.vy0, vd = t.vx1 && t.vy1;
    float du = 0.0f, dv = 0.0f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float din = __ldg(grad + p * C + c);
      if (MODE == UNFLOW_BORDER_ZERO) {
        // reference ops/backward_warp_op.cu.cc:101-131, same tap order
        float px;
        if (va) { px = __ldg(img + ia + c) * din; du -= t.w_top * px; dv -= t.w_left * px; }
        if (vc) { px = __ldg(img + ic + c) * din; du += t.w_top * px; dv -= t.xw * px; }
        if (vb) { px = __ldg(img + ib + c) * din; du -= t.yw * px; dv += t.w_left * px; }
        if (vd) { px = __ldg(img + id + c) * din; du += t.yw * px; dv += t.xw * px; }
      } else {
        const float Ia = __ldg(img + ia + c), Ib = __ldg(img + ib + c);
        const float Ic = __ldg(img + ic + c), Id = __ldg(img + id + c);
        // d(wa Ia + wb Ib + wc Ic + wd Id)/dxw and /dyw with wa=(1-xw)(1-yw) ...
        du += din * ((Ic - Ia) * t.w_top + (Id - Ib) * t.yw);
        dv += din * ((Ib - Ia) * t.w_left + (Id - Ic) * t.xw);
      }
      if (DIMAGE) {
        if (va) atomicAdd(dimg + ia + c, t.wa * din);
        if (vb) atomicAdd(dimg + ib + c, t.wb * din);
        if (vc) atomicAdd(dimg + ic + c, t.wc * din);
        if (vd) atomicAdd(dimg + id + c, t.wd * din);
      }
    }
    reinterpret_cast<float2 *>(dflow)[p] = make_float2(du, dv);
  }
}

template <int MODE>
static void launch_fwd(const float *img, const float *flow, float *out, int B, int H, int W, int C,
                       long long npix, cudaStream_t s) {
  const int grid = grid_for(npix, 256);
  switch (C) {
    case 1: backward_warp_fwd_kernel<MODE, 1><<<grid, 256, 0, s>>>(img, flow, out, B, H, W, C, npix); break;
    case 2: backward_warp_fwd_kernel<MODE, 2><<<grid, 256, 0, s>>>(img, flow, out, B, H, W, C, npix); break;
    case 3: backward_warp_fwd_kernel<MODE, 3><<<grid, 256, 0, s>>>(img, flow, out, B, H, W, C, npix); break;
    default: backward_warp_fwd_kernel<MODE, 0><<<grid, 256, 0, s>>>(img, flow, out, B, H, W, C, npix); break;
  }
}

template <int MODE, bool DIMAGE>
static void launch_bwd(const float *grad, const float *img, const float *flow, float *dflow,
                       float *dimg, int B, int H, int W, int C, long long npix, cudaStream_t s) {
  const int grid = grid_for(npix, 256);
  switch (C) {
    case 1: backward_warp_bwd_kernel<MODE, 1, DIMAGE><<<grid, 256, 0, s>>>(grad, img, flow, dflow, dimg, B, H, W, C, npix); break;
    case 2: backward_warp_bwd_kernel<MODE, 2, DIMAGE><<<grid, 256, 0, s>>>(grad, img, flow, dflow, dimg, B, H, W, C, npix); break;
    case 3: backward_warp_bwd_kernel<MODE, 3, DIMAGE><<<grid, 256, 0, s>>>(grad, img, flow, dflow, dimg, B, H, W, C, npix); break;
    default: backward_warp_bwd_kernel<MODE, 0, DIMAGE><<<grid, 256, 0, s>>>(grad, img, flow, dflow, dimg, B, H, W, C, npix); break;
  }
}

}  // namespace unflow

extern "C" int unflow_backward_warp_fwd(const float *images, const float *flows, float *out, int B,
                                        int H, int W, int C, int border_mode, void *stream) {
  using namespace unflow;
  UNFLOW_REQUIRE(B >= 0 && H >= 0 && W >= 0 && C >= 0, "backward_warp: negative dimension");
  UNFLOW_REQUIRE(border_mode == UNFLOW_BORDER_ZERO || border_mode == UNFLOW_BORDER_CLAMP ||
                     border_mode == UNFLOW_BORDER_STN,
                 "backward_warp: unknown border_mode %d", border_mode);
  const long long npix = (long long)B * H * W;
  if (npix == 0 || C == 0) return UNFLOW_OK;
  UNFLOW_REQUIRE(images && flows && out, "backward_warp: null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  if (border_mode == UNFLOW_BORDER_ZERO) launch_fwd<UNFLOW_BORDER_ZERO>(images, flows, out, B, H, W, C, npix, s);
  else if (border_mode == UNFLOW_BORDER_STN) launch_fwd<UNFLOW_BORDER_STN>(images, flows, out, B, H, W, C, npix, s);
  else launch_fwd<UNFLOW_BORDER_CLAMP>(images, flows, out, B, H, W, C, npix, s);
  count_launch();
  return check_launch("backward_warp_fwd");
}

extern "C" int unflow_backward_warp_bwd(const float *grad, const float *images, const float *flows,
                                        float *dflow, float *dimage, int B, int H, int W, int C,
                                        int border_mode, void *stream) {
  using namespace unflow;
  UNFLOW_REQUIRE(B >= 0 && H >= 0 && W >= 0 && C >= 0, "backward_warp_grad: negative dimension");
  UNFLOW_REQUIRE(border_mode == UNFLOW_BORDER_ZERO || border_mode == UNFLOW_BORDER_CLAMP,
                 "backward_warp_grad: unknown border_mode %d", border_mode);
  const long long npix = (long long)B * H * W;
  if (npix == 0) return UNFLOW_OK;
  UNFLOW_REQUIRE(grad && images && flows && dflow, "backward_warp_grad: null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  if (border_mode == UNFLOW_BORDER_ZERO) {
    if (dimage) launch_bwd<UNFLOW_BORDER_ZERO, true>(grad, images, flows, dflow, dimage, B, H, W, C, npix, s);
    else launch_bwd<UNFLOW_BORDER_ZERO, false>(grad, images, flows, dflow, dimage, B, H, W, C, npix, s);
  } else {
    if (dimage) launch_bwd<UNFLOW_BORDER_CLAMP, true>(grad, images, flows, dflow, dimage, B, H, W, C, npix, s);
    else launch_bwd<UNFLOW_BORDER_CLAMP, false>(grad, images, flows, dflow, dimage, B, H, W, C, npix, s);
  }
  count_launch();
  return check_launch("backward_warp_bwd");
}
