// split.cu -- operand preparation for the 3xTF32 tensor-core conv path.
//
// The conv / deconv stacks are dense contractions and belong on the tensor cores, but a single
// TF32 pass (10-bit mantissa) costs ~1e-3 relative on the flow fields.  Every operand is written as
// x = hi + lo with hi = round_to_tf32(x), lo = x - hi (exact in fp32) and the contraction
//     hi*hi' + hi*lo' + lo*hi'          (the lo*lo' term is ~2^-22 relative and dropped)
// is evaluated by ONE library convolution whose contraction dimension carries the three products
// side by side:  X' = [hi, hi, lo],  W' = [hi', lo', hi'];  fp32 accumulation inside the MMA.
//
// This kernel builds X' / W' in one pass and folds in everything else cuDNN would otherwise do with
// extra kernels (each of them showed up in the ncu launch list of the step):
//   * it reads the source through arbitrary strides (NCHW, channels_last, channel-sliced views of
//     concat buffers) -- no separate layout-conversion copy;
//   * it writes dense NHWC (channels_last), the layout of cuDNN's tensor-core kernels -- no
//     nchwToNhwc / nhwcToNchw transforms;
//   * it pads the channel count to a multiple of 4 with zeros (16-byte channel vectors) -- no
//     nhwcAddPaddingKernel;
//   * it applies TensorFlow's asymmetric SAME padding (top/left offsets) -- no F.pad copy;
//   * the three slabs go either side by side along C ("channel" concat: fprop / dgrad operands) or
//     along N ("batch" concat: wgrad operands), order (hi,hi,lo) or (hi,lo,hi).
// HBM-bound: 4 B read + 12 B written per (padded) element.
#include "common.cuh"

namespace unflow {

__device__ __forceinline__ float tf32_rn(float x) {
  unsigned u = __float_as_uint(x);
  u += 0xFFFu + ((u >> 13) & 1u);   // round to nearest even on the 13 dropped bits
  u &= 0xFFFFE000u;
  return __uint_as_float(u);
}

struct OperandArgs {
  const float *x;
  float *out;
  long long sN, sC, sH, sW;       // source strides (floats) of the logical [N,C,H,W] tensor
  int N, C, H, W;                 // source extents
  int No, Cp, Hp, Wp, pt, pl;     // output extents (No >= N, Cp >= C, Cp % 4 == 0) and pad offsets
  long long item_stride;          // floats between consecutive output pixels (3*Cp or Cp)
  long long slab_stride;          // floats between the three slabs (Cp or No*Hp*Wp*Cp)
  long long total;                // No*Hp*Wp*(Cp/4) work items
  const float *act;               // optional: leaky-ReLU OUTPUT of the layer (dense NHWC [N,H,W,C]);
  float slope;                    //   the source is a gradient and is multiplied by lrelu'(act)
};

template <int ORDER, bool VEC>
__global__ void __launch_bounds__(256)
conv_operand_kernel(OperandArgs a) {
  const int c4n = a.Cp >> 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < a.total;
       i += (long long)gridDim.x * blockDim.x) {
    long long t = i;
    const int c4 = (int)(t % c4n); t /= c4n;
    const int xo = (int)(t % a.Wp); t /= a.Wp;
    const int yo = (int)(t % a.Hp);
    const int n = (int)(t / a.Hp);
    const int y = yo - a.pt, x = xo - a.pl, c = c4 * 4;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (n < a.N && y >= 0 && y < a.H && x >= 0 && x < a.W) {
      const float *src = a.x + n * a.sN + y * a.sH + x * a.sW + c * a.sC;
      if (VEC) {  // sC == 1, 16-byte aligned rows, C % 4 == 0
        const float4 q = __ldg(reinterpret_cast<const float4 *>(src));
        v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (c + k < a.C) v[k] = __ldg(src + k * a.sC);
      }
      if (a.act) {  // fused leaky_relu backward: d/dy = g * (out > 0 ? 1 : slope)
        const float *ap = a.act + (((long long)n * a.H + y) * a.W + x) * a.C + c;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (c + k < a.C && __ldg(ap + k) <= 0.0f) v[k] *= a.slope;
      }
    }
    float hi[4], lo[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { hi[k] = tf32_rn(v[k]); lo[k] = v[k] - hi[k]; }
    const long long pix = ((long long)n * a.Hp + yo) * a.Wp + xo;
    float *o = a.out + pix * a.item_stride + c;
    const float4 H4 = make_float4(hi[0], hi[1], hi[2], hi[3]);
    const float4 L4 = make_float4(lo[0], lo[1], lo[2], lo[3]);
    *reinterpret_cast<float4 *>(o) = H4;
    *reinterpret_cast<float4 *>(o + a.slab_stride) = ORDER == 0 ? H4 : L4;
    *reinterpret_cast<float4 *>(o + 2 * a.slab_stride) = ORDER == 0 ? L4 : H4;
  }
}

// ---------------------------------------------------------------------------------------------
// Fused bias + leaky ReLU (in place, dense NHWC) and its bias gradient.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
bias_lrelu_kernel(float4 *__restrict__ y, const float *__restrict__ bias, long long n4, int c4n, float slope) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % c4n) * 4;
    float4 v = y[i];
    // the bias is a view into the flat parameter buffer: only 4-byte aligned
    v.x += __ldg(bias + c); v.y += __ldg(bias + c + 1); v.z += __ldg(bias + c + 2); v.w += __ldg(bias + c + 3);
    v.x = v.x > 0.f ? v.x : v.x * slope; v.y = v.y > 0.f ? v.y : v.y * slope;
    v.z = v.z > 0.f ? v.z : v.z * slope; v.w = v.w > 0.f ? v.w : v.w * slope;
    y[i] = v;
  }
}

// gb[c] += sum over pixels of g * lrelu'(act); one CTA = 32 channels x a strip of pixels.
// LINEAR: the gradient is NHWC-like (possibly a channel slice of a wider buffer), so pixel p lives
// at p*pix_stride -- no per-element div/mod (the first version spent most of its time on them).
template <bool LINEAR>
__global__ void __launch_bounds__(256)
bias_grad_lrelu_kernel(const float *__restrict__ g, long long sN, long long sC, long long sH, long long sW,
                       const float *__restrict__ act, float *__restrict__ gb, int N, int C, int H, int W,
                       float slope, int pix_per_cta, float *__restrict__ gpre, long long AP, long long GP) {
  __shared__ float red[8][32];
  const int lane = threadIdx.x & 31, row = threadIdx.x >> 5;
  const int c = blockIdx.y * 32 + lane;
  const long long npix = (long long)N * H * W;
  const long long p0 = (long long)blockIdx.x * pix_per_cta;
  const long long p1 = p0 + pix_per_cta < npix ? p0 + pix_per_cta : npix;
  float s0 = 0.f, s1 = 0.f;
  if (c < C) {
    if (LINEAR) {
      const float *gp = g + c * sC;
      const float *ap = act ? act + c : nullptr;
      long long p = p0 + row;
      for (; p + 8 < p1; p += 16) {           // two independent loads in flight per thread
        float v0 = __ldg(gp + p * sW), v1 = __ldg(gp + (p + 8) * sW);
        if (ap) {
          if (__ldg(ap + p * AP) <= 0.0f) v0 *= slope;
          if (__ldg(ap + (p + 8) * AP) <= 0.0f) v1 *= slope;
        }
        if (gpre) { gpre[p * GP + c] = v0; gpre[(p + 8) * GP + c] = v1; }
        s0 += v0; s1 += v1;
      }
      for (; p < p1; p += 8) {
        float v = __ldg(gp + p * sW);
        if (ap && __ldg(ap + p * AP) <= 0.0f) v *= slope;
        if (gpre) gpre[p * GP + c] = v;
        s0 += v;
      }
    } else {
      for (long long p = p0 + row; p < p1; p += 8) {
        const int x = (int)(p % W);
        const int y = (int)((p / W) % H);
        const long long n = p / ((long long)W * H);
        float v = __ldg(g + n * sN + y * sH + x * sW + c * sC);
        if (act && __ldg(act + p * AP + c) <= 0.0f) v *= slope;
        if (gpre) gpre[p * GP + c] = v;
        s0 += v;
      }
    }
  }
  red[row][lane] = s0 + s1;
  __syncthreads();
  if (row == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k][lane];
    atomicAdd(gb + c, t);
  }
}

// Vector form of the LINEAR case (NHWC gradient, channels contiguous, everything 16-byte aligned):
// a lane owns 4 channels (one 128-bit load per tensor and pixel), a warp 128 channels, the 8 warps of a
// CTA walk the pixel strip with two pixels in flight each.  The scalar kernel above reaches 39 % of the
// HBM roofline (one 4-byte load per lane in flight); this one moves 4x the bytes per instruction.
// Layers with fewer than 128 channels (conv1 / deconv2: 64, the largest tensors of the step) would leave
// lanes without channels: `cl` = lanes per pixel (a power of two <= 32, cl * 4 >= min(C, 128)), the other
// 32 / cl lane groups of a warp take the following pixels.
__global__ void __launch_bounds__(256)
bias_grad_lrelu_vec_kernel(const float *__restrict__ g, long long GS, const float *__restrict__ act, long long AP,
                           float *__restrict__ gpre, long long GP, float *__restrict__ gb, int C, long long npix,
                           float slope, int pix_per_cta, int cl) {
  __shared__ float4 red[8][32];
  const int lane = threadIdx.x & 31, row = threadIdx.x >> 5;
  const int sub = 32 / cl;                        // pixels per warp and step
  const int c = blockIdx.y * 128 + (lane & (cl - 1)) * 4;
  const long long p0 = (long long)blockIdx.x * pix_per_cta;
  const long long p1 = p0 + pix_per_cta < npix ? p0 + pix_per_cta : npix;
  const int step = 8 * sub;
  float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0;
  if (c < C) {                                   // C % 4 == 0: a lane's four channels are all inside or all outside
    long long p = p0 + row * sub + lane / cl;
    for (; p + step < p1; p += 2 * step) {
      float4 v0 = __ldg(reinterpret_cast<const float4 *>(g + p * GS + c));
      float4 v1 = __ldg(reinterpret_cast<const float4 *>(g + (p + step) * GS + c));
      if (act) {
        const float4 a0 = __ldg(reinterpret_cast<const float4 *>(act + p * AP + c));
        const float4 a1 = __ldg(reinterpret_cast<const float4 *>(act + (p + step) * AP + c));
        if (a0.x <= 0.f) v0.x *= slope; if (a0.y <= 0.f) v0.y *= slope; if (a0.z <= 0.f) v0.z *= slope; if (a0.w <= 0.f) v0.w *= slope;
        if (a1.x <= 0.f) v1.x *= slope; if (a1.y <= 0.f) v1.y *= slope; if (a1.z <= 0.f) v1.z *= slope; if (a1.w <= 0.f) v1.w *= slope;
      }
      if (gpre) {
        *reinterpret_cast<float4 *>(gpre + p * GP + c) = v0;
        *reinterpret_cast<float4 *>(gpre + (p + step) * GP + c) = v1;
      }
      s0.x += v0.x; s0.y += v0.y; s0.z += v0.z; s0.w += v0.w;
      s1.x += v1.x; s1.y += v1.y; s1.z += v1.z; s1.w += v1.w;
    }
    for (; p < p1; p += step) {
      float4 v = __ldg(reinterpret_cast<const float4 *>(g + p * GS + c));
      if (act) {
        const float4 a = __ldg(reinterpret_cast<const float4 *>(act + p * AP + c));
        if (a.x <= 0.f) v.x *= slope; if (a.y <= 0.f) v.y *= slope; if (a.z <= 0.f) v.z *= slope; if (a.w <= 0.f) v.w *= slope;
      }
      if (gpre) *reinterpret_cast<float4 *>(gpre + p * GP + c) = v;
      s0.x += v.x; s0.y += v.y; s0.z += v.z; s0.w += v.w;
    }
  }
  red[row][lane] = make_float4(s0.x + s1.x, s0.y + s1.y, s0.z + s1.z, s0.w + s1.w);
  __syncthreads();
  if (row == 0 && lane < cl && c < C) {
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k = 0; k < 8; ++k)
      for (int u = 0; u < sub; ++u) { const float4 r = red[k][lane + u * cl]; t.x += r.x; t.y += r.y; t.z += r.z; t.w += r.w; }
    atomicAdd(gb + c, t.x); atomicAdd(gb + c + 1, t.y); atomicAdd(gb + c + 2, t.z); atomicAdd(gb + c + 3, t.w);
  }
}

}  // namespace unflow

extern "C" int unflow_bias_lrelu(float *y, const float *bias, long long pixels, int C, float slope,
                                 void *stream) {
  using namespace unflow;
  UNFLOW_REQUIRE(pixels >= 0 && C >= 4 && C % 4 == 0, "bias_lrelu: C must be a positive multiple of 4");
  if (pixels == 0) return UNFLOW_OK;
  UNFLOW_REQUIRE(y && bias, "bias_lrelu: null pointer");
  UNFLOW_REQUIRE(((uintptr_t)y & 15) == 0, "bias_lrelu: y must be 16-byte aligned");
  const long long n4 = pixels * (C / 4);
  bias_lrelu_kernel<<<grid_for(n4, 256, 16), 256, 0, (cudaStream_t)stream>>>((float4 *)y, bias, n4, C / 4, slope);
  count_launch();
  return check_launch("bias_lrelu");
}

extern "C" int unflow_lrelu_bwd_bias(const float *g, long long sN, long long sC, long long sH, long long sW,
                                     const float *act, long long act_pitch, float *gpre, long long gpre_pitch,
                                     float *gb, int N, int C, int H, int W, float slope, void *stream);

extern "C" int unflow_bias_grad_lrelu(const float *g, long long sN, long long sC, long long sH,
                                      long long sW, const float *act, float *gb, int N, int C, int H,
                                      int W, float slope, void *stream) {
  return unflow_lrelu_bwd_bias(g, sN, sC, sH, sW, act, C, nullptr, C, gb, N, C, H, W, slope, stream);
}

extern "C" int unflow_lrelu_bwd_bias(const float *g, long long sN, long long sC, long long sH, long long sW,
                                     const float *act, long long act_pitch, float *gpre, long long gpre_pitch,
                                     float *gb, int N, int C, int H, int W, float slope, void *stream) {
  using namespace unflow;
  UNFLOW_REQUIRE(!gpre || gpre_pitch >= C, "lrelu_bwd_bias: gpre pitch smaller than C");
  UNFLOW_REQUIRE(!act || act_pitch >= C, "lrelu_bwd_bias: activation pitch smaller than C");
  UNFLOW_REQUIRE(N >= 0 && C >= 1 && H >= 1 && W >= 1, "bias_grad: bad shape");
  UNFLOW_REQUIRE(g && gb, "bias_grad: null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(gb, 0, sizeof(float) * C, s);
  if (e != cudaSuccess) { set_error("bias_grad memset: %s", cudaGetErrorString(e)); return UNFLOW_ECUDA; }
  const long long npix = (long long)N * H * W;
  if (npix == 0) return UNFLOW_OK;
  const int cblocks = ceil_div(C, 32);
  long long strips = (long long)kNumSMs * 8 / cblocks;
  if (strips < 1) strips = 1;
  int pix_per_cta = (int)((npix + strips - 1) / strips);
  if (pix_per_cta < 64) pix_per_cta = 64;
  dim3 grid(ceil_div(npix, pix_per_cta), cblocks);
  const bool linear = sH == (long long)W * sW && sN == (long long)H * sH;
  const bool vec = linear && sC == 1 && C % 4 == 0 && sW % 4 == 0 && ((uintptr_t)g & 15) == 0 &&
                   (!act || (act_pitch % 4 == 0 && ((uintptr_t)act & 15) == 0)) &&
                   (!gpre || (gpre_pitch % 4 == 0 && ((uintptr_t)gpre & 15) == 0));
  if (vec) {
    const int cb = ceil_div(C, 128);
    long long st = (long long)kNumSMs * 8 / cb;
    if (st < 1) st = 1;
    int ppc = (int)((npix + st - 1) / st);
    if (ppc < 64) ppc = 64;
    dim3 vgrid(ceil_div(npix, ppc), cb);
    int cl = 32;                                  // lanes per pixel: the smallest power of two covering min(C, 128) / 4
    while (cl > 1 && (cl / 2) * 4 >= (C < 128 ? C : 128)) cl /= 2;
    bias_grad_lrelu_vec_kernel<<<vgrid, 256, 0, s>>>(g, sW, act, act_pitch, gpre, gpre_pitch, gb, C, npix, slope, ppc, cl);
    count_launch();
    return check_launch("bias_grad_lrelu(vec)");
  }
  if (linear) bias_grad_lrelu_kernel<true><<<grid, 256, 0, s>>>(g, sN, sC, sH, sW, act, gb, N, C, H, W, slope, pix_per_cta, gpre, act_pitch, gpre_pitch);
  else bias_grad_lrelu_kernel<false><<<grid, 256, 0, s>>>(g, sN, sC, sH, sW, act, gb, N, C, H, W, slope, pix_per_cta, gpre, act_pitch, gpre_pitch);
  count_launch();
  return check_launch("bias_grad_lrelu");
}

extern "C" int unflow_conv_operand_tf32(const float *x, float *out, int N, int C, int H, int W,
                                        long long sN, long long sC, long long sH, long long sW,
                                        int N_out, int C_pad, int pad_top, int pad_bottom,
                                        int pad_left, int pad_right, int concat_batch, int order,
                                        const float *act, float slope, void *stream) {
  using namespace unflow;
  UNFLOW_REQUIRE(N >= 0 && C >= 1 && H >= 1 && W >= 1, "conv_operand: bad source shape");
  UNFLOW_REQUIRE(N_out >= N && C_pad >= C && C_pad % 4 == 0, "conv_operand: bad padded shape");
  UNFLOW_REQUIRE(pad_top >= 0 && pad_bottom >= 0 && pad_left >= 0 && pad_right >= 0, "conv_operand: negative padding");
  UNFLOW_REQUIRE(order == 0 || order == 1, "conv_operand: order must be 0 (hi,hi,lo) or 1 (hi,lo,hi)");
  if (N_out == 0) return UNFLOW_OK;
  UNFLOW_REQUIRE(x && out, "conv_operand: null pointer");
  UNFLOW_REQUIRE(((uintptr_t)out & 15) == 0, "conv_operand: output must be 16-byte aligned");
  OperandArgs a;
  a.x = x; a.out = out; a.sN = sN; a.sC = sC; a.sH = sH; a.sW = sW;
  a.N = N; a.C = C; a.H = H; a.W = W;
  a.No = N_out; a.Cp = C_pad; a.Hp = H + pad_top + pad_bottom; a.Wp = W + pad_left + pad_right;
  a.pt = pad_top; a.pl = pad_left;
  const long long pixels = (long long)N_out * a.Hp * a.Wp;
  a.item_stride = concat_batch ? C_pad : 3ll * C_pad;
  a.slab_stride = concat_batch ? pixels * C_pad : C_pad;
  a.total = pixels * (C_pad / 4);
  a.act = act; a.slope = slope;
  const bool vec = sC == 1 && C % 4 == 0 && ((uintptr_t)x & 15) == 0 && sN % 4 == 0 && sH % 4 == 0 && sW % 4 == 0;
  const int grid = grid_for(a.total, 256, 16);
  cudaStream_t s = (cudaStream_t)stream;
  if (order == 0) {
    if (vec) conv_operand_kernel<0, true><<<grid, 256, 0, s>>>(a);
    else conv_operand_kernel<0, false><<<grid, 256, 0, s>>>(a);
  } else {
    if (vec) conv_operand_kernel<1, true><<<grid, 256, 0, s>>>(a);
    else conv_operand_kernel<1, false><<<grid, 256, 0, s>>>(a);
  }
  count_launch();
  return check_launch("conv_operand_tf32");
}
