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

}  // namespace unflow

extern "C" int unflow_conv_operand_tf32(const float *x, float *out, int N, int C, int H, int W,
                                        long long sN, long long sC, long long sH, long long sW,
                                        int N_out, int C_pad, int pad_top, int pad_bottom,
                                        int pad_left, int pad_right, int concat_batch, int order,
                                        void *stream) {
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
