// downsample.cu -- box-filter downsample, NHWC.
// Replaces DownsampleKernel (reference ops/downsample_op.cu.cc:15-72).
//
// HBM-bound: algorithmic bytes = 4*B*C*(H*W + H*W/s^2).  The input row segment that feeds one
// output pixel is scale*C contiguous floats; for the shapes on the path (scale 2/4, C 1/2/3) one
// thread owns one output pixel and reads each of its `scale` row segments with 128/64-bit loads
// (consecutive threads read consecutive segments: fully coalesced), instead of the reference's one
// thread per output element with scale^2 strided scalar loads.  The sum runs in the reference's
// order (in_y outer, in_x inner) and divides by scale_x*scale_y.
#include "common.cuh"

namespace unflow {

template <int SCALE, int C>
__global__ void __launch_bounds__(256)
downsample_px_kernel(const float *__restrict__ in, float *__restrict__ out, int H, int W, long long npix_out) {
  constexpr int SEG = SCALE * C;                      // floats per row segment
  constexpr int VEC = SEG % 4 == 0 ? 4 : (SEG % 2 == 0 ? 2 : 1);
  const int oh = H / SCALE, ow = W / SCALE;
  const float div = (float)(SCALE * SCALE);
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < npix_out;
       p += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(p % ow);
    const int y = (int)((p / ow) % oh);
    const long long b = p / ((long long)ow * oh);
    const float *src = in + ((b * H + (long long)y * SCALE) * W + (long long)x * SCALE) * C;
    float sum[C];
#pragma unroll
    for (int c = 0; c < C; ++c) sum[c] = 0.0f;
#pragma unroll
    for (int iy = 0; iy < SCALE; ++iy) {
      float seg[SEG];
      const float *row = src + (long long)iy * W * C;
      if (VEC == 4) {
#pragma unroll
        for (int k = 0; k < SEG / 4; ++k) {
          const float4 v = __ldg(reinterpret_cast<const float4 *>(row) + k);
          seg[4 * k] = v.x; seg[4 * k + 1] = v.y; seg[4 * k + 2] = v.z; seg[4 * k + 3] = v.w;
        }
      } else if (VEC == 2) {
#pragma unroll
        for (int k = 0; k < SEG / 2; ++k) {
          const float2 v = __ldg(reinterpret_cast<const float2 *>(row) + k);
          seg[2 * k] = v.x; seg[2 * k + 1] = v.y;
        }
      } else {
#pragma unroll
        for (int k = 0; k < SEG; ++k) seg[k] = __ldg(row + k);
      }
#pragma unroll
      for (int ix = 0; ix < SCALE; ++ix)
#pragma unroll
        for (int c = 0; c < C; ++c) sum[c] += seg[ix * C + c];
    }
#pragma unroll
    for (int c = 0; c < C; ++c) out[p * C + c] = sum[c] / div;
  }
}

// any scale / channel count: one thread per output element
__global__ void __launch_bounds__(256)
downsample_generic_kernel(const float *__restrict__ in, float *__restrict__ out, int B, int H, int W,
                          int C, int scale, long long total) {
  const int oh = H / scale, ow = W / scale;
  const float div = (float)(scale * scale);
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    long long t = idx;
    const int c = (int)(t % C); t /= C;
    const int x = (int)(t % ow); t /= ow;
    const int y = (int)(t % oh);
    const int b = (int)(t / oh);
    const float *src = in + (((long long)b * H + (long long)y * scale) * W + (long long)x * scale) * C + c;
    float sum = 0.0f;
    for (int iy = 0; iy < scale; ++iy) {
      const float *row = src + (long long)iy * W * C;
      for (int ix = 0; ix < scale; ++ix) sum += __ldg(row + (long long)ix * C);
    }
    out[idx] = sum / div;
  }
}

template <int SCALE, int C>
static void launch_px(const float *in, float *out, int B, int H, int W, cudaStream_t s) {
  const long long npix = (long long)B * (H / SCALE) * (W / SCALE);
  downsample_px_kernel<SCALE, C><<<grid_for(npix, 256), 256, 0, s>>>(in, out, H, W, npix);
}

}  // namespace unflow

extern "C" int unflow_downsample(const float *images, float *out, int B, int H, int W, int C,
                                 int scale, void *stream) {
  using namespace unflow;
  UNFLOW_REQUIRE(B >= 0 && H >= 0 && W >= 0 && C >= 0, "downsample: negative dimension");
  UNFLOW_REQUIRE(scale >= 1, "downsample: scale must be >= 1");
  UNFLOW_REQUIRE(H % scale == 0 && W % scale == 0,
                 "Input height and width must be divisible by scale");
  const long long total = (long long)B * (H / scale) * (W / scale) * C;
  if (total == 0) return UNFLOW_OK;
  UNFLOW_REQUIRE(images && out, "downsample: null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  // vector loads need the row segments 16/8-byte aligned: base pointer aligned and (W*C) even / %4
  const bool al16 = ((uintptr_t)images & 15) == 0, al8 = ((uintptr_t)images & 7) == 0;
  const int key = scale * 10 + C;
  bool done = true;
  switch (key) {
    case 43: if (al16) launch_px<4, 3>(images, out, B, H, W, s); else done = false; break;
    case 41: if (al16) launch_px<4, 1>(images, out, B, H, W, s); else done = false; break;
    case 42: if (al16) launch_px<4, 2>(images, out, B, H, W, s); else done = false; break;
    case 23: if (al8) launch_px<2, 3>(images, out, B, H, W, s); else done = false; break;
    case 21: if (al8) launch_px<2, 1>(images, out, B, H, W, s); else done = false; break;
    case 22: if (al16) launch_px<2, 2>(images, out, B, H, W, s); else done = false; break;
    default: done = false;
  }
  if (!done)
    downsample_generic_kernel<<<grid_for(total, 256), 256, 0, s>>>(images, out, B, H, W, C, scale, total);
  count_launch();
  return check_launch("downsample");
}
