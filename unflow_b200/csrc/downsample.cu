// downsample.cu -- box-filter downsample, NHWC.
// Replaces DownsampleKernel (reference ops/downsample_op.cu.cc:15-72).
//
// HBM-bound: algorithmic bytes = 4*B*C*(H*W + H*W/s^2).  One thread per output
// pixel-channel group; the inner dimension (x*C + c) of an input row is
// contiguous, so a warp reads `scale` contiguous row segments and writes one
// contiguous segment -- fully coalesced for any C.  The sum runs in the
// reference's order (in_y outer, in_x inner) and divides by scale_x*scale_y.
#include "common.cuh"

namespace unflow {

__global__ void __launch_bounds__(256)
downsample_kernel(const float *__restrict__ in, float *__restrict__ out, int B, int H, int W,
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
  downsample_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(images, out, B, H, W, C,
                                                                           scale, total);
  count_launch();
  return check_launch("downsample");
}
