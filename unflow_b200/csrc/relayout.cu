// relayout.cu -- planar <-> interleaved copies between the two layouts the hot path uses.
//
// The correlation kernels (correlation_tiled.cu) read and write the reference op's layout, NCHW
// (ops/correlation_op.cu.cc:250-315: inputs and cost volume are [B, C, H, W]); the convolution stack around
// them (tc_conv.cu) keeps every activation NHWC inside pitch-padded concat buffers.  Four tensors per step
// cross that border (flownet.py:34-44): the two feature maps going in, the two cost volumes coming out, and
// the same four gradients on the way back -- 0.6 ms of strided library copies, zero fills and adds per step
// before this file.  A 32 x 32 tile through shared memory makes both sides of the transpose coalesced.
//
//   planar      [B][C][P]          P = H*W pixels contiguous, channel stride = P (dense NCHW)
//   interleaved [B][P][pitch]      channels contiguous, pixel pitch >= C (a channel slice of an NHWC buffer)
#include "common.cuh"

namespace unflow {

// planar -> interleaved:  dst[b][p][c] (+)= src[b][c][p]
template <bool ACC>
__global__ void __launch_bounds__(256)
planar_to_interleaved_kernel(const float *__restrict__ src, float *__restrict__ dst, int C, int P,
                             long long src_batch, long long dst_batch, long long pitch) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;          // 32 x 8
  const float *s = src + (long long)b * src_batch;
  float *d = dst + (long long)b * dst_batch;
#pragma unroll
  for (int k = ty; k < 32; k += 8) {
    const int c = c0 + k, p = p0 + tx;
    tile[k][tx] = (c < C && p < P) ? s[(long long)c * P + p] : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int k = ty; k < 32; k += 8) {
    const int p = p0 + k, c = c0 + tx;
    if (p < P && c < C) {
      float *o = d + (long long)p * pitch + c;
      *o = ACC ? *o + tile[tx][k] : tile[tx][k];
    }
  }
}

// interleaved -> planar:  dst[b][c][p] = src[b][p][c]
__global__ void __launch_bounds__(256)
interleaved_to_planar_kernel(const float *__restrict__ src, float *__restrict__ dst, int C, int P,
                             long long src_batch, long long dst_batch, long long pitch) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const float *s = src + (long long)b * src_batch;
  float *d = dst + (long long)b * dst_batch;
#pragma unroll
  for (int k = ty; k < 32; k += 8) {
    const int p = p0 + k, c = c0 + tx;
    tile[k][tx] = (p < P && c < C) ? s[(long long)p * pitch + c] : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int k = ty; k < 32; k += 8) {
    const int c = c0 + k, p = p0 + tx;
    if (c < C && p < P) d[(long long)c * P + p] = tile[tx][k];
  }
}

}  // namespace unflow

using namespace unflow;

// dst[b][p][0..C) (+)= src[b][0..C)[p]; src dense planar with `src_batch` floats between images, dst with
// pixel pitch `pitch` and `dst_batch` floats between images (so dst may be a channel AND batch slice).
extern "C" int unflow_planar_to_interleaved(const float *src, long long src_batch, float *dst, long long dst_batch,
                                            long long pitch, int B, int C, int P, int accumulate, void *stream) {
  UNFLOW_REQUIRE(src && dst, "planar_to_interleaved: null pointer");
  UNFLOW_REQUIRE(B >= 0 && C > 0 && P > 0 && pitch >= C, "planar_to_interleaved: bad extents");
  UNFLOW_REQUIRE(B <= 65535 && (C + 31) / 32 <= 65535, "planar_to_interleaved: too many images / channels");
  if (B == 0) return UNFLOW_OK;
  dim3 grid((P + 31) / 32, (C + 31) / 32, B);
  if (accumulate)
    planar_to_interleaved_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(src, dst, C, P, src_batch, dst_batch, pitch);
  else
    planar_to_interleaved_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(src, dst, C, P, src_batch, dst_batch, pitch);
  count_launch();
  return check_launch("planar_to_interleaved");
}

// dst[b][0..C)[p] = src[b][p][0..C)
extern "C" int unflow_interleaved_to_planar(const float *src, long long src_batch, long long pitch, float *dst,
                                            long long dst_batch, int B, int C, int P, void *stream) {
  UNFLOW_REQUIRE(src && dst, "interleaved_to_planar: null pointer");
  UNFLOW_REQUIRE(B >= 0 && C > 0 && P > 0 && pitch >= C, "interleaved_to_planar: bad extents");
  UNFLOW_REQUIRE(B <= 65535 && (C + 31) / 32 <= 65535, "interleaved_to_planar: too many images / channels");
  if (B == 0) return UNFLOW_OK;
  dim3 grid((P + 31) / 32, (C + 31) / 32, B);
  interleaved_to_planar_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, dst, C, P, src_batch, dst_batch, pitch);
  count_launch();
  return check_launch("interleaved_to_planar");
}
