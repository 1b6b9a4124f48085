// split.cu -- fp32 -> (TF32 hi, fp32 residual lo) operand split for the 3xTF32 conv path.
//
// The conv / deconv stacks are dense contractions and belong on the tensor cores, but a single
// TF32 pass (10-bit mantissa) costs ~1e-3 relative on the flow fields.  Writing every operand as
// x = hi + lo with hi = round_to_tf32(x), lo = x - hi (exact in fp32) and contracting
//     hi*hi' + hi*lo' + lo*hi'          (the lo*lo' term is ~2^-22 relative and dropped)
// restores fp32-level accuracy with fp32 accumulation in the MMA.  The three products are folded
// into ONE library convolution by concatenating along the contraction dimension:
//     X' = [hi, hi, lo],  W' = [hi', lo', hi']
// This kernel writes that concatenated layout directly: for each of `items` slabs of `inner`
// floats it emits three slabs (order 0: hi,hi,lo   order 1: hi,lo,hi).  HBM-bound: 4 B read +
// 12 B written per element.
#include "common.cuh"

namespace unflow {

__device__ __forceinline__ float tf32_rn(float x) {
  unsigned u = __float_as_uint(x);
  u += 0xFFFu + ((u >> 13) & 1u);   // round to nearest even on the 13 dropped bits
  u &= 0xFFFFE000u;
  return __uint_as_float(u);
}

template <int ORDER>
__global__ void __launch_bounds__(256)
split3_kernel(const float *__restrict__ x, float *__restrict__ out, long long items, long long inner,
              long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long it = i / inner, r = i - it * inner;
    const float v = __ldg(x + i);
    const float hi = tf32_rn(v);
    const float lo = v - hi;
    float *o = out + it * 3 * inner + r;
    o[0] = hi;
    o[inner] = ORDER == 0 ? hi : lo;
    o[2 * inner] = ORDER == 0 ? lo : hi;
  }
}

}  // namespace unflow

extern "C" int unflow_split3_tf32(const float *x, float *out, long long items, long long inner,
                                  int order, void *stream) {
  using namespace unflow;
  UNFLOW_REQUIRE(items >= 0 && inner >= 0, "split3: negative size");
  UNFLOW_REQUIRE(order == 0 || order == 1, "split3: order must be 0 (hi,hi,lo) or 1 (hi,lo,hi)");
  const long long total = items * inner;
  if (total == 0) return UNFLOW_OK;
  UNFLOW_REQUIRE(x && out, "split3: null pointer");
  const int grid = grid_for(total, 256, 16);
  if (order == 0) split3_kernel<0><<<grid, 256, 0, (cudaStream_t)stream>>>(x, out, items, inner, total);
  else split3_kernel<1><<<grid, 256, 0, (cudaStream_t)stream>>>(x, out, items, inner, total);
  count_launch();
  return check_launch("split3_tf32");
}
