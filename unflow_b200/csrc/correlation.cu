// correlation.cu -- cost-volume correlation: geometry, argument checks, dispatch and the
// GENERIC kernels (any kernel_size / strides / pad).  The FlowNetC configuration
// (kernel_size=1, stride_1=1, stride_2=2, pad == max_displacement) is served by the tiled
// TMA kernels in correlation_tiled.cu.
//
// Reference: ops/correlation_op.h:28-52 (geometry), ops/correlation_op.cu.cc:52-248 (kernels).
// Unlike the reference no zero-padded NHWC copies are made (its blob_rearrange_kernel2 pass,
// ops/correlation_op.cu.cc:31-49,282-293): out-of-image taps are predicated to zero.
#include <cmath>

#include "common.cuh"
#include "correlation.cuh"

namespace unflow {

int make_corr_geom(CorrGeom &g, int B, int C, int H, int W, int ks, int md, int pad, int s1, int s2) {
  UNFLOW_REQUIRE(B >= 0 && C >= 1 && H >= 1 && W >= 1, "correlation: bad input shape");
  UNFLOW_REQUIRE(ks >= 1 && ks % 2 != 0, "kernel_size must be odd");
  UNFLOW_REQUIRE(s1 >= 1 && s2 >= 1 && md >= 0 && pad >= 0, "correlation: bad attribute");
  g.B = B; g.C = C; g.H = H; g.W = W;
  g.ks = ks; g.md = md; g.pad = pad; g.s1 = s1; g.s2 = s2;
  g.kr = (ks - 1) / 2;
  g.border = md + g.kr;
  g.ngr = md / s2;
  g.ngw = 2 * g.ngr + 1;
  const int ph = H + 2 * pad, pw = W + 2 * pad;
  g.ow = (int)ceilf((float)(pw - g.border * 2) / (float)s1);
  g.oh = (int)ceilf((float)(ph - g.border * 2) / (float)s1);
  g.oc = g.ngw * g.ngw;
  // reference: OP_REQUIRES(out_width * out_height > 0) (correlation_op.cc:60-61); two negative
  // extents would pass that product test and then fail in the TF allocator, so require both > 0.
  UNFLOW_REQUIRE(g.ow > 0 && g.oh > 0, "Invalid correlation settings");
  return UNFLOW_OK;
}

// ---------------------------------------------------------------------------------------------
// Generic forward: one thread per output element, x fastest (coalesced along rows when s1 == 1).
// Sums over (j, i, c) and divides by K*K*C with a true division as the reference does (:112-114).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
corr_fwd_generic_kernel(const float *__restrict__ in0, const float *__restrict__ in1,
                        float *__restrict__ out, CorrGeom g, long long total) {
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    long long t = idx;
    const int x = (int)(t % g.ow); t /= g.ow;
    const int y = (int)(t % g.oh); t /= g.oh;
    const int tc = (int)(t % g.oc);
    const int b = (int)(t / g.oc);
    const int s2o = (tc % g.ngw - g.ngr) * g.s2;
    const int s2p = (tc / g.ngw - g.ngr) * g.s2;
    // padded coordinates of the patch corner in image 0, then shift to unpadded
    const int x1 = x * g.s1 + g.md - g.pad;
    const int y1 = y * g.s1 + g.md - g.pad;
    const float *a = in0 + (long long)b * g.C * g.H * g.W;
    const float *bb = in1 + (long long)b * g.C * g.H * g.W;
    float sum = 0.0f;
    for (int j = 0; j < g.ks; ++j)
      for (int i = 0; i < g.ks; ++i) {
        const int ya = y1 + j, xa = x1 + i;
        const int yb = ya + s2p, xb = xa + s2o;
        const bool ok = ya >= 0 && ya < g.H && xa >= 0 && xa < g.W &&
                        yb >= 0 && yb < g.H && xb >= 0 && xb < g.W;
        if (!ok) continue;  // a zero-padded operand: the product is 0
        const float *pa = a + (long long)ya * g.W + xa;
        const float *pb = bb + (long long)yb * g.W + xb;
        for (int c = 0; c < g.C; ++c)
          sum += __ldg(pa + (long long)c * g.H * g.W) * __ldg(pb + (long long)c * g.H * g.W);
      }
    out[idx] = sum / (float)(g.ks * g.ks * g.C);
  }
}

// ---------------------------------------------------------------------------------------------
// Generic backward (both gradients), one thread per input element (b,c,y,x), x fastest.
// Same window arithmetic as CorrelateDataBackward0/1 (reference :120-248); integer floor/ceil
// divisions are written directly instead of the ROUND_OFF trick.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int floor_div(int a, int b) { return a >= 0 ? a / b : -((-a + b - 1) / b); }
__device__ __forceinline__ int ceil_divi(int a, int b) { return -floor_div(-a, b); }

template <int WHICH>
__global__ void __launch_bounds__(256)
corr_bwd_generic_kernel(const float *__restrict__ gout, const float *__restrict__ other,
                        float *__restrict__ gin, CorrGeom g, long long total) {
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    long long t = idx;
    const int x = (int)(t % g.W); t /= g.W;
    const int y = (int)(t % g.H); t /= g.H;
    const int c = (int)(t % g.C);
    const int b = (int)(t / g.C);
    const int l = x + g.pad, m = y + g.pad;  // padded coordinates
    const float *oth = other + ((long long)b * g.C + c) * g.H * g.W;
    const float *go = gout + (long long)b * g.oc * g.oh * g.ow;
    float sum = 0.0f;
    for (int p = -g.ngr; p <= g.ngr; ++p)
      for (int o = -g.ngr; o <= g.ngr; ++o) {
        const int s2o = g.s2 * o, s2p = g.s2 * p;
        const int sh_x = WHICH == 0 ? 0 : s2o, sh_y = WHICH == 0 ? 0 : s2p;
        int xmin = ceil_divi(l - 2 * g.kr - g.md - sh_x, g.s1);
        int ymin = ceil_divi(m - 2 * g.kr - g.md - sh_y, g.s1);
        int xmax = floor_div(l - g.md - sh_x, g.s1);
        int ymax = floor_div(m - g.md - sh_y, g.s1);
        if (!(xmax >= 0 && ymax >= 0 && xmin <= g.ow - 1 && ymin <= g.oh - 1)) continue;
        xmin = max(0, xmin); xmax = min(g.ow - 1, xmax);
        ymin = max(0, ymin); ymax = min(g.oh - 1, ymax);
        // the other image at (l +/- s2o, m +/- s2p) in padded coords
        const int ox = (WHICH == 0 ? l + s2o : l - s2o) - g.pad;
        const int oy = (WHICH == 0 ? m + s2p : m - s2p) - g.pad;
        if (ox < 0 || ox >= g.W || oy < 0 || oy >= g.H) continue;  // zero padding
        const float v = __ldg(oth + (long long)oy * g.W + ox);
        const int op = (p + g.ngr) * g.ngw + (o + g.ngr);
        const float *gch = go + (long long)op * g.oh * g.ow;
        for (int yy = ymin; yy <= ymax; ++yy)
          for (int xx = xmin; xx <= xmax; ++xx) sum += __ldg(gch + (long long)yy * g.ow + xx) * v;
      }
    gin[idx] = sum / (float)((g.kr * 2 + 1) * (g.kr * 2 + 1) * g.C);
  }
}

int corr_fwd_generic(const float *in0, const float *in1, float *out, const CorrGeom &g, cudaStream_t s) {
  const long long total = (long long)g.B * g.oc * g.oh * g.ow;
  if (total == 0) return UNFLOW_OK;
  corr_fwd_generic_kernel<<<grid_for(total, 256, 16), 256, 0, s>>>(in0, in1, out, g, total);
  count_launch();
  return check_launch("correlation_fwd(generic)");
}

int corr_bwd_generic(const float *gout, const float *in0, const float *in1, float *g0, float *g1,
                     const CorrGeom &g, cudaStream_t s) {
  const long long total = (long long)g.B * g.C * g.H * g.W;
  if (total == 0) return UNFLOW_OK;
  corr_bwd_generic_kernel<0><<<grid_for(total, 256, 16), 256, 0, s>>>(gout, in1, g0, g, total);
  corr_bwd_generic_kernel<1><<<grid_for(total, 256, 16), 256, 0, s>>>(gout, in0, g1, g, total);
  count_launch(2);
  return check_launch("correlation_bwd(generic)");
}

}  // namespace unflow

using namespace unflow;

extern "C" int unflow_correlation_out_shape(int H, int W, int kernel_size, int max_displacement,
                                            int pad, int stride_1, int stride_2, int *out_c,
                                            int *out_h, int *out_w) {
  CorrGeom g;
  int rc = make_corr_geom(g, 1, 1, H, W, kernel_size, max_displacement, pad, stride_1, stride_2);
  if (rc != UNFLOW_OK) return rc;
  if (out_c) *out_c = g.oc;
  if (out_h) *out_h = g.oh;
  if (out_w) *out_w = g.ow;
  return UNFLOW_OK;
}

extern "C" size_t unflow_correlation_workspace_bytes(int, int, int, int, int, int, int, int, int) {
  return 0;  // no padded copies, no scratch
}

extern "C" int unflow_correlation_fwd_path(int C, int H, int W, int kernel_size,
                                           int max_displacement, int pad, int stride_1,
                                           int stride_2) {
  CorrGeom g;
  if (make_corr_geom(g, 1, C, H, W, kernel_size, max_displacement, pad, stride_1, stride_2) != UNFLOW_OK)
    return -1;
  return corr_tiled_supported(g) ? 1 : 0;
}

extern "C" int unflow_correlation_fwd(const float *in0, const float *in1, float *out, int B, int C,
                                      int H, int W, int kernel_size, int max_displacement, int pad,
                                      int stride_1, int stride_2, void *stream) {
  CorrGeom g;
  int rc = make_corr_geom(g, B, C, H, W, kernel_size, max_displacement, pad, stride_1, stride_2);
  if (rc != UNFLOW_OK) return rc;
  if (B == 0) return UNFLOW_OK;
  UNFLOW_REQUIRE(in0 && in1 && out, "correlation: null pointer");
  if (corr_tiled_supported(g)) return corr_fwd_tiled(in0, in1, out, nullptr, g, (cudaStream_t)stream);
  return corr_fwd_generic(in0, in1, out, g, (cudaStream_t)stream);
}

// Both cost volumes of the bidirectional pass (flownet.py:34-44) in ONE launch: out = corr(in0, in1),
// out_rev = corr(in1, in0), the second obtained by re-indexing the accumulators of the first
// (corr(in1,in0)[(-p,-o)](y+2p, x+2o) == corr(in0,in1)[(p,o)](y,x)); bit-identical to two launches.
// Only the FlowNetC geometry of the tiled kernel (unflow_correlation_fwd_path == 1); UNFLOW_EINVAL else.
extern "C" int unflow_correlation_fwd_bidir(const float *in0, const float *in1, float *out, float *out_rev,
                                            int B, int C, int H, int W, int kernel_size, int max_displacement,
                                            int pad, int stride_1, int stride_2, void *stream) {
  CorrGeom g;
  int rc = make_corr_geom(g, B, C, H, W, kernel_size, max_displacement, pad, stride_1, stride_2);
  if (rc != UNFLOW_OK) return rc;
  if (B == 0) return UNFLOW_OK;
  UNFLOW_REQUIRE(in0 && in1 && out && out_rev, "correlation_bidir: null pointer");
  UNFLOW_REQUIRE(corr_tiled_supported(g), "correlation_bidir: attributes / shape not served by the tiled kernel");
  return corr_fwd_tiled(in0, in1, out, out_rev, g, (cudaStream_t)stream);
}

// gout_eff = gout + re-indexed gout_rev: the gradient of both cost volumes w.r.t. (in0, in1) is then
// unflow_correlation_bwd(gout_eff, in0, in1).
extern "C" int unflow_correlation_fold_grad(const float *gout, const float *gout_rev, float *gout_eff, int B,
                                            int C, int H, int W, int kernel_size, int max_displacement, int pad,
                                            int stride_1, int stride_2, void *stream) {
  CorrGeom g;
  int rc = make_corr_geom(g, B, C, H, W, kernel_size, max_displacement, pad, stride_1, stride_2);
  if (rc != UNFLOW_OK) return rc;
  if (B == 0) return UNFLOW_OK;
  UNFLOW_REQUIRE(gout && gout_rev && gout_eff, "correlation_fold_grad: null pointer");
  UNFLOW_REQUIRE(corr_tiled_supported(g), "correlation_fold_grad: attributes / shape not served by the tiled kernel");
  UNFLOW_REQUIRE((((uintptr_t)gout | (uintptr_t)gout_rev | (uintptr_t)gout_eff) & 7) == 0, "correlation_fold_grad: pointers must be 8-byte aligned");
  return corr_fold_grad(gout, gout_rev, gout_eff, g, (cudaStream_t)stream);
}

extern "C" int unflow_correlation_bwd(const float *gout, const float *in0, const float *in1,
                                      float *g0, float *g1, int B, int C, int H, int W,
                                      int kernel_size, int max_displacement, int pad, int stride_1,
                                      int stride_2, void *stream) {
  CorrGeom g;
  int rc = make_corr_geom(g, B, C, H, W, kernel_size, max_displacement, pad, stride_1, stride_2);
  if (rc != UNFLOW_OK) return rc;
  if (B == 0) return UNFLOW_OK;
  UNFLOW_REQUIRE(gout && in0 && in1 && g0 && g1, "correlation_grad: null pointer");
  if (corr_tiled_supported(g)) return corr_bwd_tiled(gout, in0, in1, g0, g1, g, (cudaStream_t)stream);
  return corr_bwd_generic(gout, in0, in1, g0, g1, g, (cudaStream_t)stream);
}
