/*
 * oracle_ops.c -- CPU restatement of the four UnFlow custom ops.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under unflow_b200/ may link, import or
 * call this file; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and only as the checker or as
 * the timed CPU baseline, never as the product.
 *
 * The reference implements these ops as CUDA kernels only (there is no CPU
 * kernel registered), and the .cu.cc files need TensorFlow
 * headers, so they cannot be compiled here.  Each function below re-expresses
 * one reference kernel as plain C loops: one loop iteration per CUDA thread /
 * block of the reference launch, same index arithmetic, same order of the
 * floating-point operations inside a thread.  Differences that remain:
 *   - nvcc contracts a*b+c into FMA; this file is built with
 *     -ffp-contract=off (portable across host CPUs) -> <= 1 ulp per op.
 *   - forward_warp uses atomicAdd on the GPU (order undefined); here the
 *     additions happen in source-pixel order.
 *
 * Citations are file:line under /root/reference/.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ */
/* Correlation geometry: ops/correlation_op.h:28-52 (CorrelationState) */
/* ------------------------------------------------------------------ */
typedef struct {
  int pad, s1, s2, md, ks;
  int kr, border, ngr, ngw;
  int ph, pw, oh, ow, oc;
} corr_geom;

static corr_geom make_geom(int H, int W, int ks, int md, int pad, int s1, int s2) {
  corr_geom g;
  g.pad = pad; g.s1 = s1; g.s2 = s2; g.md = md; g.ks = ks;
  g.ph = H + 2 * pad;
  g.pw = W + 2 * pad;
  g.kr = (ks - 1) / 2;
  g.border = md + g.kr;
  g.ngr = md / s2;
  g.ngw = g.ngr * 2 + 1;
  g.ow = (int)ceilf((float)(g.pw - g.border * 2) / (float)s1);
  g.oh = (int)ceilf((float)(g.ph - g.border * 2) / (float)s1);
  g.oc = g.ngw * g.ngw;
  return g;
}

/* Output shape query. Returns 0 on success, 1 on invalid settings
 * (ops/correlation_op.h:16-17 odd kernel; ops/correlation_op.cc:60-61). */
int oracle_correlation_shape(int H, int W, int ks, int md, int pad, int s1, int s2,
                             int *oc, int *oh, int *ow) {
  if (ks % 2 == 0) return 1;
  corr_geom g = make_geom(H, W, ks, md, pad, s1, s2);
  if (g.ow <= 0 || g.oh <= 0) return 1; /* ref checks ow*oh > 0; both negative would crash in TF allocation */
  *oc = g.oc; *oh = g.oh; *ow = g.ow;
  return 0;
}

/* blob_rearrange_kernel2 (ops/correlation_op.cu.cc:31-49) preceded by the
 * cudaMemset (ops/correlation_op.cu.cc:282-283): NCHW -> zero padded NHWC. */
static void rearrange_pad(const float *in, float *out, int B, int C, int H, int W, int pad) {
  const int pw = W + 2 * pad, ph = H + 2 * pad;
  memset(out, 0, sizeof(float) * (size_t)B * ph * pw * C);
  for (int n = 0; n < B; ++n)
    for (int ch = 0; ch < C; ++ch)
      for (int xy = 0; xy < H * W; ++xy) {
        float v = in[((size_t)n * C + ch) * H * W + xy];
        int xpad = xy % W + pad;
        int ypad = xy / W + pad;
        out[(((size_t)n * ph + ypad) * pw + xpad) * C + ch] = v;
      }
}

/* CorrelateData (ops/correlation_op.cu.cc:52-117): one 32-thread block per
 * output pixel; lane l accumulates channels l, l+32, ...; lane 0 adds the 32
 * partial sums serially and divides by ks*ks*C. */
int oracle_correlation_fwd(const float *in0, const float *in1, float *out,
                           float *padded0, float *padded1,
                           int B, int C, int H, int W,
                           int ks, int md, int pad, int s1, int s2) {
  if (ks % 2 == 0) return 1;
  corr_geom g = make_geom(H, W, ks, md, pad, s1, s2);
  if (g.ow <= 0 || g.oh <= 0) return 1; /* ref checks ow*oh > 0; both negative would crash in TF allocation */
  rearrange_pad(in0, padded0, B, C, H, W, pad);
  rearrange_pad(in1, padded1, B, C, H, W, pad);
  const int topcount = g.ow * g.oh * g.oc;
  const int sumelems = ks * ks * C;
#pragma omp parallel for collapse(2) schedule(static)
  for (int item = 0; item < B; ++item)
    for (int by = 0; by < g.oh; ++by)
      for (int bx = 0; bx < g.ow; ++bx) {
        const int x1 = bx * s1 + md;
        const int y1 = by * s1 + md;
        for (int tc = 0; tc < g.oc; ++tc) {
          const int s2o = (tc % g.ngw - g.ngr) * s2;
          const int s2p = (tc / g.ngw - g.ngr) * s2;
          float lane_sum[32];
          for (int l = 0; l < 32; ++l) lane_sum[l] = 0.f;
          for (int j = 0; j < ks; ++j)
            for (int i = 0; i < ks; ++i)
              for (int ch = 0; ch < C; ++ch) {
                const size_t idx1 = (((size_t)item * g.ph + y1 + j) * g.pw + x1 + i) * C + ch;
                const size_t idx2 = (((size_t)item * g.ph + y1 + s2p + j) * g.pw + x1 + s2o + i) * C + ch;
                lane_sum[ch & 31] += padded0[idx1] * padded1[idx2];
              }
          float total = 0.f;
          for (int l = 0; l < 32; ++l) total += lane_sum[l];
          out[(size_t)item * topcount + ((size_t)tc * g.oh + by) * g.ow + bx] =
              total / (float)sumelems;
        }
      }
  return 0;
}

#define ROUND_OFF 50000

/* CorrelateDataBackward0 (ops/correlation_op.cu.cc:120-181) */
static void corr_bwd0(const corr_geom *g, int B, int C, int H, int W,
                      const float *padded1, const float *topdiff, float *g0) {
  const int bottomcount = C * H * W;
  const int round_off = ROUND_OFF, round_off_s1 = g->s1 * round_off;
#pragma omp parallel for collapse(2) schedule(static)
  for (int item = 0; item < B; ++item)
    for (int index = 0; index < bottomcount; ++index) {
      const int n = index % C;
      const int l = (index / C) % W + g->pad;
      const int m = (index / C / W) % H + g->pad;
      int xmin = (l - 2 * g->kr - g->md + round_off_s1 - 1) / g->s1 + 1 - round_off;
      int ymin = (m - 2 * g->kr - g->md + round_off_s1 - 1) / g->s1 + 1 - round_off;
      int xmax = (l - g->md + round_off_s1) / g->s1 - round_off;
      int ymax = (m - g->md + round_off_s1) / g->s1 - round_off;
      float sum = 0.f;
      if (xmax >= 0 && ymax >= 0 && xmin <= g->ow - 1 && ymin <= g->oh - 1) {
        if (xmin < 0) xmin = 0;
        if (xmax > g->ow - 1) xmax = g->ow - 1;
        if (ymin < 0) ymin = 0;
        if (ymax > g->oh - 1) ymax = g->oh - 1;
        for (int p = -g->ngr; p <= g->ngr; ++p)
          for (int o = -g->ngr; o <= g->ngr; ++o) {
            const int s2o = g->s2 * o, s2p = g->s2 * p;
            const float bot1 =
                padded1[(((size_t)item * g->ph + (m + s2p)) * g->pw + (l + s2o)) * C + n];
            const int op = (p + g->ngr) * g->ngw + (o + g->ngr);
            const size_t off = (size_t)item * g->oc + op;
            for (int y = ymin; y <= ymax; ++y)
              for (int x = xmin; x <= xmax; ++x)
                sum += topdiff[(off * g->oh + y) * g->ow + x] * bot1;
          }
      }
      const int sumelems = (g->kr * 2 + 1) * (g->kr * 2 + 1) * C;
      g0[(size_t)item * bottomcount + ((size_t)n * H + (m - g->pad)) * W + (l - g->pad)] =
          sum / (float)sumelems;
    }
}

/* CorrelateDataBackward1 (ops/correlation_op.cu.cc:184-248) */
static void corr_bwd1(const corr_geom *g, int B, int C, int H, int W,
                      const float *padded0, const float *topdiff, float *g1) {
  const int bottomcount = C * H * W;
  const int round_off = ROUND_OFF, round_off_s1 = g->s1 * round_off;
#pragma omp parallel for collapse(2) schedule(static)
  for (int item = 0; item < B; ++item)
    for (int index = 0; index < bottomcount; ++index) {
      const int n = index % C;
      const int l = (index / C) % W + g->pad;
      const int m = (index / C / W) % H + g->pad;
      float sum = 0.f;
      for (int p = -g->ngr; p <= g->ngr; ++p)
        for (int o = -g->ngr; o <= g->ngr; ++o) {
          const int s2o = g->s2 * o, s2p = g->s2 * p;
          int xmin = (l - 2 * g->kr - g->md - s2o + round_off_s1 - 1) / g->s1 + 1 - round_off;
          int ymin = (m - 2 * g->kr - g->md - s2p + round_off_s1 - 1) / g->s1 + 1 - round_off;
          int xmax = (l - g->md - s2o + round_off_s1) / g->s1 - round_off;
          int ymax = (m - g->md - s2p + round_off_s1) / g->s1 - round_off;
          if (xmax >= 0 && ymax >= 0 && xmin <= g->ow - 1 && ymin <= g->oh - 1) {
            if (xmin < 0) xmin = 0;
            if (xmax > g->ow - 1) xmax = g->ow - 1;
            if (ymin < 0) ymin = 0;
            if (ymax > g->oh - 1) ymax = g->oh - 1;
            const float bot0 =
                padded0[(((size_t)item * g->ph + (m - s2p)) * g->pw + (l - s2o)) * C + n];
            const int op = (p + g->ngr) * g->ngw + (o + g->ngr);
            const size_t off = (size_t)item * g->oc + op;
            for (int y = ymin; y <= ymax; ++y)
              for (int x = xmin; x <= xmax; ++x)
                sum += topdiff[(off * g->oh + y) * g->ow + x] * bot0;
          }
        }
      const int sumelems = (g->kr * 2 + 1) * (g->kr * 2 + 1) * C;
      g1[(size_t)item * bottomcount + ((size_t)n * H + (m - g->pad)) * W + (l - g->pad)] =
          sum / (float)sumelems;
    }
}

/* CorrelationGrad (ops/correlation_op.cu.cc:317-390): per-item launches of
 * Backward0 then Backward1 on the padded NHWC copies saved by the forward. */
int oracle_correlation_bwd(const float *topdiff, const float *padded0, const float *padded1,
                           float *g0, float *g1, int B, int C, int H, int W,
                           int ks, int md, int pad, int s1, int s2) {
  if (ks % 2 == 0) return 1;
  corr_geom g = make_geom(H, W, ks, md, pad, s1, s2);
  if (g.ow <= 0 || g.oh <= 0) return 1; /* ref checks ow*oh > 0; both negative would crash in TF allocation */
  corr_bwd0(&g, B, C, H, W, padded1, topdiff, g0);
  corr_bwd1(&g, B, C, H, W, padded0, topdiff, g1);
  return 0;
}

/* ------------------------------------------------------------------ */
/* BackwardWarp: ops/backward_warp_op.cu.cc:14-68 (zero outside)       */
/* ------------------------------------------------------------------ */
void oracle_backward_warp_fwd(const float *images, const float *flows, float *output,
                              int B, int H, int W, int C) {
  const int total = B * H * W;
#pragma omp parallel for schedule(static)
  for (int out_idx = 0; out_idx < total; ++out_idx) {
    int idx = out_idx;
    const int src_x = idx % W; idx /= W;
    const int src_y = idx % H;
    const int b = idx / H;
    const float x = src_x + flows[out_idx * 2];
    const float y = src_y + flows[out_idx * 2 + 1];
    const int x0 = (int)floorf(x), x1 = x0 + 1;
    const int y0 = (int)floorf(y), y1 = y0 + 1;
    const float w_right = x - x0, w_left = x1 - x;
    const float w_bottom = y - y0, w_top = y1 - y;
    for (int c = 0; c < C; ++c) {
      float sum = 0.0f;
#define IMG(iy, ix) images[c + (size_t)C * ((ix) + (size_t)W * ((iy) + (size_t)H * b))]
      if (x0 >= 0 && x0 < W && y0 >= 0 && y0 < H) sum += w_left * w_top * IMG(y0, x0);
      if (x1 >= 0 && x1 < W && y0 >= 0 && y0 < H) sum += w_right * w_top * IMG(y0, x1);
      if (x0 >= 0 && x0 < W && y1 >= 0 && y1 < H) sum += w_left * w_bottom * IMG(y1, x0);
      if (x1 >= 0 && x1 < W && y1 >= 0 && y1 < H) sum += w_right * w_bottom * IMG(y1, x1);
#undef IMG
      output[(size_t)out_idx * C + c] = sum;
    }
  }
}

/* BackwardWarpGrad: ops/backward_warp_op.cu.cc:70-138 (gradient w.r.t. flow only) */
void oracle_backward_warp_bwd(const float *grad, const float *images, const float *flows,
                              float *out_grad, int B, int H, int W, int C) {
  const int total = B * H * W;
#pragma omp parallel for schedule(static)
  for (int in_idx = 0; in_idx < total; ++in_idx) {
    int idx = in_idx;
    const int src_x = idx % W; idx /= W;
    const int src_y = idx % H;
    const int b = idx / H;
    const float x = src_x + flows[in_idx * 2];
    const float y = src_y + flows[in_idx * 2 + 1];
    const int x0 = (int)floorf(x), x1 = x0 + 1;
    const int y0 = (int)floorf(y), y1 = y0 + 1;
    const float w_right = x - x0, w_left = x1 - x;
    const float w_bottom = y - y0, w_top = y1 - y;
    float du = 0.0f, dv = 0.0f;
    for (int c = 0; c < C; ++c) {
      float px;
      const float din = grad[c + (size_t)C * in_idx];
#define IMG(iy, ix) images[c + (size_t)C * ((ix) + (size_t)W * ((iy) + (size_t)H * b))]
      if (x0 >= 0 && x0 < W && y0 >= 0 && y0 < H) { px = IMG(y0, x0) * din; du -= w_top * px; dv -= w_left * px; }
      if (x1 >= 0 && x1 < W && y0 >= 0 && y0 < H) { px = IMG(y0, x1) * din; du += w_top * px; dv -= w_right * px; }
      if (x0 >= 0 && x0 < W && y1 >= 0 && y1 < H) { px = IMG(y1, x0) * din; du -= w_bottom * px; dv += w_left * px; }
      if (x1 >= 0 && x1 < W && y1 >= 0 && y1 < H) { px = IMG(y1, x1) * din; du += w_bottom * px; dv += w_right * px; }
#undef IMG
    }
    out_grad[(size_t)in_idx * 2] = du;
    out_grad[(size_t)in_idx * 2 + 1] = dv;
  }
}

/* ------------------------------------------------------------------ */
/* ForwardWarp: ops/forward_warp_op.cu.cc:16-65                        */
/* Serial over source pixels (the GPU order of atomicAdd is undefined).*/
/* ------------------------------------------------------------------ */
void oracle_forward_warp_fwd(const float *flows, float *output, int B, int H, int W) {
  const int total = B * H * W;
  memset(output, 0, sizeof(float) * (size_t)total);
  const float dist = 2.0f;
  const float std_ = dist * 0.5f;
  const int k = (int)ceilf(dist + 2);
  for (int out_idx = 0; out_idx < total; ++out_idx) {
    int idx = out_idx;
    const int src_x = idx % W; idx /= W;
    const int src_y = idx % H;
    const int b = idx / H;
    const float tx = src_x + flows[out_idx * 2];
    const float ty = src_y + flows[out_idx * 2 + 1];
    if (floorf(tx - k) < W && floorf(tx + k) >= 0 && floorf(ty - k) < H && floorf(ty + k) >= 0) {
      const int min_n_x = tx - k > 0 ? (int)floorf(tx - k) : 0;
      const int min_n_y = ty - k > 0 ? (int)floorf(ty - k) : 0;
      const int max_n_x = tx + k < W ? (int)floorf(tx + k) : W - 1;
      const int max_n_y = ty + k < H ? (int)floorf(ty + k) : H - 1;
      const float gauss_divisor = 2 * powf(std_, 2);
      for (int n_x = min_n_x; n_x <= max_n_x; ++n_x)
        for (int n_y = min_n_y; n_y <= max_n_y; ++n_y) {
          const float x = n_x - tx, y = n_y - ty;
          const float weight = expf(-(powf(x, 2) + powf(y, 2)) / gauss_divisor);
          output[n_x + (size_t)W * (n_y + (size_t)H * b)] += weight;
        }
    }
  }
}

/* ForwardWarpGrad: ops/forward_warp_op.cu.cc:67-125 */
void oracle_forward_warp_bwd(const float *grad, const float *flows, float *out_grad,
                             int B, int H, int W) {
  const int total = B * H * W;
  const float dist = 2.0f;
  const float std_ = dist * 0.5f;
  const int k = (int)ceilf(dist + 2);
#pragma omp parallel for schedule(static)
  for (int in_idx = 0; in_idx < total; ++in_idx) {
    int idx = in_idx;
    const int src_x = idx % W; idx /= W;
    const int src_y = idx % H;
    const int b = idx / H;
    const float tx = src_x + flows[in_idx * 2];
    const float ty = src_y + flows[in_idx * 2 + 1];
    float du = 0.0f, dv = 0.0f;
    if (floorf(tx - k) < W && floorf(tx + k) >= 0 && floorf(ty - k) < H && floorf(ty + k) >= 0) {
      const int min_n_x = tx - k > 0 ? (int)floorf(tx - k) : 0;
      const int min_n_y = ty - k > 0 ? (int)floorf(ty - k) : 0;
      const int max_n_x = tx + k < W ? (int)floorf(tx + k) : W - 1;
      const int max_n_y = ty + k < H ? (int)floorf(ty + k) : H - 1;
      const float gauss_divisor = 2 * powf(std_, 2);
      for (int n_x = min_n_x; n_x <= max_n_x; ++n_x)
        for (int n_y = min_n_y; n_y <= max_n_y; ++n_y) {
          const float x = n_x - tx, y = n_y - ty;
          const float weight = expf(-(powf(x, 2) + powf(y, 2)) / gauss_divisor);
          const float din = grad[n_x + (size_t)W * (n_y + (size_t)H * b)];
          const float factor = 2 * din * weight / gauss_divisor;
          du += factor * x;
          dv += factor * y;
        }
    }
    out_grad[(size_t)in_idx * 2] = du;
    out_grad[(size_t)in_idx * 2 + 1] = dv;
  }
}

/* ------------------------------------------------------------------ */
/* Downsample: ops/downsample_op.cu.cc:15-49; shape rule and the       */
/* divisibility check ops/downsample_op.cc:37-47.                      */
/* ------------------------------------------------------------------ */
int oracle_downsample(const float *images, float *output, int B, int H, int W, int C, int scale) {
  if (scale < 1 || H % scale != 0 || W % scale != 0) return 1;
  const int oh = H / scale, ow = W / scale;
  const int total = B * oh * ow * C;
#pragma omp parallel for schedule(static)
  for (int out_idx = 0; out_idx < total; ++out_idx) {
    int idx = out_idx;
    const int c = idx % C; idx /= C;
    const int x = idx % ow; idx /= ow;
    const int y = idx % oh;
    const int b = idx / oh;
    const int scale_y = H / oh, scale_x = W / ow;
    const int min_in_y = y * scale_y, min_in_x = x * scale_x;
    const int max_in_y = min_in_y + scale_y, max_in_x = min_in_x + scale_x;
    float sum = 0.0f;
    for (int in_y = min_in_y; in_y < max_in_y; ++in_y)
      for (int in_x = min_in_x; in_x < max_in_x; ++in_x)
        sum += images[c + (size_t)C * (in_x + (size_t)W * (in_y + (size_t)H * b))];
    sum /= scale_x * scale_y;
    output[c + (size_t)C * (x + (size_t)ow * (y + (size_t)oh * b))] = sum;
  }
  return 0;
}

void oracle_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
