// wrapper.cu -- C entry points around the reference's OWN CUDA kernels (TEST INFRASTRUCTURE).
//
// /root/reference/ops/{correlation,backward_warp,forward_warp,downsample}_op.cu.cc are compiled as
// they lie (oracle/ref_ops/build.sh), against the stand-in headers of oracle/tf_stub for the few
// TensorFlow declarations they include (tensor maps, the launch-config helper, the grid-stride
// macro); the op-registration files (*_op.cc) need the whole TensorFlow op framework and are not
// compiled.  The result, oracle/_ref/libref_ops.so, runs the reference kernels on the GPU: the ground
// truth for the four custom ops (forward and gradients) and their timing baseline.  This file only
// builds the tensor maps and calls the reference's launchers; it contains no arithmetic.
#include <cuda_runtime.h>

#include "tensorflow/core/framework/tensor_types.h"
#include "correlation_op.h"   // the reference's header (CorrelationAttrs / CorrelationState)

using namespace tensorflow;
typedef Eigen::GpuDevice GPUDevice;
typedef TTypes<float, 4>::Tensor T4;
typedef TTypes<float, 4>::ConstTensor C4;

// launchers defined in the reference's .cu.cc files
void Correlation(const GPUDevice &d, C4 input_0, C4 input_1, T4 output, T4 padded_0, T4 padded_1, CorrelationState st);
void CorrelationGrad(const GPUDevice &d, C4 input_grad, C4 padded_0, C4 padded_1, T4 output_grad_0, T4 output_grad_1,
                     CorrelationState st);
void BackwardWarp(const GPUDevice &d, C4 images, C4 flows, T4 output);
void BackwardWarpGrad(const GPUDevice &d, C4 input_grad, C4 input_images, C4 flows, T4 output_grad);
void ForwardWarp(const GPUDevice &d, C4 flows, T4 output);
void ForwardWarpGrad(const GPUDevice &d, C4 input_grad, C4 flows, T4 output_grad);
void Downsample(const GPUDevice &d, C4 images, T4 output);

static T4 t4(float *p, long long a, long long b, long long c, long long e) { T4 t; t.ptr = p; t.dims[0] = a; t.dims[1] = b; t.dims[2] = c; t.dims[3] = e; return t; }
static C4 c4(const float *p, long long a, long long b, long long c, long long e) { C4 t; t.ptr = p; t.dims[0] = a; t.dims[1] = b; t.dims[2] = c; t.dims[3] = e; return t; }

static CorrelationState corr_state(int H, int W, int C, int ks, int md, int pad, int s1, int s2) {
  CorrelationAttrs a;
  a.kernel_size = ks; a.max_displacement = md; a.pad_size = pad; a.stride_1 = s1; a.stride_2 = s2;
  return CorrelationState(a, H, W, C);
}

static int last_error() { return cudaGetLastError() == cudaSuccess ? 0 : 2; }

extern "C" {
// out_shape[3] = {channels, height, width}; padded buffers are [B, H+2pad, W+2pad, C] floats each
int ref_correlation_out_shape(int C, int H, int W, int ks, int md, int pad, int s1, int s2, int *out_shape) {
  CorrelationState st = corr_state(H, W, C, ks, md, pad, s1, s2);
  out_shape[0] = st.out_channels; out_shape[1] = st.out_height; out_shape[2] = st.out_width;
  return 0;
}

int ref_correlation_fwd(const float *in0, const float *in1, float *out, float *padded0, float *padded1, int B, int C,
                        int H, int W, int ks, int md, int pad, int s1, int s2) {
  CorrelationState st = corr_state(H, W, C, ks, md, pad, s1, s2);
  Correlation(GPUDevice(0), c4(in0, B, C, H, W), c4(in1, B, C, H, W), t4(out, B, st.out_channels, st.out_height, st.out_width),
              t4(padded0, B, st.padded_height, st.padded_width, C), t4(padded1, B, st.padded_height, st.padded_width, C), st);
  return last_error();
}

int ref_correlation_bwd(const float *gout, const float *padded0, const float *padded1, float *g0, float *g1, int B, int C,
                        int H, int W, int ks, int md, int pad, int s1, int s2) {
  CorrelationState st = corr_state(H, W, C, ks, md, pad, s1, s2);
  CorrelationGrad(GPUDevice(0), c4(gout, B, st.out_channels, st.out_height, st.out_width),
                  c4(padded0, B, st.padded_height, st.padded_width, C), c4(padded1, B, st.padded_height, st.padded_width, C),
                  t4(g0, B, C, H, W), t4(g1, B, C, H, W), st);
  return last_error();
}

int ref_backward_warp_fwd(const float *images, const float *flows, float *out, int B, int H, int W, int C) {
  BackwardWarp(GPUDevice(0), c4(images, B, H, W, C), c4(flows, B, H, W, 2), t4(out, B, H, W, C));
  return last_error();
}

int ref_backward_warp_bwd(const float *grad, const float *images, const float *flows, float *dflow, int B, int H, int W, int C) {
  BackwardWarpGrad(GPUDevice(0), c4(grad, B, H, W, C), c4(images, B, H, W, C), c4(flows, B, H, W, 2), t4(dflow, B, H, W, 2));
  return last_error();
}

int ref_forward_warp_fwd(const float *flows, float *out, int B, int H, int W) {
  ForwardWarp(GPUDevice(0), c4(flows, B, H, W, 2), t4(out, B, H, W, 1));
  return last_error();
}

int ref_forward_warp_bwd(const float *grad, const float *flows, float *dflow, int B, int H, int W) {
  ForwardWarpGrad(GPUDevice(0), c4(grad, B, H, W, 1), c4(flows, B, H, W, 2), t4(dflow, B, H, W, 2));
  return last_error();
}

int ref_downsample(const float *images, float *out, int B, int H, int W, int C, int scale) {
  Downsample(GPUDevice(0), c4(images, B, H, W, C), t4(out, B, H / scale, W / scale, C));
  return last_error();
}
}
