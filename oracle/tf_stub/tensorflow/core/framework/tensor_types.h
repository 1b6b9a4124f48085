// Stand-in for the TensorFlow header of the same name (TEST INFRASTRUCTURE, see oracle/ref_ops/README).
// The reference's CUDA files use exactly this much of it: TTypes<float, 4>::{Tensor, ConstTensor}
// with .dimension(i), .data() and .size(), and Eigen::GpuDevice::stream().
#ifndef UNFLOW_TF_STUB_TENSOR_TYPES_H_
#define UNFLOW_TF_STUB_TENSOR_TYPES_H_
#include <cuda_runtime.h>

namespace Eigen {
struct GpuDevice {
  cudaStream_t stream_;
  explicit GpuDevice(cudaStream_t s = 0) : stream_(s) {}
  cudaStream_t stream() const { return stream_; }
};
struct ThreadPoolDevice {};
}  // namespace Eigen

namespace tensorflow {
template <typename T, int N>
struct StubTensorMap {
  T *ptr;
  long long dims[N];
  long long dimension(int i) const { return dims[i]; }
  T *data() const { return ptr; }
  long long size() const {
    long long n = 1;
    for (int i = 0; i < N; ++i) n *= dims[i];
    return n;
  }
};

template <typename T, int N = 1>
struct TTypes {
  typedef StubTensorMap<T, N> Tensor;
  typedef StubTensorMap<const T, N> ConstTensor;
};
}  // namespace tensorflow
#endif
