// Stand-in for tensorflow/core/util/cuda_kernel_helper.h (TF 1.x), restated from its documentation:
// the grid-stride loop macro, the launch-configuration helper, SetZero and CudaAtomicAdd.
#ifndef UNFLOW_TF_STUB_CUDA_KERNEL_HELPER_H_
#define UNFLOW_TF_STUB_CUDA_KERNEL_HELPER_H_
#include <cuda_runtime.h>
#include "tensorflow/core/framework/tensor_types.h"

#define CUDA_1D_KERNEL_LOOP(i, n) \
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < (n); i += blockDim.x * gridDim.x)

namespace tensorflow {
struct CudaLaunchConfig {
  int virtual_thread_count = -1;   // the work size
  int thread_per_block = -1;
  int block_count = -1;
};

// TF 1.x: threads/block = min(1024, device max); blocks = min(ceil(physical / tpb), #SMs) where
// physical = min(#SMs * max resident threads per SM, work).
inline CudaLaunchConfig GetCudaLaunchConfig(int work_element_count, const Eigen::GpuDevice &) {
  int dev = 0, sms = 1, threads_per_sm = 2048, max_tpb = 1024;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&threads_per_sm, cudaDevAttrMaxThreadsPerMultiProcessor, dev);
  cudaDeviceGetAttribute(&max_tpb, cudaDevAttrMaxThreadsPerBlock, dev);
  CudaLaunchConfig c;
  c.virtual_thread_count = work_element_count;
  long long physical = (long long)sms * threads_per_sm;
  if (physical > work_element_count) physical = work_element_count;
  c.thread_per_block = max_tpb < 1024 ? max_tpb : 1024;
  long long blocks = (physical + c.thread_per_block - 1) / c.thread_per_block;
  c.block_count = (int)(blocks < sms ? blocks : sms);
  if (c.block_count < 1) c.block_count = 1;
  return c;
}

template <typename T>
__global__ void SetZero(const int nthreads, T *bottom_diff) {
  CUDA_1D_KERNEL_LOOP(index, nthreads) { *(bottom_diff + index) = T(0); }
}

template <typename T>
__device__ inline T CudaAtomicAdd(T *ptr, T value) { return atomicAdd(ptr, value); }
}  // namespace tensorflow
#endif
