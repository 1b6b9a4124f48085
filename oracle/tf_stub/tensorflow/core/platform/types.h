// Stand-in for tensorflow/core/platform/types.h: the fixed-width integer names the kernels use.
#ifndef UNFLOW_TF_STUB_TYPES_H_
#define UNFLOW_TF_STUB_TYPES_H_
#include <cstdint>
namespace tensorflow {
typedef std::int32_t int32;
typedef std::int64_t int64;
typedef std::uint8_t uint8;
typedef std::uint32_t uint32;
}  // namespace tensorflow
#endif
