// Stand-in for tensorflow/core/framework/op_kernel.h: correlation_op.h declares an attribute struct
// whose constructor reads op attributes through OpKernelConstruction; the kernels never call it.
#ifndef UNFLOW_TF_STUB_OP_KERNEL_H_
#define UNFLOW_TF_STUB_OP_KERNEL_H_
#include <cmath>
using std::ceil;
namespace tensorflow {
struct Status {};
namespace errors {
inline Status InvalidArgument(const char *) { return Status(); }
}  // namespace errors
class OpKernelConstruction {
 public:
  template <typename T>
  Status GetAttr(const char *, T *) { return Status(); }
};
}  // namespace tensorflow
#define OP_REQUIRES_OK(ctx, expr) do { (void)(ctx); (void)(expr); } while (0)
#define OP_REQUIRES(ctx, cond, status) do { (void)(ctx); if (!(cond)) { (void)(status); } } while (0)
#endif
