// SYNTAX-CHECK STAND-IN for tensorflow/core/framework/op_kernel.h (TF 1.15 API subset) -- see ../../../../README.md
#ifndef SAP3D_TF_STUB_OP_KERNEL_H_
#define SAP3D_TF_STUB_OP_KERNEL_H_
#include <stddef.h>
#include <stdint.h>

#include <initializer_list>
#include <string>
#include <vector>

typedef struct CUstream_st* cudaStream_t;
namespace Eigen {
struct GpuDevice {
  const cudaStream_t& stream() const;
};
struct bfloat16_stub {
  uint16_t value;
};
}  // namespace Eigen

namespace tensorflow {
typedef int32_t int32;
typedef long long int64;
typedef unsigned long long uint64;
typedef uint8_t uint8;
typedef Eigen::bfloat16_stub bfloat16;
using std::string;

enum DataType { DT_INVALID = 0, DT_FLOAT = 1, DT_DOUBLE = 2, DT_INT32 = 3, DT_UINT8 = 4, DT_INT64 = 9, DT_BFLOAT16 = 14 };
extern const char* const DEVICE_GPU;
extern const char* const DEVICE_CPU;

class Status {
 public:
  Status();
  static Status OK();
  bool ok() const;
  const string& error_message() const;
};
namespace errors {
template <typename... Args> Status InvalidArgument(Args... args);
template <typename... Args> Status Internal(Args... args);
template <typename... Args> Status Unimplemented(Args... args);
}  // namespace errors

class StringPiece {
 public:
  const char* data() const;
  size_t size() const;
};

class TensorShape {
 public:
  TensorShape();
  TensorShape(std::initializer_list<int64> dims);
  int dims() const;
  int64 dim_size(int d) const;
  int64 num_elements() const;
  void AddDim(int64 size);
};

template <typename T> struct TTypesFlat {
  T* data() const;
  T& operator()(int64 i) const;
  int64 size() const;
};

class Tensor {
 public:
  Tensor();
  DataType dtype() const;
  const TensorShape& shape() const;
  int dims() const;
  int64 dim_size(int d) const;
  int64 NumElements() const;
  StringPiece tensor_data() const;
  template <typename T> TTypesFlat<T> flat();
  template <typename T> TTypesFlat<const T> flat() const;
  template <typename T> T& scalar();
};

class OpKernelConstruction {
 public:
  template <typename T> Status GetAttr(const char* name, T* value) const;
  void CtxFailure(const Status& s);
  void CtxFailureWithWarning(const Status& s);
};

class OpKernelContext {
 public:
  int num_inputs() const;
  const Tensor& input(int index);
  Tensor* mutable_output(int index);
  Status allocate_output(int index, const TensorShape& shape, Tensor** tensor);
  Status allocate_temp(DataType type, const TensorShape& shape, Tensor* out_temp);
  void set_output(int index, const Tensor& tensor);
  template <typename EigenDeviceType> const EigenDeviceType& eigen_device() const;
  void CtxFailure(const Status& s);
  void CtxFailureWithWarning(const Status& s);
};

class OpKernel {
 public:
  explicit OpKernel(OpKernelConstruction* context);
  virtual ~OpKernel();
  virtual void Compute(OpKernelContext* context) = 0;
};

namespace register_kernel {
class Name {
 public:
  explicit Name(const char* op);
  Name& Device(const char* device_type);
  template <typename T> Name& TypeConstraint(const char* attr_name);
  Name& HostMemory(const char* arg_name);
};
struct OpKernelRegistrar {
  template <typename F> OpKernelRegistrar(const Name& n, const char* class_name, F factory);
};
}  // namespace register_kernel
}  // namespace tensorflow

#define SAP3D_TF_STUB_CAT2(a, b) a##b
#define SAP3D_TF_STUB_CAT(a, b) SAP3D_TF_STUB_CAT2(a, b)
#define REGISTER_KERNEL_BUILDER(kernel_builder, ...)                                                              \
  static ::tensorflow::register_kernel::OpKernelRegistrar SAP3D_TF_STUB_CAT(registrar__body__, __COUNTER__)(      \
      ::tensorflow::register_kernel::kernel_builder, #__VA_ARGS__,                                                \
      [](::tensorflow::OpKernelConstruction* context) -> ::tensorflow::OpKernel* { return new __VA_ARGS__(context); })

#define OP_REQUIRES(CTX, EXP, STATUS) \
  do {                                \
    if (!(EXP)) {                     \
      (CTX)->CtxFailure((STATUS));    \
      return;                         \
    }                                 \
  } while (0)
#define OP_REQUIRES_OK(CTX, ...)                     \
  do {                                               \
    ::tensorflow::Status _s(__VA_ARGS__);            \
    if (!_s.ok()) {                                  \
      (CTX)->CtxFailureWithWarning(_s);              \
      return;                                        \
    }                                                \
  } while (0)
#define TF_RETURN_IF_ERROR(...)                      \
  do {                                               \
    ::tensorflow::Status _status = (__VA_ARGS__);    \
    if (!_status.ok()) return _status;               \
  } while (0)
#endif  // SAP3D_TF_STUB_OP_KERNEL_H_
