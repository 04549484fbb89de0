// SYNTAX-CHECK STAND-IN for tensorflow/core/framework/op.h (TF 1.15 API subset)
#ifndef SAP3D_TF_STUB_OP_H_
#define SAP3D_TF_STUB_OP_H_
#include "tensorflow/core/framework/shape_inference.h"

namespace tensorflow {
namespace register_op {
class OpDefBuilderWrapper {
 public:
  explicit OpDefBuilderWrapper(const char* name);
  OpDefBuilderWrapper& Attr(const char* spec);
  OpDefBuilderWrapper& Input(const char* spec);
  OpDefBuilderWrapper& Output(const char* spec);
  OpDefBuilderWrapper& SetIsStateful();
  OpDefBuilderWrapper& Doc(const char* text);
  OpDefBuilderWrapper& SetShapeFn(Status (*fn)(shape_inference::InferenceContext*));
};
struct OpDefBuilderReceiver {
  OpDefBuilderReceiver(const OpDefBuilderWrapper& wrapper);  // NOLINT
};
}  // namespace register_op
}  // namespace tensorflow

#define REGISTER_OP(name)                                                                       \
  static ::tensorflow::register_op::OpDefBuilderReceiver SAP3D_TF_STUB_CAT(register_op, __COUNTER__) = \
      ::tensorflow::register_op::OpDefBuilderWrapper(name)
#endif  // SAP3D_TF_STUB_OP_H_
