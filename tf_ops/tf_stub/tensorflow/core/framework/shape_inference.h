// SYNTAX-CHECK STAND-IN for tensorflow/core/framework/shape_inference.h (TF 1.15 API subset)
#ifndef SAP3D_TF_STUB_SHAPE_INFERENCE_H_
#define SAP3D_TF_STUB_SHAPE_INFERENCE_H_
#include "tensorflow/core/framework/op_kernel.h"

namespace tensorflow {
namespace shape_inference {
class DimensionHandle {
 public:
  DimensionHandle();
};
class ShapeHandle {
 public:
  ShapeHandle();
};
class DimensionOrConstant {
 public:
  DimensionOrConstant(DimensionHandle dim);  // NOLINT
  DimensionOrConstant(int64 val);            // NOLINT
};
class InferenceContext {
 public:
  static constexpr int64 kUnknownDim = -1;
  ShapeHandle input(int64 idx) const;
  int num_inputs() const;
  void set_output(int idx, ShapeHandle shape);
  DimensionHandle Dim(ShapeHandle s, int64 idx);
  static bool ValueKnown(DimensionOrConstant d);
  static int64 Value(DimensionOrConstant d);
  Status WithRank(ShapeHandle shape, int64 rank, ShapeHandle* out);
  ShapeHandle MakeShape(const std::vector<DimensionHandle>& dims);
  ShapeHandle MakeShape(std::initializer_list<DimensionOrConstant> dims);
  ShapeHandle UnknownShape();
  ShapeHandle UnknownShapeOfRank(int64 rank);
  ShapeHandle Scalar();
  ShapeHandle Vector(DimensionOrConstant dim);
  ShapeHandle Matrix(DimensionOrConstant dim1, DimensionOrConstant dim2);
  DimensionHandle MakeDim(DimensionOrConstant d);
  DimensionHandle UnknownDim();
  template <class T> Status GetAttr(const char* attr_name, T* value) const;
};
}  // namespace shape_inference
}  // namespace tensorflow
#endif  // SAP3D_TF_STUB_SHAPE_INFERENCE_H_
