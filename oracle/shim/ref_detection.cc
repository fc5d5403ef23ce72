// TEST INFRASTRUCTURE ONLY -- compiles the reference's operator/multibox_detection.cc in place (see mxnet_shim.h).
#define MXNET_OPERATOR_CONTRIB_MULTIBOX_DETECTION_INL_H_  // skip the real -inl.h (needs MXNet)
#define SHIM_PARAM MultiBoxDetectionParam
#define SHIM_OP MultiBoxDetectionOp
#define SHIM_PROP MultiBoxDetectionProp
#include "mxnet_shim.h"
#include REF_SOURCE(multibox_detection.cc)

// Glue standing in for MultiBoxDetectionOp::Forward (operator/multibox_detection-inl.h:81-107): `out = -1`, a temp
// space of the output's shape, and the call into the reference's MultiBoxDetectionForward.
extern "C" int ref_multibox_detection(const float *cls_prob, const float *loc_pred, const float *anchors, float *out,
                                      int B, int A, int C, float threshold, int clip, const float *variances,
                                      float nms_threshold, int force_suppress, int nms_topk) {
  using namespace mshadow;
  std::vector<float> temp((size_t)B * A * 7);
  for (size_t i = 0; i < (size_t)B * A * 7; ++i) out[i] = -1.f;
  Tensor<cpu, 3, float> t_out(out, {(index_t)B, (index_t)A, 7u}), t_temp(temp.data(), {(index_t)B, (index_t)A, 7u});
  Tensor<cpu, 3, float> t_prob(const_cast<float *>(cls_prob), {(index_t)B, (index_t)C, (index_t)A});
  Tensor<cpu, 2, float> t_loc(const_cast<float *>(loc_pred), {(index_t)B, (index_t)(A * 5)});
  Tensor<cpu, 2, float> t_anchor(const_cast<float *>(anchors), {(index_t)A, 4u});
  nnvm::Tuple<float> var{variances[0], variances[1], variances[2], variances[3]};
  try {
    MultiBoxDetectionForward(t_out, t_prob, t_loc, t_anchor, t_temp, threshold, clip != 0, var, nms_threshold,
                             force_suppress != 0, nms_topk);
  } catch (const shim::Error &) {
    return -1;
  }
  return 0;
}
