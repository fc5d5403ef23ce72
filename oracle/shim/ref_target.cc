// TEST INFRASTRUCTURE ONLY -- compiles the reference's operator/multibox_target.cc in place (see mxnet_shim.h).
#define MXNET_OPERATOR_CONTRIB_MULTIBOX_TARGET_INL_H_  // skip the real -inl.h (needs MXNet)
#define SHIM_PARAM MultiBoxTargetParam
#define SHIM_OP MultiBoxTargetOp
#define SHIM_PROP MultiBoxTargetProp
#include "mxnet_shim.h"
#include REF_SOURCE(multibox_target.cc)

#include <string>

namespace {
// safe_divide and the 11-plane IoU expression of operator/multibox_target-inl.h:44-50,137-161, restated as the
// scalar loop those mshadow expression templates evaluate (every plane is a stored fp32 tensor).
inline float iou_plane(const float *a, const float *g) {
  float l1 = a[0], t1 = a[1], r1 = a[2], b1 = a[3], l2 = g[0], t2 = g[1], r2 = g[2], b2 = g[3];
  float mr = r1 < r2 ? r1 : r2, ml = l1 > l2 ? l1 : l2, mb = b1 < b2 ? b1 : b2, mt = t1 > t2 ? t1 : t2;
  float dw = mr - ml, dh = mb - mt;
  float iw = 0.0f > dw ? 0.0f : dw, ih = 0.0f > dh ? 0.0f : dh;
  float inter = iw * ih;
  float a1 = (r1 - l1) * (b1 - t1), a2 = (r2 - l2) * (b2 - t2);
  float uni = a1 + a2;
  uni = uni - inter;
  if (uni == 0.0f) return 0.0f;
  return inter / uni;
}
}  // namespace

// Glue standing in for MultiBoxTargetOp::Forward (operator/multibox_target-inl.h:89-171).
// Returns 0, or -2 / -3 / -4 / -1 when one of the reference's CHECKs fires (same codes as include/dspmb.h).
extern "C" int ref_multibox_target(const float *anchors, const float *labels, const float *cls_preds,
                                   float *loc_target, float *loc_mask, float *cls_target, int B, int A, int L,
                                   int label_width, int C, float overlap_threshold, float ignore_label,
                                   float negative_mining_ratio, float negative_mining_thresh,
                                   int minimum_negative_samples, const float *variances) {
  using namespace mshadow;
  std::vector<float> temp((size_t)B * A * L);
  for (int b = 0; b < B; ++b)
    for (int j = 0; j < A; ++j)
      for (int k = 0; k < L; ++k)
        temp[((size_t)b * A + j) * L + k] = iou_plane(anchors + 4 * j, labels + ((size_t)b * L + k) * label_width + 1);
  for (size_t i = 0; i < (size_t)B * A * 5; ++i) loc_target[i] = 0.f, loc_mask[i] = 0.0f;
  for (size_t i = 0; i < (size_t)B * A; ++i) cls_target[i] = ignore_label;
  Tensor<cpu, 2, float> t_loc(loc_target, {(index_t)B, (index_t)(A * 5)}), t_mask(loc_mask, {(index_t)B, (index_t)(A * 5)});
  Tensor<cpu, 2, float> t_cls(cls_target, {(index_t)B, (index_t)A});
  Tensor<cpu, 2, float> t_anchor(const_cast<float *>(anchors), {(index_t)A, 4u});
  Tensor<cpu, 3, float> t_label(const_cast<float *>(labels), {(index_t)B, (index_t)L, (index_t)label_width});
  Tensor<cpu, 3, float> t_pred(const_cast<float *>(cls_preds), {(index_t)B, (index_t)C, (index_t)A});
  Tensor<cpu, 4, float> t_temp(temp.data(), {1u, (index_t)B, (index_t)A, (index_t)L});  // plane 0 is all the CPU code reads
  nnvm::Tuple<float> var{variances[0], variances[1], variances[2], variances[3]};
  try {
    MultiBoxTargetForward(t_loc, t_mask, t_cls, t_anchor, t_label, t_pred, t_temp, overlap_threshold, ignore_label,
                          negative_mining_ratio, negative_mining_thresh, minimum_negative_samples, var);
  } catch (const shim::Error &e) {
    const std::string w = e.what();
    if (w.find("temp.size()") != std::string::npos) return -3;
    if (w.find("negative_mining_thresh") != std::string::npos) return -4;
    if (w.find("p_label") != std::string::npos) return -2;
    return -1;
  }
  return 0;
}
