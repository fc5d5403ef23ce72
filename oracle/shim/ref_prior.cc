// TEST INFRASTRUCTURE ONLY -- compiles the reference's operator/multibox_prior.cc in place (see mxnet_shim.h).
#define MXNET_OPERATOR_CONTRIB_MULTIBOX_PRIOR_INL_H_  // skip the real -inl.h (needs MXNet)
#define SHIM_PARAM MultiBoxPriorParam
#define SHIM_OP MultiBoxPriorOp
#define SHIM_PROP MultiBoxPriorProp
#include "mxnet_shim.h"
#include REF_SOURCE(multibox_prior.cc)

// Glue standing in for MultiBoxPriorOp::Forward (operator/multibox_prior-inl.h:97-129): auto step (:119-123),
// the call into the reference's MultiBoxPriorForward, and clip_zero_one (:44-51,126-128).
extern "C" int ref_multibox_prior(float *out, int in_height, int in_width, const float *sizes, int num_sizes,
                                  const float *ratios, int num_ratios, float step_y, float step_x, float off_y,
                                  float off_x, int clip) {
  try {
    std::vector<float> sizes_(sizes, sizes + num_sizes), ratios_(ratios, ratios + num_ratios);
    std::vector<float> steps_{step_y, step_x}, offsets_{off_y, off_x};
    CHECK_GT(sizes_.size(), 0);
    CHECK_GT(ratios_.size(), 0);
    CHECK_GE(offsets_[0], 0.f);
    CHECK_LE(offsets_[0], 1.f);
    CHECK_GE(offsets_[1], 0.f);
    CHECK_LE(offsets_[1], 1.f);
    const int num_anchors = num_sizes - 1 + num_ratios;
    mshadow::Tensor<mshadow::cpu, 2, float> o(out, {(mshadow::index_t)(num_anchors * in_width * in_height), 4u});
    CHECK_GE(steps_[0] * steps_[1], 0) << "Must specify both step_y and step_x";
    if (steps_[0] <= 0 || steps_[1] <= 0) {
      steps_[0] = 1.f / in_height;
      steps_[1] = 1.f / in_width;
    }
    mshadow::MultiBoxPriorForward(o, sizes_, ratios_, in_width, in_height, steps_, offsets_);
    if (clip) {
      const size_t n = (size_t)num_anchors * in_width * in_height * 4;
      for (size_t i = 0; i < n; ++i) {
        float a = out[i];
        out[i] = a < 0.f ? 0.f : (a > 1.f ? 1.f : a);
      }
    }
    return 0;
  } catch (const shim::Error &) {
    return -1;
  }
}
