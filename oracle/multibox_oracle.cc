// TEST INFRASTRUCTURE ONLY -- not part of the product path.
//
// CPU oracle: a from-scratch restatement of the reference's *CPU* multibox operators and of the
// Cython cpu_nms, used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs as the checker.  Nothing under dspnet_b200/ may import, link or call this.
//
// Every function cites the reference file:line it follows (paths relative to /root/reference).
// Arithmetic notes:
//   * fp32 throughout, evaluated in the written order, built with -O2 -ffp-contract=off
//     -fno-fast-math so that no a*b+c is contracted into an FMA (x86-64 SSE, FLT_EVAL_METHOD=0).
//   * std::exp/std::log on float resolve to the platform libm expf/logf (glibc 2.39 here).
//   * the unqualified exp(pw * vw) of multibox_detection.cc:117-118 is taken to be expf (a CUDA
//     enabled MXNet build pulls <math.h> in, which makes the float overload visible).
//
// Parity pin: the reference ships no tests or golden vectors for this path.
// This restatement is pinned against oracle/_ref/libmultibox_ref.so, i.e. the reference's own
// operator/*.cc function bodies compiled in place from /root/reference behind a header shim
// (oracle/shim/, oracle/build_ref.py), and the resulting vectors are committed under tests/golden/.

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <utility>
#include <vector>

namespace {

// Error codes shared with include/dspmb.h (a CHECK failure in the reference aborts with
// dmlc::Error; here and in the CUDA library it maps to a negative return code).
enum {
  kOk = 0,
  kErrBadArg = -1,
  kErrLabelPadding = -2,      // multibox_target.cc:98-101 CHECK_EQ on the first padding row
  kErrMiningCandidates = -3,  // multibox_target.cc:236 CHECK_GE(temp.size(), num_negative)
  kErrMiningThresh = -4,      // multibox_target.cc:184 CHECK_GT(negative_mining_thresh, 0)
};

// ---------------------------------------------------------------------------------------------
// MultiBoxPrior -- operator/multibox_prior.cc:29-71, multibox_prior-inl.h:111-128 (auto step,
// clip) and :44-51 (clip_zero_one).
// ---------------------------------------------------------------------------------------------
inline float clip01(float a) {
  if (a < 0.f) return 0.f;
  if (a > 1.f) return 1.f;
  return a;
}

int prior_impl(float *out, int in_height, int in_width, const float *sizes, int num_sizes,
               const float *ratios, int num_ratios, float step_y, float step_x, float off_y,
               float off_x, int clip) {
  if (num_sizes <= 0 || num_ratios <= 0 || in_height <= 0 || in_width <= 0) return kErrBadArg;
  if (off_y < 0.f || off_y > 1.f || off_x < 0.f || off_x > 1.f) return kErrBadArg;  // -inl.h:90-93
  if (step_y * step_x < 0) return kErrBadArg;                                        // -inl.h:118
  if (step_y <= 0 || step_x <= 0) {  // -inl.h:119-123
    step_y = 1.f / in_height;
    step_x = 1.f / in_width;
  }
  int count = 0;
  for (int r = 0; r < in_height; ++r) {
    float center_y = (r + off_y) * step_y;
    for (int c = 0; c < in_width; ++c) {
      float center_x = (c + off_x) * step_x;
      for (int i = 0; i < num_sizes; ++i) {  // ratio = 1, every size (.cc:46-55)
        float size = sizes[i];
        float w = size * in_height / in_width / 2;
        float h = size / 2;
        float *o = out + 4 * count++;
        o[0] = center_x - w;
        o[1] = center_y - h;
        o[2] = center_x + w;
        o[3] = center_y + h;
      }
      float size = sizes[0];
      for (int j = 1; j < num_ratios; ++j) {  // ratios[1:], size = sizes[0] (.cc:57-67)
        float ratio = sqrtf(ratios[j]);
        float w = size * in_height / in_width * ratio / 2;
        float h = size / ratio / 2;
        float *o = out + 4 * count++;
        o[0] = center_x - w;
        o[1] = center_y - h;
        o[2] = center_x + w;
        o[3] = center_y + h;
      }
    }
  }
  if (clip) {
    for (int i = 0; i < 4 * count; ++i) out[i] = clip01(out[i]);
  }
  return kOk;
}

// ---------------------------------------------------------------------------------------------
// MultiBoxTarget
// ---------------------------------------------------------------------------------------------

// IoU plane temp_space[0] -- multibox_target-inl.h:137-161 (+ safe_divide :44-50).  mshadow's
// maximum/minimum are `a > b ? a : b` / `a < b ? a : b`; every plane is a stored fp32 tensor, so
// each step below is individually rounded to fp32.
inline float iou_target(const float *a, const float *g) {
  float l1 = a[0], t1 = a[1], r1 = a[2], b1 = a[3];
  float l2 = g[0], t2 = g[1], r2 = g[2], b2 = g[3];
  float mr = r1 < r2 ? r1 : r2;
  float ml = l1 > l2 ? l1 : l2;
  float mb = b1 < b2 ? b1 : b2;
  float mt = t1 > t2 ? t1 : t2;
  float dw = mr - ml;
  float dh = mb - mt;
  float iw = 0.0f > dw ? 0.0f : dw;
  float ih = 0.0f > dh ? 0.0f : dh;
  float inter = iw * ih;
  float area1 = (r1 - l1) * (b1 - t1);
  float area2 = (r2 - l2) * (b2 - t2);
  float uni = area1 + area2;
  uni = uni - inter;
  if (uni == 0.0f) return 0.0f;
  return inter / uni;
}

// multibox_target.cc:30-56
inline void assign_loc_targets(const float *anchor, const float *l, float *dst, float vx, float vy,
                               float vw, float vh) {
  float al = anchor[0], at = anchor[1], ar = anchor[2], ab = anchor[3];
  float aw = ar - al;
  float ah = ab - at;
  float ax = (al + ar) * 0.5;  // double multiply, rounded back to float (exact)
  float ay = (at + ab) * 0.5;
  float gl = l[0], gt = l[1], gr = l[2], gb = l[3], gz = l[4];
  float gw = gr - gl;
  float gh = gb - gt;
  float gx = (gl + gr) * 0.5;
  float gy = (gt + gb) * 0.5;
  dst[0] = (gx - ax) / aw / vx;
  dst[1] = (gy - ay) / ah / vy;
  dst[2] = std::log(gw / aw) / vw;
  dst[3] = std::log(gh / ah) / vh;
  dst[4] = gz / 0.1;  // float / double -> double division, rounded once to float (.cc:55)
}

struct Descend {  // multibox_target.cc:58-70, multibox_detection.cc:30-42
  float value;
  int index;
  bool operator<(const Descend &o) const { return value > o.value; }
};

struct TargetArgs {
  const float *anchors, *labels, *cls_preds;
  float *loc_target, *loc_mask, *cls_target;
  int B, A, L, label_width, C;
  float overlap_threshold, ignore_label, negative_mining_ratio, negative_mining_thresh;
  int minimum_negative_samples;  // accepted and ignored, exactly as the CPU reference does
  const float *variances;
  // optional debug planes (may be null)
  int32_t *match_gt;     // (B, A) gt index of max_matches[j].second, -1 if never computed
  float *match_iou;      // (B, A) max_matches[j].first
  int8_t *anchor_flags;  // (B, A) 1 positive / 0 negative / -1 don't care
  int32_t *stats;        // (B, 4) num_valid_gt, num_positive, num_negative, bipartite matches
};

// One image of multibox_target.cc:92-281 (+ the output initialisation of -inl.h:121-124).
int target_one_image(const TargetArgs &t, int nbatch) {
  const int A = t.A, L = t.L, W = t.label_width, C = t.C;
  const float *p_anchor = t.anchors;
  const float *p_label = t.labels + (size_t)nbatch * L * W;
  float *p_loc_target = t.loc_target + (size_t)nbatch * A * 5;
  float *p_loc_mask = t.loc_mask + (size_t)nbatch * A * 5;
  float *p_cls_target = t.cls_target + (size_t)nbatch * A;
  // -inl.h:121-123
  std::fill(p_loc_target, p_loc_target + (size_t)A * 5, 0.f);
  std::fill(p_loc_mask, p_loc_mask + (size_t)A * 5, 0.f);
  std::fill(p_cls_target, p_cls_target + A, t.ignore_label);

  int num_valid_gt = 0;  // .cc:95-105
  for (int i = 0; i < L; ++i) {
    if (p_label[i * W] == -1.0f) {
      for (int c = 1; c <= 4; ++c)
        if (p_label[i * W + c] != -1.0f) return kErrLabelPadding;
      break;
    }
    ++num_valid_gt;
  }

  std::vector<std::pair<float, int>> max_matches(A, std::pair<float, int>(-1.0f, -1));
  std::vector<char> anchor_flags(A, -1);
  int num_positive = 0, num_negative_used = 0, num_bipartite = 0;

  if (num_valid_gt > 0) {
    const int G = num_valid_gt;
    // temp_space[0][nbatch] restricted to the columns the CPU code reads (k < num_valid_gt)
    std::vector<float> overlaps((size_t)A * G);
    for (int j = 0; j < A; ++j)
      for (int k = 0; k < G; ++k) overlaps[(size_t)j * G + k] = iou_target(p_anchor + 4 * j, p_label + k * W + 1);

    std::vector<bool> gt_flags(G, false);
    // bipartite stage, .cc:113-149
    while (std::find(gt_flags.begin(), gt_flags.end(), false) != gt_flags.end()) {
      int best_anchor = -1, best_gt = -1;
      float max_overlap = 1e-6;
      for (int j = 0; j < A; ++j) {
        if (anchor_flags[j] == 1) continue;
        const float *pp = &overlaps[(size_t)j * G];
        for (int k = 0; k < G; ++k) {
          if (gt_flags[k]) continue;
          float iou = pp[k];
          if (iou > max_overlap) {
            best_anchor = j;
            best_gt = k;
            max_overlap = iou;
          }
        }
      }
      if (best_anchor == -1) break;
      max_matches[best_anchor].first = max_overlap;
      max_matches[best_anchor].second = best_gt;
      num_positive += 1;
      num_bipartite += 1;
      gt_flags[best_gt] = true;
      anchor_flags[best_anchor] = 1;
    }

    auto row_argmax = [&](int j) {  // .cc:158-166 and :206-214
      const float *pp = &overlaps[(size_t)j * G];
      int best_gt = -1;
      float max_iou = -1.0f;
      for (int k = 0; k < G; ++k) {
        float iou = pp[k];
        if (iou > max_iou) {
          best_gt = k;
          max_iou = iou;
        }
      }
      if (best_gt != -1) {
        max_matches[j].first = max_iou;
        max_matches[j].second = best_gt;
      }
    };

    if (t.overlap_threshold > 0) {  // threshold stage, .cc:151-180
      for (int j = 0; j < A; ++j) {
        if (anchor_flags[j] == 1) continue;
        row_argmax(j);
        if (max_matches[j].second != -1 && max_matches[j].first > t.overlap_threshold) {
          num_positive += 1;
          anchor_flags[j] = 1;
        }
      }
    }

    if (t.negative_mining_ratio > 0) {  // hard-negative mining, .cc:182-241
      const float *p_cls = t.cls_preds + (size_t)nbatch * C * A;
      if (!(t.negative_mining_thresh > 0)) return kErrMiningThresh;
      int num_negative = num_positive * t.negative_mining_ratio;
      if (num_negative > (A - num_positive)) num_negative = A - num_positive;
      if (num_negative > 0) {
        std::vector<Descend> temp;
        temp.reserve(A - num_positive);
        for (int j = 0; j < A; ++j) {
          if (anchor_flags[j] == 1) continue;
          if (max_matches[j].first < 0) row_argmax(j);
          if (max_matches[j].first < t.negative_mining_thresh && anchor_flags[j] == -1) {
            float max_val = p_cls[j];
            for (int k = 1; k < C; ++k) {
              float tmp = p_cls[j + (size_t)A * k];
              if (tmp > max_val) max_val = tmp;
            }
            float sum = 0.f;
            for (int k = 0; k < C; ++k) {
              float tmp = p_cls[j + (size_t)A * k];
              sum += std::exp(tmp - max_val);
            }
            float prob = std::exp(p_cls[j] - max_val) / sum;
            temp.push_back(Descend{-prob, j});
          }
        }
        if ((int)temp.size() < num_negative) return kErrMiningCandidates;
        std::stable_sort(temp.begin(), temp.end());
        for (int i = 0; i < num_negative; ++i) anchor_flags[temp[i].index] = 0;
        num_negative_used = num_negative;
      }
    } else {  // .cc:242-249
      for (int i = 0; i < A; ++i)
        if (anchor_flags[i] != 1) {
          anchor_flags[i] = 0;
          ++num_negative_used;
        }
    }

    for (int i = 0; i < A; ++i) {  // .cc:251-281
      if (anchor_flags[i] == 1) {
        const float *lab = p_label + W * max_matches[i].second;
        p_cls_target[i] = lab[0] + 1;
        for (int c = 0; c < 5; ++c) p_loc_mask[i * 5 + c] = 1;
        assign_loc_targets(p_anchor + i * 4, lab + 1, p_loc_target + i * 5, t.variances[0],
                           t.variances[1], t.variances[2], t.variances[3]);
      } else if (anchor_flags[i] == 0) {
        p_cls_target[i] = 0;
        for (int c = 0; c < 5; ++c) p_loc_mask[i * 5 + c] = 0;
      }
    }
  }

  if (t.match_gt)
    for (int j = 0; j < A; ++j) t.match_gt[(size_t)nbatch * A + j] = max_matches[j].second;
  if (t.match_iou)
    for (int j = 0; j < A; ++j) t.match_iou[(size_t)nbatch * A + j] = max_matches[j].first;
  if (t.anchor_flags)
    for (int j = 0; j < A; ++j) t.anchor_flags[(size_t)nbatch * A + j] = anchor_flags[j];
  if (t.stats) {
    int32_t *s = t.stats + 4 * nbatch;
    s[0] = num_valid_gt;
    s[1] = num_positive;
    s[2] = num_negative_used;
    s[3] = num_bipartite;
  }
  return kOk;
}

// ---------------------------------------------------------------------------------------------
// MultiBoxDetection
// ---------------------------------------------------------------------------------------------

// multibox_detection.cc:44-51
inline float overlap_det(const float *a, const float *b) {
  float w = std::max(0.f, std::min(a[2], b[2]) - std::max(a[0], b[0]));
  float h = std::max(0.f, std::min(a[3], b[3]) - std::max(a[1], b[1]));
  float i = w * h;
  float u = (a[2] - a[0]) * (a[3] - a[1]) + (b[2] - b[0]) * (b[3] - b[1]) - i;
  return u <= 0.f ? 0.f : i / u;
}

struct DetArgs {
  const float *cls_prob, *loc_pred, *anchors;
  float *out;
  int B, A, C;
  float threshold;
  int clip;
  const float *variances;
  float nms_threshold;
  int force_suppress, nms_topk;
  int32_t *valid_count;  // optional (B)
};

// One image of multibox_detection.cc:74-167 (+ `out = -1` of -inl.h:103).
void detection_one_image(const DetArgs &d, int nbatch) {
  const int A = d.A, C = d.C;
  const float vx = d.variances[0], vy = d.variances[1], vw = d.variances[2], vh = d.variances[3];
  const float *p_cls_prob = d.cls_prob + (size_t)nbatch * C * A;
  const float *p_loc_pred = d.loc_pred + (size_t)nbatch * A * 5;
  const float *p_anchor = d.anchors;
  float *p_out = d.out + (size_t)nbatch * A * 7;
  std::fill(p_out, p_out + (size_t)A * 7, -1.f);
  const bool clip = d.clip != 0;
  int valid_count = 0;
  for (int i = 0; i < A; ++i) {  // pass 1, .cc:79-128
    float score = -1;
    int id = 0;
    for (int j = 1; j < C; ++j) {
      float temp = p_cls_prob[(size_t)j * A + i];
      if (temp > score) {
        score = temp;
        id = j;
      }
    }
    if (id > 0 && score < d.threshold) id = 0;
    if (id > 0) {
      float *row = p_out + (size_t)valid_count * 7;
      row[0] = id - 1;
      row[1] = score;
      float al = p_anchor[i * 4], at = p_anchor[i * 4 + 1], ar = p_anchor[i * 4 + 2], ab = p_anchor[i * 4 + 3];
      float aw = ar - al;
      float ah = ab - at;
      float ax = (al + ar) / 2.f;
      float ay = (at + ab) / 2.f;
      const float *lp = p_loc_pred + (size_t)i * 5;
      float px = lp[0], py = lp[1], pw = lp[2], ph = lp[3], pz = lp[4];
      float ox = px * vx * aw + ax;
      float oy = py * vy * ah + ay;
      float ow = expf(pw * vw) * aw / 2;
      float oh = expf(ph * vh) * ah / 2;
      float oz = pz * 0.1;  // float * double, rounded once to float (.cc:119)
      row[2] = clip ? std::max(0.f, std::min(1.f, ox - ow)) : (ox - ow);
      row[3] = clip ? std::max(0.f, std::min(1.f, oy - oh)) : (oy - oh);
      row[4] = clip ? std::max(0.f, std::min(1.f, ox + ow)) : (ox + ow);
      row[5] = clip ? std::max(0.f, std::min(1.f, oy + oh)) : (oy + oh);
      row[6] = clip ? std::max(0.f, std::min(1.f, oz)) : (oz);
      ++valid_count;
    }
  }
  if (d.valid_count) d.valid_count[nbatch] = valid_count;
  if (valid_count < 1 || d.nms_threshold <= 0 || d.nms_threshold > 1) return;  // .cc:130

  // sort + top-k, .cc:132-151: only rows [0, nkeep) are rewritten in sorted order.
  std::vector<float> temp(p_out, p_out + (size_t)valid_count * 7);
  std::vector<Descend> sorter;
  sorter.reserve(valid_count);
  for (int i = 0; i < valid_count; ++i) sorter.push_back(Descend{p_out[i * 7 + 1], i});
  std::stable_sort(sorter.begin(), sorter.end());
  int nkeep = valid_count;
  if (d.nms_topk > 0 && d.nms_topk < nkeep) nkeep = d.nms_topk;
  for (int i = 0; i < nkeep; ++i)
    for (int j = 0; j < 7; ++j) p_out[i * 7 + j] = temp[(size_t)sorter[i].index * 7 + j];

  // greedy NMS over all valid_count rows, .cc:153-167
  for (int i = 0; i < valid_count; ++i) {
    float *ri = p_out + (size_t)i * 7;
    if (ri[0] < 0) continue;
    for (int j = i + 1; j < valid_count; ++j) {
      float *rj = p_out + (size_t)j * 7;
      if (rj[0] < 0) continue;
      if (d.force_suppress || ri[0] == rj[0]) {
        float iou = overlap_det(ri + 2, rj + 2);
        if (iou >= d.nms_threshold) rj[0] = -1;
      }
    }
  }
}

template <typename F>
void parallel_images(int B, int nthreads, F &&fn) {
  if (nthreads <= 1 || B <= 1) {
    for (int b = 0; b < B; ++b) fn(b);
    return;
  }
  nthreads = std::min(nthreads, B);
  std::vector<std::thread> pool;
  for (int t = 0; t < nthreads; ++t)
    pool.emplace_back([&, t] {
      for (int b = t; b < B; b += nthreads) fn(b);
    });
  for (auto &th : pool) th.join();
}

}  // namespace

extern "C" {

int oracle_multibox_prior(float *out, int in_height, int in_width, const float *sizes, int num_sizes,
                          const float *ratios, int num_ratios, float step_y, float step_x,
                          float off_y, float off_x, int clip) {
  return prior_impl(out, in_height, in_width, sizes, num_sizes, ratios, num_ratios, step_y, step_x,
                    off_y, off_x, clip);
}

// IoU plane exposed for stage-wise tests: out[(j * G) + k], anchors (A,4), gts (G,4).
void oracle_target_iou(const float *anchors, int A, const float *gts, int G, int gt_stride, float *out) {
  for (int j = 0; j < A; ++j)
    for (int k = 0; k < G; ++k) out[(size_t)j * G + k] = iou_target(anchors + 4 * j, gts + (size_t)k * gt_stride);
}

// nthreads > 1 runs one image per thread (the reference loop is single-threaded; this is the
// most the CPU path can do without changing the algorithm).  Returns the first non-zero status.
int oracle_multibox_target(const float *anchors, const float *labels, const float *cls_preds,
                           float *loc_target, float *loc_mask, float *cls_target, int B, int A,
                           int L, int label_width, int C, float overlap_threshold,
                           float ignore_label, float negative_mining_ratio,
                           float negative_mining_thresh, int minimum_negative_samples,
                           const float *variances, int32_t *match_gt, float *match_iou,
                           int8_t *anchor_flags, int32_t *stats, int nthreads) {
  if (B < 0 || A <= 0 || L <= 0 || label_width < 6 || C <= 0) return kErrBadArg;
  TargetArgs t{anchors, labels, cls_preds, loc_target, loc_mask, cls_target, B, A, L, label_width, C,
               overlap_threshold, ignore_label, negative_mining_ratio, negative_mining_thresh,
               minimum_negative_samples, variances, match_gt, match_iou, anchor_flags, stats};
  std::vector<int> rc(B, 0);
  parallel_images(B, nthreads, [&](int b) { rc[b] = target_one_image(t, b); });
  for (int b = 0; b < B; ++b)
    if (rc[b]) return rc[b];
  return kOk;
}

int oracle_multibox_detection(const float *cls_prob, const float *loc_pred, const float *anchors,
                              float *out, int B, int A, int C, float threshold, int clip,
                              const float *variances, float nms_threshold, int force_suppress,
                              int nms_topk, int32_t *valid_count, int nthreads) {
  if (B < 0 || A <= 0 || C <= 0) return kErrBadArg;
  DetArgs d{cls_prob, loc_pred, anchors, out, B, A, C, threshold, clip, variances, nms_threshold,
            force_suppress, nms_topk, valid_count};
  parallel_images(B, nthreads, [&](int b) { detection_one_image(d, b); });
  return kOk;
}

// cython/cpu_nms.pyx:17-68.  `order` is scores.argsort()[::-1] computed by the caller with numpy
// (numpy's default argsort is unstable, so parity is only defined for tie-free scores).
// mode 0: suppress iff (double)ovr >= thresh (cpu_nms.pyx:64, thresh is a C double)
// mode 1: suppress iff ovr > (float)thresh   (nms_kernel.cu:71 / detect/nms.py:55)
int oracle_cpu_nms(const float *dets, int ndets, int dim, const int64_t *order, double thresh, int mode,
                   int64_t *keep) {
  std::vector<float> areas(ndets);
  for (int i = 0; i < ndets; ++i) {
    const float *d = dets + (size_t)i * dim;
    areas[i] = (d[2] - d[0] + 1) * (d[3] - d[1] + 1);
  }
  std::vector<char> suppressed(ndets, 0);
  const float thresh_f = (float)thresh;
  int nkeep = 0;
  for (int _i = 0; _i < ndets; ++_i) {
    int64_t i = order[_i];
    if (suppressed[i]) continue;
    keep[nkeep++] = i;
    const float *di = dets + (size_t)i * dim;
    float ix1 = di[0], iy1 = di[1], ix2 = di[2], iy2 = di[3], iarea = areas[i];
    for (int _j = _i + 1; _j < ndets; ++_j) {
      int64_t j = order[_j];
      if (suppressed[j]) continue;
      const float *dj = dets + (size_t)j * dim;
      float xx1 = ix1 >= dj[0] ? ix1 : dj[0];
      float yy1 = iy1 >= dj[1] ? iy1 : dj[1];
      float xx2 = ix2 <= dj[2] ? ix2 : dj[2];
      float yy2 = iy2 <= dj[3] ? iy2 : dj[3];
      float tw = xx2 - xx1 + 1;
      float th = yy2 - yy1 + 1;
      float w = 0.0f >= tw ? 0.0f : tw;
      float h = 0.0f >= th ? 0.0f : th;
      float inter = w * h;
      float ovr = inter / (iarea + areas[j] - inter);
      bool sup = mode == 0 ? ((double)ovr >= thresh) : (ovr > thresh_f);
      if (sup) suppressed[j] = 1;
    }
  }
  return nkeep;
}

// Channel softmax of the class head, SURVEY.md section 8f row f1: `SoftmaxActivation(mode='channel')`
// (symbol/symbol_builder.py:161-162) and the forward of `SoftmaxOutput(multi_output=True)` (:82-84) on cls_preds
// (B, C, A).  The arithmetic lives in MXNet, which is not part of the reference tree, so this restates the softmax the
// reference itself spells out for the same tensor in multibox_target.cc:220-231 -- running maximum over the classes,
// fp32 sum of expf(x - max) in class order, one division per value -- and declares it the pin ("parity unpinned"
// against MXNet's own kernel, which evaluates the same three steps per position).
void oracle_softmax_channel(const float *x, float *out, int B, int C, int A) {
  for (int b = 0; b < B; ++b) {
    const float *xb = x + (size_t)b * C * A;
    float *ob = out + (size_t)b * C * A;
    for (int a = 0; a < A; ++a) {
      float mx = xb[a];
      for (int c = 1; c < C; ++c) {
        float t = xb[(size_t)c * A + a];
        if (t > mx) mx = t;
      }
      float sum = 0.f;
      for (int c = 0; c < C; ++c) sum += expf(xb[(size_t)c * A + a] - mx);
      for (int c = 0; c < C; ++c) ob[(size_t)c * A + a] = expf(xb[(size_t)c * A + a] - mx) / sum;
    }
  }
}

// libm probes used by tests to check the CUDA library's glibc-compatible expf/logf.
float oracle_expf(float x) { return expf(x); }
float oracle_logf(float x) { return logf(x); }
void oracle_expf_array(const float *x, float *y, long n) {
  for (long i = 0; i < n; ++i) y[i] = expf(x[i]);
}
void oracle_logf_array(const float *x, float *y, long n) {
  for (long i = 0; i < n; ++i) y[i] = logf(x[i]);
}

// cython/bbox.pyx:15-55 (bbox_overlaps_cython): float64, "+1" pixel convention, zero unless both extents > 0.
void oracle_bbox_overlaps(const double *boxes, int N, const double *query, int K, double *overlaps) {
  for (size_t i = 0; i < (size_t)N * K; ++i) overlaps[i] = 0.0;
  for (int k = 0; k < K; ++k) {
    const double *q = query + (size_t)k * 4;
    const double box_area = (q[2] - q[0] + 1) * (q[3] - q[1] + 1);
    for (int n = 0; n < N; ++n) {
      const double *b = boxes + (size_t)n * 4;
      const double iw = (b[2] < q[2] ? b[2] : q[2]) - (b[0] > q[0] ? b[0] : q[0]) + 1;
      if (iw > 0) {
        const double ih = (b[3] < q[3] ? b[3] : q[3]) - (b[1] > q[1] ? b[1] : q[1]) + 1;
        if (ih > 0) {
          const double ua = (b[2] - b[0] + 1) * (b[3] - b[1] + 1) + box_area - iw * ih;
          overlaps[(size_t)n * K + k] = iw * ih / ua;
        }
      }
    }
  }
}

}  // extern "C"
