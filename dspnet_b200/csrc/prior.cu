// MultiBoxPrior for sm_100a -- anchor generation.
//
// Reference semantics: operator/multibox_prior.cc:29-71 (CPU loop), multibox_prior-inl.h:111-128 (auto step,
// clip).  The reference GPU version (operator/multibox_prior.cu:86-99) launches one kernel per anchor kind with
// four strided scalar stores per thread; here one launch covers every kind of every feature map of a head, one
// thread per anchor, one 128-bit store per anchor, written straight into the concatenated (A, 4) tensor that
// symbol/common.py:424-432 builds with Flatten + Concat + Reshape.
//
// Bit-exactness: the per-kind half extents (size*H/W/2, size/2, size*H/W*sqrt(r)/2, size/sqrt(r)/2) depend only
// on the op parameters, so they are evaluated once on the host in the reference's own expression order
// (IEEE fp32, correctly rounded sqrtf); the per-cell centre (c + off) * step and the +-extent are evaluated on
// the device with explicitly rounded fp32 ops (no FMA contraction).
#include <math.h>

#include "common.cuh"

namespace dspmb {
namespace {

struct PriorMap {
  int height, width;
  float step_y, step_x, off_y, off_x;
  int kind_base, num_kinds;
  int anchor_base;  // first output row of this map
};

struct PriorParams {
  PriorMap maps[DSPMB_MAX_MAPS];
  float half_w[DSPMB_MAX_KINDS];
  float half_h[DSPMB_MAX_KINDS];
  int num_maps;
  int total_anchors;
  int clip;
};

__device__ __forceinline__ float clip01(float a) { return a < 0.f ? 0.f : (a > 1.f ? 1.f : a); }

__global__ void __launch_bounds__(256) prior_kernel(const __grid_constant__ PriorParams p, float4 *__restrict__ out) {
  for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < p.total_anchors; g += gridDim.x * blockDim.x) {
    int m = 0;
#pragma unroll 1
    while (m + 1 < p.num_maps && g >= p.maps[m + 1].anchor_base) ++m;
    const PriorMap &fm = p.maps[m];
    const int local = g - fm.anchor_base;
    const int cell = local / fm.num_kinds;
    const int kind = local - cell * fm.num_kinds;
    const int r = cell / fm.width;
    const int c = cell - r * fm.width;
    const float cy = fmul(fadd((float)r, fm.off_y), fm.step_y);
    const float cx = fmul(fadd((float)c, fm.off_x), fm.step_x);
    const float w = p.half_w[fm.kind_base + kind];
    const float h = p.half_h[fm.kind_base + kind];
    float4 box = make_float4(fsub(cx, w), fsub(cy, h), fadd(cx, w), fadd(cy, h));
    if (p.clip) box = make_float4(clip01(box.x), clip01(box.y), clip01(box.z), clip01(box.w));
    out[g] = box;
  }
}

int fill_map(PriorParams &p, int m, int height, int width, const float *sizes, int num_sizes, const float *ratios,
             int num_ratios, float step_y, float step_x, float off_y, float off_x, int &kinds, int &anchors) {
  // Parameter CHECKs of MultiBoxPriorOp's ctor and Forward (multibox_prior-inl.h:82-95,118) and of
  // InferShape (:178-181).
  DSPMB_REQUIRE(height > 0 && width > 0, "MultiBoxPrior: input height/width must be > 0");
  DSPMB_REQUIRE(num_sizes > 0 && num_ratios > 0, "MultiBoxPrior: need at least one size and one ratio");
  DSPMB_REQUIRE(off_y >= 0.f && off_y <= 1.f && off_x >= 0.f && off_x <= 1.f, "MultiBoxPrior: offsets must be in [0,1]");
  DSPMB_REQUIRE(step_y * step_x >= 0, "MultiBoxPrior: must specify both step_y and step_x");
  const int nk = num_sizes + num_ratios - 1;
  DSPMB_REQUIRE(kinds + nk <= DSPMB_MAX_KINDS, "MultiBoxPrior: more than %d sizes+ratios in one launch", DSPMB_MAX_KINDS);
  DSPMB_REQUIRE((long long)anchors + (long long)height * width * nk < (1ll << 31), "MultiBoxPrior: too many anchors");
  if (step_y <= 0 || step_x <= 0) {  // -inl.h:119-123
    step_y = 1.f / height;
    step_x = 1.f / width;
  }
  PriorMap &fm = p.maps[m];
  fm.height = height;
  fm.width = width;
  fm.step_y = step_y;
  fm.step_x = step_x;
  fm.off_y = off_y;
  fm.off_x = off_x;
  fm.kind_base = kinds;
  fm.num_kinds = nk;
  fm.anchor_base = anchors;
  for (int i = 0; i < num_sizes; ++i) {  // multibox_prior.cc:46-49
    const float size = sizes[i];
    p.half_w[kinds + i] = size * height / width / 2;
    p.half_h[kinds + i] = size / 2;
  }
  const float size = sizes[0];
  for (int j = 1; j < num_ratios; ++j) {  // multibox_prior.cc:58-62
    const float ratio = sqrtf(ratios[j]);
    p.half_w[kinds + num_sizes + j - 1] = size * height / width * ratio / 2;
    p.half_h[kinds + num_sizes + j - 1] = size / ratio / 2;
  }
  kinds += nk;
  anchors += height * width * nk;
  return DSPMB_OK;
}

int launch(const PriorParams &p, float *out, cudaStream_t stream) {
  DSPMB_REQUIRE(out != nullptr, "MultiBoxPrior: out is NULL");
  DSPMB_REQUIRE(((uintptr_t)out & 15) == 0, "MultiBoxPrior: out must be 16-byte aligned");
  const int threads = 256;
  int blocks = ceil_div(p.total_anchors, threads);
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  {
    ProfileScope _p(kSlotPrior, stream);
    prior_kernel<<<blocks, threads, 0, stream>>>(p, reinterpret_cast<float4 *>(out));
  }
  DSPMB_CUDA_TRY(cudaGetLastError());
  return DSPMB_OK;
}

}  // namespace
}  // namespace dspmb

using namespace dspmb;

extern "C" int dspmb_prior_f32(float *out, int in_height, int in_width, const float *sizes, int num_sizes,
                               const float *ratios, int num_ratios, float step_y, float step_x, float offset_y,
                               float offset_x, int clip, void *stream) {
  DSPMB_REQUIRE(sizes && ratios, "MultiBoxPrior: sizes/ratios are NULL");
  PriorParams p;
  int kinds = 0, anchors = 0;
  int rc = fill_map(p, 0, in_height, in_width, sizes, num_sizes, ratios, num_ratios, step_y, step_x, offset_y,
                    offset_x, kinds, anchors);
  if (rc) return rc;
  p.num_maps = 1;
  p.total_anchors = anchors;
  p.clip = clip;
  return launch(p, out, (cudaStream_t)stream);
}

extern "C" int dspmb_prior_multi_f32(float *out, int num_maps, const int *heights, const int *widths,
                                     const float *sizes, const int *num_sizes, const float *ratios,
                                     const int *num_ratios, const float *steps, const float *offsets, int clip,
                                     void *stream) {
  DSPMB_REQUIRE(num_maps > 0 && num_maps <= DSPMB_MAX_MAPS, "MultiBoxPrior: 1..%d feature maps per launch", DSPMB_MAX_MAPS);
  DSPMB_REQUIRE(heights && widths && sizes && num_sizes && ratios && num_ratios && steps && offsets,
                "MultiBoxPrior: NULL parameter array");
  PriorParams p;
  int kinds = 0, anchors = 0;
  const float *s = sizes, *r = ratios;
  for (int m = 0; m < num_maps; ++m) {
    int rc = fill_map(p, m, heights[m], widths[m], s, num_sizes[m], r, num_ratios[m], steps[2 * m], steps[2 * m + 1],
                      offsets[2 * m], offsets[2 * m + 1], kinds, anchors);
    if (rc) return rc;
    s += num_sizes[m];
    r += num_ratios[m];
  }
  p.num_maps = num_maps;
  p.total_anchors = anchors;
  p.clip = clip;
  return launch(p, out, (cudaStream_t)stream);
}
