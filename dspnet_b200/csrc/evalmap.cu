// Detection post-filter and mAP matching for sm_100a -- the consumer of MultiBoxDetection's output in evaluation
// (SURVEY.md 8f row f3): multi_solver.py:419-432 (rows with id >= 0 and score > 0.25, padded to 200) and the
// per-image TP / FP matching of evaluate/eval_metric.py:113-160.  Doing both on the device means only K rows and K
// flags per image cross PCIe instead of the (B, A, 7) operator output.
//
// Matching semantics (MApMetric.update): per image and class, detections in row order; each takes the gt of its
// class with the largest IoU (numpy argmax: first maximum, NaN counts as maximal); IoU > ovp_thresh -> TP if that gt is
// still free (and marks it), FP if it was taken, nothing if the gt is "difficult" (label column 5 > 0) and
// use_difficult is off; otherwise FP; no gt of the class -> FP.  Classes own disjoint label rows, so one pass over
// the rows in order with a `found` bit per label row reproduces the reference's class-by-class loop.  The IoU is
// evaluated like numpy does (eval_metric.py:96-105): individually rounded fp32 operations, max(x, 0), uni < 1e-12 -> 0.
#include "common.cuh"

namespace dspmb {
namespace {

__global__ void __launch_bounds__(256) det_postfilter_kernel(const float *__restrict__ out, const int *__restrict__ valid,
                                                             int A, int K, float score_thresh,
                                                             float *__restrict__ dst, int *__restrict__ counts) {
  __shared__ int scan_smem[256 / 32 + 1];
  __shared__ int carry_smem;
  const int b = blockIdx.x;
  const float *src = out + (size_t)b * A * 7;
  float *d = dst + (size_t)b * K * 7;
  const int V = valid ? min(valid[b], A) : A;
  if (threadIdx.x == 0) carry_smem = 0;
  __syncthreads();
  for (int base = 0; base < V; base += blockDim.x) {
    const int r = base + threadIdx.x;
    const int keep = (r < V && src[(size_t)r * 7] >= 0.f && src[(size_t)r * 7 + 1] > score_thresh) ? 1 : 0;
    int total;
    const int ex = block_scan_excl(keep, scan_smem, &total);
    const int carry = carry_smem;
    const int pos = carry + ex;
    if (keep && pos < K) {
#pragma unroll
      for (int c = 0; c < 7; ++c) d[(size_t)pos * 7 + c] = src[(size_t)r * 7 + c];
    }
    __syncthreads();
    if (threadIdx.x == 0) carry_smem = carry + total;
    __syncthreads();
    if (carry_smem >= K) break;
  }
  const int n = min(carry_smem, K);
  for (int q = n * 7 + threadIdx.x; q < K * 7; q += blockDim.x) d[q] = -1.f;
  if (threadIdx.x == 0) counts[b] = n;
}

// One warp per image (a CTA of 32 threads).  The image's label rows and prediction rows are staged in shared
// memory when they fit (the loop over the M predictions is sequential -- `found` carries over -- so a global-memory
// round trip per prediction would be the whole cost); `found` is one byte per label row.
template <bool kStaged>
__global__ void __launch_bounds__(32) map_match_kernel(const float *__restrict__ labels, int L, int W,
                                                       const float *__restrict__ preds, int M, int PW, float ovp_thresh,
                                                       int use_difficult, int *__restrict__ flags) {
  extern __shared__ __align__(16) unsigned char match_smem[];  // [labels L*W f32][preds M*PW f32] (staged) [found L u8]
  const int lane = threadIdx.x;
  const int b = blockIdx.x;
  const float *lab = labels + (size_t)b * L * W;
  const float *prd = preds + (size_t)b * M * PW;
  unsigned char *found = match_smem;
  if (kStaged) {
    float *sl = reinterpret_cast<float *>(match_smem);
    float *sp = sl + L * W;
    for (int q = lane; q < L * W; q += 32) sl[q] = lab[q];
    for (int q = lane; q < M * PW; q += 32) sp[q] = prd[q];
    lab = sl;
    prd = sp;
    found = reinterpret_cast<unsigned char *>(sp + M * PW);
  }
  for (int l = lane; l < L; l += 32) found[l] = 0;
  __syncwarp();
  for (int j = 0; j < M; ++j) {
    const float *p = prd + (size_t)j * PW;
    const int cid = (int)p[0];  // int(pred[0, 0]): truncation toward zero
    int flag = 0;
    if (cid >= 0) {
      const float x0 = p[2], y0 = p[3], x1 = p[4], y1 = p[5];
      const float area_x = fmul(fsub(x1, x0), fsub(y1, y0));
      // lane-local first maximum over the gts of this class, then the warp's first maximum (lowest label row)
      float best = 0.f;
      int best_l = -1;
      bool best_nan = false;
      for (int l = lane; l < L; l += 32) {
        const float *g = lab + (size_t)l * W;
        if ((int)g[0] != cid) continue;
        const float ixmin = fmaxf(g[1], x0), iymin = fmaxf(g[2], y0), ixmax = fminf(g[3], x1), iymax = fminf(g[4], y1);
        const float iw = fmaxf(fsub(ixmax, ixmin), 0.f), ih = fmaxf(fsub(iymax, iymin), 0.f);
        const float inters = fmul(iw, ih);
        const float uni = fsub(fadd(area_x, fmul(fsub(g[3], g[1]), fsub(g[4], g[2]))), inters);
        float iou = fdiv(inters, uni);
        if (uni < 1e-12f) iou = 0.f;
        const bool is_nan = iou != iou;
        // numpy argmax: the first NaN wins outright, otherwise the first strictly larger value
        if (best_l < 0 || (!best_nan && (is_nan || iou > best))) {
          best = iou;
          best_l = l;
          best_nan = is_nan;
        }
      }
#pragma unroll
      for (int m = 16; m > 0; m >>= 1) {
        const float ob = __shfl_xor_sync(kFullMask, best, m);
        const int ol = __shfl_xor_sync(kFullMask, best_l, m);
        const bool on = __shfl_xor_sync(kFullMask, (int)best_nan, m) != 0;
        if (ol < 0) continue;
        bool take;
        if (best_l < 0) take = true;
        else if (on != best_nan) take = on;                                  // a NaN beats any number ...
        else if (on) take = ol < best_l;                                     // ... the earliest NaN wins
        else take = ob > best || (ob == best && ol < best_l);               // first maximum
        if (take) {
          best = ob;
          best_l = ol;
          best_nan = on;
        }
      }
      if (best_l < 0) {
        flag = 2;  // no ground truth of this class
      } else if (!best_nan && best > ovp_thresh) {
        const bool difficult = !use_difficult && W >= 6 && lab[(size_t)best_l * W + 5] > 0.f;
        if (!difficult) {
          flag = found[best_l] ? 2 : 1;
          __syncwarp();
          if (lane == 0) found[best_l] = 1;
        }
      } else {
        flag = 2;
      }
      __syncwarp();
    }
    if (lane == 0) flags[(size_t)b * M + j] = flag;
  }
}

}  // namespace
}  // namespace dspmb

using namespace dspmb;

extern "C" int dspmb_detection_postfilter_f32(const float *out, const int32_t *valid_count, int B, int A, int K,
                                              float score_thresh, float *rows, int32_t *counts, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  DSPMB_REQUIRE(B >= 0 && A > 0 && K > 0, "detection_postfilter: bad shape B=%d A=%d K=%d", B, A, K);
  DSPMB_REQUIRE(out && rows && counts, "detection_postfilter: NULL tensor");
  if (B == 0) return DSPMB_OK;
  det_postfilter_kernel<<<B, 256, 0, stream>>>(out, valid_count, A, K, score_thresh, rows, counts);
  DSPMB_CUDA_TRY(cudaGetLastError());
  return DSPMB_OK;
}

extern "C" int dspmb_map_match_f32(const float *labels, int B, int L, int label_width, const float *preds, int M,
                                   int pred_width, float ovp_thresh, int use_difficult, int32_t *flags, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  DSPMB_REQUIRE(B >= 0 && L >= 0 && M >= 0, "map_match: bad shape B=%d L=%d M=%d", B, L, M);
  DSPMB_REQUIRE(label_width >= 5 && pred_width >= 6, "map_match: labels need >= 5 columns, predictions >= 6");
  if (B == 0 || M == 0) return DSPMB_OK;
  DSPMB_REQUIRE(preds && flags && (labels || L == 0), "map_match: NULL tensor");
  DSPMB_REQUIRE(L <= 40 * 1024, "map_match: more than %d label rows are not supported", 40 * 1024);
  const size_t staged = sizeof(float) * ((size_t)L * label_width + (size_t)M * pred_width) + (size_t)(L > 0 ? L : 1);
  if (staged <= 44 * 1024)
    map_match_kernel<true><<<B, 32, staged, stream>>>(labels, L, label_width, preds, M, pred_width, ovp_thresh, use_difficult, flags);
  else
    map_match_kernel<false><<<B, 32, (size_t)(L > 0 ? L : 1), stream>>>(labels, L, label_width, preds, M, pred_width, ovp_thresh,
                                                                      use_difficult, flags);
  DSPMB_CUDA_TRY(cudaGetLastError());
  return DSPMB_OK;
}
