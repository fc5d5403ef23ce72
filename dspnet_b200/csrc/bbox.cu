// bbox_overlaps_cython for sm_100a -- the reference's shipped IoU helper (cython/bbox.pyx:15-55).
//
// overlaps[n, k] of boxes (N, 4) and query_boxes (K, 4) in float64 with the "+1" pixel convention; 0 unless both
// the width and the height of the intersection are > 0.  Every operation is an individually rounded fp64 op in
// the reference's order (no FMA contraction): iw = (min - max) + 1, ua = (bw * bh + area_q) - iw * ih, iw * ih / ua.
//
// The output (8 N K bytes) is the only traffic that matters, so the kernel is a streaming writer: a CTA owns a
// 32 x 256 tile of the output, each thread keeps two query boxes and their areas in registers for the 32 rows, the
// row's box is a broadcast load, and consecutive threads write consecutive pairs of k (coalesced 16-byte streaming
// stores; the result is written once and not read back by the kernel).
#include "common.cuh"

namespace dspmb {
namespace {

constexpr int kBboxThreads = 128, kBboxCols = 256, kBboxRows = 32;  // 2 query columns per thread

struct QueryBox {
  double x0, y0, x1, y1, area;
};

__device__ __forceinline__ double overlap_of(double bx0, double by0, double bx1, double by1, double ab, const QueryBox &q) {
  const double iw = __dadd_rn(__dsub_rn(bx1 < q.x1 ? bx1 : q.x1, bx0 > q.x0 ? bx0 : q.x0), 1.0);
  if (!(iw > 0.0)) return 0.0;
  const double ih = __dadd_rn(__dsub_rn(by1 < q.y1 ? by1 : q.y1, by0 > q.y0 ? by0 : q.y0), 1.0);
  if (!(ih > 0.0)) return 0.0;
  const double inter = __dmul_rn(iw, ih);
  return __ddiv_rn(inter, __dsub_rn(__dadd_rn(ab, q.area), inter));
}

__global__ void __launch_bounds__(kBboxThreads) bbox_overlaps_kernel(const double *__restrict__ boxes, int N,
                                                                     const double *__restrict__ query, int K,
                                                                     double *__restrict__ out) {
  const int k = blockIdx.x * kBboxCols + threadIdx.x * 2, n0 = blockIdx.y * kBboxRows;
  QueryBox q[2];
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    const int kk = k + c < K ? k + c : 0;
    q[c].x0 = query[(size_t)kk * 4];
    q[c].y0 = query[(size_t)kk * 4 + 1];
    q[c].x1 = query[(size_t)kk * 4 + 2];
    q[c].y1 = query[(size_t)kk * 4 + 3];
    q[c].area = __dmul_rn(__dadd_rn(__dsub_rn(q[c].x1, q[c].x0), 1.0), __dadd_rn(__dsub_rn(q[c].y1, q[c].y0), 1.0));
  }
  // row starts are 16-byte aligned when K is even AND the output tensor itself is (a C-ABI caller may pass a view)
  const bool pair_store = (K & 1) == 0 && k + 1 < K && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
  const int rows = min(kBboxRows, N - n0);
  for (int r = 0; r < rows; ++r) {
    const int n = n0 + r;
    // the row's box is the same for the whole CTA: broadcast loads through the read-only path
    const double bx0 = __ldg(boxes + (size_t)n * 4), by0 = __ldg(boxes + (size_t)n * 4 + 1);
    const double bx1 = __ldg(boxes + (size_t)n * 4 + 2), by1 = __ldg(boxes + (size_t)n * 4 + 3);
    if (k >= K) continue;
    const double ab = __dmul_rn(__dadd_rn(__dsub_rn(bx1, bx0), 1.0), __dadd_rn(__dsub_rn(by1, by0), 1.0));
    const double v0 = overlap_of(bx0, by0, bx1, by1, ab, q[0]);
    double *o = out + (size_t)n * K + k;
    if (pair_store) {
      const double v1 = overlap_of(bx0, by0, bx1, by1, ab, q[1]);
      __stcs(reinterpret_cast<double2 *>(o), make_double2(v0, v1));
    } else {
      __stcs(o, v0);
      if (k + 1 < K) __stcs(o + 1, overlap_of(bx0, by0, bx1, by1, ab, q[1]));
    }
  }
}

}  // namespace
}  // namespace dspmb

using namespace dspmb;

extern "C" int dspmb_bbox_overlaps_f64(const double *boxes, int N, const double *query_boxes, int K, double *overlaps,
                                       void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  DSPMB_REQUIRE(N >= 0 && K >= 0, "bbox_overlaps: bad shape N=%d K=%d", N, K);
  if (N == 0 || K == 0) return DSPMB_OK;
  DSPMB_REQUIRE(boxes && query_boxes && overlaps, "bbox_overlaps: NULL tensor");
  DSPMB_REQUIRE(ceil_div(N, kBboxRows) <= 65535, "bbox_overlaps: more than %d boxes are not supported in one call",
                65535 * kBboxRows);
  dim3 grid(ceil_div(K, kBboxCols), ceil_div(N, kBboxRows));
  bbox_overlaps_kernel<<<grid, kBboxThreads, 0, stream>>>(boxes, N, query_boxes, K, overlaps);
  DSPMB_CUDA_TRY(cudaGetLastError());
  return DSPMB_OK;
}
