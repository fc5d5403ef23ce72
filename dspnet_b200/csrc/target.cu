// MultiBoxTarget for sm_100a.
//
// Reference semantics (CPU operator, the parity target): operator/multibox_target-inl.h:89-171 (output init
// :121-124, IoU planes :137-161) and operator/multibox_target.cc:72-284:
//   G            (:95-105)  number of leading label rows with cls != -1;
//   bipartite    (:113-149) repeatedly take the global argmax IoU (> 1e-6, first in (anchor, gt) scan order)
//                           over unmatched anchors x unmatched gts;
//   threshold    (:151-180) every other anchor: first-max gt; positive iff max IoU > overlap_threshold;
//   mining       (:182-241) num_negative = (int)(num_positive * ratio) clamped to A - num_positive; candidates are
//                           non-positive anchors with max IoU < negative_mining_thresh; stable sort by background
//                           softmax probability ascending (ties: lower anchor first); first num_negative -> 0;
//   write        (:251-281) positives: cls+1, mask 1x5, 5-wide encoding; negatives: 0; rest: ignore_label.
// The reference materialises 11 broadcast planes of B*A*L floats for the IoU (4 GB at SSD-512, B=64); its GPU
// kernels (operator/multibox_target.cu) diverge from the CPU semantics and are not followed.
//
// Structure here (memset + two launches, all images in every grid):
//   target_stream_kernel  HBM-bound.  One thread per 4 consecutive anchors; ground truths staged in shared
//                         memory; fused IoU / first-max row argmax / column argmax (packed 64-bit keys: warp
//                         shuffle max -> shared atomicMax -> one global atomicMax per CTA and gt); exact two-pass
//                         softmax of the background logit (glibc-bit-exact expf in fp64); writes loc_target,
//                         loc_mask, provisional cls_target and the 32-bit mining key of every anchor.
//   target_match_kernel   one CTA per image: lazy greedy bipartite matching on the cached column maxima (a stale
//                         maximum is an upper bound, so it is recomputed only when it reaches the top), output
//                         fix-up of the <= G matched anchors, then an MSB-first radix select (4 x 8 bit) of the
//                         num_negative smallest (probability, anchor) keys with an ordered tie pass.
#include <cooperative_groups.h>

#include "common.cuh"

namespace dspmb {
namespace {

constexpr int kStreamThreads = 128;
constexpr int kMatchThreads = 1024;
constexpr int kShortCap = 4096;  // matcher: capacity of the mining shortlist (keys + anchor ids in shared memory)
// smallest A the shortlist is used for: below, four passes over the keys cost less than the sample + compaction barriers
// (SSD-300, 8732 anchors, one image, cold L2: 59.4 us per step without, 61.5 us with)
constexpr int kShortMinAnchors = 12000;
#ifndef DSPMB_TARGET_MIN_BLOCKS
#define DSPMB_TARGET_MIN_BLOCKS 5
#endif
constexpr int kTargetMinBlocks = DSPMB_TARGET_MIN_BLOCKS;  // CTAs per SM the stream kernel is compiled for
constexpr unsigned kKeySentinel = 0xffffffffu;  // "not a mining candidate"

// Optional per-CTA phase stamps of the match kernel (dspmb_debug_target_stamps): 12 x uint64 %globaltimer values per
// image.  The pointer lives in constant memory, so the disabled check is one LDC.
__constant__ unsigned long long *g_tstamps = nullptr;
__constant__ unsigned long long *g_sstamps = nullptr;  // stream kernel: 8 stamps per CTA
#define DSPMB_SSTAMP(k)                                                                         \
  do {                                                                                          \
    if (g_sstamps && threadIdx.x == 0) {                                                        \
      unsigned long long t_;                                                                    \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                                    \
      g_sstamps[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 8 + (k)] = t_;                  \
    }                                                                                           \
  } while (0)
#define DSPMB_TSTAMP_IMG(img, k)                                                                \
  do {                                                                                          \
    if (g_tstamps && threadIdx.x == 0) {                                                        \
      unsigned long long t_;                                                                    \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                                    \
      g_tstamps[(size_t)(img) * 12 + (k)] = t_;                                                 \
    }                                                                                           \
  } while (0)
#define DSPMB_TSTAMP(k) DSPMB_TSTAMP_IMG(blockIdx.x, k)

struct TargetWorkspace {
  WsHeader *header;
  int *gcount;     // (B) valid ground truths | (B) label-padding status, written by tile 0 of every image
  int *thr_count;  // (B, Tmax) threshold-stage positives of every tile
  unsigned long long *colbest;  // (B, Tmax, L) best (iou, anchor) per tile and gt (0: none)
  unsigned *key;   // (B, A) mining keys
  int *amb_list;      // (B, A) anchors inside the pivot's error band
  unsigned *amb_key;  // (B, A) their exact keys
  size_t bytes;
};

// Nothing in the workspace needs zeroing: every CTA of the stream kernel owns its slots (the match kernel reduces
// them), so the operator is two launches and no memset node.
TargetWorkspace carve(void *base, int B, int A, int L) {
  TargetWorkspace w;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off = align_up(off + bytes, 256);
    return (char *)base + o;
  };
  w.header = (WsHeader *)take(sizeof(WsHeader));
  const size_t Tmax = (size_t)ceil_div(A, kStreamThreads);
  w.gcount = (int *)take(sizeof(int) * 2 * B);
  w.thr_count = (int *)take(sizeof(int) * B * Tmax);
  w.colbest = (unsigned long long *)take(sizeof(unsigned long long) * (size_t)B * Tmax * L);
  w.key = (unsigned *)take(sizeof(unsigned) * (size_t)B * A);
  w.amb_list = (int *)take(sizeof(int) * (size_t)B * A);
  w.amb_key = (unsigned *)take(sizeof(unsigned) * (size_t)B * A);
  w.bytes = off;
  return w;
}

struct TargetArgs {
  const float *anchors, *labels, *cls_preds;
  float *loc_target, *loc_mask, *cls_target;
  int32_t *match_out, *stats_out;
  WsHeader *header;
  int *gcount, *thr_count;
  unsigned long long *colbest;
  unsigned *key;
  int *amb_list;
  unsigned *amb_key;
  float delta;  // relative error bound of the approximate mining keys
  int B, A, L, W, C, T;
  float overlap_threshold, ignore_label, mining_ratio, mining_thresh;
  float vx, vy, vw, vh;
  int fma_build;
  int prefetch;   // stream kernel: late L2 prefetch distance in CTAs of launch order (0: none)
  int shortlist;  // matcher: mine on a shortlist of the keys below a sampled bound (DSPMB_TUNE_TARGET_SHORTLIST)
};

// 5-wide box encoding, operator/multibox_target.cc:30-56.
__device__ __forceinline__ void encode_loc(float4 an, const float *lab, const TargetArgs &a, float *dst) {
  const float aw = fsub(an.z, an.x);
  const float ah = fsub(an.w, an.y);
  const float ax = fmul(fadd(an.x, an.z), 0.5f);  // (al + ar) * 0.5 in double and back: exact either way
  const float ay = fmul(fadd(an.y, an.w), 0.5f);
  const float gl = lab[1], gt = lab[2], gr = lab[3], gb = lab[4], gz = lab[5];
  const float gw = fsub(gr, gl);
  const float gh = fsub(gb, gt);
  const float gx = fmul(fadd(gl, gr), 0.5f);
  const float gy = fmul(fadd(gt, gb), 0.5f);
  dst[0] = fdiv(fdiv(fsub(gx, ax), aw), a.vx);
  dst[1] = fdiv(fdiv(fsub(gy, ay), ah), a.vy);
  dst[2] = fdiv(libm::logf_glibc(fdiv(gw, aw), a.fma_build), a.vw);
  dst[3] = fdiv(libm::logf_glibc(fdiv(gh, ah), a.fma_build), a.vh);
  dst[4] = __double2float_rn(__ddiv_rn((double)gz, 0.1));  // DType(gz) / 0.1 is a double division
}

__device__ __forceinline__ unsigned long long col_key(float iou, int anchor) {
  // larger IoU wins, then the lower anchor index (scan order of multibox_target.cc:117-134)
  return ((unsigned long long)__float_as_uint(iou) << 32) | (unsigned)(0xffffffffu - (unsigned)anchor);
}

// Number of leading valid label rows (multibox_target.cc:95-105), computed by every warp on its own (32 label rows per
// ballot): no shared memory, no block barrier.
__device__ __forceinline__ int count_valid_gt_warp(const float *lab, int L, int W) {
  for (int l0 = 0; l0 < L; l0 += 32) {
    const int l = l0 + (int)lane_id();
    const unsigned pad = __ballot_sync(kFullMask, l < L && __ldg(lab + (size_t)l * W) == -1.0f);
    if (pad) return l0 + __ffs(pad) - 1;
  }
  return L;
}

// exp(d) for d <= 0 in fp32 through MUFU.EX2: the product d*log2(e) is rounded once (absolute error <= 2^-18 for
// |d*log2e| < 128, i.e. a relative error of the result <= 2.7e-6) and ex2.approx has a relative error <= 2^-22.
__device__ __forceinline__ float exp_approx(float d) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(fmul(d, 1.4426950408889634f)));
  return r;
}
// exp(x - mx) with the subtraction folded into one fused multiply-add: ex2(x*log2e - mx*log2e), `nml` = -mx*log2e
// rounded once.  The argument differs from (x - mx)*log2e by <= 2^-24 |mx*log2e| + 2^-24 |arg| (two roundings of
// magnitudes <= |mx| log2e each); for the |logits| <= 64 a softmax sees that is an absolute error <= 1.1e-5 in the
// exponent, i.e. a relative error <= 7.7e-6 of the result -- inside the bound delta the match kernel's band assumes.
// Only used while |mx| <= 64 (the caller checks), where the rounding of nml is <= 2^-18.
__device__ __forceinline__ float exp_approx_fma(float x, float nml) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(__fmaf_rn(x, 1.4426950408889634f, nml)));
  return r;
}

// The reference's background probability, bit for bit (multibox_target.cc:220-231): running max, sequential fp32 sum
// of glibc expf, one division.  Used for the few anchors near the selection pivot and for denormal-range values.
template <bool kFma>
__device__ __noinline__ float exact_bg_prob(const float *p_cls, int j, int A, int C) {
  float mx = __ldg(p_cls + j);
  for (int k = 1; k < C; ++k) {
    const float t = __ldg(p_cls + j + (size_t)A * k);
    if (t > mx) mx = t;
  }
  float sum = 0.f;
  for (int k = 0; k < C; ++k)
    sum = fadd(sum, libm::expf_glibc_t<kFma>(fsub(__ldg(p_cls + j + (size_t)A * k), mx), libm::GlobalExpTab()));
  return fdiv(libm::expf_glibc_t<kFma>(fsub(__ldg(p_cls + j), mx), libm::GlobalExpTab()), sum);
}

// The same value computed by a whole warp for ONE anchor (every lane passes the same j and receives the result): the
// C logits are loaded and exponentiated in parallel, only the fp32 sum keeps the reference's class order (one
// shuffle + add per class).  Used when only a handful of anchors need the exact value, where the serial routine's
// 2 C dependent loads and expf calls would sit on the critical path of the match kernel.
template <bool kFma>
__device__ __forceinline__ float exact_bg_prob_warp(const float *p_cls, int j, int A, int C) {
  const int lane = (int)lane_id();
  float mx = __int_as_float(0xff800000);
  for (int k = lane; k < C; k += 32) {
    const float t = __ldg(p_cls + j + (size_t)A * k);
    if (t > mx) mx = t;
  }
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) {
    const float o = __shfl_xor_sync(kFullMask, mx, m);
    if (o > mx) mx = o;
  }
  float sum = 0.f, e0 = 0.f;
  for (int k0 = 0; k0 < C; k0 += 32) {
    const int k = k0 + lane;
    const float e = k < C ? libm::expf_glibc_t<kFma>(fsub(__ldg(p_cls + j + (size_t)A * k), mx), libm::GlobalExpTab()) : 0.f;
    if (k0 == 0) e0 = __shfl_sync(kFullMask, e, 0);
    const int n = min(32, C - k0);
    for (int i = 0; i < n; ++i) sum = fadd(sum, __shfl_sync(kFullMask, e, i));
  }
  return fdiv(e0, sum);
}

// IoU of the target operator with the division skipped for disjoint boxes: inter == 0 gives iou == 0 both through
// safe_divide's union == 0 branch and through 0 / union (multibox_target-inl.h:44-50,153-161).
__device__ __forceinline__ float iou_target_fast(float4 a, float area_a, float4 g, float area_g) {
  const float mr = a.z < g.z ? a.z : g.z;
  const float ml = a.x > g.x ? a.x : g.x;
  const float mb = a.w < g.w ? a.w : g.w;
  const float mt = a.y > g.y ? a.y : g.y;
  const float dw = fsub(mr, ml);
  const float dh = fsub(mb, mt);
  if (!(dw > 0.0f) || !(dh > 0.0f)) {
    // iw or ih is 0 (or -0 * x): inter = +-0 -> iou = 0, except NaN inputs, which the reference does not see
    return 0.0f;
  }
  const float inter = fmul(dw, dh);
  const float uni = fsub(fadd(area_a, area_g), inter);
  if (uni == 0.0f) return 0.0f;
  return fdiv(inter, uni);
}

// VEC: anchors per thread (4 with 128-bit accesses, 1 for unaligned shapes).  NC > 0: compile-time class count, the
// thread keeps its NC x VEC logits in registers (one HBM read, no re-load for the second softmax pass); NC == 0:
// runtime class count, second pass re-reads through L1/L2.  kFma: which glibc build of expf/logf to reproduce.
template <int VEC, int NC, bool kFma>
__global__ void __launch_bounds__(kStreamThreads, kTargetMinBlocks) target_stream_kernel(const __grid_constant__ TargetArgs a) {
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  __shared__ int sm_pos;
  float4 *sm_gt = reinterpret_cast<float4 *>(dyn_smem);                                  // [L]
  unsigned long long *sm_col = reinterpret_cast<unsigned long long *>(sm_gt + a.L);       // [L]
  float *sm_garea = reinterpret_cast<float *>(sm_col + a.L);                              // [L]

  // the match kernel is launched as a programmatic dependent: let its CTAs take the SMs this grid's tail frees
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int b = blockIdx.y, t = blockIdx.x;
  constexpr int kTile = kStreamThreads * VEC;
  const int A = a.A, W = a.W;
  const int i0 = t * kTile + threadIdx.x * VEC;
  const float *lab = a.labels + (size_t)b * a.L * W;
  const bool active = i0 < A;
  const bool mining = a.mining_ratio > 0.f;
  const int C = NC > 0 ? NC : a.C;

  // ---- issue every global load of this thread first.  The label loads go out BEFORE the 21 logit loads: the warp
  // needs them first (valid-gt count, gt staging), and queued behind the logits they were the kernel's largest stall.
  // Thread k < L fetches label row k speculatively, every lane the class fields of rows lane, lane + 32, .. + 96.
  float4 gsp = make_float4(0.f, 0.f, 0.f, 0.f);
  if ((int)threadIdx.x < a.L) {
    const float *row = lab + (size_t)threadIdx.x * W;
    gsp = make_float4(__ldg(row + 1), __ldg(row + 2), __ldg(row + 3), __ldg(row + 4));
  }
  float cfield[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int l = r * 32 + (int)lane_id();
    cfield[r] = l < a.L ? __ldg(lab + (size_t)l * W) : 0.f;
  }
  // ---- logits (register-resident when NC > 0) and anchors ----
  const float *cp = a.cls_preds + (size_t)b * C * A + i0;
  float xr[NC > 0 ? NC : 1][VEC];
  float4 an[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) an[v] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (active) {
    if (NC > 0 && mining) {
      // one 64-bit add of the row stride IN BYTES per load: indexing cp + c * A made the compiler carry a 64-bit
      // element index and scale it for every row (four address instructions per load, 9 % of the kernel)
      const char *pc = reinterpret_cast<const char *>(cp);
      const size_t row_bytes = (size_t)A * sizeof(float);
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        if constexpr (VEC == 4) {
          const float4 q = ld_stream_f4(reinterpret_cast<const float *>(pc));
          xr[c][0] = q.x, xr[c][1] = q.y, xr[c][2] = q.z, xr[c][3] = q.w;
        } else if constexpr (VEC == 2) {
          const float2 q = ld_stream_f2(reinterpret_cast<const float *>(pc));
          xr[c][0] = q.x, xr[c][1] = q.y;
        } else {
          xr[c][0] = ld_stream_f1(reinterpret_cast<const float *>(pc));
        }
        pc += row_bytes;
      }
    }
#pragma unroll
    for (int v = 0; v < VEC; ++v) an[v] = __ldg(reinterpret_cast<const float4 *>(a.anchors) + i0 + v);
  }

  DSPMB_SSTAMP(0);
  int G = -1;  // number of leading valid label rows (multibox_target.cc:95-105), every warp on its own
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int l = r * 32 + (int)lane_id();
    const unsigned pad = __ballot_sync(kFullMask, l < a.L && cfield[r] == -1.0f);
    if (G < 0 && pad) G = r * 32 + __ffs(pad) - 1;
  }
  if (G < 0) G = a.L <= 128 ? a.L : 128 + count_valid_gt_warp(lab + (size_t)128 * W, a.L - 128, W);
  DSPMB_SSTAMP(1);
  if (t == 0 && threadIdx.x == 0) {
    a.gcount[b] = G;
    int bad = DSPMB_OK;
    if (G < a.L) {  // CHECK_EQ on the first padding row, multibox_target.cc:98-101
      const float *row = lab + (size_t)G * W;
      if (row[1] != -1.0f || row[2] != -1.0f || row[3] != -1.0f || row[4] != -1.0f) bad = DSPMB_ERR_LABEL_PADDING;
    }
    a.gcount[a.B + b] = bad;                   // latched into the header by the match kernel
    if (b == 0) a.header->status = DSPMB_OK;   // nobody else touches the header during this launch
  }
  for (int k = threadIdx.x; k < G; k += blockDim.x) {
    float4 g = gsp;  // row k == threadIdx.x was fetched up front
    if (k >= (int)blockDim.x) {
      const float *row = lab + (size_t)k * W;
      g = make_float4(row[1], row[2], row[3], row[4]);
    }
    sm_gt[k] = g;
    sm_garea[k] = fmul(fsub(g.z, g.x), fsub(g.w, g.y));
    sm_col[k] = 0ull;
  }
  if (threadIdx.x == 0) sm_pos = 0;
  __syncthreads();
  DSPMB_SSTAMP(2);
  if (a.prefetch > 0 && threadIdx.x == 64 && mining) {
    // late L2 prefetch (DSPMB_TUNE_TARGET_PREFETCH): this CTA's own loads have been issued long ago; ask L2 for the
    // logits tile of the CTA `prefetch` launches ahead so that DRAM has work while the resident CTAs compute
    const int lin = b * (int)gridDim.x + t + a.prefetch;
    const int pb = lin / (int)gridDim.x, pt = lin - pb * (int)gridDim.x;
    if (pb < (int)gridDim.y) {
      const int pbegin = pt * kTile, prow = min(kTile, A - pbegin);
      const float *pp = a.cls_preds + (size_t)pb * C * A + pbegin;
      if (prow > 0 && (prow & 3) == 0 && (reinterpret_cast<uintptr_t>(pp) & 15) == 0 && (A & 3) == 0)
        for (int c = 0; c < C; ++c)
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(pp + (size_t)c * A), "r"(prow * 4) : "memory");
    }
  }

  float best_iou[VEC], area[VEC];
  int best_k[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    best_iou[v] = -1.0f;
    best_k[v] = -1;
    area[v] = fmul(fsub(an[v].z, an[v].x), fsub(an[v].w, an[v].y));
  }

  // Bounding box of the warp's anchors (consecutive anchors are neighbouring cells of one feature map): a gt that
  // does not reach into it has IoU 0 with every anchor of the warp, which only matters for gt 0 (the row's first
  // maximum starts at iou 0, gt 0) -- three quarters of the (warp, gt) pairs leave the loop after four compares.
  float4 bb = make_float4(__int_as_float(0x7f800000), __int_as_float(0x7f800000), __int_as_float(0xff800000),
                          __int_as_float(0xff800000));
  if (active) {
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      bb.x = fminf(bb.x, an[v].x);
      bb.y = fminf(bb.y, an[v].y);
      bb.z = fmaxf(bb.z, an[v].z);
      bb.w = fmaxf(bb.w, an[v].w);
    }
  }
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) {
    bb.x = fminf(bb.x, __shfl_xor_sync(kFullMask, bb.x, m));
    bb.y = fminf(bb.y, __shfl_xor_sync(kFullMask, bb.y, m));
    bb.z = fmaxf(bb.z, __shfl_xor_sync(kFullMask, bb.z, m));
    bb.w = fmaxf(bb.w, __shfl_xor_sync(kFullMask, bb.w, m));
  }

  // ---- fused IoU + row first-max + column max (multibox_target-inl.h:137-161, .cc:113-134,158-166) ----
  // The warp first lists the gts that reach its bounding box (32 gts per ballot, gt order kept), then walks the list
  // two gts at a time: the four IoUs of an iteration (two anchors x two gts) are independent, so their divisions and
  // the two warp-wide column maxima overlap instead of queueing behind a branch per gt.
  unsigned short *wlist = reinterpret_cast<unsigned short *>(sm_garea + a.L) + warp_id() * a.L;
  int nlist = 0;
  bool gt0_listed = false;
  for (int k0 = 0; k0 < G; k0 += 32) {
    const int k = k0 + (int)lane_id();
    bool reach = false;
    if (k < G) {
      const float4 g = sm_gt[k];
      reach = bb.z > g.x && g.z > bb.x && bb.w > g.y && g.w > bb.y;
    }
    const unsigned m = __ballot_sync(kFullMask, reach);
    if (reach) wlist[nlist + __popc(m & ((1u << lane_id()) - 1u))] = (unsigned short)k;
    if (k0 == 0) gt0_listed = (m & 1u) != 0u;
    nlist += __popc(m);
  }
  __syncwarp();
  if (G > 0 && !gt0_listed) {  // gt 0 has IoU 0 with every anchor of the warp: the row's first maximum so far
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      best_iou[v] = 0.0f;
      best_k[v] = 0;
    }
  }
  for (int i = 0; i < nlist; i += 2) {
    const bool two = i + 1 < nlist;
    const int k0 = wlist[i], k1 = two ? (int)wlist[i + 1] : k0;
    const float4 g0 = sm_gt[k0], g1 = sm_gt[k1];
    const float ga0 = sm_garea[k0], ga1 = sm_garea[k1];
    unsigned long long tkey0 = 0ull, tkey1 = 0ull;
    if (active) {
      float iou0[VEC], iou1[VEC];
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        iou0[v] = iou_target_fast(an[v], area[v], g0, ga0);
        iou1[v] = iou_target_fast(an[v], area[v], g1, ga1);
      }
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        if (iou0[v] > best_iou[v]) {
          best_iou[v] = iou0[v];
          best_k[v] = k0;
        }
        if (two && iou1[v] > best_iou[v]) {
          best_iou[v] = iou1[v];
          best_k[v] = k1;
        }
        if (iou0[v] > 1e-6f) {
          const unsigned long long ck = col_key(iou0[v], i0 + v);
          tkey0 = ck > tkey0 ? ck : tkey0;
        }
        if (two && iou1[v] > 1e-6f) {
          const unsigned long long ck = col_key(iou1[v], i0 + v);
          tkey1 = ck > tkey1 ? ck : tkey1;
        }
      }
    }
    const bool any0 = __any_sync(kFullMask, tkey0 != 0ull), any1 = __any_sync(kFullMask, tkey1 != 0ull);
    if (any0) {
      const unsigned long long wk = warp_max_u64(tkey0);
      if (lane_id() == 0) atomicMax(&sm_col[k0], wk);
    }
    if (any1) {
      const unsigned long long wk = warp_max_u64(tkey1);
      if (lane_id() == 0) atomicMax(&sm_col[k1], wk);
    }
  }

  DSPMB_SSTAMP(3);
  // ---- threshold-stage positives + mining keys ----
  bool pos[VEC];
  unsigned key[VEC];
  int npos = 0;
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    pos[v] = active && G > 0 && a.overlap_threshold > 0.f && best_k[v] >= 0 && best_iou[v] > a.overlap_threshold;
    npos += pos[v];
    key[v] = kKeySentinel;
  }
  if (mining && G > 0 && active) {
    // Mining key = background softmax probability (multibox_target.cc:220-231).  The stream kernel computes it in
    // fp32 with MUFU.EX2 (relative error << a.delta, see approx_softmax_bg); the match kernel selects on these keys
    // and re-evaluates with the bit-exact glibc expf only the handful of anchors whose key lies within the error
    // band of the selection pivot, so the chosen SET is exactly the reference's.  Probabilities small enough for
    // the approximation to lose relative accuracy (denormal range) are evaluated exactly right here.
    float mx[VEC], sum[VEC], p0[VEC];
    bool cand[VEC];
    bool any_cand = false;
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      cand[v] = !pos[v] && best_iou[v] < a.mining_thresh;
      any_cand |= cand[v];
      sum[v] = 0.f;
    }
    if (any_cand) {
      if constexpr (NC > 0) {
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
          mx[v] = xr[0][v];
          p0[v] = xr[0][v];
        }
#pragma unroll
        for (int c = 1; c < NC; ++c)
#pragma unroll
          for (int v = 0; v < VEC; ++v)
            mx[v] = fmaxf(mx[v], xr[c][v]);
        float nml[VEC];
        bool small = true;
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
          nml[v] = -fmul(mx[v], 1.4426950408889634f);
          small = small && fabsf(mx[v]) <= 64.0f;
        }
        if (small) {  // the error bound of exp_approx_fma needs |mx| <= 64; larger logits take the two-step form
#pragma unroll
          for (int c = 0; c < NC; ++c)
#pragma unroll
            for (int v = 0; v < VEC; ++v) sum[v] = fadd(sum[v], exp_approx_fma(xr[c][v], nml[v]));
        } else {
#pragma unroll
          for (int c = 0; c < NC; ++c)
#pragma unroll
            for (int v = 0; v < VEC; ++v) sum[v] = fadd(sum[v], exp_approx(fsub(xr[c][v], mx[v])));
        }
      } else {
#pragma unroll 4
        for (int c = 0; c < C; ++c) {
          float x[VEC];
          if constexpr (VEC == 4) {
            const float4 q = __ldg(reinterpret_cast<const float4 *>(cp + (size_t)c * A));
            x[0] = q.x, x[1] = q.y, x[2] = q.z, x[3] = q.w;
          } else if constexpr (VEC == 2) {
            const float2 q = __ldg(reinterpret_cast<const float2 *>(cp + (size_t)c * A));
            x[0] = q.x, x[1] = q.y;
          } else {
            x[0] = __ldg(cp + (size_t)c * A);
          }
#pragma unroll
          for (int v = 0; v < VEC; ++v) {
            if (c == 0) {
              mx[v] = x[v];
              p0[v] = x[v];
            } else if (x[v] > mx[v]) {
              mx[v] = x[v];
            }
          }
        }
#pragma unroll 4
        for (int c = 0; c < C; ++c) {
          float x[VEC];
          if constexpr (VEC == 4) {
            const float4 q = __ldg(reinterpret_cast<const float4 *>(cp + (size_t)c * A));
            x[0] = q.x, x[1] = q.y, x[2] = q.z, x[3] = q.w;
          } else if constexpr (VEC == 2) {
            const float2 q = __ldg(reinterpret_cast<const float2 *>(cp + (size_t)c * A));
            x[0] = q.x, x[1] = q.y;
          } else {
            x[0] = __ldg(cp + (size_t)c * A);
          }
#pragma unroll
          for (int v = 0; v < VEC; ++v) sum[v] = fadd(sum[v], exp_approx(fsub(x[v], mx[v])));
        }
      }
#pragma unroll
      for (int v = 0; v < VEC; ++v)
        if (cand[v]) {
          const float d0 = fsub(p0[v], mx[v]);
          float prob;
          if (d0 < -80.0f) {
            prob = exact_bg_prob<kFma>(a.cls_preds + (size_t)b * C * A, i0 + v, A, C);  // rare
          } else {
            prob = fdiv(exp_approx(d0), sum[v]);
          }
          key[v] = __float_as_uint(prob);
        }
    }
  }

  DSPMB_SSTAMP(4);
  // ---- outputs ----
  bool any_pos = false;
#pragma unroll
  for (int v = 0; v < VEC; ++v) any_pos |= pos[v];
  const bool zero_warp = (VEC == 2 || VEC == 4) && __all_sync(kFullMask, active && !any_pos) &&
                         ((reinterpret_cast<uintptr_t>(a.loc_target) | reinterpret_cast<uintptr_t>(a.loc_mask)) & 15) == 0 &&
                         (((size_t)b * A + (i0 - (int)lane_id() * VEC)) & 3) == 0;
  if (active) {
    const size_t row0 = (size_t)b * A + i0;
    float ct[VEC];
    float lt[VEC * 5], lm[VEC * 5];
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      if (pos[v]) {
        const float *lrow = lab + (size_t)best_k[v] * W;
        ct[v] = fadd(lrow[0], 1.0f);
        encode_loc(an[v], lrow, a, lt + v * 5);
#pragma unroll
        for (int c = 0; c < 5; ++c) lm[v * 5 + c] = 1.0f;
      } else {
        ct[v] = (G > 0 && !mining) ? 0.0f : a.ignore_label;  // multibox_target.cc:242-249 vs -inl.h:123
#pragma unroll
        for (int c = 0; c < 5; ++c) {
          lt[v * 5 + c] = 0.f;
          lm[v * 5 + c] = 0.f;
        }
      }
    }
    float *plt = a.loc_target + row0 * 5, *plm = a.loc_mask + row0 * 5;
    // Most warps hold no positive anchor at all (99 % of the anchors are not matched): their 32 x VEC x 5 zeros of
    // loc_target and loc_mask are one contiguous block each, written as full 128-bit lines by consecutive lanes instead
    // of 40-byte pieces per lane (a fifth of the sectors through L1 / TEX).  `zero_warp` is set before the divergent
    // part: the whole warp is active, lies inside the image and its block starts 16-byte aligned.
    if (zero_warp) {
      const size_t wrow0 = (size_t)b * A + (i0 - (int)lane_id() * VEC);  // first anchor of the warp
      float *zt = a.loc_target + wrow0 * 5, *zm = a.loc_mask + wrow0 * 5;
      const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
      constexpr int kVec4 = 32 * VEC * 5 / 4;  // float4 per warp and tensor
#pragma unroll
      for (int q = (int)lane_id(); q < kVec4; q += 32) {
        st_stream_f4(zt + 4 * q, z4);
        st_stream_f4(zm + 4 * q, z4);
      }
    }
    if constexpr (VEC == 4) {
      if (!zero_warp) {
#pragma unroll
        for (int q = 0; q < 5; ++q) {
          st_stream_f4(plt + 4 * q, make_float4(lt[4 * q], lt[4 * q + 1], lt[4 * q + 2], lt[4 * q + 3]));
          st_stream_f4(plm + 4 * q, make_float4(lm[4 * q], lm[4 * q + 1], lm[4 * q + 2], lm[4 * q + 3]));
        }
      }
      *reinterpret_cast<float4 *>(a.cls_target + row0) = make_float4(ct[0], ct[1], ct[2], ct[3]);
      *reinterpret_cast<uint4 *>(a.key + row0) = make_uint4(key[0], key[1], key[2], key[3]);
      if (a.match_out)
        *reinterpret_cast<int4 *>(a.match_out + row0) =
            make_int4(pos[0] ? best_k[0] : -1, pos[1] ? best_k[1] : -1, pos[2] ? best_k[2] : -1, pos[3] ? best_k[3] : -1);
    } else if constexpr (VEC == 2) {
      if (!zero_warp) {
#pragma unroll
        for (int q = 0; q < 5; ++q) {
          reinterpret_cast<float2 *>(plt)[q] = make_float2(lt[2 * q], lt[2 * q + 1]);
          reinterpret_cast<float2 *>(plm)[q] = make_float2(lm[2 * q], lm[2 * q + 1]);
        }
      }
      *reinterpret_cast<float2 *>(a.cls_target + row0) = make_float2(ct[0], ct[1]);
      *reinterpret_cast<uint2 *>(a.key + row0) = make_uint2(key[0], key[1]);
      if (a.match_out) *reinterpret_cast<int2 *>(a.match_out + row0) = make_int2(pos[0] ? best_k[0] : -1, pos[1] ? best_k[1] : -1);
    } else {
#pragma unroll
      for (int c = 0; c < 5; ++c) {
        plt[c] = lt[c];
        plm[c] = lm[c];
      }
      a.cls_target[row0] = ct[0];
      a.key[row0] = key[0];
      if (a.match_out) a.match_out[row0] = pos[0] ? best_k[0] : -1;
    }
  }

  // ---- publish the CTA's column maxima and positive count ----
  npos = warp_sum_i32(npos);
  if (lane_id() == 0 && npos) atomicAdd(&sm_pos, npos);
  __syncthreads();
  if (threadIdx.x == 0) a.thr_count[(size_t)b * a.T + t] = sm_pos;
  unsigned long long *gcol = a.colbest + ((size_t)b * a.T + t) * a.L;
  for (int k = threadIdx.x; k < G; k += blockDim.x) gcol[k] = sm_col[k];
  DSPMB_SSTAMP(6);
}

// ----------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long block_max_u64(unsigned long long v, unsigned long long *smem) {
  v = warp_max_u64(v);
  const unsigned w = warp_id(), l = lane_id(), nw = blockDim.x >> 5;
  __syncthreads();
  if (l == 0) smem[w] = v;
  __syncthreads();
  if (w == 0) {
    unsigned long long x = l < nw ? smem[l] : 0ull;
    x = warp_max_u64(x);
    if (l == 0) smem[0] = x;
  }
  __syncthreads();
  return smem[0];
}

enum { kStateDone = 0, kStateRecompute = 1 };

template <bool kKeysInSmem>
__global__ void __launch_bounds__(kMatchThreads) target_match_kernel(const __grid_constant__ TargetArgs a) {
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  __shared__ unsigned hist[256];
  __shared__ int sm_state, sm_arg, sm_nmatch, sm_dup, sm_carry, sm_thr;
  __shared__ unsigned sm_prefix;
  __shared__ int sm_need;
  constexpr int kLowBuckets = 2 * kMatchThreads, kPivotList = 1024;
  constexpr int kSampleKeys = 2 * kMatchThreads;
  constexpr int kSampleShift = 19;  // sample buckets = float bits >> 19: 16 per octave, p <= 1 stays below bucket 2033
  __shared__ unsigned bucket[kLowBuckets], plist[kPivotList];
  __shared__ int scan_smem[kMatchThreads / 32 + 1];
  __shared__ int sm_np, sm_pb, sm_before;
  __shared__ unsigned sm_kmin, sm_kmax;
  __shared__ unsigned wmin_smem[kMatchThreads / 32];

  const int b = blockIdx.x;
  const int A = a.A, L = a.L, W = a.W;
  // The kernel is a PROGRAMMATIC dependent of the stream kernel (which triggers on entry): its CTAs become resident
  // while the last stream CTAs drain, and everything up to griddepcontrol.wait below touches only the operator's
  // INPUTS (labels) and shared memory -- the valid-gt count is recomputed from the labels (every warp on its own, as
  // in the stream kernel) instead of being read back from the stream kernel's gcount[].
  const float *lab = a.labels + (size_t)b * L * W;
  const int G = count_valid_gt_warp(lab, L, W);
  int32_t *stats = a.stats_out ? a.stats_out + 4 * b : nullptr;
  // dynamic smem: [gt: L float4][col: L u64][m_anchor: L int][m_gt: L int][ord: L int][slist: L int][done: L u8 (padded)][bits: ceil(A/32) u32]
  float4 *sm_gt = reinterpret_cast<float4 *>(dyn_smem);
  unsigned long long *sm_col = reinterpret_cast<unsigned long long *>(sm_gt + L);
  int *m_anchor = reinterpret_cast<int *>(sm_col + L);
  int *m_gt = m_anchor + L;
  int *ord = m_gt + L;  // gts with a candidate, sorted by cached key (sequential bipartite path)
  int *slist = ord + ((L + 3) & ~3);  // gts whose cached maximum has gone stale (batched recompute), then sort scratch
  unsigned char *done = reinterpret_cast<unsigned char *>(slist + ((L + 3) & ~3));  // keeps `skeys` 16-byte aligned
  unsigned *bits = reinterpret_cast<unsigned *>(done + ((L + 15) / 16) * 16);
  const int nwords = (A + 31) / 32;
  unsigned *skeys = bits + ((nwords + 3) & ~3);  // [A] staged mining keys (kKeysInSmem)

  const float4 *anchors = reinterpret_cast<const float4 *>(a.anchors);
  for (int k = threadIdx.x; k < G; k += blockDim.x) {
    const float *row = lab + (size_t)k * W;
    sm_gt[k] = make_float4(row[1], row[2], row[3], row[4]);
    sm_col[k] = 0ull;
    done[k] = 0;
  }
  for (int w = threadIdx.x; w < nwords; w += blockDim.x) bits[w] = 0u;
  if (threadIdx.x == 0) {
    sm_nmatch = 0;
    sm_dup = 0;
    sm_thr = 0;
    sm_pb = -1;              // sample bucket of the shortlist bound
  }
  for (int i = threadIdx.x; i < kLowBuckets; i += blockDim.x) bucket[i] = 0u;  // sample histogram of the shortlist
  __syncthreads();  // shared-memory initialisation complete -- still in front of the wait, where it costs nothing
  asm volatile("griddepcontrol.wait;" ::: "memory");  // the stream kernel's outputs are complete and visible from here
  DSPMB_TSTAMP(0);
  if (G == 0) {  // multibox_target.cc:107 -- outputs stay at their initial values
    if (threadIdx.x == 0) {
      if (a.gcount[a.B + b] != DSPMB_OK) atomicMin(&a.header->status, a.gcount[a.B + b]);
      if (stats) stats[0] = stats[1] = stats[2] = stats[3] = 0;
    }
    return;
  }
  const unsigned *gkeys = a.key + (size_t)b * A;
  const bool async_keys = kKeysInSmem && a.mining_ratio > 0.f && (A & 3) == 0 && (reinterpret_cast<uintptr_t>(gkeys) & 15) == 0;
  {  // column maxima and positive counts of the image's tiles (each written by its own stream CTA, no atomics there)
    const unsigned long long *tcol = a.colbest + (size_t)b * a.T * L;
    // One warp per gt, lanes over the tiles, warp-wide maximum: no shared-memory atomics (a 64-bit atomicMax is a CAS
    // loop, and with one (tile, gt) pair per thread ~96 of them hit every sm_col[k]).
    const int nw = (int)(blockDim.x >> 5);
    for (int k = (int)warp_id(); k < G; k += nw) {
      unsigned long long v = 0ull;
#pragma unroll 4
      for (int t = (int)lane_id(); t < a.T; t += 32) {
        const unsigned long long ck = tcol[(size_t)t * L + k];
        v = ck > v ? ck : v;
      }
      v = warp_max_u64(v);
      if (lane_id() == 0) sm_col[k] = v;
    }
    if ((int)warp_id() == nw - 1) {
      int pos = 0;
      for (int t = (int)lane_id(); t < a.T; t += 32) pos += a.thr_count[(size_t)b * a.T + t];
      pos = warp_sum_i32(pos);
      if (lane_id() == 0) sm_thr = pos;
    }
  }
  // The mining keys are only needed after the matching, but their 100 KB take two microseconds to arrive: start the
  // copy into shared memory now (cp.async: no registers, nobody waits) and let it fly beside the bipartite stage; the
  // fix-up below patches the matched anchors' keys in the staged copy.  Issued BEHIND the column loads above: those
  // are the critical path, and six 16-byte copies per thread in front of them kept them waiting in the load queue.
  if (async_keys) {
    for (int j = threadIdx.x; j < (A >> 2); j += blockDim.x)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(skeys + 4 * j)),
                   "l"(gkeys + 4 * j)
                   : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  if (threadIdx.x == (blockDim.x >> 1) && a.gcount[a.B + b] != DSPMB_OK) atomicMin(&a.header->status, a.gcount[a.B + b]);
  __syncthreads();
  DSPMB_TSTAMP(1);

  // ---- bipartite stage (multibox_target.cc:113-149) ----
  // Fast path: if the cached best anchors of the gts are pairwise distinct, the greedy loop never meets a stale
  // maximum -- it visits the gts in descending key order and takes each one's cached anchor -- so all gts with a
  // candidate (key != 0) are assigned at once.  Otherwise the sequential lazy-greedy loop below runs.
  {
    int clash = 0;
    for (int k = threadIdx.x; k < G; k += blockDim.x) {
      const unsigned long long ck = sm_col[k];
      if (ck == 0ull) continue;
      const unsigned mine = (unsigned)(ck & 0xffffffffull);
      for (int kk = 0; kk < G; ++kk)
        if (kk != k && sm_col[kk] != 0ull && (unsigned)(sm_col[kk] & 0xffffffffull) == mine) clash = 1;
    }
    const bool distinct = __syncthreads_or(clash) == 0;
    if (distinct) {
      if (warp_id() == 0) {  // ordered compaction by ballots: no queue on one shared-memory counter
        int base = 0;
        for (int k0 = 0; k0 < G; k0 += 32) {
          const int k = k0 + (int)lane_id();
          const unsigned long long ck = k < G ? sm_col[k] : 0ull;
          const unsigned m = __ballot_sync(kFullMask, ck != 0ull);
          if (ck != 0ull) {
            const int j = (int)(0xffffffffu - (unsigned)(ck & 0xffffffffull));
            const int n = base + __popc(m & ((1u << lane_id()) - 1u));
            m_anchor[n] = j;
            m_gt[n] = k;
            atomicOr(&bits[j >> 5], 1u << (j & 31));
          }
          base += __popc(m);
        }
        if (lane_id() == 0) sm_nmatch = base;
      }
      __syncthreads();
    }
    if (!distinct) {
      // Sequential lazy greedy on a SORTED list: the gts with a candidate, ordered by cached key (descending; equal
      // keys -> lower gt index first, the reference's scan order).  The head of the list is the global maximum of
      // the cached keys.  If its anchor is still free the pair is final (cached keys are upper bounds of the true
      // column maxima over the unmatched anchors); otherwise the column is recomputed by the whole CTA and the gt is
      // re-inserted further down.  One thread walks the list, so a step without a recompute costs a few shared
      // memory accesses instead of a warp-wide arg-max over all gts.
      if (threadIdx.x == 0) sm_arg = 0;
      __syncthreads();
      for (int k = threadIdx.x; k < G; k += blockDim.x) {
        const unsigned long long ck = sm_col[k];
        if (ck == 0ull) continue;
        int r = 0;
        for (int kk = 0; kk < G; ++kk) {
          const unsigned long long o = sm_col[kk];
          r += (o > ck || (o == ck && kk < k)) ? 1 : 0;
        }
        ord[r] = k;
        atomicAdd(&sm_arg, 1);
      }
      __syncthreads();
      int n_ord = sm_arg;  // keys that are 0 rank behind every candidate and are not listed
      int head = 0;        // both are kept identical in every thread
      __syncthreads();     // everyone has read sm_arg before thread 0 reuses it below
      while (true) {
        if (threadIdx.x == 0) {
          int state = kStateDone;
          while (head < n_ord) {
            const int k = ord[head];
            const unsigned long long ck = sm_col[k];
            const int j = (int)(0xffffffffu - (unsigned)(ck & 0xffffffffull));
            if ((bits[j >> 5] >> (j & 31)) & 1u) {
              state = kStateRecompute;  // cached maximum points at an anchor that has been taken since
              break;
            }
            const int n = sm_nmatch;
            m_anchor[n] = j;
            m_gt[n] = k;
            sm_nmatch = n + 1;
            bits[j >> 5] |= 1u << (j & 31);
            ++head;
          }
          sm_state = state;
          sm_arg = head;
        }
        __syncthreads();
        if (sm_state == kStateDone) break;
        head = sm_arg;
        // Batched recompute.  Not only the head gt: EVERY listed gt whose cached best anchor has been taken by now is
        // stale (gts crowd around the same anchors, so they go stale together -- with 200 gts one at a time meant ~25
        // passes over the anchor table, 5 us each).  One pass over the unmatched anchors refreshes all of them: the
        // stale gts are listed, their keys reset, and every anchor is tried against the list (overlap pre-test, IoU,
        // shared-memory atomicMax only when it improves the column).
        if (threadIdx.x == 0) sm_dup = 0;
        __syncthreads();
        for (int pos = head + (int)threadIdx.x; pos < n_ord; pos += blockDim.x) {
          const int k = ord[pos];
          const int j = (int)(0xffffffffu - (unsigned)(sm_col[k] & 0xffffffffull));
          if ((bits[j >> 5] >> (j & 31)) & 1u) {
            slist[atomicAdd(&sm_dup, 1)] = k;
            sm_col[k] = 0ull;
          }
        }
        __syncthreads();
        const int nstale = sm_dup;
        // four anchors per thread in flight: the anchor table lives in L2, one dependent load per iteration would
        // put its latency on this loop twelve to twenty-four times
        for (int j0 = threadIdx.x; j0 < A; j0 += 4 * blockDim.x) {
          float4 an4[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int j = j0 + u * (int)blockDim.x;
            an4[u] = j < A ? __ldg(anchors + j) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int j = j0 + u * (int)blockDim.x;
            if (j >= A || ((bits[j >> 5] >> (j & 31)) & 1u)) continue;
            const float4 an = an4[u];
            for (int si = 0; si < nstale; ++si) {
              const int k = slist[si];
              const float4 g = sm_gt[k];
              // disjoint boxes have inter == 0 (or NaN), never > 1e-6: skip the IoU arithmetic and its division
              if (!(an.z > g.x && g.z > an.x && an.w > g.y && g.w > an.y)) continue;
              const float iou = iou_target(an, g);
              if (iou > 1e-6f) {
                const unsigned long long ck = col_key(iou, j);
                if (ck > sm_col[k]) atomicMax(&sm_col[k], ck);
              }
            }
          }
        }
        __syncthreads();
        // re-rank the rest of the list by (key descending, gt ascending); gts without a remaining candidate drop out
        const int nrem = n_ord - head;
        if (threadIdx.x == 0) sm_dup = 0;
        __syncthreads();
        for (int i = threadIdx.x; i < nrem; i += blockDim.x) {
          const int k = ord[head + i];
          const unsigned long long ck = sm_col[k];
          if (ck == 0ull) continue;
          int r = 0;
          for (int x = 0; x < nrem; ++x) {
            const int kk = ord[head + x];
            const unsigned long long o = sm_col[kk];
            r += (o > ck || (o == ck && kk < k)) ? 1 : 0;
          }
          slist[r] = k;  // the stale list is dead: scratch for the new order
          atomicAdd(&sm_dup, 1);
        }
        __syncthreads();
        const int nkeep = sm_dup;
        for (int i = threadIdx.x; i < nkeep; i += blockDim.x) ord[head + i] = slist[i];
        n_ord = head + nkeep;
        __syncthreads();
      }
      if (threadIdx.x == 0) sm_dup = 0;
      __syncthreads();
    }
  }
  const int nmatch = sm_nmatch;
  DSPMB_TSTAMP(2);
  if (async_keys) {
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();  // every thread's share of the staged keys is in shared memory before the fix-up patches it
  }

  // ---- fix up the bipartite-matched anchors (they override whatever the threshold stage wrote) ----
  // a group of lanes per matched anchor (16, or 4 when there are more than 64 matches so that one round covers
  // them): the G IoUs of its row are spread over the group
  const int gs = nmatch > (int)(blockDim.x >> 4) ? 4 : 16;
  for (int q = (int)threadIdx.x / gs; q < nmatch; q += (int)blockDim.x / gs) {
    const int sub = threadIdx.x & (gs - 1);
    const int j = m_anchor[q], k = m_gt[q];
    const float4 an = __ldg(anchors + j);
    float max_iou = -1.0f;
    for (int kk = sub; kk < G; kk += gs) {
      const float4 g = sm_gt[kk];
      // a disjoint gt has IoU 0 (G > 0, so the row maximum is at least that): no arithmetic, no division
      const bool reach = an.z > g.x && g.z > an.x && an.w > g.y && g.w > an.y;
      const float iou = reach ? iou_target(an, g) : 0.0f;
      if (iou > max_iou) max_iou = iou;
    }
    const unsigned gmask = (gs == 16 ? 0xffffu : 0xfu) << (lane_id() & ~(unsigned)(gs - 1));  // this anchor's group
    for (int m = gs >> 1; m > 0; m >>= 1) {
      const float o = __shfl_xor_sync(gmask, max_iou, m);
      if (o > max_iou) max_iou = o;
    }
    if (sub != 0) continue;
    done[q] = (a.overlap_threshold > 0.f && max_iou > a.overlap_threshold) ? 1 : 0;  // counted by the stream kernel
    const size_t row = (size_t)b * A + j;
    const float *lrow = lab + (size_t)k * W;
    float enc[5];
    encode_loc(an, lrow, a, enc);
#pragma unroll
    for (int c = 0; c < 5; ++c) {
      a.loc_target[row * 5 + c] = enc[c];
      a.loc_mask[row * 5 + c] = 1.0f;
    }
    a.cls_target[row] = fadd(lrow[0], 1.0f);
    a.key[row] = kKeySentinel;
    if (async_keys) skeys[j] = kKeySentinel;
    if (a.match_out) a.match_out[row] = k;
  }
  __syncthreads();
  DSPMB_TSTAMP(3);
  int ndup = 0;  // matched anchors the threshold stage had already counted (flags instead of a contended counter)
  for (int q = (int)lane_id(); q < nmatch; q += 32) ndup += done[q];
  ndup = warp_sum_i32(ndup);
  const int num_positive = sm_thr + nmatch - ndup;

  // ---- hard-negative mining (multibox_target.cc:182-241) ----
  int num_negative = 0;
  if (a.mining_ratio > 0.f) {
    num_negative = (int)fmul((float)num_positive, a.mining_ratio);
    if (num_negative > A - num_positive) num_negative = A - num_positive;
    if (num_negative < 0) num_negative = 0;
  }
  if (stats && threadIdx.x == 0) {
    stats[0] = G;
    stats[1] = num_positive;
    stats[2] = a.mining_ratio > 0.f ? num_negative : A - num_positive;
    stats[3] = nmatch;
  }
  if (num_negative <= 0) return;

  // Stage the keys in shared memory (the fix-up above has already replaced the matched anchors' keys) unless the
  // asynchronous copy has done it.
  if (kKeysInSmem && !async_keys) {
    if ((A & 3) == 0 && (reinterpret_cast<uintptr_t>(gkeys) & 15) == 0) {  // 128-bit loads, all in flight at once
      const uint4 *g4 = reinterpret_cast<const uint4 *>(gkeys);
      uint4 *s4 = reinterpret_cast<uint4 *>(skeys);
#pragma unroll 8
      for (int j = threadIdx.x; j < (A >> 2); j += blockDim.x) s4[j] = g4[j];
    } else {
#pragma unroll 8
      for (int j = threadIdx.x; j < A; j += blockDim.x) skeys[j] = gkeys[j];
    }
  }
  __syncthreads();  // the staged keys are read by other threads from here on (racecheck: staging vs. the min/max pass)
  DSPMB_TSTAMP(4);

  // ---- shortlist ----
  // The wanted keys are the num_negative SMALLEST of A (a per cent or two of the image), yet the selection below used
  // to walk all A keys four times (range, histogram, pivot bucket, final pass): 60 % of this kernel's instructions on
  // a single SM.  Instead: a 2048-key sample of the staged keys gives an upper bound U of the pivot (the sample's
  // quantile of num_negative/A plus four standard deviations, rounded up to a bucket edge), ONE pass over the keys
  // compacts everything <= U into a shortlist (key, anchor), and range / histogram / pivot / final pass then run on
  // the shortlist alone.  U is only a guess -- the count of the shortlist verifies it: fewer than num_negative
  // entries, an overflow of the list, or (below) an ambiguity band that reaches beyond U fall back to the full walk.
  unsigned *sl_key = skeys + ((A + 3) & ~3);                       // [kShortCap] (kKeysInSmem only)
  int *sl_idx = reinterpret_cast<int *>(sl_key + kShortCap);       // [kShortCap]
  unsigned U = 0u, kmin_short = 0u;
  int n_short = 0;
  bool use_short = false;
  if (kKeysInSmem && a.shortlist && (A & 3) == 0 && A >= kShortMinAnchors) {
    // Sample histogram on the float bits themselves (sign 0, 8 exponent and 4 mantissa bits: 16 buckets per octave of
    // the probability, 2032 buckets up to p = 1): no range pass, and the bound only has to be roughly right.
    // bucket[] was zeroed at kernel start.
    const float m = (float)num_negative * (float)kSampleKeys / (float)A;  // expected sample keys below the pivot
    const int r_s = (int)(m + 4.0f * sqrtf(m) + 4.0f) + 1;
    for (int s = threadIdx.x; s < kSampleKeys; s += blockDim.x) {
      const unsigned kv = skeys[(int)(((long long)s * A) / kSampleKeys)];
      if (kv != kKeySentinel) atomicAdd(&bucket[min(kv >> kSampleShift, (unsigned)kLowBuckets - 1u)], 1u);
    }
    __syncthreads();
    const unsigned c0 = bucket[2 * threadIdx.x], c1 = bucket[2 * threadIdx.x + 1];
    bucket[2 * threadIdx.x] = 0u;  // ready for the selection's histogram (nobody else touches these two)
    bucket[2 * threadIdx.x + 1] = 0u;
    int total;
    const int ex = block_scan_excl((int)(c0 + c1), scan_smem, &total);
    if (ex < r_s && ex + (int)c0 >= r_s) sm_pb = 2 * threadIdx.x;
    else if (ex + (int)c0 < r_s && ex + (int)(c0 + c1) >= r_s) sm_pb = 2 * threadIdx.x + 1;
    __syncthreads();
    const int P = sm_pb;  // -1 (set at kernel start): the sample holds fewer than r_s candidates
    DSPMB_TSTAMP(8);
    if (P >= 0 && P < kLowBuckets - 1) {
      U = (((unsigned)P + 1u) << kSampleShift) - 1u;  // upper edge of the bucket; < 2^30, far below the sentinel
      // Compaction: every thread counts its hits first (no communication), ONE warp scan and ONE shared-memory atomic
      // per warp reserve the slots, then the thread walks its keys again and stores.  (A scan + atomic per 128 keys
      // made this pass a 2.7 us chain of dependent shuffles and atomics.)  The smallest key of the image falls out.
      const uint4 *s4 = reinterpret_cast<const uint4 *>(skeys);
      const int n4 = A >> 2;
      int c = 0;
      unsigned mymin = kKeySentinel;
#pragma unroll 3
      for (int j = threadIdx.x; j < n4; j += blockDim.x) {
        const uint4 kv = s4[j];
        c += (kv.x <= U ? 1 : 0) + (kv.y <= U ? 1 : 0) + (kv.z <= U ? 1 : 0) + (kv.w <= U ? 1 : 0);
        mymin = min(min(mymin, kv.x), min(min(kv.y, kv.z), kv.w));
      }
      const int incl = warp_scan_incl(c);
      mymin = __reduce_min_sync(kFullMask, mymin);
      // warp totals through shared memory and one redundant 32-wide scan per warp: 32 warps queueing on one
      // shared-memory atomic (and a second one for the minimum) cost more than the pass itself
      if (lane_id() == 31) {
        scan_smem[warp_id()] = incl;
        wmin_smem[warp_id()] = mymin;
      }
      __syncthreads();
      const int nw = (int)(blockDim.x >> 5);  // <= 32
      const int wt = (int)lane_id() < nw ? scan_smem[lane_id()] : 0;
      const int wincl = warp_scan_incl(wt);
      const int wbase = __shfl_sync(kFullMask, wincl - wt, (int)warp_id());
      n_short = __shfl_sync(kFullMask, wincl, 31);
      kmin_short = __reduce_min_sync(kFullMask, (int)lane_id() < nw ? wmin_smem[lane_id()] : kKeySentinel);
      int pos = wbase + incl - c;
      if (c) {
#pragma unroll 3
        for (int j = threadIdx.x; j < n4; j += blockDim.x) {
          const uint4 kv = s4[j];
          if (kv.x <= U) {
            if (pos < kShortCap) sl_key[pos] = kv.x, sl_idx[pos] = 4 * j;
            ++pos;
          }
          if (kv.y <= U) {
            if (pos < kShortCap) sl_key[pos] = kv.y, sl_idx[pos] = 4 * j + 1;
            ++pos;
          }
          if (kv.z <= U) {
            if (pos < kShortCap) sl_key[pos] = kv.z, sl_idx[pos] = 4 * j + 2;
            ++pos;
          }
          if (kv.w <= U) {
            if (pos < kShortCap) sl_key[pos] = kv.w, sl_idx[pos] = 4 * j + 3;
            ++pos;
          }
        }
      }
      __syncthreads();
      use_short = n_short >= num_negative && n_short <= kShortCap;
      DSPMB_TSTAMP(9);
    }
  }

  float *ct = a.cls_target + (size_t)b * A;
  int *amb_list = a.amb_list + (size_t)b * A;
  unsigned *amb_key = a.amb_key + (size_t)b * A;
  for (int attempt = use_short ? 0 : 1; attempt < 2; ++attempt) {
    const bool sh = attempt == 0;                                          // CTA-uniform
    const unsigned *mk = sh ? sl_key : (kKeysInSmem ? skeys : gkeys);     // the keys this attempt selects from
    const int mn = sh ? n_short : A;
    __syncthreads();  // (second attempt: everybody has left the first one)
    // MSB-first radix select of the num_negative-th smallest key.  Non-candidates carry the sentinel 0xffffffff
    // (larger than the bits of any probability), so they are only reached if there are too few candidates.
    if (threadIdx.x == 0) {
      sm_prefix = 0u;
      sm_need = num_negative;
    }
    // Fast path: one 2048-bucket histogram over the range the candidate keys actually span (the bit patterns of
    // positive floats order like the floats) replaces the four 8-bit radix passes: the bucket that holds the
    // num_negative-th smallest key follows from a prefix sum, and the key itself from ranking the members of that one
    // bucket.  A bucket too crowded to rank (massive ties) or too few candidates leave it to the radix passes below.
    bool have_q = false;
    if (kKeysInSmem) {
      unsigned kmin = kKeySentinel, kmax = 0u;
      if (sh) {
        // range known from the compaction (smallest key of the image .. U); its per-thread re-zeroing of bucket[]
        // and the barriers since then stand in for the zeroing below
        kmin = kmin_short;
        kmax = U;
        if (threadIdx.x == 0) sm_np = 0;
      } else {
        for (int j = threadIdx.x; j < mn; j += blockDim.x) {
          const unsigned kv = mk[j];
          if (kv != kKeySentinel) {
            kmin = min(kmin, kv);
            kmax = max(kmax, kv);
          }
        }
        kmin = __reduce_min_sync(kFullMask, kmin);
        kmax = __reduce_max_sync(kFullMask, kmax);
        if (threadIdx.x == 0) {
          sm_kmin = kKeySentinel;
          sm_kmax = 0u;
          sm_np = 0;
        }
        for (int i = threadIdx.x; i < kLowBuckets; i += blockDim.x) bucket[i] = 0u;
        __syncthreads();
        if (lane_id() == 0) {
          atomicMin(&sm_kmin, kmin);
          atomicMax(&sm_kmax, kmax);
        }
        __syncthreads();
        kmin = sm_kmin;
        kmax = sm_kmax;
      }
      if (kmin != kKeySentinel) {  // CTA-uniform: at least one candidate
        // smallest shift with ((kmax - kmin) >> shift) < kLowBuckets (= 2^11): bit length of the range minus 11
        const int shift = max(0, 32 - __clz((int)(kmax - kmin)) - 11);
        static_assert(kLowBuckets == 2048, "shift formula assumes 2^11 buckets");
        for (int j = threadIdx.x; j < mn; j += blockDim.x) {
          const unsigned kv = mk[j];
          if (kv != kKeySentinel) atomicAdd(&bucket[(kv - kmin) >> shift], 1u);
        }
        __syncthreads();
        const unsigned c0 = bucket[2 * threadIdx.x], c1 = bucket[2 * threadIdx.x + 1];
        int total;
        const int ex = block_scan_excl((int)(c0 + c1), scan_smem, &total);  // total = number of candidates
        if (ex < num_negative && ex + (int)c0 >= num_negative) {
          sm_pb = 2 * threadIdx.x;
          sm_before = ex;
        } else if (ex + (int)c0 < num_negative && ex + (int)(c0 + c1) >= num_negative) {
          sm_pb = 2 * threadIdx.x + 1;
          sm_before = ex + (int)c0;
        }
        __syncthreads();
        DSPMB_TSTAMP(10);
        if (num_negative <= total) {  // otherwise the radix path reports the missing candidates
          const unsigned P = (unsigned)sm_pb;
          const int r = num_negative - sm_before;  // 1-based rank of the wanted key inside bucket P
          const int cP = (int)bucket[P];
          if (cP <= kPivotList) {
            for (int j = threadIdx.x; j < mn; j += blockDim.x) {
              const unsigned kv = mk[j];
              if (kv != kKeySentinel && ((kv - kmin) >> shift) == P) plist[atomicAdd(&sm_np, 1)] = kv;
            }
            __syncthreads();
            for (int i = threadIdx.x; i < cP; i += blockDim.x) {
              const unsigned v = plist[i];
              int lo = 0, hi = 0;
              for (int x = 0; x < cP; ++x) {
                lo += plist[x] < v ? 1 : 0;
                hi += plist[x] <= v ? 1 : 0;
              }
              if (lo < r && r <= hi) sm_prefix = v;  // equal keys write the same value
            }
            __syncthreads();
            have_q = true;
          }
        }
      }
    }
    for (int pass = 0; pass < (have_q ? 0 : 4); ++pass) {
      const int shift = 24 - 8 * pass;
      const unsigned mask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
      for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0u;
      __syncthreads();
      const unsigned prefix = sm_prefix;
      // warp-aggregated histogram update (probabilities cluster, so plain atomics would serialise)
      auto hist_add = [&](unsigned kv, bool in_range) {
        const bool hit = in_range && (kv & mask) == prefix;
        const unsigned active = __ballot_sync(kFullMask, hit);
        if (active == 0u) return;
        // the digit of the first hit lane is counted for the whole warp with one ballot (probabilities cluster: in the
        // first pass that is nearly every lane); the other lanes add themselves, and their digits rarely coincide
        const unsigned digit = (kv >> shift) & 0xffu;
        const int leader = __ffs(active) - 1;
        const unsigned d0 = __shfl_sync(kFullMask, digit, leader);
        const unsigned same = __ballot_sync(kFullMask, hit && digit == d0);
        if ((int)lane_id() == leader) atomicAdd(&hist[d0], (unsigned)__popc(same));
        else if (hit && digit != d0) atomicAdd(&hist[digit], 1u);
      };
      if (kKeysInSmem && !sh && (A & 3) == 0) {  // four keys per thread and iteration: one LDS.128, four independent updates
        const uint4 *s4 = reinterpret_cast<const uint4 *>(mk);
        const int n4 = A >> 2;
        for (int base = 0; base < n4; base += blockDim.x) {
          const int j = base + threadIdx.x;
          const bool in = j < n4;
          const uint4 kv = in ? s4[j] : make_uint4(0u, 0u, 0u, 0u);
          hist_add(kv.x, in);
          hist_add(kv.y, in);
          hist_add(kv.z, in);
          hist_add(kv.w, in);
        }
      } else {
        for (int base = 0; base < mn; base += blockDim.x) {
          const int j = base + threadIdx.x;
          hist_add(j < mn ? mk[j] : kKeySentinel, j < mn);
        }
      }
      __syncthreads();
      if (warp_id() == 0) {  // digit search: 8 bins per lane + warp scan
        const unsigned lane = lane_id();
        unsigned h[8];
        int tot = 0;
#pragma unroll
        for (int d = 0; d < 8; ++d) {
          h[d] = hist[lane * 8 + d];
          tot += (int)h[d];
        }
        const int incl = warp_scan_incl(tot);
        const int excl = incl - tot;
        const int need = sm_need;  // 1 <= need <= number of keys matching the prefix
        __syncwarp();  // every lane has read sm_need before one lane overwrites it
        if (excl < need && incl >= need) {
          int acc = excl;
#pragma unroll
          for (int d = 0; d < 8; ++d) {
            if (acc + (int)h[d] >= need) {
              sm_need = need - acc;
              sm_prefix = prefix | ((unsigned)(lane * 8 + d) << shift);
              break;
            }
            acc += (int)h[d];
          }
        }
      }
      __syncthreads();
    }
    DSPMB_TSTAMP(5);
    const unsigned qkey = sm_prefix;  // the num_negative-th smallest APPROXIMATE key
    if (qkey == kKeySentinel) {
      // the num_negative-th smallest key is a non-candidate: CHECK_GE(temp.size(), num_negative) fails
      // (multibox_target.cc:236)
      if (threadIdx.x == 0) atomicMin(&a.header->status, DSPMB_ERR_MINING_CANDIDATES);
      return;
    }
    // Keys q approximate the reference's probabilities p with |q - p| <= delta * p, so the exact pivot lies in
    // [Q/(1+delta), Q/(1-delta)]: keys below Q(1-3 delta) are certainly selected, keys above Q(1+3 delta) certainly
    // not, and only the band in between is re-evaluated exactly and ranked by (p, anchor) like the stable sort.
    const float Q = __uint_as_float(qkey);
    const unsigned klo = __float_as_uint(fmul(Q, 1.0f - 3.0f * a.delta));
    const unsigned khi = __float_as_uint(fmul(Q, 1.0f + 3.0f * a.delta));
    if (sh && khi > U) continue;  // the band reaches past the shortlist's bound: keys beyond U could belong to it
    if (threadIdx.x == 0) {
      sm_carry = 0;  // surely selected
      sm_dup = 0;    // ambiguous
    }
    __syncthreads();
    int n_in_local = 0;
    for (int base = 0; base < mn; base += blockDim.x) {
      const int i = base + threadIdx.x;
      const unsigned kv = i < mn ? mk[i] : kKeySentinel;
      const int j = (sh && i < mn) ? sl_idx[i] : i;
      if (kv < klo) {
        ct[j] = 0.0f;
        ++n_in_local;
      }
      const bool amb = kv >= klo && kv <= khi;
      const unsigned m = __ballot_sync(kFullMask, amb);
      if (m) {
        int wbase = 0;
        if (lane_id() == 0) wbase = atomicAdd(&sm_dup, __popc(m));
        wbase = __shfl_sync(kFullMask, wbase, 0);
        if (amb) amb_list[wbase + __popc(m & ((1u << lane_id()) - 1u))] = j;
      }
    }
    n_in_local = warp_sum_i32(n_in_local);
    if (lane_id() == 0) scan_smem[warp_id()] = n_in_local;  // per-warp partials, summed by every warp after the barrier
    __syncthreads();
    DSPMB_TSTAMP(6);
    const int n_amb = sm_dup;
    const int need = num_negative - warp_sum_i32(lane_id() < (blockDim.x >> 5) ? scan_smem[lane_id()] : 0);
    if (need < 0 || need > n_amb) {  // would mean the error bound was violated
      if (threadIdx.x == 0) atomicMin(&a.header->status, DSPMB_ERR_INTERNAL);
      return;
    }
    if (need == n_amb) {  // the whole band is selected (typically: the pivot is alone in it) -- nothing to rank
      for (int q = threadIdx.x; q < n_amb; q += blockDim.x) ct[amb_list[q]] = 0.0f;
      DSPMB_TSTAMP(7);
      return;
    }
    const float *p_cls = a.cls_preds + (size_t)b * a.C * A;
    if (n_amb <= 4 * (int)(blockDim.x >> 5)) {  // a handful: one warp each (latency); many: one thread each (throughput)
      for (int q = (int)warp_id(); q < n_amb; q += (int)(blockDim.x >> 5)) {
        const int j = amb_list[q];
        const float pe = a.fma_build ? exact_bg_prob_warp<true>(p_cls, j, A, a.C) : exact_bg_prob_warp<false>(p_cls, j, A, a.C);
        if (lane_id() == 0) amb_key[q] = __float_as_uint(pe);
      }
    } else {
      for (int q = threadIdx.x; q < n_amb; q += blockDim.x) {
        const int j = amb_list[q];
        const float pe = a.fma_build ? exact_bg_prob<true>(p_cls, j, A, a.C) : exact_bg_prob<false>(p_cls, j, A, a.C);
        amb_key[q] = __float_as_uint(pe);
      }
    }
    __syncthreads();
    for (int q = threadIdx.x; q < n_amb; q += blockDim.x) {
      const unsigned kq = amb_key[q];
      const int jq = amb_list[q];
      int rank = 0;
      for (int i = 0; i < n_amb; ++i) {
        const unsigned ki = amb_key[i];
        rank += (ki < kq || (ki == kq && amb_list[i] < jq)) ? 1 : 0;
      }
      if (rank < need) ct[jq] = 0.0f;
    }
    DSPMB_TSTAMP(7);
    return;
  }
}


// ----------------------------------------------------------------------------------------------------
// Cluster version of the matcher (opt-in, DSPMB_TUNE_TARGET_PIPELINE = 1; measured slower than the single-CTA kernel, see
// include/dspmb.h): one thread-block CLUSTER of kMatchCluster CTAs per image instead of
// one 1024-thread CTA.  The single-CTA kernel is pure latency (0.43 waves, four passes over the image's A mining keys
// by one CTA); here every CTA owns a slice of A / kMatchCluster anchors and their keys in its shared memory, and the
// CTAs talk through distributed shared memory (cluster.map_shared_rank) and the hardware cluster barrier (~0.2 us):
//   * reduction of the stream kernel's per-tile column maxima: tiles dealt round-robin, CTA 0 combines the partials;
//   * bipartite stage: CTA 0 runs the (mostly one-shot) greedy assignment; when a cached column maximum is stale, all
//     CTAs recompute that column over their anchor slices and hand their partial maxima to CTA 0;
//   * fix-up of the matched anchors by CTA 0, which also patches their keys in the owners' shared memory;
//   * mining: MSB-first radix select with 11-bit digits taken straight from the float bits of the probability
//     (p <= 1 keeps bit 30 clear, so bits 29..19 are digit one): every CTA histograms its slice, every CTA sums the
//     kMatchCluster histograms through DSMEM and finds the pivot bin redundantly (no broadcast round); as soon as the
//     pivot bin holds <= 256 keys they are pushed into CTA 0's list and ranked there; the final pass (selected /
//     ambiguous band) again runs on the slices, and CTA 0 re-evaluates the few ambiguous anchors bit-exactly.
constexpr int kMatchCluster = 8;
constexpr int kClusterThreads = 256;
constexpr int kClusterSliceMax = 12288;   // keys per CTA kept in shared memory (48 KB)
constexpr int kDigitBins = 2048;
constexpr int kPivotCap = 2048;   // pivot-bin keys ranked directly, one per thread of the cluster
static_assert(kPivotCap == kMatchCluster * kClusterThreads, "one pivot-bin key per thread of the cluster");

// Warp-aggregated histogram update of one key per lane: background probabilities cluster (most of an image's keys
// share one or two digits), so plain shared-memory atomics would serialise.  The digit of the first hit lane is
// counted for the whole warp with one ballot; the lanes with another digit add themselves.
__device__ __forceinline__ void hist_add_warp(unsigned *h, unsigned digit, bool hit) {
  const unsigned act = __ballot_sync(kFullMask, hit);
  if (act == 0u) return;
  const int leader = __ffs(act) - 1;
  const unsigned d0 = __shfl_sync(kFullMask, digit, leader);
  const unsigned same = __ballot_sync(kFullMask, hit && digit == d0);
  if ((int)lane_id() == leader) atomicAdd(&h[d0], (unsigned)__popc(same));
  else if (hit && digit != d0) atomicAdd(&h[digit], 1u);
}

struct ClusterCtl {
  int state, k, num_negative, err;
  unsigned qkey;
  int have_q;
};

template <bool kFma>
__global__ void __cluster_dims__(kMatchCluster, 1, 1) __launch_bounds__(kClusterThreads)
    target_match_cluster_kernel(const __grid_constant__ TargetArgs a) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  __shared__ unsigned long long red_smem[kClusterThreads / 32];
  __shared__ unsigned long long sm_part[kMatchCluster];
  __shared__ unsigned hist[2][kDigitBins], hsum[kDigitBins];
  __shared__ unsigned plist[kPivotCap];
  __shared__ int sm_npatch, sm_lcnt, sm_off;
  __shared__ int scan_smem[kClusterThreads / 32 + 1];
  __shared__ ClusterCtl ctl;
  __shared__ int sm_arg, sm_nmatch, sm_dup, sm_thr, sm_thr_part, sm_np, sm_carry, sm_amb, sm_pb, sm_before, sm_cp, sm_total;

  const int rank = (int)cluster.block_rank();
  const int b = blockIdx.x / kMatchCluster;
  const int A = a.A, L = a.L, W = a.W;
  if (rank == 0) DSPMB_TSTAMP_IMG(b, 0);
  const int G = a.gcount[b];
  int32_t *stats = a.stats_out ? a.stats_out + 4 * b : nullptr;
  if (rank == 0 && threadIdx.x == 0 && a.gcount[a.B + b] != DSPMB_OK) atomicMin(&a.header->status, a.gcount[a.B + b]);
  if (G == 0) {  // multibox_target.cc:107 -- outputs stay at their initial values (uniform over the cluster)
    if (rank == 0 && threadIdx.x == 0 && stats) stats[0] = stats[1] = stats[2] = stats[3] = 0;
    return;
  }
  // slices: CTA r owns anchors [r * slice, min(A, (r + 1) * slice)), slice a multiple of 32
  const int slice = (((A + kMatchCluster - 1) / kMatchCluster) + 31) & ~31;
  const int s0 = min(A, rank * slice), s1 = min(A, s0 + slice);
  const int nwords_all = (A + 31) / 32, nwords_slice = slice / 32;
  // dynamic smem (same layout in every CTA, so that DSMEM offsets coincide):
  //   [gt: L float4][col: L u64][colp: L u64][m_anchor: L int][m_gt: L int][ord: L int][bits_all: ceil(A/32) u32]
  //   [bits: slice/32 u32][skeys: slice u32][patch_bin: L u16][patch_owner: L u16]
  float4 *sm_gt = reinterpret_cast<float4 *>(dyn_smem);
  unsigned long long *sm_col = reinterpret_cast<unsigned long long *>(sm_gt + L);
  unsigned long long *sm_colp = sm_col + L;
  int *m_anchor = reinterpret_cast<int *>(sm_colp + L);
  int *m_gt = m_anchor + L;
  int *ord = m_gt + L;
  unsigned *bits_all = reinterpret_cast<unsigned *>(ord + ((L + 3) & ~3));
  unsigned *bits = bits_all + ((nwords_all + 3) & ~3);
  unsigned *skeys = bits + ((nwords_slice + 3) & ~3);
  unsigned short *patch_bin = reinterpret_cast<unsigned short *>(skeys + slice);  // matched anchors: first digit of their
  unsigned short *patch_owner = patch_bin + L;                                     // key and the CTA that held it
  const int kPatchCap = L;  // at most G <= L anchors are matched

  const float *lab = a.labels + (size_t)b * L * W;
  const float4 *anchors = reinterpret_cast<const float4 *>(a.anchors);
  const bool mining = a.mining_ratio > 0.f;
  // ---- phase 0 (every CTA): gts, this CTA's share of the tile reductions, its key slice ----
  for (int k = threadIdx.x; k < G; k += blockDim.x) {
    const float *row = lab + (size_t)k * W;
    sm_gt[k] = make_float4(row[1], row[2], row[3], row[4]);
    sm_col[k] = 0ull;
    sm_colp[k] = 0ull;
  }
  if (rank == 0)
    for (int w = threadIdx.x; w < nwords_all; w += blockDim.x) bits_all[w] = 0u;
  if (threadIdx.x == 0) {
    sm_nmatch = 0;
    sm_dup = 0;
    sm_thr = 0;
    sm_thr_part = 0;
    sm_np = 0;
    sm_carry = 0;
    sm_amb = 0;
    ctl.state = kStateDone;
    ctl.num_negative = 0;
    ctl.err = 0;
    ctl.have_q = 0;
    sm_npatch = 0;
    sm_lcnt = 0;
  }
  if (mining)
    for (int i = threadIdx.x; i < kDigitBins; i += blockDim.x) hist[0][i] = 0u;
  if (mining) {
    const unsigned *gkeys = a.key + (size_t)b * A;
    if ((A & 3) == 0 && (reinterpret_cast<uintptr_t>(gkeys) & 15) == 0) {
      const uint4 *g4 = reinterpret_cast<const uint4 *>(gkeys + s0);
      uint4 *s4 = reinterpret_cast<uint4 *>(skeys);
#pragma unroll 4
      for (int j = threadIdx.x; j < ((s1 - s0) >> 2); j += blockDim.x) s4[j] = g4[j];
    } else {
#pragma unroll 4
      for (int j = threadIdx.x; j < s1 - s0; j += blockDim.x) skeys[j] = gkeys[s0 + j];
    }
  }
  __syncthreads();
  if (mining) {
    // first digit (bits 29..19) of this slice's keys.  The histogram does not depend on num_negative, so it is built
    // here, off the critical path; the keys of the anchors the bipartite stage matches later are taken out again
    // through the patch list.
    for (int base = 0; base < s1 - s0; base += blockDim.x) {
      const int j = base + (int)threadIdx.x;
      const unsigned kv = j < s1 - s0 ? skeys[j] : kKeySentinel;
      hist_add_warp(hist[0], (kv >> 19) & (unsigned)(kDigitBins - 1), kv != kKeySentinel);
    }
  }
  {
    const unsigned long long *tcol = a.colbest + (size_t)b * a.T * L;
    const int ntl = (a.T - rank + kMatchCluster - 1) / kMatchCluster;  // tiles rank, rank + CS, ...
    for (int idx = threadIdx.x; idx < ntl * G; idx += blockDim.x) {
      const int tl = idx / G, k = idx - tl * G;
      const unsigned long long ck = tcol[(size_t)(rank + tl * kMatchCluster) * L + k];
      if (ck) atomicMax(&sm_colp[k], ck);
    }
    int pos = 0;
    for (int t = rank + (int)threadIdx.x * kMatchCluster; t < a.T; t += (int)blockDim.x * kMatchCluster)
      pos += a.thr_count[(size_t)b * a.T + t];
    pos = warp_sum_i32(pos);
    if (lane_id() == 0 && pos) atomicAdd(&sm_thr_part, pos);
  }
  cluster.sync();  // #1: partial maxima / counts, every key slice and its first-digit histogram are in place
  if (mining && rank != 0) {
    // cluster total of the first-digit histograms, while CTA 0 is matching (it copies the result afterwards)
    for (int i = threadIdx.x; i < kDigitBins; i += blockDim.x) {
      unsigned t = 0u;
#pragma unroll
      for (int r = 0; r < kMatchCluster; ++r) t += cluster.map_shared_rank(hist[0], r)[i];
      hsum[i] = t;
      if (rank == 1) hist[1][i] = t;  // the copy CTA 0 picks up (hsum itself is patched in place later)
    }
  }
  if (rank == 0) {
    for (int k = threadIdx.x; k < G; k += blockDim.x) {
      unsigned long long m = 0ull;
#pragma unroll
      for (int r = 0; r < kMatchCluster; ++r) {
        const unsigned long long o = *cluster.map_shared_rank(&sm_colp[k], r);
        m = o > m ? o : m;
      }
      sm_col[k] = m;
    }
    if (threadIdx.x < kMatchCluster) atomicAdd(&sm_thr, *cluster.map_shared_rank(&sm_thr_part, threadIdx.x));
    __syncthreads();
    DSPMB_TSTAMP_IMG(b, 1);
  }

  // ---- bipartite stage (multibox_target.cc:113-149): CTA 0 decides, everybody recomputes stale columns ----
  bool distinct = true;
  int n_ord = 0, head = 0;
  if (rank == 0) {
    int clash = 0;
    for (int k = threadIdx.x; k < G; k += blockDim.x) {
      const unsigned long long ck = sm_col[k];
      if (ck == 0ull) continue;
      const unsigned mine = (unsigned)(ck & 0xffffffffull);
      for (int kk = 0; kk < G; ++kk)
        if (kk != k && sm_col[kk] != 0ull && (unsigned)(sm_col[kk] & 0xffffffffull) == mine) clash = 1;
    }
    distinct = __syncthreads_or(clash) == 0;
    if (distinct) {
      // cached best anchors pairwise distinct: the greedy loop never meets a stale maximum, assign all at once
      for (int k = threadIdx.x; k < G; k += blockDim.x) {
        const unsigned long long ck = sm_col[k];
        if (ck == 0ull) continue;
        const int j = (int)(0xffffffffu - (unsigned)(ck & 0xffffffffull));
        const int n = atomicAdd(&sm_nmatch, 1);
        m_anchor[n] = j;
        m_gt[n] = k;
      }
      __syncthreads();
    } else {
      // sorted list of the gts with a candidate (cached key descending, lower gt first), walked by one thread
      if (threadIdx.x == 0) sm_arg = 0;
      __syncthreads();
      for (int k = threadIdx.x; k < G; k += blockDim.x) {
        const unsigned long long ck = sm_col[k];
        if (ck == 0ull) continue;
        int r = 0;
        for (int kk = 0; kk < G; ++kk) {
          const unsigned long long o = sm_col[kk];
          r += (o > ck || (o == ck && kk < k)) ? 1 : 0;
        }
        ord[r] = k;
        atomicAdd(&sm_arg, 1);
      }
      __syncthreads();
      n_ord = sm_arg;
      __syncthreads();
    }
  }
  while (true) {
    if (rank == 0) {
      if (!distinct && threadIdx.x == 0) {
        int state = kStateDone;
        while (head < n_ord) {
          const int k = ord[head];
          const unsigned long long ck = sm_col[k];
          const int j = (int)(0xffffffffu - (unsigned)(ck & 0xffffffffull));
          if ((bits_all[j >> 5] >> (j & 31)) & 1u) {
            state = kStateRecompute;  // cached maximum points at an anchor that has been taken since
            break;
          }
          const int n = sm_nmatch;
          m_anchor[n] = j;
          m_gt[n] = k;
          sm_nmatch = n + 1;
          bits_all[j >> 5] |= 1u << (j & 31);
          ++head;
        }
        ctl.state = state;
        ctl.k = state == kStateRecompute ? ord[head] : 0;
        sm_arg = head;
      }
      __syncthreads();
      if (ctl.state == kStateDone) break;  // CTA 0 leaves the loop WITHOUT the barrier: it arrives at (a) after the fix-up
      head = sm_arg;
    }
    cluster.sync();  // (a) recompute round: ctl of CTA 0 is final
    const int state = *cluster.map_shared_rank(&ctl.state, 0);
    if (state == kStateDone) break;  // ranks != 0 only (CTA 0 broke out above)
    const int k = *cluster.map_shared_rank(&ctl.k, 0);
    // this CTA's slice of the matched-anchor bitmap, then the column maximum of gt k over its unmatched anchors
    {
      const unsigned *rb = cluster.map_shared_rank(bits_all, 0) + (s0 >> 5);
      for (int w = threadIdx.x; w < ((s1 - s0 + 31) >> 5); w += blockDim.x) bits[w] = rb[w];
    }
    __syncthreads();
    const float4 g = sm_gt[k];
    unsigned long long tkey = 0ull;
    for (int j0 = s0 + (int)threadIdx.x; j0 < s1; j0 += 4 * (int)blockDim.x) {
      float4 an4[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = j0 + u * (int)blockDim.x;
        an4[u] = j < s1 ? __ldg(anchors + j) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = j0 + u * (int)blockDim.x;
        if (j >= s1 || ((bits[(j - s0) >> 5] >> ((j - s0) & 31)) & 1u)) continue;
        const float4 an = an4[u];
        if (!(an.z > g.x && g.z > an.x && an.w > g.y && g.w > an.y)) continue;
        const float iou = iou_target(an, g);
        if (iou > 1e-6f) {
          const unsigned long long ck = col_key(iou, j);
          tkey = ck > tkey ? ck : tkey;
        }
      }
    }
    const unsigned long long pm = block_max_u64(tkey, red_smem);
    if (threadIdx.x == 0) *cluster.map_shared_rank(&sm_part[rank], 0) = pm;
    cluster.sync();  // (b) the partial maxima are in CTA 0
    if (rank == 0) {
      unsigned long long bm = 0ull;
#pragma unroll
      for (int r = 0; r < kMatchCluster; ++r) bm = sm_part[r] > bm ? sm_part[r] : bm;
      // re-insert gt k behind the entries that still rank before it (a gt without candidate leaves the list)
      int base = head + 1;
      while (true) {
        const int pos = base + (int)threadIdx.x;
        int e = -1;
        bool before = false;
        if (pos < n_ord) {
          e = ord[pos];
          const unsigned long long ek = sm_col[e];
          before = bm == 0ull || ek > bm || (ek == bm && e < k);
        }
        const int moved = __syncthreads_count(before);
        if (before) ord[pos - 1] = e;
        base += moved;
        if (moved < (int)blockDim.x) break;
      }
      if (threadIdx.x == 0) {
        ord[base - 1] = k;
        sm_col[k] = bm;
      }
      if (bm == 0ull) --n_ord;
      __syncthreads();
    }
  }

  int num_negative = 0;
  if (rank == 0) {
    const int nmatch = sm_nmatch;
    DSPMB_TSTAMP_IMG(b, 2);
    // ---- fix up the bipartite-matched anchors; their mining keys are taken out of the owners' slices ----
    const int gs = nmatch > (int)(blockDim.x >> 4) ? 4 : 16;
    for (int q = (int)threadIdx.x / gs; q < nmatch; q += (int)blockDim.x / gs) {
      const int sub = threadIdx.x & (gs - 1);
      const int j = m_anchor[q], k = m_gt[q];
      const float4 an = __ldg(anchors + j);
      float max_iou = -1.0f;
      for (int kk = sub; kk < G; kk += gs) {
        const float4 g = sm_gt[kk];
        const bool reach = an.z > g.x && g.z > an.x && an.w > g.y && g.w > an.y;
        const float iou = reach ? iou_target(an, g) : 0.0f;
        if (iou > max_iou) max_iou = iou;
      }
      const unsigned gmask = (gs == 16 ? 0xffffu : 0xfu) << (lane_id() & ~(unsigned)(gs - 1));
      for (int m = gs >> 1; m > 0; m >>= 1) {
        const float o = __shfl_xor_sync(gmask, max_iou, m);
        if (o > max_iou) max_iou = o;
      }
      if (sub != 0) continue;
      if (a.overlap_threshold > 0.f && max_iou > a.overlap_threshold) atomicAdd(&sm_dup, 1);  // counted by the stream kernel
      const size_t row = (size_t)b * A + j;
      const float *lrow = lab + (size_t)k * W;
      float enc[5];
      encode_loc(an, lrow, a, enc);
#pragma unroll
      for (int c = 0; c < 5; ++c) {
        a.loc_target[row * 5 + c] = enc[c];
        a.loc_mask[row * 5 + c] = 1.0f;
      }
      a.cls_target[row] = fadd(lrow[0], 1.0f);
      if (mining) {
        const int owner = j / slice;
        const unsigned old = atomicExch(cluster.map_shared_rank(&skeys[j - owner * slice], owner), kKeySentinel);
        if (old != kKeySentinel) {
          const int n = atomicAdd(&sm_npatch, 1);
          if (n < kPatchCap) {
            patch_bin[n] = (unsigned short)((old >> 19) & (unsigned)(kDigitBins - 1));
            patch_owner[n] = (unsigned short)owner;
          }
        }
      }
      if (a.match_out) a.match_out[row] = k;
    }
    __syncthreads();
    DSPMB_TSTAMP_IMG(b, 3);
    DSPMB_TSTAMP_IMG(b, 4);
    const int num_positive = sm_thr + nmatch - sm_dup;
    if (mining) {  // multibox_target.cc:186-189
      num_negative = (int)fmul((float)num_positive, a.mining_ratio);
      if (num_negative > A - num_positive) num_negative = A - num_positive;
      if (num_negative < 0) num_negative = 0;
    }
    if (threadIdx.x == 0) {
      if (stats) {
        stats[0] = G;
        stats[1] = num_positive;
        stats[2] = mining ? num_negative : A - num_positive;
        stats[3] = nmatch;
      }
      ctl.state = kStateDone;
      ctl.num_negative = num_negative;
    }
    cluster.sync();  // (a), final round: state == done, num_negative and the patched key slices are visible
  }
  num_negative = *cluster.map_shared_rank(&ctl.num_negative, 0);
  if (num_negative <= 0) {
    cluster.sync();  // nobody leaves while its shared memory may still be read
    return;
  }

  // ---- hard-negative mining (multibox_target.cc:182-241): MSB-first select, 11 + 11 + 8 bits ----
  const int nloc = s1 - s0;
  unsigned prefix = 0u, prefix_mask = 0u;
  int need = num_negative;
  bool exact_key = false;
  int cp = 0;
  const int npatch = *cluster.map_shared_rank(&sm_npatch, 0);
  const unsigned short *rpb = cluster.map_shared_rank(patch_bin, 0), *rpo = cluster.map_shared_rank(patch_owner, 0);
  for (int level = 0; level < 3; ++level) {
    const int shift = level == 0 ? 19 : (level == 1 ? 8 : 0);
    const int nbins = level == 2 ? 256 : kDigitBins;
    unsigned *h = hist[level & 1];
    if (level == 0 && npatch <= kPatchCap) {
      // first digit: histogram and cluster total were prepared before num_negative was known; take the matched
      // anchors' keys out of the total
      if (rank == 0) {
        const unsigned *rs = cluster.map_shared_rank(hist[1], 1);
        for (int i = threadIdx.x; i < kDigitBins; i += blockDim.x) hsum[i] = rs[i];
      }
      __syncthreads();
      for (int i = threadIdx.x; i < npatch; i += blockDim.x) atomicSub(&hsum[rpb[i]], 1u);
      __syncthreads();
    } else {
      if (level == 1) cluster.sync();  // CTA 0 has copied the first-level total out of CTA 1's hist[1]
      for (int i = threadIdx.x; i < nbins; i += blockDim.x) h[i] = 0u;
      __syncthreads();
      for (int base = 0; base < nloc; base += blockDim.x) {
        const int j = base + (int)threadIdx.x;
        const unsigned kv = j < nloc ? skeys[j] : kKeySentinel;
        hist_add_warp(h, (kv >> shift) & (unsigned)(nbins - 1), kv != kKeySentinel && (kv & prefix_mask) == prefix);
      }
      cluster.sync();  // every CTA's histogram of this level is complete
      // every CTA sums the histograms on its own (same result everywhere): coalesced DSMEM reads
      for (int i = threadIdx.x; i < nbins; i += blockDim.x) {
        unsigned t = 0u;
#pragma unroll
        for (int r = 0; r < kMatchCluster; ++r) t += cluster.map_shared_rank(h, r)[i];
        hsum[i] = t;
      }
      __syncthreads();
    }
    const int per = nbins / (int)blockDim.x;  // 8 or 1 consecutive bins per thread
    unsigned c[8];
    int tot = 0;
    for (int i = 0; i < per; ++i) c[i] = hsum[threadIdx.x * per + i];
    for (int i = 0; i < per; ++i) tot += (int)c[i];
    int total;
    const int ex = block_scan_excl(tot, scan_smem, &total);
    if (ex < need && ex + tot >= need) {
      int acc = ex;
      for (int i = 0; i < per; ++i) {
        if (acc + (int)c[i] >= need) {
          sm_pb = (int)threadIdx.x * per + i;
          sm_before = acc;
          sm_cp = (int)c[i];
          break;
        }
        acc += (int)c[i];
      }
    }
    if (threadIdx.x == 0) sm_total = total;
    __syncthreads();
    if (level == 0 && sm_total < need) {
      // fewer candidates than num_negative: CHECK_GE(temp.size(), num_negative) fails (multibox_target.cc:236)
      if (rank == 0 && threadIdx.x == 0) atomicMin(&a.header->status, DSPMB_ERR_MINING_CANDIDATES);
      cluster.sync();
      return;
    }
    prefix |= (unsigned)sm_pb << shift;
    prefix_mask |= (unsigned)(nbins - 1) << shift;
    need -= sm_before;
    cp = sm_cp;
    __syncthreads();  // sm_pb / sm_before / sm_cp are re-written by the next level
    if (level == 2) {
      exact_key = true;  // all 30 bits fixed: the pivot key itself
      break;
    }
    if (cp <= kPivotCap) break;  // few enough to rank directly, one key per thread of the cluster
  }
  unsigned qkey = prefix;
  if (!exact_key) {
    // The pivot bin's keys are copied into EVERY CTA's list at identical positions: CTA r's keys start behind those
    // of the lower ranks (their counts are the histogram entries, minus patched keys at the first level); CTA r then
    // ranks entries [256 r, 256 r + 256) against the whole list, and the thread that holds the need-th smallest key
    // tells everybody.
    const int level_done = prefix_mask == (0x7ffu << 19) ? 0 : 1;
    const unsigned pbin = (prefix >> (level_done == 0 ? 19 : 8)) & (unsigned)(kDigitBins - 1);
    if (threadIdx.x < 32) {
      int cnt = 0;
      if ((int)lane_id() < rank) {
        cnt = (int)cluster.map_shared_rank(hist[level_done & 1], (int)lane_id())[pbin];
        if (level_done == 0 && npatch <= kPatchCap)
          for (int i = 0; i < npatch; ++i) cnt -= (rpb[i] == pbin && rpo[i] == lane_id()) ? 1 : 0;
      }
      cnt = warp_sum_i32(cnt);
      if (lane_id() == 0) sm_off = cnt;
    }
    __syncthreads();
    const int off = sm_off;
    for (int j = threadIdx.x; j < nloc; j += blockDim.x) {
      const unsigned kv = skeys[j];
      if (kv != kKeySentinel && (kv & prefix_mask) == prefix) {
        const int pos = off + atomicAdd(&sm_lcnt, 1);
#pragma unroll
        for (int r = 0; r < kMatchCluster; ++r) cluster.map_shared_rank(plist, r)[pos] = kv;
      }
    }
    cluster.sync();  // every CTA holds the complete list
    {
      const int i = rank * (int)blockDim.x + (int)threadIdx.x;
      if (i < cp) {
        const unsigned v = plist[i];
        int lo = 0, hi = 0;
        for (int x = 0; x < cp; ++x) {
          const unsigned o = plist[x];
          lo += o < v ? 1 : 0;
          hi += o <= v ? 1 : 0;
        }
        if (lo < need && need <= hi) {  // equal keys write the same value
#pragma unroll
          for (int r = 0; r < kMatchCluster; ++r) *cluster.map_shared_rank(&ctl.qkey, r) = v;
        }
      }
    }
    cluster.sync();  // qkey is final everywhere
    qkey = ctl.qkey;
  }
  if (rank == 0) DSPMB_TSTAMP_IMG(b, 5);
  // Keys q approximate the reference's probabilities p with |q - p| <= delta * p: keys below Q(1-3 delta) are certainly
  // selected, keys above Q(1+3 delta) certainly not, the band in between is re-evaluated exactly (CTA 0, below).
  const float Q = __uint_as_float(qkey);
  const unsigned klo = __float_as_uint(fmul(Q, 1.0f - 3.0f * a.delta));
  const unsigned khi = __float_as_uint(fmul(Q, 1.0f + 3.0f * a.delta));
  float *ct = a.cls_target + (size_t)b * A;
  int *amb_list = a.amb_list + (size_t)b * A;
  unsigned *amb_key = a.amb_key + (size_t)b * A;
  {
    int *ramb = cluster.map_shared_rank(&sm_amb, 0);
    int n_in_local = 0;
    for (int base = 0; base < nloc; base += blockDim.x) {
      const int j = base + threadIdx.x;
      const unsigned kv = j < nloc ? skeys[j] : kKeySentinel;
      if (kv < klo) {
        ct[s0 + j] = 0.0f;
        ++n_in_local;
      }
      const bool amb = kv >= klo && kv <= khi;
      const unsigned m = __ballot_sync(kFullMask, amb);
      if (m) {
        int wbase = 0;
        if (lane_id() == 0) wbase = atomicAdd(ramb, __popc(m));
        wbase = __shfl_sync(kFullMask, wbase, 0);
        if (amb) amb_list[wbase + __popc(m & ((1u << lane_id()) - 1u))] = s0 + j;
      }
    }
    n_in_local = warp_sum_i32(n_in_local);
    if (lane_id() == 0 && n_in_local) atomicAdd(cluster.map_shared_rank(&sm_carry, 0), n_in_local);
  }
  __threadfence();  // the ambiguous list (global memory) is read by CTA 0 after the barrier
  cluster.sync();
  if (rank != 0) return;
  DSPMB_TSTAMP_IMG(b, 6);
  const int n_amb = sm_amb;
  const int need2 = num_negative - sm_carry;
  if (need2 < 0 || need2 > n_amb) {  // would mean the error bound was violated
    if (threadIdx.x == 0) atomicMin(&a.header->status, DSPMB_ERR_INTERNAL);
    return;
  }
  const float *p_cls = a.cls_preds + (size_t)b * a.C * A;
  if (n_amb <= 4 * (int)(blockDim.x >> 5)) {  // a handful: one warp each (latency); many: one thread each (throughput)
    for (int q = (int)warp_id(); q < n_amb; q += (int)(blockDim.x >> 5)) {
      const int j = __ldcg(amb_list + q);
      const float pe = exact_bg_prob_warp<kFma>(p_cls, j, A, a.C);
      if (lane_id() == 0) amb_key[q] = __float_as_uint(pe);
    }
  } else {
    for (int q = threadIdx.x; q < n_amb; q += blockDim.x) {
      const int j = __ldcg(amb_list + q);
      const float pe = exact_bg_prob<kFma>(p_cls, j, A, a.C);
      amb_key[q] = __float_as_uint(pe);
    }
  }
  __syncthreads();
  for (int q = threadIdx.x; q < n_amb; q += blockDim.x) {
    const unsigned kq = amb_key[q];
    const int jq = __ldcg(amb_list + q);
    int rnk = 0;
    for (int i = 0; i < n_amb; ++i) {
      const unsigned ki = amb_key[i];
      rnk += (ki < kq || (ki == kq && __ldcg(amb_list + i) < jq)) ? 1 : 0;
    }
    if (rnk < need2) ct[jq] = 0.0f;
  }
  DSPMB_TSTAMP_IMG(b, 7);
}

}  // namespace
}  // namespace dspmb

using namespace dspmb;

extern "C" int dspmb_debug_target_stamps(unsigned long long *device_buffer) {
  DSPMB_CUDA_TRY(cudaMemcpyToSymbol(g_tstamps, &device_buffer, sizeof(device_buffer)));
  return DSPMB_OK;
}

extern "C" int dspmb_debug_target_stream_stamps(unsigned long long *device_buffer) {
  DSPMB_CUDA_TRY(cudaMemcpyToSymbol(g_sstamps, &device_buffer, sizeof(device_buffer)));
  return DSPMB_OK;
}

extern "C" size_t dspmb_target_workspace_bytes(int B, int A, int L, int C) {
  (void)C;
  if (B <= 0 || A <= 0 || L <= 0) return 0;
  return carve(nullptr, B, A, L).bytes;
}

extern "C" int dspmb_target_f32(const float *anchors, const float *labels, const float *cls_preds,
                                float *loc_target, float *loc_mask, float *cls_target, int B, int A, int L,
                                int label_width, int C, float overlap_threshold, float ignore_label,
                                float negative_mining_ratio, float negative_mining_thresh,
                                int minimum_negative_samples, const float *variances, int32_t *match_out,
                                int32_t *stats_out, void *workspace, size_t workspace_bytes, void *stream_) {
  (void)minimum_negative_samples;  // declared but never read by the CPU operator (multibox_target.cc:182-241)
  // Shape CHECKs of MultiBoxTargetProp::InferShape (multibox_target-inl.h:213-238).
  DSPMB_REQUIRE(B >= 0 && A > 0 && L > 0 && C > 0, "MultiBoxTarget: bad shape B=%d A=%d L=%d C=%d", B, A, L, C);
  DSPMB_REQUIRE(label_width >= 6, "MultiBoxTarget: label width must be >= 6 [cls,xmin,ymin,xmax,ymax,dist]");
  DSPMB_REQUIRE(anchors && labels && cls_preds && loc_target && loc_mask && cls_target && variances,
                "MultiBoxTarget: NULL tensor");
  DSPMB_REQUIRE(B <= 65535, "MultiBoxTarget: batch > 65535 not supported in one call");
  DSPMB_REQUIRE(((uintptr_t)anchors & 15) == 0, "MultiBoxTarget: anchors must be 16-byte aligned");
  if (negative_mining_ratio > 0.f && !(negative_mining_thresh > 0.f)) {  // multibox_target.cc:184
    set_error("MultiBoxTarget: negative_mining_thresh must be > 0 when mining is enabled");
    return DSPMB_ERR_MINING_THRESH;
  }
  if (B == 0) return DSPMB_OK;
  const size_t need = carve(nullptr, B, A, L).bytes;
  if (!workspace || workspace_bytes < need || ((uintptr_t)workspace & 255)) {
    set_error("MultiBoxTarget: workspace must be 256-byte aligned and >= %zu bytes (got %zu)", need, workspace_bytes);
    return DSPMB_ERR_WORKSPACE;
  }
  TargetWorkspace w = carve(workspace, B, A, L);

  struct {
    const void *p[9];
    int i[5];
    float f[8];
  } key;
  memset(&key, 0, sizeof(key));
  key.p[0] = anchors, key.p[1] = labels, key.p[2] = cls_preds, key.p[3] = loc_target, key.p[4] = loc_mask;
  key.p[5] = cls_target, key.p[6] = match_out, key.p[7] = stats_out, key.p[8] = workspace;
  key.i[0] = B, key.i[1] = A, key.i[2] = L, key.i[3] = label_width, key.i[4] = C;
  key.f[0] = overlap_threshold, key.f[1] = ignore_label, key.f[2] = negative_mining_ratio, key.f[3] = negative_mining_thresh;
  for (int k = 0; k < 4; ++k) key.f[4 + k] = variances[k];

  return graph_cached_launch(&key, sizeof(key), (cudaStream_t)stream_, [&](const LaunchCtx &ctx) -> int {
  cudaStream_t stream = ctx.stream;

  const uintptr_t align_or = (uintptr_t)cls_preds | (uintptr_t)loc_target | (uintptr_t)loc_mask |
                             (uintptr_t)cls_target | (uintptr_t)match_out;
  const bool vec4 = (A % 4 == 0) && (align_or & 15) == 0;
  // register-resident variants: 2 anchors/thread (87 registers => 5 CTAs/SM; compiled for 6 it spills and measured
  // 48.4 us against 47.4) unless knob 0 says otherwise
  int tvec = (vec4 && (C == 21 || C == 9) && tuning(DSPMB_TUNE_DET_STREAM_VARIANT) != 4) ? 2 : 4;
  // Small batches are latency bound by the CTAs of images with many ground truths (every warp of a large-anchor tile
  // walks the whole gt list, two IoUs per anchor and gt): up to about two waves of CTAs (measured: 8 and 16 images
  // of SSD-512), one anchor per thread shortens that chain and still fills the machine.
  const int small_mode = tuning(DSPMB_TUNE_TARGET_SMALL);  // 0 never, 1 (default) by size, 2 always
  if (vec4 && (C == 21 || C == 9) && tvec == 2 &&
      (small_mode == 2 || (small_mode == 1 && (long long)B * ceil_div(A, 2 * kStreamThreads) <= 12LL * kNumSMs)))
    tvec = 1;
  const int tile = kStreamThreads * (vec4 ? tvec : 1);

  TargetArgs ta;
  ta.anchors = anchors;
  ta.labels = labels;
  ta.cls_preds = cls_preds;
  ta.loc_target = loc_target;
  ta.loc_mask = loc_mask;
  ta.cls_target = cls_target;
  ta.match_out = match_out;
  ta.stats_out = stats_out;
  ta.header = w.header;
  ta.gcount = w.gcount;
  ta.thr_count = w.thr_count;
  ta.colbest = w.colbest;
  ta.key = w.key;
  ta.amb_list = w.amb_list;
  ta.amb_key = w.amb_key;
  ta.delta = 1e-4f > (2e-7f * C + 5e-5f) ? 1e-4f : (2e-7f * C + 5e-5f);
  ta.B = B;
  ta.A = A;
  ta.L = L;
  ta.W = label_width;
  ta.C = C;
  ta.T = ceil_div(A, tile);
  ta.overlap_threshold = overlap_threshold;
  ta.ignore_label = ignore_label;
  ta.mining_ratio = negative_mining_ratio;
  ta.mining_thresh = negative_mining_thresh;
  ta.vx = variances[0];
  ta.vy = variances[1];
  ta.vw = variances[2];
  ta.vh = variances[3];
  ta.fma_build = libm_fma_mode();
  ta.shortlist = tuning(DSPMB_TUNE_TARGET_SHORTLIST);
  ta.prefetch = tuning(DSPMB_TUNE_TARGET_PREFETCH);

  const size_t smem1 = (sizeof(float4) + sizeof(unsigned long long) + sizeof(float)) * (size_t)L +
                       sizeof(unsigned short) * (size_t)L * (kStreamThreads / 32);
  const size_t smem2 = (sizeof(float4) + sizeof(unsigned long long) + 2 * sizeof(int)) * (size_t)L +
                       2 * sizeof(int) * (size_t)((L + 3) & ~3) +
                       (size_t)((L + 15) / 16) * 16 + sizeof(unsigned) * (size_t)((((A + 31) / 32) + 3) & ~3);
  const bool keys_in_smem = smem2 + sizeof(unsigned) * (size_t)A <= 170 * 1024;
  const size_t smem2_total = smem2 + (keys_in_smem ? sizeof(unsigned) * (size_t)((A + 3) & ~3) + 2 * sizeof(unsigned) * (size_t)kShortCap : 0);
  DSPMB_REQUIRE(smem1 <= 48 * 1024, "MultiBoxTarget: more than %d label slots are not supported", 48 * 1024 / 36);
  DSPMB_REQUIRE(smem2 <= 190 * 1024, "MultiBoxTarget: A=%d / L=%d exceed the matcher's shared memory", A, L);
  DSPMB_ENSURE_DYN_SMEM(target_match_kernel<true>, 204 * 1024);  // + ~14 KB static
  DSPMB_ENSURE_DYN_SMEM(target_match_kernel<false>, 204 * 1024);
  const int phases = tuning(DSPMB_TUNE_PHASES);
  dim3 grid1(ta.T, B);
  if (phases & 1) {
    ProfileScope _p(kSlotTargetStream, stream);
    const bool fma = ta.fma_build != 0;
#define DSPMB_LAUNCH_TS(V, N)                                                          \
  do {                                                                                 \
    if (fma)                                                                           \
      target_stream_kernel<V, N, true><<<grid1, kStreamThreads, smem1, stream>>>(ta);  \
    else                                                                               \
      target_stream_kernel<V, N, false><<<grid1, kStreamThreads, smem1, stream>>>(ta); \
  } while (0)
    if (vec4 && C == 21 && tvec == 1)
      DSPMB_LAUNCH_TS(1, 21);
    else if (vec4 && C == 9 && tvec == 1)
      DSPMB_LAUNCH_TS(1, 9);
    else if (vec4 && C == 21 && tvec == 2)
      DSPMB_LAUNCH_TS(2, 21);
    else if (vec4 && C == 9 && tvec == 2)
      DSPMB_LAUNCH_TS(2, 9);
    else if (vec4 && C == 21)
      DSPMB_LAUNCH_TS(4, 21);
    else if (vec4 && C == 9)
      DSPMB_LAUNCH_TS(4, 9);
    else if (vec4)
      DSPMB_LAUNCH_TS(4, 0);
    else
      DSPMB_LAUNCH_TS(1, 0);
#undef DSPMB_LAUNCH_TS
    ++ctx.launches;
  }
  DSPMB_CUDA_TRY(cudaGetLastError());
  // cluster matcher: key slices and the matched-anchor bitmap must fit in the CTAs' shared memory
  const int cslice = (((A + kMatchCluster - 1) / kMatchCluster) + 31) & ~31;
  const size_t smem3 = (sizeof(float4) + 2 * sizeof(unsigned long long) + 2 * sizeof(int)) * (size_t)L +
                       sizeof(int) * (size_t)((L + 3) & ~3) + sizeof(unsigned) * (size_t)((((A + 31) / 32) + 3) & ~3) +
                       sizeof(unsigned) * (size_t)(((cslice / 32) + 3) & ~3) + sizeof(unsigned) * (size_t)cslice +
                       2 * sizeof(unsigned short) * (size_t)L;
  const bool use_cluster = tuning(DSPMB_TUNE_TARGET_PIPELINE) != 0 && cslice <= kClusterSliceMax && smem3 <= 150 * 1024;
  if ((phases & 2) && use_cluster) {
    ProfileScope _p(kSlotTargetMatch, stream);
    DSPMB_ENSURE_DYN_SMEM(target_match_cluster_kernel<true>, 150 * 1024);
    DSPMB_ENSURE_DYN_SMEM(target_match_cluster_kernel<false>, 150 * 1024);
    if (ta.fma_build)
      target_match_cluster_kernel<true><<<B * kMatchCluster, kClusterThreads, smem3, stream>>>(ta);
    else
      target_match_cluster_kernel<false><<<B * kMatchCluster, kClusterThreads, smem3, stream>>>(ta);
    ++ctx.launches;
  } else if (phases & 2) {
    ProfileScope _p(kSlotTargetMatch, stream);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(B);
    cfg.blockDim = dim3(kMatchThreads);
    cfg.dynamicSmemBytes = smem2_total;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    // programmatic launch pays when the stream kernel has a tail to hide behind (SSD-512: -0.6 us at 64 images, -1.6 us
    // at 8); with less than one CTA per SM there is none, and the cold single-image step measured 3 us SLOWER
    attr[0].val.programmaticStreamSerializationAllowed =
        (tuning(DSPMB_TUNE_TARGET_PDL) && (long long)ta.T * B > kNumSMs) ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (keys_in_smem)
      DSPMB_CUDA_TRY(cudaLaunchKernelEx(&cfg, target_match_kernel<true>, ta));
    else
      DSPMB_CUDA_TRY(cudaLaunchKernelEx(&cfg, target_match_kernel<false>, ta));
    ++ctx.launches;
  }
  DSPMB_CUDA_TRY(cudaGetLastError());
  return DSPMB_OK;
  });
}
