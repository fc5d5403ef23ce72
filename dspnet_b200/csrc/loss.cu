// Forward of the training graph right behind MultiBoxTarget in ONE streaming pass:
//   cls_prob = SoftmaxOutput(cls_preds, cls_target, ignore_label=-1, use_ignore, multi_output, normalization='valid')
//              -> its forward value is the channel softmax of cls_preds (B, C, A)      symbol/symbol_builder.py:82-84
//   loc_loss = MakeLoss(smooth_l1(loc_target_mask * (loc_preds - loc_target), scalar=1.0))              :85-88
// and the statistics MultiBoxMetric.update takes from those tensors on the host (train/metric.py:27-46):
//   valid_count = #(cls_target >= 0), CrossEntropy sum = sum over those of -log(prob[label] + eps), SmoothL1 sum =
//   sum(loc_loss)  -- plus the number of loc_loss elements > 0, the count MakeLoss(normalization='valid') divides its
//   gradient by.  normalization only acts in Backward; the forward values do not depend on it.
// The reference runs SoftmaxOutput, the broadcast multiply / subtract / smooth_l1 chain and the metric's numpy code
// as separate passes over cls_preds (132 MB at SSD-512, batch 64) and the 69 MB of targets; here every tensor is read
// once, and the metric needs 32 bytes per image instead of the (B, C, A) probabilities on the host.
// The softmax arithmetic lives in MXNet (not part of the reference tree): it is the softmax the reference itself
// spells out in multibox_target.cc:220-231 (running maximum, fp32 sum of expf(x - max) in class order, one division),
// restated by oracle_softmax_channel and evaluated here with the glibc-bit-exact expf, so cls_prob and loc_loss are
// bit-identical to the oracle; the three sums are accumulated in fp64 in a fixed order (per-CTA partials, reduced in
// tile order by the last CTA of an image), so they are reproducible run to run.
#include "common.cuh"

namespace dspmb {
namespace {

constexpr int kLossThreads = 128;

struct LossArgs {
  const float *cls_preds, *loc_preds, *loc_target, *loc_mask, *cls_target;
  float *cls_prob, *loc_loss;
  double *stats;       // (B, 4): valid count, cross-entropy sum, smooth-L1 sum, # loc_loss elements > 0
  double *partial;     // (B, T, 4) per-CTA partial sums
  unsigned *arrived;   // (B) CTAs of the image that have published their partials (zeroed by a memset node per call)
  int A, C, T;
  float eps;
  int fma_build;
};

__device__ __forceinline__ float smooth_l1_unit(float d) {  // mx.symbol.smooth_l1(scalar=1.0)
  const float ad = fabsf(d);
  return ad < 1.0f ? fmul(0.5f, fmul(d, d)) : fsub(ad, 0.5f);
}


// Masked smooth-L1 of one warp's block of the (B, A*5) tensors: the loss is element-wise, so the warp walks its
// 32 x VEC anchors x 5 contiguous floats as full 128-bit lines (lane l takes float4 l, l + 32, ...) instead of every
// lane reading its own 80-byte piece in five 16-byte steps, which moves every 32-byte sector through L1 twice.
template <int VEC>
__device__ __forceinline__ void warp_loc_loss(const LossArgs &a, size_t warp_row0, int warp_anchors, double &s_l1, double &s_pos) {
  const float *pp = a.loc_preds + warp_row0 * 5, *pt = a.loc_target + warp_row0 * 5, *pm = a.loc_mask + warp_row0 * 5;
  float *po = a.loc_loss ? a.loc_loss + warp_row0 * 5 : nullptr;
  const int n4 = warp_anchors * 5 / 4;  // warp_anchors is a multiple of 4
#pragma unroll
  for (int q0 = 0; q0 < 32 * VEC * 5 / 4; q0 += 32) {
    const int q = q0 + (int)lane_id();
    if (q < n4) {
      const float4 p4 = ld_stream_f4(pp + 4 * q), t4 = ld_stream_f4(pt + 4 * q), m4 = ld_stream_f4(pm + 4 * q);
      float4 l4;
      l4.x = smooth_l1_unit(fmul(m4.x, fsub(p4.x, t4.x)));
      l4.y = smooth_l1_unit(fmul(m4.y, fsub(p4.y, t4.y)));
      l4.z = smooth_l1_unit(fmul(m4.z, fsub(p4.z, t4.z)));
      l4.w = smooth_l1_unit(fmul(m4.w, fsub(p4.w, t4.w)));
      if (po) st_stream_f4(po + 4 * q, l4);
      s_l1 += (double)l4.x + (double)l4.y + (double)l4.z + (double)l4.w;
      s_pos += (l4.x > 0.f) + (l4.y > 0.f) + (l4.z > 0.f) + (l4.w > 0.f);
    }
  }
}

template <int VEC, int NC>
__global__ void __launch_bounds__(kLossThreads) multibox_loss_kernel(const __grid_constant__ LossArgs a) {
  __shared__ double red[kLossThreads / 32][4];
  __shared__ bool last;
  const int b = blockIdx.y, t = blockIdx.x, A = a.A;
  const int C = NC > 0 ? NC : a.C;
  const int i0 = (t * kLossThreads + (int)threadIdx.x) * VEC;
  double s_valid = 0.0, s_ce = 0.0, s_l1 = 0.0, s_pos = 0.0;
  if constexpr (VEC == 4) {
    const int w0 = (t * kLossThreads + (int)warp_id() * 32) * VEC;  // first anchor of this warp
    if (w0 < A) warp_loc_loss<VEC>(a, (size_t)b * A + w0, min(32 * VEC, A - w0), s_l1, s_pos);
  }
  if (i0 < A) {
    const size_t row0 = (size_t)b * A + i0;
    // ---- localisation loss: 5 values per anchor, VEC anchors = 5 * VEC consecutive floats ----
    {
      const float *pp = a.loc_preds + row0 * 5, *pt = a.loc_target + row0 * 5, *pm = a.loc_mask + row0 * 5;
      float *po = a.loc_loss ? a.loc_loss + row0 * 5 : nullptr;
      if constexpr (VEC == 4) {
        (void)pp, (void)pt, (void)pm, (void)po;  // walked warp-wide above (coalesced)
      } else {
        for (int q = 0; q < 5 * VEC; ++q) {
          const float l = smooth_l1_unit(fmul(pm[q], fsub(pp[q], pt[q])));
          if (po) po[q] = l;
          s_l1 += (double)l;
          s_pos += l > 0.f;
        }
      }
    }
    // ---- channel softmax + cross-entropy of the labelled class ----
    const float *cp = a.cls_preds + (size_t)b * C * A + i0;
    float *op = a.cls_prob ? a.cls_prob + (size_t)b * C * A + i0 : nullptr;
    float lab[VEC];
    if constexpr (VEC == 4) {
      const float4 l4 = ld_stream_f4(a.cls_target + row0);
      lab[0] = l4.x, lab[1] = l4.y, lab[2] = l4.z, lab[3] = l4.w;
    } else {
      lab[0] = a.cls_target[row0];
    }
    float mx[VEC], sum[VEC], plab[VEC];
    if constexpr (NC > 0 && VEC == 4) {
      float4 x[NC];
#pragma unroll
      for (int c = 0; c < NC; ++c) x[c] = ld_stream_f4(cp + (size_t)c * A);
      mx[0] = x[0].x, mx[1] = x[0].y, mx[2] = x[0].z, mx[3] = x[0].w;
#pragma unroll
      for (int c = 1; c < NC; ++c) {
        if (x[c].x > mx[0]) mx[0] = x[c].x;
        if (x[c].y > mx[1]) mx[1] = x[c].y;
        if (x[c].z > mx[2]) mx[2] = x[c].z;
        if (x[c].w > mx[3]) mx[3] = x[c].w;
      }
#pragma unroll
      for (int v = 0; v < 4; ++v) sum[v] = 0.f, plab[v] = 0.f;
#pragma unroll
      for (int c = 0; c < NC; ++c) {  // x[c] is overwritten by its exponential
        x[c].x = libm::expf_glibc(fsub(x[c].x, mx[0]), a.fma_build);
        x[c].y = libm::expf_glibc(fsub(x[c].y, mx[1]), a.fma_build);
        x[c].z = libm::expf_glibc(fsub(x[c].z, mx[2]), a.fma_build);
        x[c].w = libm::expf_glibc(fsub(x[c].w, mx[3]), a.fma_build);
        sum[0] = fadd(sum[0], x[c].x);
        sum[1] = fadd(sum[1], x[c].y);
        sum[2] = fadd(sum[2], x[c].z);
        sum[3] = fadd(sum[3], x[c].w);
      }
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        const float4 p4 = make_float4(fdiv(x[c].x, sum[0]), fdiv(x[c].y, sum[1]), fdiv(x[c].z, sum[2]), fdiv(x[c].w, sum[3]));
        if (op) st_stream_f4(op + (size_t)c * A, p4);
        if (lab[0] == (float)c) plab[0] = p4.x;
        if (lab[1] == (float)c) plab[1] = p4.y;
        if (lab[2] == (float)c) plab[2] = p4.z;
        if (lab[3] == (float)c) plab[3] = p4.w;
      }
    } else {
      for (int v = 0; v < VEC; ++v) {
        mx[v] = cp[v];
        for (int c = 1; c < C; ++c) {
          const float xv = cp[(size_t)c * A + v];
          if (xv > mx[v]) mx[v] = xv;
        }
        sum[v] = 0.f;
        for (int c = 0; c < C; ++c) sum[v] = fadd(sum[v], libm::expf_glibc(fsub(cp[(size_t)c * A + v], mx[v]), a.fma_build));
        plab[v] = 0.f;
        for (int c = 0; c < C; ++c) {
          const float p = fdiv(libm::expf_glibc(fsub(cp[(size_t)c * A + v], mx[v]), a.fma_build), sum[v]);
          if (op) op[(size_t)c * A + v] = p;
          if (lab[v] == (float)c) plab[v] = p;
        }
      }
    }
#pragma unroll
    for (int v = 0; v < VEC; ++v)
      if (lab[v] >= 0.f) {  // label = cls_label.flatten(); mask = where(label >= 0), train/metric.py:37-38
        s_valid += 1.0;
        s_ce -= (double)libm::logf_glibc(fadd(plab[v], a.eps), a.fma_build);
      }
  }
  // ---- per-CTA partials in a fixed order, then the image's last CTA adds the partials up in tile order ----
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) {
    s_valid += __shfl_xor_sync(kFullMask, s_valid, m);
    s_ce += __shfl_xor_sync(kFullMask, s_ce, m);
    s_l1 += __shfl_xor_sync(kFullMask, s_l1, m);
    s_pos += __shfl_xor_sync(kFullMask, s_pos, m);
  }
  if (lane_id() == 0) {
    red[warp_id()][0] = s_valid;
    red[warp_id()][1] = s_ce;
    red[warp_id()][2] = s_l1;
    red[warp_id()][3] = s_pos;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    double v = 0.0;
    for (int w = 0; w < kLossThreads / 32; ++w) v += red[w][threadIdx.x];
    a.partial[((size_t)b * a.T + t) * 4 + threadIdx.x] = v;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned n = atomicAdd(&a.arrived[b], 1u);
    last = n == (unsigned)a.T - 1u;
  }
  __syncthreads();
  if (last && threadIdx.x < 4) {
    __threadfence();
    double v = 0.0;
    for (int tt = 0; tt < a.T; ++tt) v += __ldcg(a.partial + ((size_t)b * a.T + tt) * 4 + threadIdx.x);
    a.stats[(size_t)b * 4 + threadIdx.x] = v;
  }
}


// Statistics-only variant (cls_prob not requested): the cross-entropy needs the probability of the LABELLED class of
// the anchors with a label only -- a few percent of them when hard-negative mining is on -- and every probability
// costs C bit-exact expf evaluations in fp64, which is what the kernel above spends its time on (33 M of them at
// SSD-512, batch 64).  Here a CTA first streams its loc tensors and labels (the logits are not touched), lists its
// labelled anchors in shared memory, and then evaluates the softmax of only those, kLossBatch anchors at a time with
// one (anchor, class) pair per thread.  Same arithmetic, same fixed-order fp64 sums.
constexpr int kLossBatch = 64;
template <int VEC>
__global__ void __launch_bounds__(kLossThreads) multibox_metric_kernel(const __grid_constant__ LossArgs a) {
  __shared__ double red[kLossThreads / 32][4];
  __shared__ bool last;
  __shared__ unsigned short list[kLossThreads * VEC];
  __shared__ unsigned char llab[kLossThreads * VEC];
  __shared__ int scan_smem[kLossThreads / 32 + 1];
  __shared__ float cmx[kLossBatch];
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  float *es = reinterpret_cast<float *>(dyn_smem);  // [kLossBatch][C]
  const int b = blockIdx.y, t = blockIdx.x, A = a.A, C = a.C;
  const int tile0 = t * kLossThreads * VEC;
  const int i0 = tile0 + (int)threadIdx.x * VEC;
  double s_valid = 0.0, s_ce = 0.0, s_l1 = 0.0, s_pos = 0.0;
  float lab[VEC];
  int nmine = 0;
#pragma unroll
  for (int v = 0; v < VEC; ++v) lab[v] = -1.f;
  if constexpr (VEC == 4) {
    const int w0 = tile0 + (int)warp_id() * 32 * VEC;  // first anchor of this warp
    if (w0 < A) warp_loc_loss<VEC>(a, (size_t)b * A + w0, min(32 * VEC, A - w0), s_l1, s_pos);
  }
  if (i0 < A) {
    const size_t row0 = (size_t)b * A + i0;
    const float *pp = a.loc_preds + row0 * 5, *pt = a.loc_target + row0 * 5, *pm = a.loc_mask + row0 * 5;
    float *po = a.loc_loss ? a.loc_loss + row0 * 5 : nullptr;
    if constexpr (VEC == 4) {
      (void)pp, (void)pt, (void)pm, (void)po;  // the loc tensors are walked warp-wide above
      const float4 l4 = ld_stream_f4(a.cls_target + row0);
      lab[0] = l4.x, lab[1] = l4.y, lab[2] = l4.z, lab[3] = l4.w;
    } else {
      for (int q = 0; q < 5 * VEC; ++q) {
        const float l = smooth_l1_unit(fmul(pm[q], fsub(pp[q], pt[q])));
        if (po) po[q] = l;
        s_l1 += (double)l;
        s_pos += l > 0.f;
      }
      lab[0] = a.cls_target[row0];
    }
#pragma unroll
    for (int v = 0; v < VEC; ++v) nmine += lab[v] >= 0.f;
  }
  s_valid = (double)nmine;
  int nlist;
  int pos = block_scan_excl(nmine, scan_smem, &nlist);
#pragma unroll
  for (int v = 0; v < VEC; ++v)
    if (lab[v] >= 0.f) {
      list[pos] = (unsigned short)(threadIdx.x * VEC + v);
      // a label outside [0, C) indexes outside the probabilities in the reference too (numpy would raise): class 0 here
      llab[pos] = (unsigned char)((lab[v] < (float)C && lab[v] < 256.f) ? (int)lab[v] : 0);
      ++pos;
    }
  __syncthreads();
  const float *cp = a.cls_preds + (size_t)b * C * A + tile0;
  for (int c0 = 0; c0 < nlist; c0 += kLossBatch) {
    const int nb = min(kLossBatch, nlist - c0);
    for (int q = threadIdx.x; q < nb * C; q += kLossThreads) {  // logits of the batch: class-major reads
      const int c = q / nb, k = q - c * nb;
      es[k * C + c] = __ldg(cp + (size_t)c * A + list[c0 + k]);
    }
    __syncthreads();
    if ((int)threadIdx.x < nb) {
      const float *x = es + threadIdx.x * C;
      float mx = x[0];
      for (int c = 1; c < C; ++c)
        if (x[c] > mx) mx = x[c];
      cmx[threadIdx.x] = mx;
    }
    __syncthreads();
    for (int q = threadIdx.x; q < nb * C; q += kLossThreads) {
      const int k = q / C;
      es[q] = libm::expf_glibc(fsub(es[q], cmx[k]), a.fma_build);
    }
    __syncthreads();
    if ((int)threadIdx.x < nb) {
      const float *e = es + threadIdx.x * C;
      float sum = 0.f;
      for (int c = 0; c < C; ++c) sum = fadd(sum, e[c]);
      const float p = fdiv(e[llab[c0 + threadIdx.x]], sum);
      s_ce -= (double)libm::logf_glibc(fadd(p, a.eps), a.fma_build);
    }
    __syncthreads();
  }
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) {
    s_valid += __shfl_xor_sync(kFullMask, s_valid, m);
    s_ce += __shfl_xor_sync(kFullMask, s_ce, m);
    s_l1 += __shfl_xor_sync(kFullMask, s_l1, m);
    s_pos += __shfl_xor_sync(kFullMask, s_pos, m);
  }
  if (lane_id() == 0) {
    red[warp_id()][0] = s_valid;
    red[warp_id()][1] = s_ce;
    red[warp_id()][2] = s_l1;
    red[warp_id()][3] = s_pos;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    double v = 0.0;
    for (int w = 0; w < kLossThreads / 32; ++w) v += red[w][threadIdx.x];
    a.partial[((size_t)b * a.T + t) * 4 + threadIdx.x] = v;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = atomicAdd(&a.arrived[b], 1u) == (unsigned)a.T - 1u;
  __syncthreads();
  if (last && threadIdx.x < 4) {
    __threadfence();
    double v = 0.0;
    for (int tt = 0; tt < a.T; ++tt) v += __ldcg(a.partial + ((size_t)b * a.T + tt) * 4 + threadIdx.x);
    a.stats[(size_t)b * 4 + threadIdx.x] = v;
  }
}

}  // namespace
}  // namespace dspmb

using namespace dspmb;

static size_t loss_layout(int B, int A, size_t *partial_off) {
  const size_t T = (size_t)ceil_div(A, kLossThreads);  // bound for every VEC
  const size_t arrived = align_up(sizeof(unsigned) * (size_t)B, 256);
  if (partial_off) *partial_off = arrived;
  return arrived + align_up(sizeof(double) * (size_t)B * T * 4, 256);
}

extern "C" size_t dspmb_multibox_loss_workspace_bytes(int B, int A) {
  if (B <= 0 || A <= 0) return 0;
  return loss_layout(B, A, nullptr);
}

extern "C" int dspmb_multibox_loss_f32(const float *cls_preds, const float *loc_preds, const float *loc_target,
                                       const float *loc_mask, const float *cls_target, float *cls_prob, float *loc_loss,
                                       double *stats, int B, int A, int C, float eps, void *workspace,
                                       size_t workspace_bytes, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  DSPMB_REQUIRE(B >= 0 && A > 0 && C > 0, "MultiBoxLoss: bad shape B=%d A=%d C=%d", B, A, C);
  DSPMB_REQUIRE(cls_preds && loc_preds && loc_target && loc_mask && cls_target && stats, "MultiBoxLoss: NULL tensor");
  DSPMB_REQUIRE(B <= 65535, "MultiBoxLoss: batch > 65535 not supported in one call");
  if (B == 0) return DSPMB_OK;
  size_t partial_off = 0;
  const size_t need = loss_layout(B, A, &partial_off);
  if (!workspace || workspace_bytes < need || ((uintptr_t)workspace & 255)) {
    set_error("MultiBoxLoss: workspace must be 256-byte aligned and >= %zu bytes (got %zu)", need, workspace_bytes);
    return DSPMB_ERR_WORKSPACE;
  }
  const uintptr_t align_or = (uintptr_t)cls_preds | (uintptr_t)loc_preds | (uintptr_t)loc_target | (uintptr_t)loc_mask |
                             (uintptr_t)cls_target | (uintptr_t)cls_prob | (uintptr_t)loc_loss;
  const bool vec4 = (A % 4 == 0) && (align_or & 15) == 0;
  LossArgs la;
  la.cls_preds = cls_preds;
  la.loc_preds = loc_preds;
  la.loc_target = loc_target;
  la.loc_mask = loc_mask;
  la.cls_target = cls_target;
  la.cls_prob = cls_prob;
  la.loc_loss = loc_loss;
  la.stats = stats;
  la.arrived = (unsigned *)workspace;
  la.partial = (double *)((char *)workspace + partial_off);
  la.A = A;
  la.C = C;
  la.T = ceil_div(A, kLossThreads * (vec4 ? 4 : 1));
  la.eps = eps;
  la.fma_build = libm_fma_mode();
  DSPMB_CUDA_TRY(cudaMemsetAsync(workspace, 0, sizeof(unsigned) * (size_t)B, stream));  // arrival counters
  dim3 grid(la.T, B);
  {
    ProfileScope _p(kSlotLoss, stream);
    const size_t smem_es = sizeof(float) * (size_t)kLossBatch * C;
    if (!cls_prob && smem_es <= 40 * 1024) {  // statistics (and loc_loss) only: softmax of the labelled anchors alone
      if (vec4)
        multibox_metric_kernel<4><<<grid, kLossThreads, smem_es, stream>>>(la);
      else
        multibox_metric_kernel<1><<<grid, kLossThreads, smem_es, stream>>>(la);
    } else if (vec4 && C == 21)
      multibox_loss_kernel<4, 21><<<grid, kLossThreads, 0, stream>>>(la);
    else if (vec4 && C == 9)
      multibox_loss_kernel<4, 9><<<grid, kLossThreads, 0, stream>>>(la);
    else if (vec4)
      multibox_loss_kernel<4, 0><<<grid, kLossThreads, 0, stream>>>(la);
    else
      multibox_loss_kernel<1, 0><<<grid, kLossThreads, 0, stream>>>(la);
  }
  DSPMB_CUDA_TRY(cudaGetLastError());
  note_launches(2);  // memset node + kernel
  return DSPMB_OK;
}
