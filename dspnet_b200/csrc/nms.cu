// Standalone greedy NMS for sm_100a -- replaces the py-faster-rcnn helpers the reference ships:
//   cpu_nms  cython/cpu_nms.pyx:17-68   (suppress iff double(iou) >= thresh; the parity target)
//   gpu_nms  cython/gpu_nms.pyx:16-31 + _nms / nms_kernel, cython/nms_kernel.cu:34-144 (suppress iff iou > thresh)
//   nms      detect/nms.py:24-58         (numpy, same rule as gpu_nms)
// All use the pixel "+1" IoU convention and process boxes in descending score order.
//
// The reference's gpu path argsorts on the host, cudaMallocs per call, computes the FULL N x N/64 mask (the
// lower-triangle early-out is commented out, nms_kernel.cu:39), copies the mask to the host and sweeps it there.
// Here everything stays on the device and on the caller's stream.  Default (DSPMB_TUNE_NMS_PIPELINE = 1), work
// O(N x kept) instead of O(N^2) and workspace O(N) instead of the N x N/64 mask:
//   sort                N <= 16384: one CTA, bitonic network in shared memory; larger: multi-CTA bitonic (every CTA
//                       sorts a 4096-key tile in shared memory, the wide strides are one compare-exchange pass over
//                       global memory each, the narrow ones are finished per tile in shared memory again);
//   nms_gather_kernel   sorted boxes as float4 + precomputed fp32 areas (cpu_nms.pyx:24) + optional class;
//   then, per chunk of 1024 boxes in score order (greedy NMS only ever needs a box compared with the boxes KEPT before
//   it -- 21 k of 200 k on the benchmark sweep):
//     nms_cull_kernel     every CTA stages a slab of 128 kept boxes in shared memory and tests the chunk against it
//                         (four chained compares on x2+1 / y2+1 decide "cannot overlap"; the few pairs that pass take
//                         the reference's IoU and its threshold rule); a suppressed box gets its dead flag;
//     nms_resolve_kernel  one CTA: ordered compaction of the chunk's survivors, their upper-triangular bit mask in
//                         shared memory, greedy resolve on 64-bit words, survivors appended to the kept list (boxes
//                         for the later chunks' cull + original indices = the output, already in score order).
// DSPMB_TUNE_NMS_PIPELINE = 0 keeps the first implementation (full mask, kept for A/B runs and as a second opinion in
// the tests):
//   nms_sort_kernel     64-bit keys (~score | ~index) sorted by a bitonic network (shared memory up to 16K keys,
//                       global above) => descending score, ties to the higher index like a stable
//                       argsort()[::-1];
//   nms_gather_kernel   sorted boxes as float4 + precomputed fp32 areas (cpu_nms.pyx:24) + optional class;
//   nms_mask_kernel     upper-triangular 64 x 64 tiles only, column tile staged in shared memory;
//   nms_scan_kernel     one CTA: 64-row chunks; the diagonal words are resolved serially in registers, the rows
//                       of the surviving boxes are OR-ed into a shared-memory bit vector with all loads in flight
//                       at once; then an ordered compaction writes the kept original indices.
#include <cooperative_groups.h>

#include "common.cuh"

namespace dspmb {
namespace {

constexpr int kSortThreads = 1024;
constexpr int kSortSmemKeys = 16384;
constexpr int kScanThreads = 1024;

struct NmsWorkspace {
  unsigned long long *keys;  // npad
  int *order;                // N sorted position -> original row
  float4 *box;               // N (sorted)
  float *area;               // N
  float *cls;                // N
  unsigned long long *mask;  // N x W
  size_t bytes;
};

inline int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

NmsWorkspace carve(void *base, int N) {
  NmsWorkspace w;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off = align_up(off + bytes, 256);
    return (char *)base + o;
  };
  const size_t W = (size_t)ceil_div(N, 64);
  w.keys = (unsigned long long *)take(sizeof(unsigned long long) * (size_t)next_pow2(N < 2 ? 2 : N));
  w.order = (int *)take(sizeof(int) * (size_t)N);
  w.box = (float4 *)take(sizeof(float4) * (size_t)N);
  w.area = (float *)take(sizeof(float) * (size_t)N);
  w.cls = (float *)take(sizeof(float) * (size_t)N);
  w.mask = (unsigned long long *)take(sizeof(unsigned long long) * (size_t)N * W);
  w.bytes = off;
  return w;
}

__device__ void bitonic_sort_u64(unsigned long long *keys, int n) {
  for (int k = 2; k <= n; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int q = threadIdx.x; q < (n >> 1); q += blockDim.x) {
        const int lo = ((q & ~(j - 1)) << 1) | (q & (j - 1));
        const int hi = lo | j;
        const bool up = (lo & k) == 0;
        const unsigned long long x = keys[lo], y = keys[hi];
        if ((x > y) == up) {
          keys[lo] = y;
          keys[hi] = x;
        }
      }
      __syncthreads();
    }
}

__global__ void __launch_bounds__(kSortThreads) nms_sort_kernel(const float *__restrict__ dets, int N, int dim,
                                                                 int npad, unsigned long long *gkeys,
                                                                 int *__restrict__ order) {
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  unsigned long long *keys = npad <= kSortSmemKeys ? reinterpret_cast<unsigned long long *>(dyn_smem) : gkeys;
  for (int i = threadIdx.x; i < npad; i += blockDim.x) {
    unsigned long long k = ~0ull;
    if (i < N) k = ((unsigned long long)(~float_order_key(dets[(size_t)i * dim + 4])) << 32) | (0xffffffffu - (unsigned)i);
    keys[i] = k;
  }
  __syncthreads();
  bitonic_sort_u64(keys, npad);
  for (int i = threadIdx.x; i < N; i += blockDim.x) order[i] = (int)(0xffffffffu - (unsigned)(keys[i] & 0xffffffffull));
}

__global__ void nms_gather_kernel(const float *__restrict__ dets, int N, int dim, int class_col,
                                  const int *__restrict__ order, float4 *__restrict__ box, float *__restrict__ area,
                                  float *__restrict__ cls) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const int src = order ? order[i] : i;
  const float *d = dets + (size_t)src * dim;
  const float x1 = d[0], y1 = d[1], x2 = d[2], y2 = d[3];
  box[i] = make_float4(x1, y1, x2, y2);
  area[i] = fmul(fadd(fsub(x2, x1), 1.0f), fadd(fsub(y2, y1), 1.0f));  // cpu_nms.pyx:24
  if (class_col >= 0) cls[i] = d[class_col];
}

template <bool kUseClass>
__global__ void __launch_bounds__(64) nms_mask_kernel(int N, int W, double thresh, int mode,
                                                      const float4 *__restrict__ box, const float *__restrict__ area,
                                                      const float *__restrict__ cls, unsigned long long *__restrict__ mask) {
  const int rt = blockIdx.y, ct = blockIdx.x;
  if (ct < rt) return;  // only the upper triangle is ever read by the scan
  __shared__ float4 sbox[64];
  __shared__ float sarea[64];
  __shared__ float scls[64];
  const int col_n = min(64, N - ct * 64);
  if ((int)threadIdx.x < col_n) {
    const int c = ct * 64 + threadIdx.x;
    sbox[threadIdx.x] = box[c];
    sarea[threadIdx.x] = area[c];
    if (kUseClass) scls[threadIdx.x] = cls[c];
  }
  __syncthreads();
  const int r = rt * 64 + threadIdx.x;
  if (r >= N) return;
  const float4 b = box[r];
  const float ar = area[r];
  const float cr = kUseClass ? cls[r] : 0.f;
  const float thresh_f = (float)thresh;
  unsigned long long bits = 0ull;
  const int start = rt == ct ? threadIdx.x + 1 : 0;
  for (int j = start; j < col_n; ++j) {
    if (kUseClass && scls[j] != cr) continue;
    const float iou = iou_plus1(b, ar, sbox[j], sarea[j]);
    const bool sup = mode == 0 ? ((double)iou >= thresh) : (iou > thresh_f);
    if (sup) bits |= 1ull << j;
  }
  mask[(size_t)r * W + ct] = bits;
}

__global__ void __launch_bounds__(kScanThreads) nms_scan_kernel(int N, int W, const unsigned long long *__restrict__ mask,
                                                                const int *__restrict__ order, int32_t *__restrict__ keep,
                                                                int32_t *__restrict__ num_keep) {
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  unsigned long long *remv = reinterpret_cast<unsigned long long *>(dyn_smem);  // [W]
  __shared__ unsigned long long diag[64];
  __shared__ unsigned long long sm_alive;
  __shared__ int scan_smem[kScanThreads / 32 + 1];
  __shared__ int sm_carry;
  for (int w = threadIdx.x; w < W; w += blockDim.x) remv[w] = 0ull;
  __syncthreads();
  for (int c = 0; c < W; ++c) {
    const int m = min(64, N - c * 64);
    if ((int)threadIdx.x < m) diag[threadIdx.x] = mask[(size_t)(c * 64 + threadIdx.x) * W + c];
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned long long cur = remv[c], alive = 0ull;
      for (int t = 0; t < m; ++t)
        if (!((cur >> t) & 1ull)) {
          alive |= 1ull << t;
          cur |= diag[t];
        }
      remv[c] = cur;
      sm_alive = alive;
    }
    __syncthreads();
    const unsigned long long alive = sm_alive;
    for (int w = c + 1 + threadIdx.x; w < W; w += blockDim.x) {
      unsigned long long acc = 0ull, rem = alive;
      while (rem) {
        const int t = __ffsll((long long)rem) - 1;
        rem &= rem - 1;
        acc |= mask[(size_t)(c * 64 + t) * W + w];
      }
      remv[w] |= acc;
    }
    __syncthreads();
  }
  // ordered compaction of the survivors (sorted position order == score order)
  if (threadIdx.x == 0) sm_carry = 0;
  __syncthreads();
  for (int base = 0; base < N; base += blockDim.x) {
    const int i = base + threadIdx.x;
    const int kept = (i < N && !((remv[i >> 6] >> (i & 63)) & 1ull)) ? 1 : 0;
    int total;
    const int ex = block_scan_excl(kept, scan_smem, &total);
    const int carry = sm_carry;
    if (kept) keep[carry + ex] = order ? order[i] : i;
    __syncthreads();
    if (threadIdx.x == 0) sm_carry = carry + total;
    __syncthreads();
  }
  if (threadIdx.x == 0) *num_keep = sm_carry;
}


// ====================================================================================================
// Chunked pipeline (default)
// ====================================================================================================
constexpr int kChunk = 1024;        // boxes per greedy step
constexpr int kSlab = 128;          // kept boxes per cull CTA
constexpr int kCullThreads = 1024;  // one chunk box per thread (32 warps keep the issue slots of the SM busy)
constexpr int kTileKeys = 4096;     // keys per CTA of the multi-CTA bitonic sort

struct NmsPipe {
  unsigned long long *keys;  // npad sort keys
  int *order;                // N sorted position -> original row
  float4 *box;               // N sorted boxes
  float *area, *cls;         // N
  unsigned char *dead;       // N: suppressed by an earlier kept box
  float4 *kbox;              // kept boxes, score order (x2 / y2 are stored as x2 + 1 / y2 + 1 rounded up, see cull)
  float4 *kraw;              // kept boxes as given
  float *karea, *kcls;
  int *nkept;                // [0] number of kept boxes so far
  size_t bytes;
};

NmsPipe carve_pipe(void *base, int N) {
  NmsPipe w;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off = align_up(off + bytes, 256);
    return (char *)base + o;
  };
  const size_t npad = (size_t)next_pow2(N < 2 ? 2 : N);
  w.keys = (unsigned long long *)take(sizeof(unsigned long long) * npad);
  w.order = (int *)take(sizeof(int) * (size_t)N);
  w.box = (float4 *)take(sizeof(float4) * (size_t)N);
  w.area = (float *)take(sizeof(float) * (size_t)N);
  w.cls = (float *)take(sizeof(float) * (size_t)N);
  w.dead = (unsigned char *)take((size_t)N);
  w.kbox = (float4 *)take(sizeof(float4) * (size_t)N);
  w.kraw = (float4 *)take(sizeof(float4) * (size_t)N);
  w.karea = (float *)take(sizeof(float) * (size_t)N);
  w.kcls = (float *)take(sizeof(float) * (size_t)N);
  w.nkept = (int *)take(256);
  w.bytes = off;
  return w;
}

// ---- multi-CTA bitonic sort of npad (power of two, >= kTileKeys) 64-bit keys, ascending ----
__global__ void __launch_bounds__(1024) nms_keys_kernel(const float *__restrict__ dets, int N, int dim, int npad,
                                                        unsigned long long *__restrict__ keys, unsigned char *__restrict__ dead,
                                                        int *__restrict__ nkept) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) *nkept = 0;
  if (i < N) dead[i] = 0;
  if (i >= npad) return;
  unsigned long long k = ~0ull;
  if (i < N) k = ((unsigned long long)(~float_order_key(dets[(size_t)i * dim + 4])) << 32) | (0xffffffffu - (unsigned)i);
  keys[i] = k;
}

// compare-exchange steps j = j_hi .. 1 of merge stage k on a tile of kTileKeys keys held in shared memory; with
// full != 0 the tile is sorted from scratch first (stages 2 .. kTileKeys)
__global__ void __launch_bounds__(1024) nms_bitonic_tile_kernel(unsigned long long *__restrict__ keys, int k_stage, int full) {
  __shared__ unsigned long long sk[kTileKeys];
  const int base = blockIdx.x * kTileKeys;
  for (int i = threadIdx.x; i < kTileKeys; i += blockDim.x) sk[i] = keys[base + i];
  __syncthreads();
  for (int k = full ? 2 : k_stage; k <= k_stage; k <<= 1) {
    for (int j = min(k >> 1, kTileKeys >> 1); j > 0; j >>= 1) {
      for (int q = threadIdx.x; q < (kTileKeys >> 1); q += blockDim.x) {
        const int lo = ((q & ~(j - 1)) << 1) | (q & (j - 1));
        const int hi = lo | j;
        const bool up = ((base + lo) & k) == 0;
        const unsigned long long x = sk[lo], y = sk[hi];
        if ((x > y) == up) {
          sk[lo] = y;
          sk[hi] = x;
        }
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < kTileKeys; i += blockDim.x) keys[base + i] = sk[i];
}

// one compare-exchange pass with stride j >= kTileKeys of merge stage k over global memory
__global__ void __launch_bounds__(256) nms_bitonic_global_kernel(unsigned long long *__restrict__ keys, int npad, int k, int j) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= (npad >> 1)) return;
  const int lo = ((q & ~(j - 1)) << 1) | (q & (j - 1));
  const int hi = lo | j;
  const bool up = (lo & k) == 0;
  const unsigned long long x = keys[lo], y = keys[hi];
  if ((x > y) == up) {
    keys[lo] = y;
    keys[hi] = x;
  }
}

__global__ void nms_order_kernel(const unsigned long long *__restrict__ keys, int N, int *__restrict__ order) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) order[i] = (int)(0xffffffffu - (unsigned)(keys[i] & 0xffffffffull));
}

// single-CTA variant for N <= kSortSmemKeys: keys, sort and order in one launch (also clears the pipeline's state)
__global__ void __launch_bounds__(kSortThreads) nms_sort_small_kernel(const float *__restrict__ dets, int N, int dim, int npad,
                                                                       int class_col, int *__restrict__ order,
                                                                       unsigned char *__restrict__ dead, int *__restrict__ nkept,
                                                                       float4 *__restrict__ box, float *__restrict__ area,
                                                                       float *__restrict__ cls) {
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  unsigned long long *keys = reinterpret_cast<unsigned long long *>(dyn_smem);
  // the first resolve kernel is a chained launch (launch_chained): let its cluster become resident beside this one CTA
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (threadIdx.x == 0) *nkept = 0;
  for (int i = threadIdx.x; i < npad; i += blockDim.x) {
    unsigned long long k = ~0ull;
    if (i < N) {
      k = ((unsigned long long)(~float_order_key(dets[(size_t)i * dim + 4])) << 32) | (0xffffffffu - (unsigned)i);
      dead[i] = 0;
    }
    keys[i] = k;
  }
  __syncthreads();
  bitonic_sort_u64(keys, npad);
  for (int i = threadIdx.x; i < N; i += blockDim.x) {  // order + the gather of nms_gather_kernel, no second launch
    const int src = (int)(0xffffffffu - (unsigned)(keys[i] & 0xffffffffull));
    order[i] = src;
    const float *d = dets + (size_t)src * dim;
    const float x1 = d[0], y1 = d[1], x2 = d[2], y2 = d[3];
    box[i] = make_float4(x1, y1, x2, y2);
    area[i] = fmul(fadd(fsub(x2, x1), 1.0f), fadd(fsub(y2, y1), 1.0f));  // cpu_nms.pyx:24
    if (class_col >= 0) cls[i] = d[class_col];
  }
}

__global__ void nms_clear_kernel(int N, unsigned char *__restrict__ dead, int *__restrict__ nkept) {  // presorted input
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) *nkept = 0;
  if (i < N) dead[i] = 0;
}

// The reference's suppression rule on one pair (cpu_nms.pyx:57-65 / nms_kernel.cu:24-32,71); a = the kept box.
__device__ __forceinline__ bool nms_suppresses(float4 a, float area_a, float4 b, float area_b, double thresh, float thresh_f,
                                               int mode) {
  const float iou = iou_plus1(a, area_a, b, area_b);
  return mode == 0 ? ((double)iou >= thresh) : (iou > thresh_f);
}

// Dense evaluation of a warp's candidate pairs.  Every lane holds `cand`, the columns (bits) of ITS row that passed the
// cheap filter; a warp that ran the exact test inside the filter loop would take that divergent branch whenever ANY
// lane has a hit -- with 4 % of the pairs overlapping that is 3 iterations out of 4, two lanes busy.  Instead the
// (row lane, column bit) pairs are compacted into a per-warp queue and `test(row_lane, bit)` runs on full warps.
constexpr int kQueue = 256;
template <typename F>
__device__ __forceinline__ void warp_dense_pairs(unsigned cand, unsigned short *queue, F test) {
  const unsigned lane = lane_id();
  const int cnum = __popc(cand);
  const int incl = warp_scan_incl(cnum);
  const int total = __shfl_sync(kFullMask, incl, 31);
  for (int base = 0; base < total; base += kQueue) {
    int pos = incl - cnum - base;
    for (unsigned m = cand; m; m &= m - 1, ++pos)
      if (pos >= 0 && pos < kQueue) queue[pos] = (unsigned short)((lane << 5) | (unsigned)(__ffs(m) - 1));
    __syncwarp();
    const int qn = min(kQueue, total - base);
    for (int q = (int)lane; q < qn; q += 32) {
      const unsigned e = queue[q];
      test((int)(e >> 5), (int)(e & 31u));
    }
    __syncwarp();
  }
}

// Filter bits of one row against 32 staged columns (x1, y1, RU(x2 + 1), RU(y2 + 1)): four chained compares and one
// predicated OR per pair, no branch.  "Cannot overlap" folds the +1 of the pixel convention in, rounded UP
// (x2p = RU(x2 + 1) >= x2 + 1, so x2p_j < x1_i implies RN(RN(xx2 - xx1) + 1) <= 0, i.e. the reference's w is 0):
// conservative, never drops a pair the reference would suppress as long as the threshold is positive.
__device__ __forceinline__ unsigned filter_bits32(float4 rowp, const float4 *colp) {
  unsigned cand = 0u;
#pragma unroll
  for (int jj = 0; jj < 32; ++jj) {
    const float4 k = colp[jj];
    asm("{\n\t.reg .pred p;\n\t"
        "setp.ge.f32 p, %1, %2;\n\t"
        "setp.ge.and.f32 p, %3, %4, p;\n\t"
        "setp.ge.and.f32 p, %5, %6, p;\n\t"
        "setp.ge.and.f32 p, %7, %8, p;\n\t"
        "@p or.b32 %0, %0, %9;\n\t}"
        : "+r"(cand)
        : "f"(k.z), "f"(rowp.x), "f"(rowp.z), "f"(k.x), "f"(k.w), "f"(rowp.y), "f"(rowp.w), "f"(k.y), "r"(1u << jj));
  }
  return cand;
}

// Chunk [c0, c0 + kChunk) against the kept boxes [0, nk): CTA s (grid-stride) stages kept boxes [128 s, 128 s + 128)
// and the chunk in shared memory; warp w owns chunk boxes [128 w, 128 w + 128), 32 at a time (lane = box): filter bits
// against the slab 32 kept boxes at a time, then the dense exact tests; a suppressed box gets its dead flag.
template <bool kUseClass>
__global__ void __launch_bounds__(kCullThreads) nms_cull_kernel(int N, int c0, double thresh, int mode, int no_filter,
                                                                const float4 *__restrict__ box, const float *__restrict__ area,
                                                                const float *__restrict__ cls, const float4 *__restrict__ kbox,
                                                                const float4 *__restrict__ kraw, const float *__restrict__ karea,
                                                                const float *__restrict__ kcls, const int *__restrict__ nkept,
                                                                unsigned char *__restrict__ dead) {
  __shared__ float4 sb[kSlab], sr[kSlab];        // kept: filter form / as given
  __shared__ float sa[kSlab], sc[kSlab];
  __shared__ float4 cb[kChunk];                  // chunk boxes as given
  __shared__ float ca[kChunk], cc[kChunk];
  __shared__ unsigned char cdead[kChunk];
  __shared__ unsigned short queues[kCullThreads / 32][kQueue];
  // chained launch (launch_chained below): the CTAs of this grid were made resident while the previous kernel of the
  // chunk loop was still running; everything they read is that kernel's output, so they wait for it here, and only
  // then let the NEXT kernel of the chain become resident (one successor pending at a time)
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int nk = *nkept;
  if ((int)blockIdx.x * kSlab >= nk) return;
  const float thresh_f = (float)thresh;
  for (int q = threadIdx.x; q < kChunk; q += blockDim.x) {
    const int i = c0 + q;
    const bool on = i < N && dead[i] == 0;
    cb[q] = on ? box[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    ca[q] = on ? area[i] : 0.f;
    if (kUseClass) cc[q] = on ? cls[i] : 0.f;
    cdead[q] = on ? 0 : 1;
  }
  const unsigned warp = warp_id(), lane = lane_id();
  unsigned short *queue = queues[warp];
  for (int s0 = (int)blockIdx.x * kSlab; s0 < nk; s0 += (int)gridDim.x * kSlab) {
    const int ns = min(kSlab, nk - s0);
    __syncthreads();
    if ((int)threadIdx.x < kSlab) {
      const bool on = (int)threadIdx.x < ns;
      // padding entries can reach nothing: x1 = +inf fails "x2p_row >= x1"
      sb[threadIdx.x] = on ? kbox[s0 + threadIdx.x] : make_float4(__int_as_float(0x7f800000), 0.f, 0.f, 0.f);
      if (on) {
        sr[threadIdx.x] = kraw[s0 + threadIdx.x];
        sa[threadIdx.x] = karea[s0 + threadIdx.x];
        if (kUseClass) sc[threadIdx.x] = kcls[s0 + threadIdx.x];
      }
    }
    __syncthreads();
    constexpr int kRounds = kChunk / kCullThreads;  // 4 x 32 boxes per warp
    for (int u = 0; u < kRounds; ++u) {
      const int r0 = ((int)warp * kRounds + u) << 5;  // first chunk box of this round
      const int r = r0 + (int)lane;
      if (__all_sync(kFullMask, cdead[r] != 0)) continue;
      const float4 b = cb[r];
      const float4 bp = make_float4(b.x, b.y, __fadd_ru(b.z, 1.0f), __fadd_ru(b.w, 1.0f));
      for (int jw = 0; jw < ns; jw += 32) {
        unsigned cand = no_filter ? (ns - jw >= 32 ? 0xffffffffu : (1u << (ns - jw)) - 1u) : filter_bits32(bp, sb + jw);
        if (cdead[r]) cand = 0u;
        warp_dense_pairs(cand, queue, [&](int rl, int jj) {
          const int rr = r0 + rl, j = jw + jj;
          if (cdead[rr]) return;
          if (kUseClass && sc[j] != cc[rr]) return;
          if (nms_suppresses(sr[j], sa[j], cb[rr], ca[rr], thresh, thresh_f, mode)) cdead[rr] = 1;
        });
      }
    }
  }
  __syncthreads();
  for (int q = threadIdx.x; q < kChunk; q += blockDim.x) {
    const int i = c0 + q;
    if (i < N && cdead[q] && dead[i] == 0) dead[i] = 1;
  }
}

// Greedy resolve of the chunk's survivors and append to the kept list.  One thread-block CLUSTER of kResolveCluster
// CTAs: every CTA compacts the chunk (redundantly, it is 1024 flags) and takes every kResolveCluster-th strip of
// pair-test units; hits are OR-ed into the mask of CTA 0 through distributed shared memory (they are rare: a box
// suppresses a fraction of a box on average), and after one cluster barrier CTA 0 resolves and appends alone.  With
// one CTA the 528 units of a full first chunk took 30 us -- at N = 1000 that was most of the call.
constexpr int kResolveThreads = 1024;
// Launch as a programmatic dependent of the previous kernel in the stream (DSPMB_TUNE_NMS_PDL): the chunk loop of the
// standalone NMS is a chain of up to ~400 short dependent kernels, and a plain launch pays the grid launch latency
// between every two of them.  The kernels launched this way execute griddepcontrol.wait before they touch memory.
template <typename... KArgs, typename... Args>
static cudaError_t launch_chained(bool chained, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                  Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = chained ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

constexpr int kResolveCluster = 8;
template <bool kUseClass>
__global__ void __cluster_dims__(kResolveCluster, 1, 1) __launch_bounds__(kResolveThreads) nms_resolve_kernel(int N, int c0, double thresh, int mode, int no_filter,
                                                                      const float4 *__restrict__ box, const float *__restrict__ area,
                                                                      const float *__restrict__ cls, const int *__restrict__ order,
                                                                      const unsigned char *__restrict__ dead, float4 *__restrict__ kbox,
                                                                      float4 *__restrict__ kraw, float *__restrict__ karea,
                                                                      float *__restrict__ kcls, int *__restrict__ nkept,
                                                                      int32_t *__restrict__ keep, int32_t *__restrict__ num_keep,
                                                                      int last) {
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  // [sbox: kChunk float4][sbp: kChunk float4][sarea][scls][sidx: kChunk int][mask: S x W u64, S <= kChunk]
  float4 *sbox = reinterpret_cast<float4 *>(dyn_smem);
  float4 *sbp = sbox + kChunk;
  float *sarea = reinterpret_cast<float *>(sbp + kChunk);
  float *scls = sarea + kChunk;
  int *sidx = reinterpret_cast<int *>(scls + kChunk);
  unsigned long long *mask = reinterpret_cast<unsigned long long *>(sidx + kChunk);
  __shared__ int scan_smem[kResolveThreads / 32 + 1];
  __shared__ unsigned long long rowany[kChunk / 64], remv_sm[kChunk / 64], diagany[kChunk / 64];
  __shared__ unsigned short queues[kResolveThreads / 32][kQueue];
  asm volatile("griddepcontrol.wait;" ::: "memory");  // chained launch, see nms_cull_kernel
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  const int crank = (int)cluster.block_rank();
  const float thresh_f = (float)thresh;
  const int i = c0 + (int)threadIdx.x;
  const int alive = (i < N && dead[i] == 0) ? 1 : 0;
  int S;
  const int pos = block_scan_excl(alive, scan_smem, &S);
  if (alive) {
    const float4 b = box[i];
    sbox[pos] = b;
    sbp[pos] = make_float4(b.x, b.y, __fadd_ru(b.z, 1.0f), __fadd_ru(b.w, 1.0f));
    sarea[pos] = area[i];
    if (kUseClass) scls[pos] = cls[i];
    sidx[pos] = order ? order[i] : i;
  }
  if (threadIdx.x < kChunk / 64) {
    rowany[threadIdx.x] = 0ull;
    diagany[threadIdx.x] = 0ull;
  }
  __syncthreads();
  const int nk0 = *nkept;
  if (S > 0) {
    const int W = (S + 63) >> 6;
    // mask[r * W + w] bit t: survivor r suppresses survivor 64 w + t (> r).  Units of 32 rows x 32 columns (lane = row)
    // of the upper triangle, dealt round-robin to the warps: branch-free filter bits, then the dense exact tests
    // (warp_dense_pairs); hits are OR-ed into the 32-bit halves of the mask words with shared-memory atomics.
    const int ngroups = (S + 31) >> 5;
    const unsigned warp = warp_id(), lane = lane_id(), nwarps = blockDim.x >> 5;
    unsigned *mask32 = reinterpret_cast<unsigned *>(cluster.map_shared_rank(mask, 0));       // CTA 0 owns the mask
    unsigned *rowany32 = reinterpret_cast<unsigned *>(cluster.map_shared_rank(rowany, 0));
    unsigned *diagany32 = reinterpret_cast<unsigned *>(cluster.map_shared_rank(diagany, 0));
    unsigned short *queue = queues[warp];
    if (crank == 0)
      for (int q2 = threadIdx.x; q2 < S * W; q2 += blockDim.x) mask[q2] = 0ull;
    for (int q2 = S + (int)threadIdx.x; q2 < (ngroups << 5); q2 += blockDim.x)  // padding columns reach nothing
      sbp[q2] = make_float4(__int_as_float(0x7f800000), 0.f, 0.f, 0.f);
    cluster.sync();  // CTA 0's mask is cleared before anybody ORs into it
    const int nunits = ngroups * (ngroups + 1) / 2;
    for (int u = crank * (int)nwarps + (int)warp; u < nunits; u += kResolveCluster * (int)nwarps) {
      int rg = 0, rem = u;  // unit u -> (row group, column group >= row group), row-major over the triangle
      while (rem >= ngroups - rg) {
        rem -= ngroups - rg;
        ++rg;
      }
      const int cg = rg + rem;
      const int r = (rg << 5) + (int)lane, jb = cg << 5;
      unsigned cand = 0u;
      if (r < S) {
        cand = no_filter ? 0xffffffffu : filter_bits32(sbp[r], sbp + jb);
        if (S - jb < 32) cand &= (1u << (S - jb)) - 1u;
        if (cg == rg) cand &= lane == 31 ? 0u : ~0u << (lane + 1);  // j > r
      }
      warp_dense_pairs(cand, queue, [&](int rl, int jj) {
        const int rr = (rg << 5) + rl, j = jb + jj;
        if (kUseClass && scls[j] != scls[rr]) return;
        if (nms_suppresses(sbox[rr], sarea[rr], sbox[j], sarea[j], thresh, thresh_f, mode)) {
          atomicOr(&mask32[rr * 2 * W + cg], 1u << jj);
          atomicOr(&rowany32[rg], 1u << rl);
          if ((cg >> 1) == (rg >> 1)) atomicOr(&diagany32[rg], 1u << rl);  // suppresses inside its own 64-row block
        }
      });
    }
    cluster.sync();  // every hit has landed in CTA 0
    if (crank != 0) return;
    if (warp == 0) {
      // lane w owns word w of the removed set; per 64-row chunk only the rows that suppress something are visited
      unsigned long long remv = 0ull;
      for (int c = 0; c < W; ++c) {
        // Only rows that suppress something inside their own 64-row block form a serial chain (bit t of `cur` can only
        // be set by a row < t of the same block or by an earlier block); the other rows of the block that suppress
        // anything at all are live iff their bit is clear once the chain is through.
        unsigned long long cur = __shfl_sync(kFullMask, remv, c);
        for (unsigned long long cand = diagany[c]; cand; cand &= cand - 1) {
          const int t = __ffsll((long long)cand) - 1;
          if (!((cur >> t) & 1ull)) cur |= mask[((c << 6) + t) * W + c];
        }
        const unsigned long long live = rowany[c] & ~cur;
        if ((int)lane == c) remv = cur;
        if ((int)lane > c && (int)lane < W) {
          unsigned long long acc = 0ull;
          for (unsigned long long rem = live; rem; rem &= rem - 1) {
            const int t = __ffsll((long long)rem) - 1;
            acc |= mask[((c << 6) + t) * W + lane];
          }
          remv |= acc;
        }
      }
      if ((int)lane < W) remv_sm[lane] = remv;
    }
    __syncthreads();
  }
  else {
    cluster.sync();  // (same barrier count on the S == 0 path)
    cluster.sync();
    if (crank != 0) return;
  }
  // ordered append of the survivors that stay
  const int q = (int)threadIdx.x;
  const int stays = (q < S && !((remv_sm[q >> 6] >> (q & 63)) & 1ull)) ? 1 : 0;
  int added;
  const int kp = block_scan_excl(stays, scan_smem, &added);
  if (stays) {
    const int d = nk0 + kp;
    kbox[d] = sbp[q];
    kraw[d] = sbox[q];
    karea[d] = sarea[q];
    if (kUseClass) kcls[d] = scls[q];
    keep[d] = sidx[q];
  }
  if (threadIdx.x == 0) {
    *nkept = nk0 + added;
    if (last) *num_keep = nk0 + added;
  }
}

}  // namespace
}  // namespace dspmb

using namespace dspmb;

// The chunked pipeline needs O(N) bytes; the full-mask implementation (DSPMB_TUNE_NMS_PIPELINE = 0) N^2 / 8 more, so
// its workspace is only promised up to 32768 boxes (beyond that the knob is ignored).
constexpr int kFullMaskMaxN = 32768;
extern "C" size_t dspmb_nms_workspace_bytes(int N) {
  if (N <= 0) return 256;
  const size_t pipe = carve_pipe(nullptr, N).bytes;
  const size_t full = N <= kFullMaskMaxN ? carve(nullptr, N).bytes : 0;
  return pipe > full ? pipe : full;
}

static int nms_pipeline(const float *dets, int N, int dim, double thresh, int mode, int class_col, int presorted,
                        int32_t *keep, int32_t *num_keep, void *workspace, cudaStream_t stream) {
  NmsPipe w = carve_pipe(workspace, N);
  const int *order = nullptr;
  bool gathered = false;
  int launches = 0;
  if (!presorted) {
    const int npad = next_pow2(N < 2 ? 2 : N);
    ProfileScope _p(kSlotNmsSort, stream);
    if (npad <= kSortSmemKeys) {
      DSPMB_ENSURE_DYN_SMEM(nms_sort_small_kernel, kSortSmemKeys * 8);
      nms_sort_small_kernel<<<1, kSortThreads, sizeof(unsigned long long) * npad, stream>>>(
          dets, N, dim, npad, class_col, w.order, w.dead, w.nkept, w.box, w.area, w.cls);
      gathered = true;
      ++launches;
    } else {
      nms_keys_kernel<<<ceil_div(npad, 1024), 1024, 0, stream>>>(dets, N, dim, npad, w.keys, w.dead, w.nkept);
      const int tiles = npad / kTileKeys;
      nms_bitonic_tile_kernel<<<tiles, 1024, 0, stream>>>(w.keys, kTileKeys, 1);
      launches += 2;
      for (int k = 2 * kTileKeys; k <= npad; k <<= 1) {
        for (int j = k >> 1; j >= kTileKeys; j >>= 1, ++launches)
          nms_bitonic_global_kernel<<<ceil_div(npad >> 1, 256), 256, 0, stream>>>(w.keys, npad, k, j);
        nms_bitonic_tile_kernel<<<tiles, 1024, 0, stream>>>(w.keys, k, 0);
        ++launches;
      }
      nms_order_kernel<<<ceil_div(N, 256), 256, 0, stream>>>(w.keys, N, w.order);
      ++launches;
    }
    DSPMB_CUDA_TRY(cudaGetLastError());
    order = w.order;
  } else {
    nms_clear_kernel<<<ceil_div(N, 256), 256, 0, stream>>>(N, w.dead, w.nkept);
    ++launches;
    DSPMB_CUDA_TRY(cudaGetLastError());
  }
  if (!gathered) {
    ProfileScope _p(kSlotNmsGather, stream);
    nms_gather_kernel<<<ceil_div(N, 256), 256, 0, stream>>>(dets, N, dim, class_col, order, w.box, w.area, w.cls);
    ++launches;
  }
  DSPMB_CUDA_TRY(cudaGetLastError());
  // a non-positive threshold suppresses pairs that do not overlap at all: no geometric filter then
  const int no_filter = mode == 0 ? !(thresh > 0.0) : !((float)thresh >= 0.f);
  const size_t smem_res = (size_t)kChunk * (16 + 16 + 4 + 4 + 4) + sizeof(unsigned long long) * (size_t)kChunk * (kChunk / 64);
  DSPMB_ENSURE_DYN_SMEM(nms_resolve_kernel<true>, smem_res);
  DSPMB_ENSURE_DYN_SMEM(nms_resolve_kernel<false>, smem_res);
  const bool chained = tuning(DSPMB_TUNE_NMS_PDL) != 0;
  for (int c0 = 0; c0 < N; c0 += kChunk) {
    if (c0 > 0) {
      // at most c0 boxes can have been kept so far; CTAs whose slab lies beyond the actual count return at once
      int grid = ceil_div(c0, kSlab);
      if (grid > 4 * kNumSMs) grid = 4 * kNumSMs;
      ProfileScope _p(kSlotNmsTile, stream);
      if (class_col >= 0)
        DSPMB_CUDA_TRY(launch_chained(chained, nms_cull_kernel<true>, dim3(grid), dim3(kCullThreads), 0, stream, N, c0, thresh, mode,
                                      no_filter, (const float4 *)w.box, (const float *)w.area, (const float *)w.cls,
                                      (const float4 *)w.kbox, (const float4 *)w.kraw, (const float *)w.karea,
                                      (const float *)w.kcls, (const int *)w.nkept, w.dead));
      else
        DSPMB_CUDA_TRY(launch_chained(chained, nms_cull_kernel<false>, dim3(grid), dim3(kCullThreads), 0, stream, N, c0, thresh, mode,
                                      no_filter, (const float4 *)w.box, (const float *)w.area, (const float *)w.cls,
                                      (const float4 *)w.kbox, (const float4 *)w.kraw, (const float *)w.karea,
                                      (const float *)w.kcls, (const int *)w.nkept, w.dead));
    }
    const int last = c0 + kChunk >= N;
    launches += c0 > 0 ? 2 : 1;
    ProfileScope _p(kSlotNmsReduce, stream);
    if (class_col >= 0)
      DSPMB_CUDA_TRY(launch_chained(chained, nms_resolve_kernel<true>, dim3(kResolveCluster), dim3(kResolveThreads), smem_res, stream, N, c0,
                                    thresh, mode, no_filter, (const float4 *)w.box, (const float *)w.area, (const float *)w.cls,
                                    (const int *)order, (const unsigned char *)w.dead, w.kbox, w.kraw, w.karea, w.kcls, w.nkept,
                                    keep, num_keep, last));
    else
      DSPMB_CUDA_TRY(launch_chained(chained, nms_resolve_kernel<false>, dim3(kResolveCluster), dim3(kResolveThreads), smem_res, stream, N, c0,
                                    thresh, mode, no_filter, (const float4 *)w.box, (const float *)w.area, (const float *)w.cls,
                                    (const int *)order, (const unsigned char *)w.dead, w.kbox, w.kraw, w.karea, w.kcls, w.nkept,
                                    keep, num_keep, last));
  }
  DSPMB_CUDA_TRY(cudaGetLastError());
  note_launches(launches);
  return DSPMB_OK;
}

extern "C" int dspmb_nms_f32(const float *dets, int N, int dim, double thresh, int mode, int class_col,
                             int presorted, int32_t *keep, int32_t *num_keep, void *workspace,
                             size_t workspace_bytes, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  DSPMB_REQUIRE(N >= 0 && dim >= 5, "nms: need N >= 0 and dim >= 5 (got N=%d dim=%d)", N, dim);
  DSPMB_REQUIRE(num_keep != nullptr, "nms: num_keep is NULL");
  DSPMB_REQUIRE(class_col < dim && (class_col < 0 || class_col >= 5), "nms: class_col must be < 0 or in [5, dim)");
  DSPMB_REQUIRE(mode == 0 || mode == 1, "nms: mode must be 0 (cpu_nms, >=) or 1 (gpu_nms, >)");
  if (N == 0) {
    DSPMB_CUDA_TRY(cudaMemsetAsync(num_keep, 0, sizeof(int32_t), stream));
    return DSPMB_OK;
  }
  DSPMB_REQUIRE(dets && keep, "nms: NULL tensor");
  const size_t need = dspmb_nms_workspace_bytes(N);
  if (!workspace || workspace_bytes < need || ((uintptr_t)workspace & 255)) {
    set_error("nms: workspace must be 256-byte aligned and >= %zu bytes (got %zu)", need, workspace_bytes);
    return DSPMB_ERR_WORKSPACE;
  }
  if (tuning(DSPMB_TUNE_NMS_PIPELINE) != 0 || N > kFullMaskMaxN)
    return nms_pipeline(dets, N, dim, thresh, mode, class_col, presorted, keep, num_keep, workspace, stream);
  NmsWorkspace w = carve(workspace, N);
  const int W = ceil_div(N, 64);
  DSPMB_REQUIRE((size_t)W * 8 <= 200 * 1024, "nms: more than %d boxes are not supported", 200 * 1024 / 8 * 64);
  DSPMB_ENSURE_DYN_SMEM(nms_sort_kernel, kSortSmemKeys * 8);
  DSPMB_ENSURE_DYN_SMEM(nms_scan_kernel, 200 * 1024);
  const int *order = nullptr;
  if (!presorted) {
    const int npad = next_pow2(N < 2 ? 2 : N);
    const size_t smem = npad <= kSortSmemKeys ? sizeof(unsigned long long) * npad : 0;
    {
    ProfileScope _p(kSlotNmsSort, stream);
    nms_sort_kernel<<<1, kSortThreads, smem, stream>>>(dets, N, dim, npad, w.keys, w.order);
  }
    DSPMB_CUDA_TRY(cudaGetLastError());
    order = w.order;
  }
  {
    ProfileScope _p(kSlotNmsGather, stream);
    nms_gather_kernel<<<ceil_div(N, 256), 256, 0, stream>>>(dets, N, dim, class_col, order, w.box, w.area, w.cls);
  }
  DSPMB_CUDA_TRY(cudaGetLastError());
  dim3 grid(W, W);
  {
    ProfileScope _p(kSlotNmsMask, stream);
    if (class_col >= 0)
    nms_mask_kernel<true><<<grid, 64, 0, stream>>>(N, W, thresh, mode, w.box, w.area, w.cls, w.mask);
  else
    nms_mask_kernel<false><<<grid, 64, 0, stream>>>(N, W, thresh, mode, w.box, w.area, w.cls, w.mask);
  }
  DSPMB_CUDA_TRY(cudaGetLastError());
  {
    ProfileScope _p(kSlotNmsScan, stream);
    nms_scan_kernel<<<1, kScanThreads, sizeof(unsigned long long) * W, stream>>>(N, W, w.mask, order, keep, num_keep);
  }
  DSPMB_CUDA_TRY(cudaGetLastError());
  return DSPMB_OK;
}

extern "C" int dspmb_nms_host(int *keep_out, int *num_out, const float *boxes_host, int boxes_num, int boxes_dim,
                              float nms_overlap_thresh, int device_id) {
  DSPMB_REQUIRE(keep_out && num_out && (boxes_host || boxes_num == 0), "nms_host: NULL pointer");
  DSPMB_CUDA_TRY(cudaSetDevice(device_id));  // _set_device, cython/nms_kernel.cu:80-89
  if (boxes_num == 0) {
    *num_out = 0;
    return DSPMB_OK;
  }
  const size_t in_bytes = sizeof(float) * (size_t)boxes_num * boxes_dim;
  const size_t ws_bytes = dspmb_nms_workspace_bytes(boxes_num);
  char *dev = nullptr;
  const size_t in_off = 0, keep_off = align_up(in_bytes, 256), num_off = keep_off + align_up(sizeof(int) * (size_t)boxes_num, 256);
  const size_t ws_off = num_off + 256;
  DSPMB_CUDA_TRY(cudaMalloc(&dev, ws_off + ws_bytes));
  int rc = DSPMB_OK;
  cudaError_t e = cudaMemcpy(dev + in_off, boxes_host, in_bytes, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    rc = dspmb_nms_f32((const float *)(dev + in_off), boxes_num, boxes_dim, (double)nms_overlap_thresh, 1, -1, 1,
                       (int32_t *)(dev + keep_off), (int32_t *)(dev + num_off), dev + ws_off, ws_bytes, nullptr);
    if (rc == DSPMB_OK) e = cudaMemcpy(num_out, dev + num_off, sizeof(int), cudaMemcpyDeviceToHost);
    if (rc == DSPMB_OK && e == cudaSuccess)
      e = cudaMemcpy(keep_out, dev + keep_off, sizeof(int) * (size_t)*num_out, cudaMemcpyDeviceToHost);
  }
  cudaFree(dev);
  if (rc != DSPMB_OK) return rc;
  if (e != cudaSuccess) return cuda_fail(e, "nms_host copy");
  return DSPMB_OK;
}
