// Standalone greedy NMS for sm_100a -- replaces the py-faster-rcnn helpers the reference ships:
//   cpu_nms  cython/cpu_nms.pyx:17-68   (suppress iff double(iou) >= thresh; the parity target)
//   gpu_nms  cython/gpu_nms.pyx:16-31 + _nms / nms_kernel, cython/nms_kernel.cu:34-144 (suppress iff iou > thresh)
//   nms      detect/nms.py:24-58         (numpy, same rule as gpu_nms)
// All use the pixel "+1" IoU convention and process boxes in descending score order.
//
// The reference's gpu path argsorts on the host, cudaMallocs per call, computes the FULL N x N/64 mask (the
// lower-triangle early-out is commented out, nms_kernel.cu:39), copies the mask to the host and sweeps it there.
// Here everything stays on the device and on the caller's stream:
//   nms_sort_kernel     64-bit keys (~score | ~index) sorted by a bitonic network (shared memory up to 16K keys,
//                       global above) => descending score, ties to the higher index like a stable
//                       argsort()[::-1];
//   nms_gather_kernel   sorted boxes as float4 + precomputed fp32 areas (cpu_nms.pyx:24) + optional class;
//   nms_mask_kernel     upper-triangular 64 x 64 tiles only, column tile staged in shared memory;
//   nms_scan_kernel     one CTA: 64-row chunks; the diagonal words are resolved serially in registers, the rows
//                       of the surviving boxes are OR-ed into a shared-memory bit vector with all loads in flight
//                       at once; then an ordered compaction writes the kept original indices.
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

}  // namespace
}  // namespace dspmb

using namespace dspmb;

extern "C" size_t dspmb_nms_workspace_bytes(int N) {
  if (N <= 0) return 256;
  return carve(nullptr, N).bytes;
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
  const size_t need = carve(nullptr, N).bytes;
  if (!workspace || workspace_bytes < need || ((uintptr_t)workspace & 255)) {
    set_error("nms: workspace must be 256-byte aligned and >= %zu bytes (got %zu)", need, workspace_bytes);
    return DSPMB_ERR_WORKSPACE;
  }
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
