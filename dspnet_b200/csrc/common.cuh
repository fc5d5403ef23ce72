// Shared host/device helpers of libdspmb (sm_100a only).
#ifndef DSPMB_COMMON_CUH_
#define DSPMB_COMMON_CUH_

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>
#include <functional>

#include "../../include/dspmb.h"
#include "libm_compat.h"

namespace dspmb {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs; grids are sized against this
constexpr unsigned kFullMask = 0xffffffffu;

// ---- host-side error plumbing -------------------------------------------------------------------------
void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what);
void note_launches(int n);  // kernels the calling thread's last operator call enqueued (dspmb_last_launch_count)
int libm_fma_mode();  // 0 / 1, resolved from dspmb_set_libm_mode / host CPU flags
int tuning(int knob);  // current value of a DSPMB_TUNE_* knob

#define DSPMB_CUDA_TRY(expr)                              \
  do {                                                    \
    cudaError_t _e = (expr);                              \
    if (_e != cudaSuccess) return cuda_fail(_e, #expr);   \
  } while (0)

#define DSPMB_REQUIRE(cond, ...)    \
  do {                              \
    if (!(cond)) {                  \
      set_error(__VA_ARGS__);       \
      return DSPMB_ERR_BAD_ARG;     \
    }                               \
  } while (0)

// ---- per-kernel event timing (capi.cu) ----------------------------------------------------------------
enum KernelSlot {
  kSlotPrior = 0,
  kSlotDetStream,
  kSlotDetSort,
  kSlotDetNms,
  kSlotTargetStream,
  kSlotTargetMatch,
  kSlotNmsSort,
  kSlotNmsGather,
  kSlotNmsMask,
  kSlotNmsScan,
  kSlotDetCompact,
  kSlotDetPair,
  kSlotDetResolve,
  kSlotTargetFixup,
  kSlotNmsTile,
  kSlotNmsReduce,
  kSlotSoftmaxDet,
  kSlotLoss,
  kNumKernelSlots
};
extern std::atomic<bool> g_profile_on;
void profile_mark(int slot, cudaStream_t stream, bool begin);
struct ProfileScope {  // brackets one kernel launch when profiling is enabled
  int slot;
  cudaStream_t stream;
  ProfileScope(int s, cudaStream_t st) : slot(s), stream(st) {
    if (g_profile_on) profile_mark(slot, stream, true);
  }
  ~ProfileScope() {
    if (g_profile_on) profile_mark(slot, stream, false);
  }
};

// ---- opt-in to more than 48 KB of dynamic shared memory (capi.cu) ----------------------------------------
// cudaFuncAttributeMaxDynamicSharedMemorySize is a property of a kernel ON ONE DEVICE: it is set once per (kernel,
// device) -- `done` is the call site's bit mask of devices that have it -- and setting it twice is harmless, so the
// check needs no lock.
int ensure_dyn_smem(const void *kernel, int bytes, std::atomic<unsigned long long> &done);
#define DSPMB_ENSURE_DYN_SMEM(kernel, bytes)                                             \
  do {                                                                                   \
    static std::atomic<unsigned long long> dyn_smem_done_{0ull};                         \
    const int rc_ = ensure_dyn_smem((const void *)(kernel), (int)(bytes), dyn_smem_done_); \
    if (rc_ != DSPMB_OK) return rc_;                                                     \
  } while (0)

// ---- CUDA-graph cache for the multi-launch operators (capi.cu) ------------------------------------------
// `launch(s)` enqueues an operator's kernels on stream s.  A call whose key (every pointer, shape and parameter
// that reaches a kernel argument) is seen for the second time is captured once on a private stream and replayed
// with one cudaGraphLaunch from then on, which removes the per-launch gaps of a 4-kernel step.  Bypassed while
// per-kernel profiling is on, while `stream` is itself being captured, or with DSPMB_TUNE_GRAPH_CACHE = 0.
// An operator whose launches form a fork/join (detection v2: the sort kernel and the pair-test kernel both depend
// only on the stream kernel) receives a side stream and two events while it is being captured by the cache: work
// enqueued on `side` between fork() and join() becomes a parallel branch of the graph.  Outside our own capture
// (first sighting, profiling, a caller's capture) the side streams are null and fork()/join() are no-ops: everything
// runs in order on `stream`.
struct LaunchCtx {
  static constexpr int kSides = 2;
  cudaStream_t stream = nullptr;
  cudaStream_t side[kSides] = {nullptr, nullptr};
  cudaEvent_t ev_fork = nullptr, ev_join[kSides] = {nullptr, nullptr};
  mutable bool used[kSides] = {false, false};  // side streams that joined the capture
  mutable int launches = 0;  // kernels (and memset nodes) the operator enqueued; read back by dspmb_last_launch_count
  bool forked() const { return side[0] != nullptr; }
  cudaStream_t branch(int i) const { return side[i] ? side[i] : stream; }
  int fork() const;  // both side streams wait for everything enqueued on `stream` so far
  int fork_side(int i) const;  // side stream i waits for everything enqueued on `stream` so far
  int join() const;  // `stream` waits for both side streams
};
int graph_cached_launch(const void *key, size_t key_len, cudaStream_t stream,
                        const std::function<int(const LaunchCtx &)> &launch);

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// Workspace header shared by target/detection: word 0 = latched data-dependent status.
struct WsHeader {
  int status;
  int pad[3];
};

// ---- device helpers -----------------------------------------------------------------------------------
#ifdef __CUDACC__

// fp32 primitives that must round exactly like the reference's scalar C++ (no FMA contraction).
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }

// Streaming (read-once) 128-bit global load that does not allocate in L1.
__device__ __forceinline__ float4 ld_stream_f4(const float *p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}
__device__ __forceinline__ float2 ld_stream_f2(const float *p) {
  float2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ float ld_stream_f1(const float *p) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
// Streaming 128-bit store (write-once outputs).
__device__ __forceinline__ void st_stream_f4(float *p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ unsigned warp_id() { return threadIdx.x >> 5; }

__device__ __forceinline__ unsigned long long shfl_xor_u64(unsigned long long v, int m) {
  return __shfl_xor_sync(kFullMask, v, m);
}
__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long v) {
  // two REDUX instructions: maximum of the high words, then of the low words among the lanes that hold it
  const unsigned hi = (unsigned)(v >> 32), lo = (unsigned)v;
  const unsigned mhi = __reduce_max_sync(kFullMask, hi);
  const unsigned mlo = __reduce_max_sync(kFullMask, hi == mhi ? lo : 0u);
  return ((unsigned long long)mhi << 32) | mlo;
}
__device__ __forceinline__ int warp_sum_i32(int v) {
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(kFullMask, v, m);
  return v;
}
// Inclusive warp scan.
__device__ __forceinline__ int warp_scan_incl(int v) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int o = __shfl_up_sync(kFullMask, v, d);
    if ((int)lane_id() >= d) v += o;
  }
  return v;
}

// Exclusive block scan of one int per thread; `smem` needs blockDim.x/32 + 1 ints.  Returns the exclusive
// prefix of the calling thread and the block total through *total.  Contains two __syncthreads().
__device__ __forceinline__ int block_scan_excl(int v, int *smem, int *total) {
  const int incl = warp_scan_incl(v);
  const unsigned w = warp_id(), l = lane_id();
  const unsigned nw = (blockDim.x + 31) >> 5;
  if (l == 31) smem[w] = incl;
  __syncthreads();
  if (w == 0) {
    int s = l < nw ? smem[l] : 0;
    int si = warp_scan_incl(s);
    if (l < nw) smem[l] = si - s;
    if (l == 31) smem[nw] = si;
  }
  __syncthreads();
  const int base = smem[w];
  *total = smem[nw];
  return base + incl - v;
}

// Order-preserving map float -> uint32 (ascending); -0.0 is folded onto +0.0 so that keys compare exactly
// like the reference's `a > b` on floats (no NaNs expected on this path).
__device__ __forceinline__ uint32_t float_order_key(float f) {
  uint32_t u = __float_as_uint(f);
  if (u == 0x80000000u) u = 0;
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// IoU of MultiBoxTarget (operator/multibox_target-inl.h:153-161 with safe_divide :44-50): every step is an
// individually rounded fp32 op, min/max are the mshadow ternaries.
__device__ __forceinline__ float iou_target(float4 a, float4 g) {
  const float mr = a.z < g.z ? a.z : g.z;
  const float ml = a.x > g.x ? a.x : g.x;
  const float mb = a.w < g.w ? a.w : g.w;
  const float mt = a.y > g.y ? a.y : g.y;
  const float dw = fsub(mr, ml);
  const float dh = fsub(mb, mt);
  const float iw = 0.0f > dw ? 0.0f : dw;
  const float ih = 0.0f > dh ? 0.0f : dh;
  const float inter = fmul(iw, ih);
  const float area1 = fmul(fsub(a.z, a.x), fsub(a.w, a.y));
  const float area2 = fmul(fsub(g.z, g.x), fsub(g.w, g.y));
  const float uni = fsub(fadd(area1, area2), inter);
  if (uni == 0.0f) return 0.0f;
  return fdiv(inter, uni);
}

// IoU of MultiBoxDetection's NMS (operator/multibox_detection.cc:44-51).
__device__ __forceinline__ float iou_detection(float4 a, float4 b) {
  const float w = fmaxf(0.f, fsub(fminf(a.z, b.z), fmaxf(a.x, b.x)));
  const float h = fmaxf(0.f, fsub(fminf(a.w, b.w), fmaxf(a.y, b.y)));
  const float i = fmul(w, h);
  const float u = fsub(fadd(fmul(fsub(a.z, a.x), fsub(a.w, a.y)), fmul(fsub(b.z, b.x), fsub(b.w, b.y))), i);
  return u <= 0.f ? 0.f : fdiv(i, u);
}

// IoU of the Cython NMS helpers, pixel "+1" convention (cython/cpu_nms.pyx:57-63, nms_kernel.cu:24-32).
// area_* are (x2 - x1 + 1) * (y2 - y1 + 1) precomputed per box exactly as cpu_nms.pyx:24 does.
__device__ __forceinline__ float iou_plus1(float4 a, float area_a, float4 b, float area_b) {
  const float xx1 = a.x >= b.x ? a.x : b.x;
  const float yy1 = a.y >= b.y ? a.y : b.y;
  const float xx2 = a.z <= b.z ? a.z : b.z;
  const float yy2 = a.w <= b.w ? a.w : b.w;
  const float tw = fadd(fsub(xx2, xx1), 1.0f);
  const float th = fadd(fsub(yy2, yy1), 1.0f);
  const float w = 0.0f >= tw ? 0.0f : tw;
  const float h = 0.0f >= th ? 0.0f : th;
  const float inter = fmul(w, h);
  return fdiv(inter, fsub(fadd(area_a, area_b), inter));
}

#endif  // __CUDACC__
}  // namespace dspmb
#endif  // DSPMB_COMMON_CUH_
