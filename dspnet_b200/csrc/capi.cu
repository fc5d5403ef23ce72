// Library-wide pieces of the C ABI: error state, libm-compatibility mode, status read-back, self-test hooks.
#include <stdarg.h>
#include <string.h>

#include <mutex>
#include <vector>

#include "common.cuh"

namespace dspmb {

static thread_local char g_error[512] = "";
static thread_local int g_last_launches = 0;
static std::atomic<int> g_libm_mode{-1};

void note_launches(int n) { g_last_launches = n; }

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what) {
  set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  return DSPMB_ERR_CUDA;
}

// glibc's ifunc resolvers for expf/logf pick the FMA build iff the CPU has usable FMA and AVX2
// (sysdeps/x86_64/fpu/multiarch/ifunc-fma.h); mirror that test for the host this process runs on.
static int detect_host_libm_mode() {
#if defined(__x86_64__)
  __builtin_cpu_init();
  return (__builtin_cpu_supports("fma") && __builtin_cpu_supports("avx2")) ? 1 : 0;
#else
  return 0;
#endif
}

constexpr int kNumTuning = DSPMB_NUM_TUNING;
// knobs may be set from one thread while another launches: relaxed atomics (each call reads a knob once)
static std::atomic<int> g_tuning[kNumTuning] = {{2}, {320}, {1024}, {8192}, {31}, {1}, {1}, {0}, {1}, {600}, {0}, {1}, {2}, {1}, {1}, {1}, {1}, {1}};
static const int kTuningMax[kNumTuning] = {256, 320, 1024, 8192, 31, 1, 1, 1, 1, 1 << 20, 1 << 20, 4, 2, 2, 1, 1, 1, 1};
int tuning(int knob) { return g_tuning[knob].load(std::memory_order_relaxed); }

int libm_fma_mode() {
  int m = g_libm_mode.load(std::memory_order_relaxed);
  if (m < 0) {
    m = detect_host_libm_mode();
    g_libm_mode.store(m, std::memory_order_relaxed);
  }
  return m;
}

std::atomic<bool> g_profile_on{false};
namespace {
struct ProfileRecord {
  int slot;
  cudaEvent_t begin, end;
};
std::mutex g_profile_mutex;
std::vector<ProfileRecord> g_profile_records;
const char *const kSlotNames[kNumKernelSlots] = {
    "prior_kernel",        "det_stream_kernel", "det_sort_kernel", "det_nms_kernel",  "target_stream_kernel",
    "target_match_kernel", "nms_sort_kernel",   "nms_gather_kernel", "nms_mask_kernel", "nms_scan_kernel",
    "det_compact_kernel",  "det_pair_kernel",   "det_resolve_kernel", "target_fixup_kernel", "nms_cull_kernel",
    "nms_resolve_kernel",  "softmax_det_kernel", "multibox_loss_kernel"};
}  // namespace

void profile_mark(int slot, cudaStream_t stream, bool begin) {
  std::lock_guard<std::mutex> lock(g_profile_mutex);
  if (begin) {
    ProfileRecord r;
    r.slot = slot;
    cudaEventCreate(&r.begin);
    cudaEventCreate(&r.end);
    cudaEventRecord(r.begin, stream);
    g_profile_records.push_back(r);
  } else {
    for (size_t i = g_profile_records.size(); i-- > 0;)
      if (g_profile_records[i].slot == slot) {
        cudaEventRecord(g_profile_records[i].end, stream);
        break;
      }
  }
}

namespace {
struct GraphEntry {
  std::vector<unsigned char> key;
  cudaGraphExec_t exec = nullptr;
  bool uncacheable = false;
  int device = 0;
  int launches = 0;
  unsigned long long last_use = 0;
};
constexpr size_t kGraphCacheEntries = 32;
constexpr int kMaxDevices = 64;
std::mutex g_graph_mutex;
std::vector<GraphEntry> g_graph_cache;
unsigned long long g_graph_clock = 0;
struct CaptureSet {  // private streams / events of one device, used only under g_graph_mutex
  cudaStream_t main = nullptr, side[LaunchCtx::kSides] = {nullptr, nullptr};
  cudaEvent_t ev_fork = nullptr, ev_join[LaunchCtx::kSides] = {nullptr, nullptr};
};
CaptureSet g_capture[kMaxDevices];
}  // namespace

int LaunchCtx::fork() const {
  if (!forked()) return DSPMB_OK;
  DSPMB_CUDA_TRY(cudaEventRecord(ev_fork, stream));
  for (int i = 0; i < kSides; ++i) {
    DSPMB_CUDA_TRY(cudaStreamWaitEvent(side[i], ev_fork, 0));
    used[i] = true;
  }
  return DSPMB_OK;
}
int LaunchCtx::fork_side(int i) const {
  if (!forked()) return DSPMB_OK;
  DSPMB_CUDA_TRY(cudaEventRecord(ev_fork, stream));
  DSPMB_CUDA_TRY(cudaStreamWaitEvent(side[i], ev_fork, 0));
  used[i] = true;
  return DSPMB_OK;
}
int LaunchCtx::join() const {
  if (!forked()) return DSPMB_OK;
  for (int i = 0; i < kSides; ++i) {
    if (!used[i]) continue;  // a side stream that never joined the capture cannot be recorded on
    DSPMB_CUDA_TRY(cudaEventRecord(ev_join[i], side[i]));
    DSPMB_CUDA_TRY(cudaStreamWaitEvent(stream, ev_join[i], 0));
  }
  return DSPMB_OK;
}

int ensure_dyn_smem(const void *kernel, int bytes, std::atomic<unsigned long long> &done) {
  int dev = 0;
  DSPMB_CUDA_TRY(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) {
    set_error("device ordinal %d is outside the supported range [0, 64)", dev);
    return DSPMB_ERR_BAD_ARG;
  }
  const unsigned long long bit = 1ull << dev;
  if (!(done.load(std::memory_order_acquire) & bit)) {
    DSPMB_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    done.fetch_or(bit, std::memory_order_release);
  }
  return DSPMB_OK;
}

int graph_cached_launch(const void *key_, size_t key_len, cudaStream_t stream,
                        const std::function<int(const LaunchCtx &)> &launch) {
  LaunchCtx direct;
  direct.stream = stream;
  struct Publish {  // whichever context ran last tells this thread how many kernels the call enqueued
    const LaunchCtx &c;
    ~Publish() { g_last_launches = c.launches; }
  } publish{direct};
  if (!tuning(DSPMB_TUNE_GRAPH_CACHE) || g_profile_on) return launch(direct);
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(stream, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone) {
    cudaGetLastError();
    return launch(direct);
  }
  int dev = 0;
  DSPMB_CUDA_TRY(cudaGetDevice(&dev));
  if (dev < 0 || dev >= kMaxDevices) return launch(direct);
  // the key also pins the device and every path-selection knob
  std::vector<unsigned char> key(key_len + sizeof(int) * (kNumTuning + 2));
  memcpy(key.data(), key_, key_len);
  int extra[kNumTuning + 2];
  for (int k = 0; k < kNumTuning; ++k) extra[k] = tuning(k);
  extra[kNumTuning] = dev;
  extra[kNumTuning + 1] = libm_fma_mode();
  memcpy(key.data() + key_len, extra, sizeof(extra));

  std::lock_guard<std::mutex> lock(g_graph_mutex);
  GraphEntry *hit = nullptr;
  for (auto &e : g_graph_cache)
    if (e.key == key) {
      hit = &e;
      break;
    }
  if (!hit) {  // first sighting: remember the key, run directly
    if (g_graph_cache.size() >= kGraphCacheEntries) {
      size_t lru = 0;
      for (size_t i = 1; i < g_graph_cache.size(); ++i)
        if (g_graph_cache[i].last_use < g_graph_cache[lru].last_use) lru = i;
      if (g_graph_cache[lru].exec) {
        // rare; the evicted graph may still be in flight on ITS device (not necessarily the current one)
        if (g_graph_cache[lru].device != dev) cudaSetDevice(g_graph_cache[lru].device);
        cudaDeviceSynchronize();
        cudaGraphExecDestroy(g_graph_cache[lru].exec);
        if (g_graph_cache[lru].device != dev) cudaSetDevice(dev);
      }
      g_graph_cache.erase(g_graph_cache.begin() + lru);
    }
    GraphEntry e;
    e.key = std::move(key);
    e.device = dev;
    e.last_use = ++g_graph_clock;
    g_graph_cache.push_back(std::move(e));
    return launch(direct);
  }
  hit->last_use = ++g_graph_clock;
  if (hit->uncacheable) return launch(direct);
  if (!hit->exec) {  // second sighting: capture on the private streams of this device
    CaptureSet &cs = g_capture[dev];
    if (!cs.main) {
      DSPMB_CUDA_TRY(cudaStreamCreateWithFlags(&cs.main, cudaStreamNonBlocking));
      DSPMB_CUDA_TRY(cudaEventCreateWithFlags(&cs.ev_fork, cudaEventDisableTiming));
      for (int i = 0; i < LaunchCtx::kSides; ++i) {
        // highest priority: the few latency-bound CTAs of a side branch take freed SM slots before the next wave of
        // the HBM-bound kernel on the main branch does
        int prio_lo = 0, prio_hi = 0;
        DSPMB_CUDA_TRY(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        DSPMB_CUDA_TRY(cudaStreamCreateWithPriority(&cs.side[i], cudaStreamNonBlocking, prio_hi));
        DSPMB_CUDA_TRY(cudaEventCreateWithFlags(&cs.ev_join[i], cudaEventDisableTiming));
      }
    }
    if (cudaStreamBeginCapture(cs.main, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
      cudaGetLastError();
      hit->uncacheable = true;
      return launch(direct);
    }
    LaunchCtx ctx;
    ctx.stream = cs.main;
    ctx.ev_fork = cs.ev_fork;
    for (int i = 0; i < LaunchCtx::kSides; ++i) {
      ctx.side[i] = cs.side[i];
      ctx.ev_join[i] = cs.ev_join[i];
    }
    const int rc = launch(ctx);
    hit->launches = ctx.launches;
    cudaGraph_t graph = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(cs.main, &graph);
    if (rc != DSPMB_OK || ce != cudaSuccess || !graph) {
      cudaGetLastError();
      if (graph) cudaGraphDestroy(graph);
      hit->uncacheable = true;
      return rc != DSPMB_OK ? rc : launch(direct);
    }
    cudaGraphExec_t exec = nullptr;
    const cudaError_t ie = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ie != cudaSuccess || !exec) {
      cudaGetLastError();
      hit->uncacheable = true;
      return launch(direct);
    }
    hit->exec = exec;
  }
  DSPMB_CUDA_TRY(cudaGraphLaunch(hit->exec, stream));
  direct.launches = hit->launches;
  return DSPMB_OK;
}

namespace {
__global__ void test_expf_kernel(const float *x, float *y, long n, int fma_build) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
    y[i] = libm::expf_glibc(x[i], fma_build != 0);
}
__global__ void test_logf_kernel(const float *x, float *y, long n, int fma_build) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
    y[i] = libm::logf_glibc(x[i], fma_build != 0);
}
}  // namespace
}  // namespace dspmb

using namespace dspmb;

extern "C" int dspmb_version(void) { return DSPMB_VERSION; }

extern "C" int dspmb_last_launch_count(void) { return g_last_launches; }

extern "C" const char *dspmb_last_error(void) { return g_error; }

extern "C" int dspmb_set_libm_mode(int mode) {
  const int m = mode < 0 ? detect_host_libm_mode() : (mode ? 1 : 0);
  g_libm_mode.store(m, std::memory_order_relaxed);
  return m;
}

extern "C" int dspmb_set_tuning(int knob, int value) {
  if (knob < 0 || knob >= kNumTuning) return -1;
  const int v = value < 0 ? 0 : (value > kTuningMax[knob] ? kTuningMax[knob] : value);
  return g_tuning[knob].exchange(v, std::memory_order_relaxed);
}

extern "C" int dspmb_status(const void *workspace, void *stream) {
  if (!workspace) {
    set_error("dspmb_status: workspace is NULL");
    return DSPMB_ERR_WORKSPACE;
  }
  int status = 0;
  DSPMB_CUDA_TRY(cudaMemcpyAsync(&status, workspace, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  DSPMB_CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
  switch (status) {
    case DSPMB_OK:
      break;
    case DSPMB_ERR_LABEL_PADDING:
      set_error("MultiBoxTarget: first padding label row is not all -1 (multibox_target.cc:98-101)");
      break;
    case DSPMB_ERR_MINING_CANDIDATES:
      set_error("MultiBoxTarget: fewer mining candidates than num_negative (multibox_target.cc:236)");
      break;
    default:
      set_error("device status %d", status);
  }
  return status;
}

extern "C" int dspmb_profile_enable(int on) {
  g_profile_on = on != 0;
  return on != 0 ? 1 : 0;
}

extern "C" int dspmb_profile_read(float *ms, int *launches, int max_slots) {
  std::lock_guard<std::mutex> lock(g_profile_mutex);
  for (ProfileRecord &r : g_profile_records) {
    float t = 0.f;
    if (cudaEventSynchronize(r.end) == cudaSuccess && cudaEventElapsedTime(&t, r.begin, r.end) == cudaSuccess &&
        r.slot < max_slots) {
      if (ms) ms[r.slot] += t;
      if (launches) launches[r.slot] += 1;
    }
    cudaEventDestroy(r.begin);
    cudaEventDestroy(r.end);
  }
  g_profile_records.clear();
  return kNumKernelSlots;
}

extern "C" const char *dspmb_profile_kernel_name(int slot) {
  return slot >= 0 && slot < kNumKernelSlots ? kSlotNames[slot] : "";
}

extern "C" int dspmb_test_expf(const float *x, float *y, long n, void *stream) {
  test_expf_kernel<<<kNumSMs * 4, 256, 0, (cudaStream_t)stream>>>(x, y, n, libm_fma_mode());
  DSPMB_CUDA_TRY(cudaGetLastError());
  return DSPMB_OK;
}

extern "C" int dspmb_test_logf(const float *x, float *y, long n, void *stream) {
  test_logf_kernel<<<kNumSMs * 4, 256, 0, (cudaStream_t)stream>>>(x, y, n, libm_fma_mode());
  DSPMB_CUDA_TRY(cudaGetLastError());
  return DSPMB_OK;
}
