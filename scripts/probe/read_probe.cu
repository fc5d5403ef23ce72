// Standalone bandwidth probe (not part of the library): what can a kernel that READS one detection batch worth of
// bytes (66 / 88 / 104 MB) reach on this GPU, launched back to back on buffers that rotate through more than the L2?
// Variants: grid-stride 128-bit loads (persistent and one-tile-per-CTA), cp.async.bulk (TMA 1-D) double-buffered
// persistent ring, the same with a 22 MB streaming write beside the read (the operator's `out` fill).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o read_probe read_probe.cu ; run on the GPU box.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ float4 ldg_stream(const float4 *p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

// grid-stride, UNROLL independent 128-bit loads per thread and iteration
template <int UNROLL>
__global__ void __launch_bounds__(256) ldg_kernel(const float4 *__restrict__ src, size_t n4, float *sink, float4 *wr, size_t w4) {
  float acc = 0.f;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + (UNROLL - 1) * stride < n4; i += UNROLL * stride) {
    float4 v[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) v[u] = ldg_stream(src + i + u * stride);
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) acc += v[u].x + v[u].y + v[u].z + v[u].w;
  }
  for (; i < n4; i += stride) { float4 v = ldg_stream(src + i); acc += v.x + v.y + v.z + v.w; }
  if (wr) for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < w4; j += stride) __stcs(wr + j, make_float4(-1.f, -1.f, -1.f, -1.f));
  if (acc == 123.456f) sink[0] = acc;
}

// one contiguous tile of `tile4` float4 per CTA (the shape of the operator's stream kernel), 128 threads
__global__ void __launch_bounds__(128) tile_kernel(const float4 *__restrict__ src, size_t n4, int tile4, float *sink) {
  const size_t base = (size_t)blockIdx.x * tile4;
  float acc = 0.f;
  float4 v[10];
#pragma unroll
  for (int u = 0; u < 10; ++u) { const size_t j = base + threadIdx.x + u * 128; v[u] = (u * 128 + (int)threadIdx.x < tile4 && j < n4) ? ldg_stream(src + j) : make_float4(0, 0, 0, 0); }
#pragma unroll
  for (int u = 0; u < 10; ++u) acc += v[u].x + v[u].y + v[u].z + v[u].w;
  if (acc == 123.456f) sink[0] = acc;
}

// persistent cp.async.bulk ring: STAGES tiles of TILE bytes in flight per CTA
template <int TILE, int STAGES>
__global__ void __launch_bounds__(128) bulk_kernel(const char *__restrict__ src, size_t bytes, float *sink) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bar[STAGES];
  const size_t ntiles = bytes / TILE;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(&bar[s])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](size_t tile, int s) {
    const unsigned b = (unsigned)__cvta_generic_to_shared(&bar[s]);
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem + (size_t)s * TILE);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(TILE) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d), "l"(src + tile * TILE), "r"(TILE), "r"(b) : "memory");
  };
  size_t t = blockIdx.x;
  if (threadIdx.x == 0)
    for (int s = 0; s < STAGES; ++s) if (t + (size_t)s * gridDim.x < ntiles) issue(t + (size_t)s * gridDim.x, s);
  float acc = 0.f;
  int it = 0;
  for (; t < ntiles; t += gridDim.x, ++it) {
    const int s = it % STAGES;
    const unsigned ph = (it / STAGES) & 1;
    const unsigned b = (unsigned)__cvta_generic_to_shared(&bar[s]);
    unsigned ok = 0;
    while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(b), "r"(ph) : "memory");
    const float4 *p = reinterpret_cast<const float4 *>(smem + (size_t)s * TILE);
    for (int j = threadIdx.x; j < TILE / 16; j += 128) { const float4 v = p[j]; acc += v.x + v.y + v.z + v.w; }
    __syncthreads();
    const size_t nt = t + (size_t)STAGES * gridDim.x;
    if (threadIdx.x == 0 && nt < ntiles) issue(nt, s);
  }
  if (acc == 123.456f) sink[0] = acc;
}

int main() {
  const int R = 4, K = 200;
  const size_t sizes[] = {66u << 20, 88u << 20, 104u << 20, 264u << 20};
  const size_t maxb = sizes[3];
  char *buf[R];
  for (int r = 0; r < R; ++r) { CK(cudaMalloc(&buf[r], maxb)); CK(cudaMemset(buf[r], 1, maxb)); }
  float *sink; CK(cudaMalloc(&sink, 64));
  float4 *wr; CK(cudaMalloc(&wr, 4 * (size_t)(22u << 20)));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaFuncSetAttribute(bulk_kernel<20480, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CK(cudaFuncSetAttribute(bulk_kernel<20480, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CK(cudaFuncSetAttribute(bulk_kernel<32768, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  auto timeit = [&](const char *name, size_t bytes, size_t extra_w, auto launch) {
    for (int i = 0; i < 20; ++i) launch(i % R);
    CK(cudaDeviceSynchronize());
    float best = 1e9f;
    for (int rep = 0; rep < 3; ++rep) {
      CK(cudaEventRecord(e0));
      for (int i = 0; i < K; ++i) launch(i % R);
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    const double us = best * 1e3 / K;
    printf("%-44s %4zu MB read%s  %7.2f us  %7.0f GB/s\n", name, bytes >> 20, extra_w ? " + 22 MB written" : "", us, (bytes + extra_w) / us * 1e-3);
  };
  for (size_t bytes : sizes) {
    const size_t n4 = bytes / 16;
    for (int cps : {2, 4, 8}) {
      char nm[96];
      snprintf(nm, sizeof nm, "ldg x4 persistent, %d CTAs/SM x 256 thr", cps);
      timeit(nm, bytes, 0, [&](int r) { ldg_kernel<4><<<148 * cps, 256>>>((const float4 *)buf[r], n4, sink, nullptr, 0); });
      snprintf(nm, sizeof nm, "ldg x8 persistent, %d CTAs/SM x 256 thr", cps);
      timeit(nm, bytes, 0, [&](int r) { ldg_kernel<8><<<148 * cps, 256>>>((const float4 *)buf[r], n4, sink, nullptr, 0); });
    }
    timeit("ldg x8 persistent 8 CTAs/SM + out fill", bytes, 22u << 20, [&](int r) { ldg_kernel<8><<<148 * 8, 256>>>((const float4 *)buf[r], n4, sink, wr + (size_t)r * ((22u << 20) / 16), (22u << 20) / 16); });
    timeit("one 20 KB tile per CTA (128 thr)", bytes, 0, [&](int r) { tile_kernel<<<(unsigned)((n4 + 1279) / 1280), 128>>>((const float4 *)buf[r], n4, 1280, sink); });
    timeit("bulk ring 20 KB x 2 stages, 4 CTAs/SM", bytes, 0, [&](int r) { bulk_kernel<20480, 2><<<148 * 4, 128, 2 * 20480>>>(buf[r], bytes, sink); });
    timeit("bulk ring 20 KB x 4 stages, 2 CTAs/SM", bytes, 0, [&](int r) { bulk_kernel<20480, 4><<<148 * 2, 128, 4 * 20480>>>(buf[r], bytes, sink); });
    timeit("bulk ring 32 KB x 3 stages, 2 CTAs/SM", bytes, 0, [&](int r) { bulk_kernel<32768, 3><<<148 * 2, 128, 3 * 32768>>>(buf[r], bytes, sink); });
    timeit("cudaMemsetAsync (write only)", bytes, 0, [&](int r) { CK(cudaMemsetAsync(buf[r], 1, bytes)); });
  }
  return 0;
}
