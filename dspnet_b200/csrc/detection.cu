// MultiBoxDetection for sm_100a.
//
// Reference semantics (CPU operator, the parity target): operator/multibox_detection-inl.h:81-107 (`out = -1`)
// and operator/multibox_detection.cc:53-169:
//   pass 1 (:79-128)  per anchor ascending: (score,id) = max/argmax over classes 1..C-1 (strict >, from -1),
//                     score < threshold => background; survivors are decoded (variances, expf, clip) and
//                     appended in ANCHOR ORDER at row valid_count++;
//   sort   (:130-151) if valid_count >= 1 and 0 < nms_threshold <= 1: stable sort by score descending;
//                     only rows [0, nkeep) (nkeep = min(V, nms_topk>0 ? nms_topk : V)) are rewritten in sorted
//                     order, rows [nkeep, V) KEEP their pass-1 content;
//   NMS    (:153-167) greedy over all V rows in that order, same class (or force_suppress), iou >= threshold
//                     overwrites only the id with -1.
// The reference's own GPU kernel (operator/multibox_detection.cu:52-207) is one block per image with an
// atomicAdd compaction (non-deterministic order), a global-memory merge sort and one __syncthreads per NMS
// candidate, and it diverges from the CPU semantics; it is not followed.
//
// Structure here (three launches, all images in every grid; a repeated call replays them as one CUDA graph):
//   det_stream_kernel  HBM-bound: streams cls_prob (B,C,A) with 128-bit no-allocate loads, 4 anchors/thread
//   (reg / TMA         (register-resident variant: every load of a thread in flight before the first compare),
//    variants)         fills `out` with -1 (128-bit stores), decodes the survivors and stages their finished pass-1
//                      rows, order keys, classes and boxes in shared memory (block scan => anchor order), then
//                      writes them as contiguous runs into the tile's slots together with the tile's count.
//   det_sort_kernel    grid (B, 1 + parts), two roles that write disjoint rows:
//                      sort role (one CTA per image): rank base of every tile = prefix of the tile counts; keys of the
//                      V survivors staged in rank order; MSB-first radix select of the nms_topk best (ties resolved
//                      in rank order through ballot-count tables); register/shuffle bitonic sort of the selection;
//                      those rows gathered straight from the slots into the head [0, nkeep).
//                      rank role (the other CTAs): copy the tiles' runs to their pass-1 positions [nkeep, V).
//                      (Ranking inside the stream kernel with a decoupled look-back was measured instead: it made
//                      that kernel 10 us slower because the CTAs idle on the look-back round trips; a separate rank
//                      launch cost 8 us.)
//   det_nms_kernel     one CTA per (image, class) segment (one per image with force_suppress): ordered member
//                      list from ballot-count tables, boxes staged in shared memory, then
//                        n <= 320: upper-triangular bit mask built by 32x32 warp units in two phases (branch-free
//                                  4-compare overlap bits; the few overlapping pairs are queued and take the
//                                  division-free threshold test with its exact fallback band on dense lanes) + a
//                                  word-serial resolve that only visits rows that suppress something;
//                        larger:   64-row chunks: ballot mask + serial resolve + parallel sweep of the later rows.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace dspmb {
namespace {

constexpr int kStreamThreads = 128;
constexpr int kRegThreads = 128;      // CTA size of the register-resident stream kernel
constexpr int kSortThreads = 1024;
constexpr int kNmsThreads = 256;
constexpr int kKeySmemMax = 28672;    // slot keys (T * tile) that fit in shared memory next to the sort keys
constexpr int kNmsMaskRows = 320;     // NMS segments up to this size use the shared-memory bit mask (5 words/row)
static_assert(kNmsMaskRows <= 320, "unit_tab holds (row group, column group) in 4 bits each, 55 units");
constexpr int kNmsSmemRows = 1024;    // rows of a larger NMS segment staged in shared memory

// Optional device-side timeline (dspmb_debug_trace): per kernel the earliest CTA start and the latest CTA end in
// %globaltimer nanoseconds, so that the overlap of the graph's branches can be read off without a profiler.
__constant__ unsigned long long *g_trace = nullptr;  // constant bank: the disabled check costs one LDC, no global load
struct TraceScope {
  int slot;
  __device__ __forceinline__ explicit TraceScope(int s) : slot(s) {
    if (g_trace && threadIdx.x == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      atomicMin(g_trace + 2 * slot, t);
    }
  }
  __device__ __forceinline__ ~TraceScope() {
    if (g_trace && threadIdx.x == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      atomicMax(g_trace + 2 * slot + 1, t);
    }
  }
};

// Per-CTA phase stamps of the pair kernel (dspmb_debug_trace with a second buffer): 10 x uint64 per CTA, written once.
__constant__ unsigned long long *g_stamps = nullptr;
__device__ __forceinline__ unsigned long long now_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define DSPMB_STAMP(k)                                                                          \
  do {                                                                                          \
    if (g_stamps && threadIdx.x == 0)                                                           \
      g_stamps[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 10 + (k)] = now_ns();            \
  } while (0)

__device__ __forceinline__ void trace_point(int slot) {  // latest time any CTA passed this point
  if (g_trace && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    atomicMax(g_trace + 2 * slot + 1, t);
    atomicMin(g_trace + 2 * slot, t);
  }
}

struct DetWorkspace {
  WsHeader *header;
  int *tile_count;            // (B, T) survivors per tile
  float *slot_rows;           // (B, Apad, 7) per-tile compacted pass-1 rows
  unsigned *slot_keys;        // (B, Apad)
  unsigned short *slot_cls;   // (B, cls_stride)
  float4 *slot_box;           // (B, Apad)
  int *valid;                 // (B) V
  int *nms_rows;              // (B) rows taking part in NMS (0 = skipped)
  int *cursor;                // (B) allocator of seg_list regions for large segments
  unsigned *keys;             // (B, Apad) order key of every surviving row, by rank (large A only)
  int *tile_base;             // (B, T) rank base of every tile (large T only)
  unsigned short *row_cls;    // (B, Apad) class of every output row
  float4 *row_box;            // (B, A) box of every output row
  int *seg_list;              // (B, A) member rows of large segments
  float4 *seg_box;            // (B, A) boxes of large segments in segment order
  unsigned char *seg_dead;    // (B, A)
  float *seg_area;            // (B, A) areas of large segments
  unsigned long long *sort_keys;  // (B, npad) spill for selections larger than the shared-memory budget
  // stream -> {sort || pair tests} pipeline (det_pair_kernel)
  float4 *cbox;               // (B, Apad) boxes of a tile's survivors grouped by class (rank order inside a class)
  unsigned short *crank;      // (B, cls_stride) their index inside the tile's run
  unsigned *tile_cls;         // (B, kV2ClsPad, Tmax) per (class, tile): offset | count << 16 inside the tile's run
  int *head_rank;             // (B, Apad) pass-1 rank of the row sorted to head position q
  int2 *head_list;            // (B, kHeadCap) (head position, pass-1 rank) of the head rows grouped by class, position order
  int *head_off;              // (B, kV2ClsPad + 1) start of every class in head_list
  size_t bytes;
};

constexpr int kV2ClsPad = 32;    // foreground classes the fork/join pipeline supports
constexpr int kV2MaxTiles = 512; // tiles per image whose bases fit its shared-memory tables
constexpr int kHeadCap = 1024;   // largest nms_topk whose head the sort kernel hands over as per-class lists


inline int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

// Layout is a pure function of (B, A, C); T and Apad are bounded with the smallest tile (128 anchors).
DetWorkspace carve(void *base, int B, int A, int C, int tiles_min = 0, size_t slots_min = 0) {
  DetWorkspace w;
  // the head-fed stream kernel cuts tiles at scale boundaries: a few more tiles and slots than A alone implies
  const size_t Tdef = (size_t)ceil_div(A, kStreamThreads);
  const size_t Tmax = Tdef > (size_t)tiles_min ? Tdef : (size_t)tiles_min;
  const size_t Adef = (((size_t)A + 3) & ~(size_t)3) + 4 * kStreamThreads;  // multiple of 4: 128-bit key loads
  const size_t Apad = Adef > slots_min ? Adef : slots_min;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off = align_up(off + bytes, 256);
    return (char *)base + o;
  };
  w.header = (WsHeader *)take(sizeof(WsHeader));
  w.tile_count = (int *)take(sizeof(int) * B * Tmax);
  w.slot_rows = (float *)take(sizeof(float) * 7 * B * Apad);
  w.slot_keys = (unsigned *)take(sizeof(unsigned) * B * Apad);
  w.slot_cls = (unsigned short *)take(sizeof(unsigned short) * B * ((Apad + 7) & ~(size_t)7));
  w.slot_box = (float4 *)take(sizeof(float4) * B * Apad);
  w.valid = (int *)take(sizeof(int) * B);
  w.nms_rows = (int *)take(sizeof(int) * B);
  w.cursor = (int *)take(sizeof(int) * B);
  w.keys = (unsigned *)take(sizeof(unsigned) * B * Apad);
  w.tile_base = (int *)take(sizeof(int) * B * Tmax);
  w.row_cls = (unsigned short *)take(sizeof(unsigned short) * B * ((Apad + 7) & ~(size_t)7));
  w.row_box = (float4 *)take(sizeof(float4) * (size_t)B * A);
  w.seg_list = (int *)take(sizeof(int) * (size_t)B * A);
  w.seg_box = (float4 *)take(sizeof(float4) * (size_t)B * A);
  w.seg_dead = (unsigned char *)take((size_t)B * A);
  w.seg_area = (float *)take(sizeof(float) * (size_t)B * A);
  w.sort_keys = (unsigned long long *)take(sizeof(unsigned long long) * (size_t)B * next_pow2(A));
  w.cbox = (float4 *)take(sizeof(float4) * B * Apad);
  w.crank = (unsigned short *)take(sizeof(unsigned short) * B * ((Apad + 7) & ~(size_t)7));
  w.tile_cls = (unsigned *)take(sizeof(unsigned) * (size_t)B * kV2ClsPad * Tmax);
  w.head_rank = (int *)take(sizeof(int) * B * Apad);
  w.head_list = (int2 *)take(sizeof(int2) * (size_t)B * kHeadCap);
  w.head_off = (int *)take(sizeof(int) * (size_t)B * (kV2ClsPad + 1));
  w.bytes = off;
  return w;
}

struct StreamArgs {
  const float *cls_prob, *loc_pred, *anchors;
  float *out;
  int *tile_count;            // (B, T) survivors per tile
  float *slot_rows;           // (B, Apad, 7) pass-1 rows of the survivors, compacted per tile (slot = tile_begin + k)
  unsigned *slot_keys;        // (B, Apad) their order keys
  unsigned short *slot_cls;   // (B, cls_stride)
  float4 *slot_box;           // (B, Apad)
  float4 *cbox;               // class-grouped copies for the pair-test kernel (fork/join pipeline only)
  unsigned short *crank;
  unsigned *tile_cls;
  int A, C, T, Apad, cls_stride;
  float threshold;
  int clip;
  float vx, vy, vw, vh;
  int fma_build;
  int prefetch;               // L2 prefetch distance in CTAs of launch order (0: none)
};

__device__ __forceinline__ float clip01(float v) {
  // std::max(0, std::min(1, v)) of multibox_detection.cc:121-125
  const float m = v < 1.f ? v : 1.f;
  return 0.f < m ? m : 0.f;
}

// Shared-memory staging of a tile's surviving rows.  Their ranks are consecutive, so rows, keys, classes and boxes
// each form ONE contiguous run in global memory: the CTA stages them and writes the runs with coalesced stores
// (per-row scattered 4-byte stores cost ten times the L2 write traffic).
template <int kRows>
struct RowStage {
  float rows[kRows * 7];
  float4 box[kRows];
  unsigned keys[kRows];
  unsigned short cls[kRows];
};

// Decode inputs of a tile's survivors in rank order (register-resident stream kernel).
template <int kRows>
struct DecodeStage {
  float4 an[kRows];
  float lp[5][kRows];
  float score[kRows];
  unsigned short id[kRows];
};

// Final pass-1 row of one surviving anchor (multibox_detection.cc:89-127), staged at its index inside the tile.
template <int kRows>
__device__ __forceinline__ void stage_row(const StreamArgs &a, RowStage<kRows> &sm, int local, int id, float score,
                                          float4 an, const float *lp) {
  const float aw = fsub(an.z, an.x), ah = fsub(an.w, an.y);
  const float ax = fdiv(fadd(an.x, an.z), 2.f), ay = fdiv(fadd(an.y, an.w), 2.f);
  const float ox = fadd(fmul(fmul(lp[0], a.vx), aw), ax);
  const float oy = fadd(fmul(fmul(lp[1], a.vy), ah), ay);
  const float ow = fdiv(fmul(libm::expf_glibc(fmul(lp[2], a.vw), a.fma_build), aw), 2.f);
  const float oh = fdiv(fmul(libm::expf_glibc(fmul(lp[3], a.vh), a.fma_build), ah), 2.f);
  const float oz = __double2float_rn(__dmul_rn((double)lp[4], 0.1));
  float x1 = fsub(ox, ow), y1 = fsub(oy, oh), x2 = fadd(ox, ow), y2 = fadd(oy, oh), z = oz;
  if (a.clip) {
    x1 = clip01(x1);
    y1 = clip01(y1);
    x2 = clip01(x2);
    y2 = clip01(y2);
    z = clip01(z);
  }
  float *o = sm.rows + local * 7;
  o[0] = (float)(id - 1);
  o[1] = score;
  o[2] = x1;
  o[3] = y1;
  o[4] = x2;
  o[5] = y2;
  o[6] = z;
  sm.keys[local] = ~float_order_key(score);  // ascending key == descending score
  sm.cls[local] = (unsigned short)(id - 1);
  sm.box[local] = make_float4(x1, y1, x2, y2);
}

// Coalesced write-out of the tile's `total` staged rows into the tile's slots (slot = tile_begin + k).  The rank of
// a row is only known once every earlier tile of the image has been counted; det_sort_kernel moves the runs to
// their final positions afterwards, so this kernel never waits on another CTA.  Call after a block barrier.
template <int kRows>
__device__ __forceinline__ void flush_rows(const StreamArgs &a, const RowStage<kRows> &sm, int b, int tile_begin, int total) {
  float *o = a.slot_rows + ((size_t)b * a.Apad + tile_begin) * 7;
  for (int q = threadIdx.x; q < total * 7; q += blockDim.x) o[q] = sm.rows[q];
  unsigned *gk = a.slot_keys + (size_t)b * a.Apad + tile_begin;
  unsigned short *gc = a.slot_cls + (size_t)b * a.cls_stride + tile_begin;
  float4 *gb = a.slot_box + (size_t)b * a.Apad + tile_begin;
  for (int q = threadIdx.x; q < total; q += blockDim.x) {
    gk[q] = sm.keys[q];
    gc[q] = sm.cls[q];
    gb[q] = sm.box[q];
  }
}

template <int VEC>
__global__ void __launch_bounds__(kStreamThreads) det_stream_kernel(const __grid_constant__ StreamArgs a) {
  __shared__ int scan_smem[kStreamThreads / 32 + 1];
  __shared__ RowStage<kStreamThreads * VEC> sm_rows;
  const int b = blockIdx.y, t = blockIdx.x;
  constexpr int kTile = kStreamThreads * VEC;
  const int tile_begin = t * kTile;
  const int i0 = tile_begin + threadIdx.x * VEC;
  const int A = a.A;
  const float *cp = a.cls_prob + (size_t)b * a.C * A;

  // ---- `out = -1` for this tile's rows (multibox_detection-inl.h:103) ----
  {
    float *ob = a.out + ((size_t)b * A + tile_begin) * 7;
    const int rows = min(kTile, A - tile_begin);
    const int nfl = rows * 7;
    if constexpr (VEC == 4) {
      const float4 m1 = make_float4(-1.f, -1.f, -1.f, -1.f);
      for (int q = threadIdx.x * 4; q < nfl; q += kStreamThreads * 4) *reinterpret_cast<float4 *>(ob + q) = m1;
    } else {
      for (int q = threadIdx.x; q < nfl; q += kStreamThreads) ob[q] = -1.f;
    }
  }

  // ---- class max / argmax over the foreground channels ----
  float score[VEC];
  int id[VEC];
#pragma unroll
  for (int k = 0; k < VEC; ++k) {
    score[k] = -1.f;
    id[k] = 0;
  }
  if (i0 < A) {
#pragma unroll 5
    for (int j = 1; j < a.C; ++j) {
      float v[VEC];
      if constexpr (VEC == 4) {
        const float4 q = ld_stream_f4(cp + (size_t)j * A + i0);
        v[0] = q.x;
        v[1] = q.y;
        v[2] = q.z;
        v[3] = q.w;
      } else {
        v[0] = ld_stream_f1(cp + (size_t)j * A + i0);
      }
#pragma unroll
      for (int k = 0; k < VEC; ++k)
        if (v[k] > score[k]) {
          score[k] = v[k];
          id[k] = j;
        }
    }
  }
  int nvalid = 0;
#pragma unroll
  for (int k = 0; k < VEC; ++k) {
    if (id[k] > 0 && score[k] < a.threshold) id[k] = 0;
    nvalid += id[k] > 0;
  }

  // ---- ordered compaction: rank = survivors before this anchor in the image ----
  int total;
  int pos = block_scan_excl(nvalid, scan_smem, &total);
  if (threadIdx.x == 0) a.tile_count[(size_t)b * a.T + t] = total;
  if (nvalid) {
    const float *loc = a.loc_pred + (size_t)b * A * 5;
#pragma unroll
    for (int k = 0; k < VEC; ++k)
      if (id[k] > 0) {
        const int i = i0 + k;
        const float4 an = __ldg(reinterpret_cast<const float4 *>(a.anchors) + i);
        const float *lsrc = loc + (size_t)i * 5;
        const float lp[5] = {__ldg(lsrc), __ldg(lsrc + 1), __ldg(lsrc + 2), __ldg(lsrc + 3), __ldg(lsrc + 4)};
        stage_row(a, sm_rows, pos, id[k], score[k], an, lp);
        ++pos;
      }
  }
  __syncthreads();  // the staged rows are complete
  flush_rows(a, sm_rows, b, tile_begin, total);
}

// ----------------------------------------------------------------------------------------------------
// Register-resident variant for a compile-time number of foreground classes NFG (20 = VOC, 8 = Cityscapes).
// Every load of the thread -- NFG class rows, the 4 anchors and the 4 loc_pred rows (20 floats) -- is issued
// before the first compare, so a thread has (NFG + 9) x 16 B in flight, the decode has no dependent memory round
// trip, and the ~150 registers/thread cap residency at 3 CTAs/SM: the grid runs in several waves whose load and
// store/decode phases overlap instead of all CTAs being resident and in lockstep.
template <int NFG, int kThreads>
__global__ void __launch_bounds__(kThreads) det_stream_reg_kernel(const __grid_constant__ StreamArgs a) {
  __shared__ int scan_smem[kThreads / 32 + 1];
  __shared__ RowStage<kThreads * 4> sm_rows;
  __shared__ DecodeStage<kThreads * 4> sm_in;
  const int b = blockIdx.y, t = blockIdx.x;
  constexpr int kTile = kThreads * 4;
  const int tile_begin = t * kTile;
  const int i0 = tile_begin + threadIdx.x * 4;
  const int A = a.A;
  const bool active = i0 < A;
  const float *cp = a.cls_prob + ((size_t)b * a.C + 1) * A + i0;

  float4 q[NFG], an[4], lp[5];
  if (active) {
#pragma unroll
    for (int j = 0; j < NFG; ++j) q[j] = ld_stream_f4(cp + (size_t)j * A);
    const float4 *ap = reinterpret_cast<const float4 *>(a.anchors) + i0;
#pragma unroll
    for (int k = 0; k < 4; ++k) an[k] = __ldg(ap + k);
    const float *lsrc = a.loc_pred + ((size_t)b * A + i0) * 5;
#pragma unroll
    for (int k = 0; k < 5; ++k) lp[k] = ld_stream_f4(lsrc + 4 * k);
  }
  {  // `out = -1` for this tile's rows (multibox_detection-inl.h:103)
    float *ob = a.out + ((size_t)b * A + tile_begin) * 7;
    const int nfl = min(kTile, A - tile_begin) * 7;
    const float4 m1 = make_float4(-1.f, -1.f, -1.f, -1.f);
    for (int x = threadIdx.x * 4; x < nfl; x += kThreads * 4) *reinterpret_cast<float4 *>(ob + x) = m1;
  }
  float score[4] = {-1.f, -1.f, -1.f, -1.f};
  int id[4] = {0, 0, 0, 0};
  if (active) {
#pragma unroll
    for (int j = 0; j < NFG; ++j) {
      const float v[4] = {q[j].x, q[j].y, q[j].z, q[j].w};
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (v[k] > score[k]) {
          score[k] = v[k];
          id[k] = j + 1;
        }
    }
  }
  int nvalid = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (id[k] > 0 && score[k] < a.threshold) id[k] = 0;
    nvalid += id[k] > 0;
  }
  int total;
  int pos = block_scan_excl(nvalid, scan_smem, &total);
  if (threadIdx.x == 0) a.tile_count[(size_t)b * a.T + t] = total;
  // Survivors are ~1 in 5 anchors, so decoding them where they sit would run the fp64-expf decode four times per
  // warp with a fifth of the lanes busy.  Their inputs are parked in shared memory in rank order instead (a few
  // predicated stores), and the decode runs once on dense lanes.
  if (nvalid) {
    const float lf[20] = {lp[0].x, lp[0].y, lp[0].z, lp[0].w, lp[1].x, lp[1].y, lp[1].z, lp[1].w, lp[2].x, lp[2].y,
                          lp[2].z, lp[2].w, lp[3].x, lp[3].y, lp[3].z, lp[3].w, lp[4].x, lp[4].y, lp[4].z, lp[4].w};
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (id[k] > 0) {
        sm_in.an[pos] = an[k];
#pragma unroll
        for (int c = 0; c < 5; ++c) sm_in.lp[c][pos] = lf[5 * k + c];
        sm_in.score[pos] = score[k];
        sm_in.id[pos] = (unsigned short)id[k];
        ++pos;
      }
  }
  __syncthreads();
  for (int j = threadIdx.x; j < total; j += kThreads) {
    const float l5[5] = {sm_in.lp[0][j], sm_in.lp[1][j], sm_in.lp[2][j], sm_in.lp[3][j], sm_in.lp[4][j]};
    stage_row(a, sm_rows, j, (int)sm_in.id[j], sm_in.score[j], sm_in.an[j], l5);
  }
  __syncthreads();  // the staged rows are complete
  flush_rows(a, sm_rows, b, tile_begin, total);
}

// ----------------------------------------------------------------------------------------------------
// Persistent, TMA-fed variant of the stream kernel (the default when A % 4 == 0).
//
// grid = #SMs x CTAs/SM; every CTA walks tiles of kPipeTile anchors (round-robin over image x tile) through a
// kStages-deep ring of shared-memory stages.  One elected thread feeds the ring with 1-D bulk async copies
// (cp.async.bulk ... mbarrier::complete_tx: the TMA engine, no registers, no per-thread address math): the C-1
// foreground class rows of the tile (kPipeTile * 4 B each, contiguous in the (B,C,A) layout) and the tile's
// loc_pred rows (kPipeTile * 20 B, contiguous).  The copy of tile i + kStages - 1 is issued before tile i is
// consumed, so HBM stays busy while the CTA does its argmax / compaction / decode / stores -- in the plain kernel
// all CTAs are resident at once and run their load and compute phases in lockstep, which idles DRAM during
// the tails.  The block barrier that ends a tile doubles as the "stage free" signal.
constexpr int kPipeThreads = 128;
constexpr int kPipeVec = 2;
constexpr int kPipeTile = kPipeThreads * kPipeVec;  // 256 anchors

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
  unsigned ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// One 2-D TMA tile copy (cp.async.bulk.tensor, SASS UTMALDG): box {kTile columns, NFG rows} of the (B*C, A) view of
// cls_prob at element coordinates (x = first anchor, y = first class row), completion on an mbarrier.
__device__ __forceinline__ void tma_tile_2d(void *dst, const CUtensorMap *tmap, int x, int y, unsigned long long *bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                   smem_u32(dst)),
               "l"(tmap), "r"(x), "r"(y), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void bulk_prefetch_l2(const void *src, unsigned bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}

struct PipeArgs {
  StreamArgs s;
  int num_tiles;    // B * T
  int stages;       // ring depth (2..4)
  int stage_floats; // floats per stage: (C-1) * kPipeTile class values + kPipeTile * 5 loc values
};

__global__ void __launch_bounds__(kPipeThreads) det_stream_tma_kernel(const __grid_constant__ PipeArgs p) {
  extern __shared__ __align__(128) unsigned char dyn_smem[];
  __shared__ __align__(8) unsigned long long full_bar[4];
  __shared__ int scan_smem[kPipeThreads / 32 + 1];
  __shared__ RowStage<kPipeTile> sm_rows;
  const StreamArgs &a = p.s;
  const int A = a.A, T = a.T, nfg = a.C - 1;
  float *ring = reinterpret_cast<float *>(dyn_smem);

  auto issue = [&](int tile_idx, int stage) {  // one thread: arm the barrier, then the bulk copies of one tile
    const int b = tile_idx / T, t = tile_idx - b * T;
    const int tile_begin = t * kPipeTile;
    const int rows = min(kPipeTile, A - tile_begin);
    float *dst = ring + (size_t)stage * p.stage_floats;
    mbar_expect_tx(&full_bar[stage], (unsigned)(rows * 4 * nfg + rows * 20));
    const float *cp = a.cls_prob + ((size_t)b * a.C + 1) * A + tile_begin;
    for (int j = 0; j < nfg; ++j) bulk_g2s(dst + j * kPipeTile, cp + (size_t)j * A, rows * 4, &full_bar[stage]);
    bulk_g2s(dst + nfg * kPipeTile, a.loc_pred + ((size_t)b * A + tile_begin) * 5, rows * 20, &full_bar[stage]);
  };

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) mbar_init(&full_bar[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages - 1; ++s) {
      const int idx = blockIdx.x + s * gridDim.x;
      if (idx < p.num_tiles) issue(idx, s);
    }
  }

  int it = 0;
  for (int tile_idx = blockIdx.x; tile_idx < p.num_tiles; tile_idx += gridDim.x, ++it) {
    const int stage = it % p.stages;
    // keep the ring full: the stage freed by the previous iteration receives tile it + stages - 1
    if (threadIdx.x == 0) {
      const int ahead = tile_idx + (p.stages - 1) * gridDim.x;
      if (ahead < p.num_tiles) issue(ahead, (it + p.stages - 1) % p.stages);
    }
    const int b = tile_idx / T, t = tile_idx - b * T;
    const int tile_begin = t * kPipeTile;
    const int rows = min(kPipeTile, A - tile_begin);
    const int l0 = threadIdx.x * kPipeVec;  // first anchor of this thread inside the tile

    // `out = -1` for this tile's rows (multibox_detection-inl.h:103); independent of the loads
    {
      float *ob = a.out + ((size_t)b * A + tile_begin) * 7;
      const int nfl = rows * 7;
      const float4 m1 = make_float4(-1.f, -1.f, -1.f, -1.f);
      for (int q = threadIdx.x * 4; q < nfl; q += kPipeThreads * 4) *reinterpret_cast<float4 *>(ob + q) = m1;
    }

    mbar_wait(&full_bar[stage], (unsigned)((it / p.stages) & 1));
    const float *cls = ring + (size_t)stage * p.stage_floats;
    const float *locs = cls + nfg * kPipeTile;

    float score[kPipeVec] = {-1.f, -1.f};
    int id[kPipeVec] = {0, 0};
    if (l0 < rows) {
#pragma unroll 4
      for (int j = 0; j < nfg; ++j) {
        const float2 v = *reinterpret_cast<const float2 *>(cls + j * kPipeTile + l0);
        if (v.x > score[0]) {
          score[0] = v.x;
          id[0] = j + 1;
        }
        if (v.y > score[1]) {
          score[1] = v.y;
          id[1] = j + 1;
        }
      }
    }
    int nvalid = 0;
#pragma unroll
    for (int k = 0; k < kPipeVec; ++k) {
      if (id[k] > 0 && score[k] < a.threshold) id[k] = 0;
      nvalid += id[k] > 0;
    }
    int total;
    int pos = block_scan_excl(nvalid, scan_smem, &total);
    if (threadIdx.x == 0) a.tile_count[(size_t)b * a.T + t] = total;
    if (nvalid) {
#pragma unroll
      for (int k = 0; k < kPipeVec; ++k)
        if (id[k] > 0) {
          const float4 an = __ldg(reinterpret_cast<const float4 *>(a.anchors) + tile_begin + l0 + k);
          stage_row(a, sm_rows, pos, id[k], score[k], an, locs + (l0 + k) * 5);
          ++pos;
        }
    }
    __syncthreads();  // the staged rows are complete
    flush_rows(a, sm_rows, b, tile_begin, total);
    __syncthreads();  // every thread is done with this stage and the row staging -> both may be refilled
  }
}

// ----------------------------------------------------------------------------------------------------
// One-tile-per-CTA variant fed by the TMA engine.  The register-resident kernel spends its load phase throttled by
// the load/store unit (29 LDG.128 per thread queue up behind each other) and holds ~140 registers per thread.
// Here one thread issues the tile's NFG class rows, its loc_pred rows and its anchors as 1-D bulk async copies
// (cp.async.bulk ... mbarrier::complete_tx) into shared memory; the CTA fills its `out = -1` rows while the bytes
// are in flight, then consumes the tile from shared memory.  ~40 registers and ~32 KB of shared memory per CTA
// give 7 CTAs/SM whose load and compute phases interleave freely (the hardware CTA scheduler replaces a software
// pipeline).  The row staging aliases the class rows, which are dead once the arg-max is done.
template <int NFG, int kTile, bool kLean = false>
struct BulkSmem {
  union {
    float cls[NFG][kTile];       // class rows of the tile (bulk copies)
    RowStage<kTile> rows;        // finished rows of the survivors (after the arg-max)
  } u;
  // kLean: loc_pred and the anchors are NOT staged -- only the survivors (one anchor in five) need them, and they
  // fetch their 5 + 4 floats from global memory themselves; the CTA then needs 22.7 KB instead of 31.7 KB of shared
  // memory, 9 CTAs fit on an SM instead of 7, and 9 KB less per tile go through the TMA engine.  `loc` shrinks to the
  // scratch the class table of the fork/join pipeline needs.
  float loc[kLean ? 192 : kTile * 5];
  float4 anc[kLean ? 1 : kTile];
  float score[kTile];            // survivors in rank order
  unsigned short idx[kTile], id[kTile];
};

// Second half of a stream CTA, shared by the cls_prob-fed and the head-fed kernels: `total` survivors are listed in
// anchor order in sm.score / sm.idx / sm.id; they are decoded on dense lanes, staged and flushed into the tile's slots
// (slot_begin + k), and -- fork/join pipeline -- copied once more grouped by class for the pair-test kernel.
template <int NFG, int kThreads, int kVec, bool kV2, bool kLean = false, bool kLocPlane = false, typename Smem>
__device__ __forceinline__ void finish_tile(const StreamArgs &a, Smem &sm, const int b, const int t, const int slot_begin,
                                            const int total, const int anchor_begin = 0) {
  __syncthreads();  // survivor list complete; the class rows are dead from here on
  for (int j = threadIdx.x; j < total; j += kThreads) {
    const int l = sm.idx[j];
    if constexpr (kLean) {
      const float *lp = a.loc_pred + ((size_t)b * a.A + anchor_begin + l) * 5;
      const float l5[5] = {__ldg(lp), __ldg(lp + 1), __ldg(lp + 2), __ldg(lp + 3), __ldg(lp + 4)};
      const float4 an = __ldg(reinterpret_cast<const float4 *>(a.anchors) + anchor_begin + l);
      stage_row(a, sm.u.rows, j, (int)sm.id[j], sm.score[j], an, l5);
    } else if constexpr (kLocPlane) {  // loc staged as [value][plane-order index] (head-fed kernel)
      // (gathering the survivors' loc values from the global planes instead -- five cold 32-byte sectors per survivor in
      // the decode phase -- measured 60.4 instead of 51.7 us)
      constexpr int kT = kThreads * kVec;
      const int q = sm.idxq[j];
      const float l5[5] = {sm.loc[q], sm.loc[kT + q], sm.loc[2 * kT + q], sm.loc[3 * kT + q], sm.loc[4 * kT + q]};
      stage_row(a, sm.u.rows, j, (int)sm.id[j], sm.score[j], sm.anc[l], l5);
    } else {
      const float l5[5] = {sm.loc[l * 5], sm.loc[l * 5 + 1], sm.loc[l * 5 + 2], sm.loc[l * 5 + 3], sm.loc[l * 5 + 4]};
      stage_row(a, sm.u.rows, j, (int)sm.id[j], sm.score[j], sm.anc[l], l5);
    }
  }
  __syncthreads();  // the staged rows are complete
  flush_rows(a, sm.u.rows, b, slot_begin, total);
  if constexpr (kV2) {
    // Class-grouped copy of the tile's boxes for the pair-test kernel: a stable counting sort of the <= kTile
    // survivors by class.  Lanes of a warp that hold the same class find each other with MATCH.ANY; the lowest of
    // them records the group's size per (round, warp); a 32-lane pass turns the table into offsets.  The table
    // aliases the loc_pred stage, which is dead once the rows are staged.
    static_assert(NFG <= kV2ClsPad, "tile_cls holds kV2ClsPad classes");
    constexpr int kWarps = kThreads / 32, kParts = kVec * kWarps;
    static_assert(sizeof(sm.loc) >= (kParts * 32 + 34) * sizeof(unsigned short) + kParts * sizeof(unsigned),
                  "class table does not fit in the loc stage");
    unsigned short(*cnt)[32] = reinterpret_cast<unsigned short(*)[32]>(sm.loc);
    unsigned short *coff = reinterpret_cast<unsigned short *>(sm.loc) + kParts * 32;  // [33]
    unsigned *present = reinterpret_cast<unsigned *>(coff + 34);                      // [kParts], 4-byte aligned
    const unsigned lane = lane_id(), warp = warp_id();
    int within[kVec], cc[kVec];
#pragma unroll
    for (int r = 0; r < kVec; ++r) {
      const int j = r * kThreads + (int)threadIdx.x;
      const bool valid = j < total;
      cc[r] = valid ? (int)sm.u.rows.cls[j] : 0xffff;
      const unsigned m = __match_any_sync(kFullMask, cc[r]);
      within[r] = __popc(m & ((1u << lane) - 1u));
      const bool leader = valid && within[r] == 0;
      if (leader) cnt[r * kWarps + warp][cc[r]] = (unsigned short)__popc(m);
      const unsigned pres = __reduce_or_sync(kFullMask, leader ? 1u << cc[r] : 0u);
      if (lane == 0) present[r * kWarps + warp] = pres;
    }
    __syncthreads();
    if (warp == 0) {
      int run = 0;
#pragma unroll
      for (int q = 0; q < kParts; ++q) {
        const int v = (present[q] >> lane) & 1u ? (int)cnt[q][lane] : 0;
        cnt[q][lane] = (unsigned short)run;
        run += v;
      }
      const int off = warp_scan_incl(run) - run;
      coff[lane] = (unsigned short)off;
      if ((int)lane < NFG) a.tile_cls[((size_t)b * kV2ClsPad + lane) * a.T + t] = (unsigned)off | ((unsigned)run << 16);
    }
    __syncthreads();
    float4 *gb = a.cbox + (size_t)b * a.Apad + slot_begin;
    unsigned short *gr = a.crank + (size_t)b * a.cls_stride + slot_begin;
#pragma unroll
    for (int r = 0; r < kVec; ++r) {
      const int j = r * kThreads + (int)threadIdx.x;
      if (j < total) {
        const int pos = (int)coff[cc[r]] + (int)cnt[r * kWarps + warp][cc[r]] + within[r];
        gb[pos] = sm.u.rows.box[j];
        gr[pos] = (unsigned short)j;
      }
    }
  }
}

template <int NFG, int kThreads, int kVec, bool kV2, bool kLean = false, bool kTensor = false>
__global__ void __launch_bounds__(kThreads) det_stream_bulk_kernel(const __grid_constant__ StreamArgs a,
                                                                   const __grid_constant__ CUtensorMap tmap) {
  constexpr int kTile = kThreads * kVec;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // the sort kernel may be a programmatic dependent
  TraceScope trace_(0);
  extern __shared__ __align__(128) unsigned char bulk_smem_raw[];
  BulkSmem<NFG, kTile, kLean> &sm = *reinterpret_cast<BulkSmem<NFG, kTile, kLean> *>(bulk_smem_raw);
  __shared__ __align__(8) unsigned long long full_bar;
  __shared__ int scan_smem[kThreads / 32 + 1];
  const int b = blockIdx.y, t = blockIdx.x;
  const int A = a.A;
  const int tile_begin = t * kTile;
  const int rows = min(kTile, A - tile_begin);  // multiple of 4 (A % 4 == 0)

  if (threadIdx.x == 0) {
    mbar_init(&full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if constexpr (kTensor) {
      // ONE tensor-map copy for the whole [NFG x kTile] class tile; columns beyond A are out of bounds of the map and
      // arrive as zeros, so the transaction count is always the full box
      static_assert(kLean, "the tensor-map variant stages the class rows only");
      mbar_expect_tx(&full_bar, (unsigned)(kTile * 4 * NFG));
      tma_tile_2d(&sm.u.cls[0][0], &tmap, tile_begin, b * a.C + 1, &full_bar);
    } else {
    mbar_expect_tx(&full_bar, (unsigned)(rows * 4 * NFG + (kLean ? 0 : rows * 20 + rows * 16)));
    const float *cp = a.cls_prob + ((size_t)b * a.C + 1) * A + tile_begin;
#pragma unroll 4
    for (int j = 0; j < NFG; ++j) bulk_g2s(&sm.u.cls[j][0], cp + (size_t)j * A, rows * 4, &full_bar);
    }
    if constexpr (!kLean) {
      bulk_g2s(sm.loc, a.loc_pred + ((size_t)b * A + tile_begin) * 5, rows * 20, &full_bar);
      bulk_g2s(sm.anc, a.anchors + (size_t)tile_begin * 4, rows * 16, &full_bar);
    }
  }
  {  // `out = -1` for this tile's rows (multibox_detection-inl.h:103) while the copies are in flight
    float *ob = a.out + ((size_t)b * A + tile_begin) * 7;
    const int nfl = rows * 7;
    const float4 m1 = make_float4(-1.f, -1.f, -1.f, -1.f);
    for (int x = threadIdx.x * 4; x < nfl; x += kThreads * 4) *reinterpret_cast<float4 *>(ob + x) = m1;
  }
  __syncthreads();  // the barrier initialisation is visible to every waiter
  mbar_wait(&full_bar, 0u);
  if (a.prefetch > 0 && threadIdx.x == 32) {
    // The CTAs of a wave start together, so they also load together and compute together: DRAM is saturated while the
    // wave's tiles arrive and IDLE while it computes (2.3 waves x ~3 us of a 24 us kernel).  Now that this CTA's own
    // tile has landed, one thread asks L2 for the class tile of the CTA `prefetch` launches ahead -- the one that takes
    // this CTA's place on the SM (prefetch = CTAs resident on the GPU) -- so that DRAM works through the compute
    // phase and the successor's TMA copy is an L2 hit.  (Issued at CTA start instead, the prefetch only doubled the
    // queue in front of the CTA's own tile: measured +2.5 us.)  No registers, no shared memory, nothing to wait for.
    const int lin = b * (int)gridDim.x + t + a.prefetch;
    const int pb = lin / (int)gridDim.x, pt = lin - pb * (int)gridDim.x;
    if (pb < (int)gridDim.y) {
      const int pbegin = pt * kTile;
      if constexpr (kTensor) {
        asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(&tmap), "r"(pbegin),
                     "r"(pb * a.C + 1)
                     : "memory");
      } else {
        const int prow = min(kTile, A - pbegin);
        const float *cp = a.cls_prob + ((size_t)pb * a.C + 1) * A + pbegin;
#pragma unroll 4
        for (int j = 0; j < NFG; ++j) bulk_prefetch_l2(cp + (size_t)j * A, prow * 4);
        if constexpr (!kLean) bulk_prefetch_l2(a.loc_pred + ((size_t)pb * A + pbegin) * 5, prow * 20);
      }
    }
  }

  const int l0 = threadIdx.x * kVec;
  float score[kVec];
  int id[kVec];
#pragma unroll
  for (int k = 0; k < kVec; ++k) {
    score[k] = -1.f;
    id[k] = 0;
  }
  if (l0 < rows) {
#pragma unroll
    for (int j = 0; j < NFG; ++j) {
      float v[kVec];
      if constexpr (kVec == 4) {
        const float4 q = *reinterpret_cast<const float4 *>(&sm.u.cls[j][l0]);
        v[0] = q.x, v[1] = q.y, v[2] = q.z, v[3] = q.w;
      } else {
        const float2 q = *reinterpret_cast<const float2 *>(&sm.u.cls[j][l0]);
        v[0] = q.x, v[1] = q.y;
      }
#pragma unroll
      for (int k = 0; k < kVec; ++k)
        if (v[k] > score[k]) {
          score[k] = v[k];
          id[k] = j + 1;
        }
    }
  }
  int nvalid = 0;
#pragma unroll
  for (int k = 0; k < kVec; ++k) {
    if (id[k] > 0 && score[k] < a.threshold) id[k] = 0;
    nvalid += id[k] > 0;
  }
  int total;
  int pos = block_scan_excl(nvalid, scan_smem, &total);
  if (threadIdx.x == 0) a.tile_count[(size_t)b * a.T + t] = total;
#pragma unroll
  for (int k = 0; k < kVec; ++k)
    if (id[k] > 0) {
      sm.score[pos] = score[k];
      sm.idx[pos] = (unsigned short)(l0 + k);
      sm.id[pos] = (unsigned short)id[k];
      ++pos;
    }
  finish_tile<NFG, kThreads, kVec, kV2, kLean>(a, sm, b, t, tile_begin, total, tile_begin);
}

// ----------------------------------------------------------------------------------------------------
// Head-fed stream kernel: MultiBoxDetection straight from the per-scale conv outputs.
// The reference graph (symbol/common.py:399-432, symbol/symbol_builder.py:156-165) runs, per scale, transpose ->
// Flatten, then Concat -> Reshape -> transpose -> SoftmaxActivation(mode='channel') and only then MultiBoxDetection:
// four full passes over the class tensor (66 MB at SSD-512, batch 32) before the operator reads it a fifth time.
// Here a CTA owns a tile of whole cells of one scale (<= 256 anchors): it reads the C logits and 5 loc values of
// every anchor of the tile directly from the NCHW heads (a thread walks one (cell, anchor-in-cell) pair, so that
// consecutive threads read consecutive cells of one channel plane: coalesced), keeps the logits in shared memory and
// never materialises cls_prob.
//   * every anchor: channel softmax in fp32 with MUFU.EX2 (relative error < 3e-5) -> largest foreground probability;
//     anchors whose approximate score is below threshold (1 - 1e-4) are below the threshold exactly as well;
//   * the others (the survivors and a thin band) are re-evaluated bit-exactly on dense lanes: glibc expf of every
//     class (one (candidate, class) pair per thread), fp32 sum in class order, one division -- the softmax
//     multibox_target.cc:220-231 spells out, which oracle_softmax_channel restates -- then the first-maximum arg-max on
//     the rounded probabilities and the exact threshold test of multibox_detection.cc:79-100;
//   * from there on the tile is finished exactly like one of the cls_prob-fed kernel (finish_tile).
constexpr int kMaxScales = 8;
constexpr int kHeadThreads = 128, kHeadTile = 256, kHeadBatch = 64;
struct HeadsArgs {
  const float *cls[kMaxScales], *loc[kMaxScales];   // (B, na*C, H, W) / (B, na*5, H, W)
  int hw[kMaxScales], na[kMaxScales], cpt[kMaxScales];  // cells per map, anchors per cell, cells per tile
  int tile_off[kMaxScales + 1];                     // first tile of every scale
  int anchor_off[kMaxScales];                       // first anchor of every scale
  int nscales;
};

template <int C>
struct HeadSmem {
  struct {
    RowStage<kHeadTile> rows;          // finished rows of the survivors
  } u;
  float loc[kHeadTile * 5];            // [value][plane-order index]
  float4 anc[kHeadTile];
  float score[kHeadTile];              // approximate score per anchor, then exact score per survivor (rank order)
  unsigned short idx[kHeadTile], id[kHeadTile];
  unsigned short idxq[kHeadTile];      // plane-order index of survivor j (loc is staged as [value][plane index])
  unsigned short cand[kHeadTile];      // candidates of the exact phase, anchor order
  float es[kHeadBatch][C];             // exact exponentials of one batch of candidates
  float cmx[kHeadBatch];
  float fscore[kHeadTile];             // exact score of every candidate (candidate order)
  unsigned short fid[kHeadTile];       // its class, 0 = below the threshold
};

template <int C, bool kV2>
__global__ void __launch_bounds__(kHeadThreads) det_stream_heads_kernel(const __grid_constant__ StreamArgs a,
                                                                        const __grid_constant__ HeadsArgs h) {
  constexpr int NFG = C - 1;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // the sort kernel may be a programmatic dependent
  TraceScope trace_(0);
  extern __shared__ __align__(128) unsigned char head_smem_raw[];
  HeadSmem<C> &sm = *reinterpret_cast<HeadSmem<C> *>(head_smem_raw);
  __shared__ int scan_smem[kHeadThreads / 32 + 1];
  const int b = blockIdx.y, t = blockIdx.x;
  int k = 0;
  while (k + 1 < h.nscales && t >= h.tile_off[k + 1]) ++k;
  const int na = h.na[k], HW = h.hw[k];
  const int cell0 = (t - h.tile_off[k]) * h.cpt[k];
  const int ncells = min(h.cpt[k], HW - cell0);
  const int n = ncells * na;                       // anchors of this tile
  const int a0 = h.anchor_off[k] + cell0 * na;     // first anchor of the tile
  const float *cls = h.cls[k] + (size_t)b * na * C * HW + cell0;
  const float *loc = h.loc[k] + (size_t)b * na * 5 * HW + cell0;

  {  // `out = -1` for this tile's rows (multibox_detection-inl.h:103)
    float *ob = a.out + ((size_t)b * a.A + a0) * 7;
    if (((uintptr_t)ob & 15) == 0 && (n & 3) == 0) {
      const float4 m1 = make_float4(-1.f, -1.f, -1.f, -1.f);
      for (int x = threadIdx.x * 4; x < n * 7; x += kHeadThreads * 4) *reinterpret_cast<float4 *>(ob + x) = m1;
    } else {
      for (int x = threadIdx.x; x < n * 7; x += kHeadThreads) ob[x] = -1.f;
    }
  }
  for (int j = threadIdx.x; j < n; j += kHeadThreads)
    sm.anc[j] = __ldg(reinterpret_cast<const float4 *>(a.anchors) + a0 + j);
  const bool all_cand = !(a.threshold > 1e-30f);
  const float thr_lo = fmul(a.threshold, 1.0f - 1e-4f);
  // ---- loads + approximate softmax ----
  // The logits of the tile live in shared memory as xs[class][ai * ncells + cell] ("plane order": the order of the
  // NCHW heads, so global loads and shared stores are both contiguous over the cells); anchor j = cell * na + ai.
  // A thread takes two neighbouring cells of one anchor-in-cell plane (64-bit loads) when the tile allows it.
  const bool pairs = ((ncells | cell0 | HW) & 1) == 0 && (((uintptr_t)h.cls[k] | (uintptr_t)h.loc[k]) & 7) == 0;
  if (pairs) {
    const int half = ncells >> 1;
    for (int p = threadIdx.x; p < half * na; p += kHeadThreads) {
      const int ai = p / half, cell = (p - ai * half) * 2;
      const float *pc = cls + (size_t)ai * C * HW + cell;
      float2 x[C];
#pragma unroll
      for (int c = 0; c < C; ++c) {
        x[c] = ld_stream_f2(pc);
        pc += HW;
      }
      const int q = ai * ncells + cell;  // plane-order index of the first of the two anchors
      const float *pl = loc + (size_t)ai * 5 * HW + cell;
      const int j0 = cell * na + ai, j1 = j0 + na;
#pragma unroll
      for (int d = 0; d < 5; ++d) {
        *reinterpret_cast<float2 *>(&sm.loc[d * kHeadTile + q]) = ld_stream_f2(pl);  // plane order: no bank conflicts
        pl += HW;
      }
      float2 mx = x[0], best = x[1];
#pragma unroll
      for (int c = 1; c < C; ++c) {
        mx.x = fmaxf(mx.x, x[c].x);
        mx.y = fmaxf(mx.y, x[c].y);
        if (c > 1) {
          best.x = fmaxf(best.x, x[c].x);
          best.y = fmaxf(best.y, x[c].y);
        }
      }
      float2 sum = make_float2(0.f, 0.f);
#pragma unroll
      for (int c = 0; c < C; ++c) {
        float e0, e1;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(fmul(fsub(x[c].x, mx.x), 1.4426950408889634f)));
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(fmul(fsub(x[c].y, mx.y), 1.4426950408889634f)));
        sum.x = fadd(sum.x, e0);
        sum.y = fadd(sum.y, e1);
      }
      float b0, b1;
      asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(b0) : "f"(fmul(fsub(best.x, mx.x), 1.4426950408889634f)));
      asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(b1) : "f"(fmul(fsub(best.y, mx.y), 1.4426950408889634f)));
      sm.score[j0] = fdiv(b0, sum.x);
      sm.score[j1] = fdiv(b1, sum.y);
    }
  } else {
    for (int p = threadIdx.x; p < n; p += kHeadThreads) {
      const int ai = p / ncells, cell = p - ai * ncells;
      const int j = cell * na + ai;  // anchor index inside the tile (cell-major, symbol/common.py:399-400)
      const float *pc = cls + (size_t)ai * C * HW + cell;
      float x[C];
#pragma unroll
      for (int c = 0; c < C; ++c) {
        x[c] = ld_stream_f1(pc);
        pc += HW;
      }
      const float *pl = loc + (size_t)ai * 5 * HW + cell;
#pragma unroll
      for (int d = 0; d < 5; ++d) {
        sm.loc[d * kHeadTile + p] = ld_stream_f1(pl);
        pl += HW;
      }

      float mx = x[0];
#pragma unroll
      for (int c = 1; c < C; ++c) mx = fmaxf(mx, x[c]);
      float sum = 0.f, best = x[1];
#pragma unroll
      for (int c = 0; c < C; ++c) {
        float e;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fmul(fsub(x[c], mx), 1.4426950408889634f)));
        sum = fadd(sum, e);
        if (c > 1) best = fmaxf(best, x[c]);
      }
      float eb;
      asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(eb) : "f"(fmul(fsub(best, mx), 1.4426950408889634f)));
      sm.score[j] = fdiv(eb, sum);
    }
  }
  __syncthreads();
  // ---- candidates of the exact phase, in anchor order (two consecutive anchors per thread) ----
  const int l0 = threadIdx.x * 2;
  int ncand_mine = 0;
  bool cnd[2];
#pragma unroll
  for (int v = 0; v < 2; ++v) {
    cnd[v] = l0 + v < n && (all_cand || sm.score[l0 + v] >= thr_lo);
    ncand_mine += cnd[v];
  }
  int ncand;
  int cpos = block_scan_excl(ncand_mine, scan_smem, &ncand);
#pragma unroll
  for (int v = 0; v < 2; ++v)
    if (cnd[v]) sm.cand[cpos++] = (unsigned short)(l0 + v);
  __syncthreads();
  // ---- exact phase, kHeadBatch candidates at a time ----
  for (int c0 = 0; c0 < ncand; c0 += kHeadBatch) {
    const int nb = min(kHeadBatch, ncand - c0);
    if ((int)threadIdx.x < nb) {
      // the candidate's logits come from global memory again (L2 hits: the tile was read a microsecond ago) -- keeping
      // all 21 x 256 logits of the tile in shared memory for the one anchor in five that needs them cost half the
      // occupancy
      const int j = sm.cand[c0 + threadIdx.x];
      const int cell = j / na, ai = j - cell * na;
      const float *pc = cls + (size_t)ai * C * HW + cell;
      float *x = sm.es[threadIdx.x];
      float mx = __int_as_float(0xff800000);
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const float v = __ldg(pc + (size_t)c * HW);
        x[c] = v;
        if (c == 0 || v > mx) mx = v;
      }
      sm.cmx[threadIdx.x] = mx;
    }
    __syncthreads();
    for (int q = threadIdx.x; q < nb * C; q += kHeadThreads) {
      const int kk = q / C, c = q - kk * C;
      sm.es[kk][c] = libm::expf_glibc(fsub(sm.es[kk][c], sm.cmx[kk]), a.fma_build);
    }
    __syncthreads();
    if ((int)threadIdx.x < nb) {
      const float *e = sm.es[threadIdx.x];
      float sum = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c) sum = fadd(sum, e[c]);
      // first maximum of the rounded probabilities over the foreground classes (strict >, multibox_detection.cc:84-91):
      // RN(e/sum) is monotone in e, so it is the first class whose quotient equals the quotient of the largest e
      float emax = e[1];
      int cbest = 1;
#pragma unroll
      for (int c = 2; c < C; ++c)
        if (e[c] > emax) {
          emax = e[c];
          cbest = c;
        }
      const float pmax = fdiv(emax, sum);
      for (int c = 1; c < cbest; ++c)
        if (e[c] >= fmul(emax, 0.999999f) && fdiv(e[c], sum) == pmax) {
          cbest = c;
          break;
        }
      sm.fscore[c0 + threadIdx.x] = pmax;
      sm.fid[c0 + threadIdx.x] = (unsigned short)((pmax < a.threshold) ? 0 : cbest);
    }
    __syncthreads();
  }
  // ---- survivors in anchor order ----
  int nvalid = 0;
  bool ok[2];
#pragma unroll
  for (int v = 0; v < 2; ++v) {
    const int q = l0 + v;
    ok[v] = q < ncand && sm.fid[q] != 0;
    nvalid += ok[v];
  }
  int total;
  int pos = block_scan_excl(nvalid, scan_smem, &total);
  if (threadIdx.x == 0) a.tile_count[(size_t)b * a.T + t] = total;
  float sc[2];
  unsigned short ci[2], cj[2], cqv[2];
#pragma unroll
  for (int v = 0; v < 2; ++v)
    if (ok[v]) {
      sc[v] = sm.fscore[l0 + v];
      ci[v] = sm.fid[l0 + v];
      cj[v] = sm.cand[l0 + v];
      const int cell = cj[v] / na;
      cqv[v] = (unsigned short)((cj[v] - cell * na) * ncells + cell);
    }
  __syncthreads();  // sm.score is about to change meaning (approximate per anchor -> exact per survivor)
#pragma unroll
  for (int v = 0; v < 2; ++v)
    if (ok[v]) {
      sm.score[pos] = sc[v];
      sm.idx[pos] = cj[v];
      sm.idxq[pos] = cqv[v];
      sm.id[pos] = ci[v];
      ++pos;
    }
  finish_tile<NFG, kHeadThreads, kHeadTile / kHeadThreads, kV2, false, true>(a, sm, b, t, t * kHeadTile, total);
}

// ----------------------------------------------------------------------------------------------------
// Bitonic sort of n (power of two) 64-bit keys, ascending, by the whole CTA.  `keys` may point to shared or
// global memory.  Shared-memory bandwidth bound (32 B per compare-exchange): only used for selections of more than 512 keys.
__device__ void bitonic_sort_u64(unsigned long long *keys, int n) {
  for (int k = 2; k <= n; k <<= 1) {
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
}

// Bitonic sort of 128 * KPT 64-bit keys (ascending; entries >= nkeep are padding) held in the registers of the
// first four warps.  Element index = lane | reg << 5 | warp << (5 + log2 KPT): exchanges at distance < 32 are warp
// shuffles, distances 32 .. 16 KPT are register-to-register inside a thread, and only the two highest index bits
// (three steps of the whole sort) go through shared memory behind a 128-thread named barrier.  Every step runs KPT
// independent compare-exchanges per thread, so the shuffle latency is pipelined instead of exposed.
constexpr int kRegSortThreads = 128;
template <int KPT>
__device__ __forceinline__ void bitonic_sort_regs(unsigned long long *ssel, int nkeep) {
  constexpr int LR = KPT == 1 ? 0 : (KPT == 2 ? 1 : (KPT == 4 ? 2 : 3));
  constexpr int S = kRegSortThreads * KPT;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int idx0 = lane | (w << (5 + LR));
  unsigned long long k[KPT];
#pragma unroll
  for (int q = 0; q < KPT; ++q) {
    const int idx = idx0 | (q << 5);
    k[q] = idx < nkeep ? ssel[idx] : ~0ull;
  }
#pragma unroll 1
  for (int K = 2; K <= S; K <<= 1) {
#pragma unroll 1
    for (int j = K >> 1; j > 0; j >>= 1) {
      if (j < 32) {
#pragma unroll
        for (int q = 0; q < KPT; ++q) {
          const int idx = idx0 | (q << 5);
          const unsigned long long other = __shfl_xor_sync(kFullMask, k[q], j);
          const bool take_min = ((idx & j) == 0) == ((idx & K) == 0);
          k[q] = (take_min == (other < k[q])) ? other : k[q];
        }
      } else if (j < 32 * KPT) {
#pragma unroll
        for (int r = 0; r < LR; ++r) {
          if (j == (32 << r)) {
#pragma unroll
            for (int q = 0; q < KPT; ++q) {
              if (q & (1 << r)) continue;
              const int idx = idx0 | (q << 5);
              const bool up = (idx & K) == 0;
              const unsigned long long x = k[q], y = k[q | (1 << r)];
              const bool swap = (x > y) == up;
              k[q] = swap ? y : x;
              k[q | (1 << r)] = swap ? x : y;
            }
          }
        }
      } else {
#pragma unroll
        for (int q = 0; q < KPT; ++q) ssel[idx0 | (q << 5)] = k[q];
        asm volatile("bar.sync 1, %0;" ::"n"(kRegSortThreads) : "memory");
        unsigned long long other[KPT];
#pragma unroll
        for (int q = 0; q < KPT; ++q) other[q] = ssel[(idx0 | (q << 5)) ^ j];
        asm volatile("bar.sync 1, %0;" ::"n"(kRegSortThreads) : "memory");
#pragma unroll
        for (int q = 0; q < KPT; ++q) {
          const int idx = idx0 | (q << 5);
          const bool take_min = ((idx & j) == 0) == ((idx & K) == 0);
          k[q] = (take_min == (other[q] < k[q])) ? other[q] : k[q];
        }
      }
    }
  }
#pragma unroll
  for (int q = 0; q < KPT; ++q) ssel[idx0 | (q << 5)] = k[q];
}

// Warp-aggregated shared-memory histogram increment: the bin of the first participating lane is counted for the
// whole warp with one ballot (scores cluster: in the first pass that is nearly every lane), the other lanes add
// themselves -- their bins rarely coincide.  (__match_any_sync costs more than both once the bins are spread.)
__device__ __forceinline__ void hist_add(unsigned *hist, unsigned bin, bool pred) {
  const unsigned active = __ballot_sync(kFullMask, pred);
  if (active == 0u) return;
  const int leader = __ffs(active) - 1;
  const unsigned b0 = __shfl_sync(kFullMask, bin, leader);
  const unsigned same = __ballot_sync(kFullMask, pred && bin == b0);
  if ((int)lane_id() == leader) atomicAdd(&hist[b0], (unsigned)__popc(same));
  else if (pred && bin != b0) atomicAdd(&hist[bin], 1u);
}

// Unordered append of the lanes with `pred` to a shared list through one atomic per warp.
__device__ __forceinline__ int warp_append(int *counter, bool pred) {
  const unsigned m = __ballot_sync(kFullMask, pred);
  if (m == 0u) return 0;
  int base = 0;
  if (lane_id() == 0) base = atomicAdd(counter, __popc(m));
  base = __shfl_sync(kFullMask, base, 0);
  return base + __popc(m & ((1u << lane_id()) - 1u));
}

struct SortArgs {
  float *out;
  const int *tile_count;            // (B, T) survivors per tile (stream kernel)
  const float *slot_rows;           // per-tile slots written by the stream kernel
  const unsigned *slot_keys;
  const unsigned short *slot_cls;
  const float4 *slot_box;
  unsigned *keys;                   // (B, Apad) rank-ordered keys, only when they do not fit in shared memory
  int *tile_base;                   // (B, T) exclusive prefix of tile_count, only when T > kSortSmemTiles
  int *valid;                       // (B) V
  int *nms_rows, *cursor;
  unsigned short *row_cls;
  float4 *row_box;
  unsigned long long *sort_keys;
  int *head_rank;                   // (B, Apad) pass-1 rank of every head row (fork/join pipeline), or null
  int2 *head_list;                  // per-class head lists (fork/join pipeline with 0 < nms_topk <= kHeadCap), or null
  int *head_off;
  const unsigned *tile_cls;         // (B, kV2ClsPad, T) per (class, tile) offset | count << 16 (fork/join pipeline)
  int nfg, mask_rows;
  int *valid_count_out;
  WsHeader *header;
  int A, T, tile, Apad, cls_stride, npad_max, niter_max;
  int sel_cap, rank_parts, debug, tail_copy;
  float nms_threshold;
  int force_suppress, nms_topk;
};

constexpr int kSortSmemTiles = 1024;  // tile bases kept in shared memory up to this many tiles per image

// Rank role of det_sort_kernel (blockIdx.y >= 1): moves the runs of a slice of the image's tiles from their slots to
// their final pass-1 positions -- rank base of tile t = survivors in the tiles before it, so the reference's
// anchor-ordered compaction (multibox_detection.cc:93-127) is reproduced without any CTA waiting on another.
// Rows below nkeep are skipped: the sort role of the same launch writes the sorted head there.
__device__ void det_rank_role(const SortArgs &a, int b, int part, int *red) {
  const int T = a.T;
  const int *cnt = a.tile_count + (size_t)b * T;
  const int per = (T + a.rank_parts - 1) / a.rank_parts;
  const int t0 = part * per, t1 = min(T, t0 + per);
  if (t0 >= t1) return;
  int s_all = 0, s_before = 0;
  for (int u = threadIdx.x; u < T; u += blockDim.x) {
    const int c = cnt[u];
    s_all += c;
    if (u < t0) s_before += c;
  }
  s_all = warp_sum_i32(s_all);
  s_before = warp_sum_i32(s_before);
  const unsigned warp = warp_id(), lane = lane_id(), nwarps = blockDim.x >> 5;
  if (lane == 0) {
    red[warp] = s_all;
    red[32 + warp] = s_before;
  }
  __syncthreads();
  const int V = warp_sum_i32(lane < nwarps ? red[lane] : 0);
  const int base0 = warp_sum_i32(lane < nwarps ? red[32 + lane] : 0);
  const bool do_sort = V >= 1 && a.nms_threshold > 0.f && a.nms_threshold <= 1.f;
  int nkeep = 0;
  if (do_sort) nkeep = (a.nms_topk > 0 && a.nms_topk < V) ? a.nms_topk : V;
  if (nkeep >= V) return;
  // wpt warps per tile (as many as the CTA has to spare), loads batched four deep: the copies are latency bound
  const int ntiles = t1 - t0;
  int wpt = (int)nwarps / ntiles;
  wpt = wpt < 1 ? 1 : (wpt > 8 ? 8 : wpt);
  const int sub = (int)warp % wpt, step = wpt * 32;
  for (int t = t0 + (int)warp / wpt; t < t1; t += (int)nwarps / wpt) {
    int sb = 0;
    for (int u = t0 + (int)lane; u < t; u += 32) sb += cnt[u];
    const int base = base0 + warp_sum_i32(sb);
    const int n = cnt[t];
    const int skip = min(n, max(0, nkeep - base));
    if (skip >= n) continue;
    const size_t slot0 = (size_t)b * a.Apad + (size_t)t * a.tile;
    const float *src = a.slot_rows + slot0 * 7;
    float *dst = a.out + ((size_t)b * a.A + base) * 7;
    const int end = n * 7;
    for (int q = skip * 7 + sub * 32 + (int)lane; q < end; q += 4 * step) {
      float v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k] = q + k * step < end ? src[q + k * step] : 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (q + k * step < end) dst[q + k * step] = v[k];
    }
    const unsigned short *sc = a.slot_cls + (size_t)b * a.cls_stride + (size_t)t * a.tile;
    const float4 *sbx = a.slot_box + slot0;
    unsigned short *dc = a.row_cls + (size_t)b * a.cls_stride + base;
    float4 *db = a.row_box + (size_t)b * a.A + base;
    for (int q = skip + sub * 32 + (int)lane; q < n; q += 2 * step) {
      const bool two = q + step < n;
      const unsigned short c0 = sc[q], c1 = two ? sc[q + step] : (unsigned short)0;
      const float4 b0 = sbx[q], b1 = two ? sbx[q + step] : b0;
      dc[q] = c0;
      db[q] = b0;
      if (two) {
        dc[q + step] = c1;
        db[q + step] = b1;
      }
    }
  }
}

// grid (B, 1 + rank_parts).  blockIdx.y == 0 is the sort role, one CTA per image: it ranks the image's tiles
// (prefix of the tile counts), stages the keys of the V survivors in rank order, radix-selects the nkeep best, sorts
// them and writes those rows -- gathered straight from the slots -- to rows [0, nkeep) of the output
// (multibox_detection.cc:132-151).  The other CTAs of the image move rows [nkeep, V) (det_rank_role); the two roles
// write disjoint rows, so they need no ordering between them.
template <bool kKeysInSmem>
__global__ void __launch_bounds__(kSortThreads) det_sort_kernel(const __grid_constant__ SortArgs a) {
  extern __shared__ __align__(128) unsigned char dyn_smem[];
  __shared__ int scan_smem[kSortThreads / 32 + 1];
  __shared__ unsigned hist256[256];
  __shared__ int carry_smem, sm_need, sm_count, sm_eq_total;
  __shared__ unsigned sm_prefix;
  __shared__ int sm_tbase[kSortSmemTiles];
  constexpr int kBuckets = 2 * kSortThreads, kBucketCap = 1024;
  __shared__ unsigned bucket[kBuckets];
  __shared__ unsigned long long stage_a[kBucketCap];
  __shared__ unsigned sm_kmin, sm_kmax;
  __shared__ int sm_pivot, sm_need_rows;
  const int b = blockIdx.x;  // image in x: the sort-role CTAs (y == 0) of all images are scheduled first
  // Programmatic dependent launch, twice.  This kernel may itself be a programmatic dependent of the stream kernel
  // (DSPMB_TUNE_DET_SORT_PDL; the stream kernel triggers on entry): its CTAs then take the SMs the stream kernel's tail
  // frees and wait here until that grid has completed and flushed -- a no-op under a plain launch.  Only THEN does it
  // release its own dependent: the pair-test kernel enqueued behind it reads the stream kernel's outputs before its
  // own griddepcontrol.wait, so it must not start before the stream kernel is complete; it runs beside the sort and
  // only waits for the sort before its resolve.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  TraceScope trace_(1);
  if (blockIdx.y != 0) {
    if (a.debug & 8) return;  // timing experiments only
    det_rank_role(a, b, (int)blockIdx.y - 1, reinterpret_cast<int *>(hist256));
    return;
  }
  const int A = a.A, T = a.T;
  // rank base of every tile (block scan over the tile counts) and V
  const int *cnt = a.tile_count + (size_t)b * T;
  int *tbase = T <= kSortSmemTiles ? sm_tbase : a.tile_base + (size_t)b * T;
  if (threadIdx.x == 0) {
    carry_smem = 0;
    sm_need_rows = 0;
  }
  __syncthreads();
  if (a.tail_copy && (int)warp_id() < a.nfg) {
    // Fork/join pipeline: does any class of this image exceed the pair kernel's shared-memory mask?  Only then does
    // that kernel fall back to the NMS on final-order rows, which reads row_cls / row_box of the tail (copied below).
    const unsigned *tc = a.tile_cls + ((size_t)b * kV2ClsPad + warp_id()) * T;
    int members = 0;
    for (int t = (int)lane_id(); t < T; t += 32) members += (int)(tc[t] >> 16);
    members = warp_sum_i32(members);
    if (lane_id() == 0 && members > a.mask_rows) sm_need_rows = 1;
  }
  for (int base = 0; base < T; base += blockDim.x) {
    const int i = base + threadIdx.x;
    const int v = i < T ? cnt[i] : 0;
    int total;
    const int ex = block_scan_excl(v, scan_smem, &total);
    const int carry = carry_smem;
    if (i < T) tbase[i] = carry + ex;
    __syncthreads();
    if (threadIdx.x == 0) carry_smem = carry + total;
    __syncthreads();
  }
  const int V = carry_smem;
  const bool do_sort = V >= 1 && a.nms_threshold > 0.f && a.nms_threshold <= 1.f;  // multibox_detection.cc:130
  __syncthreads();
  if (threadIdx.x == 0) {
    if (b == 0) a.header->status = DSPMB_OK;
    if (a.valid_count_out) a.valid_count_out[b] = V;
    a.valid[b] = V;
    a.nms_rows[b] = do_sort ? V : 0;
    a.cursor[b] = 0;
    sm_prefix = 0u;
    sm_count = 0;
    sm_eq_total = 0;
    carry_smem = 0;
    sm_kmin = 0xffffffffu;
    sm_kmax = 0u;
    sm_pivot = 0;
  }
  if (!do_sort) return;  // pass-1 rows are the final output
  // dynamic smem: [sel: sel_cap u64][skeys: round4(A) u32 (optional)][wtab: niter_max*32 int]
  unsigned long long *ssel = reinterpret_cast<unsigned long long *>(dyn_smem);
  unsigned *skeys = reinterpret_cast<unsigned *>(ssel + a.sel_cap);
  int *wtab = reinterpret_cast<int *>(skeys + (kKeysInSmem ? ((A + 3) & ~3) : 0));
  unsigned *gkeys = a.keys + (size_t)b * a.Apad;
  auto key_at = [&](int p) -> unsigned { return kKeysInSmem ? skeys[p] : gkeys[p]; };
  const unsigned warp = warp_id(), lane = lane_id();
  const int niter = (V + (int)blockDim.x - 1) / (int)blockDim.x;

  {  // keys of the survivors in rank order: one warp per tile, four tiles and two 32-key chunks per tile in flight
     // (the staging is one memory round trip for the usual tile occupancy); n = tbase[t + 1] - tbase[t]
    unsigned *dstk = kKeysInSmem ? skeys : gkeys;
    const unsigned *srck = a.slot_keys + (size_t)b * a.Apad;
    const int nwarps = (int)(blockDim.x >> 5);
    for (int tb4 = (int)warp; tb4 < T; tb4 += nwarps * 4) {
      int n[4], tb[4];
      unsigned v[4][2];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int t = tb4 + i * nwarps;
        tb[i] = t < T ? tbase[t] : 0;
        n[i] = t < T ? (t + 1 < T ? tbase[t + 1] : V) - tb[i] : 0;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int k = (int)lane + 32 * c;
          v[i][c] = k < n[i] ? srck[(size_t)t * a.tile + k] : 0u;
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int k = (int)lane + 32 * c;
          if (k < n[i]) dstk[tb[i] + k] = v[i][c];
        }
        const int t = tb4 + i * nwarps;
        for (int k = (int)lane + 64; k < n[i]; k += 32) dstk[tb[i] + k] = srck[(size_t)t * a.tile + k];
      }
    }
  }
  int nkeep = V;
  if (a.nms_topk > 0 && a.nms_topk < nkeep) nkeep = a.nms_topk;  // multibox_detection.cc:142-145
  int npad = 2;
  while (npad < nkeep) npad <<= 1;
  unsigned long long *sel = npad > a.sel_cap ? a.sort_keys + (size_t)b * a.npad_max : ssel;
  if (threadIdx.x == 0) sm_need = nkeep;
  __syncthreads();

  trace_point(9);
  // ---- bucket sort of the head (the usual case: keys in shared memory, a head of at most kBucketCap rows) ----
  // One 2048-bin histogram over the image's key range replaces the four 8-bit radix-select passes AND the bitonic
  // sort: the exclusive prefix of the histogram is both the selection (bins below the pivot bin are in, the pivot
  // bin is cut by rank) and the sorted position of every bin; the keys of bins <= pivot are scattered to their bins
  // (any order inside a bin) and every key then finds its place by counting the smaller keys of ITS bin -- all keys
  // in parallel, so a crowded bin (scores piling up near 1.0) costs its size, not its square.  A pivot bin that
  // would overflow the staging buffers falls back to the radix select + bitonic path below.
  bool sorted_done = false;
  if (kKeysInSmem && npad <= a.sel_cap && nkeep <= kBucketCap && !(a.debug & 16)) {
    unsigned kmin = 0xffffffffu, kmax = 0u;
    for (int p = threadIdx.x; p < V; p += blockDim.x) {
      const unsigned kv = skeys[p];
      kmin = min(kmin, kv);
      kmax = max(kmax, kv);
    }
    kmin = __reduce_min_sync(kFullMask, kmin);
    kmax = __reduce_max_sync(kFullMask, kmax);
    if (lane == 0) {
      atomicMin(&sm_kmin, kmin);
      atomicMax(&sm_kmax, kmax);
    }
    for (int i = threadIdx.x; i < kBuckets; i += blockDim.x) bucket[i] = 0u;
    __syncthreads();
    kmin = sm_kmin;
    int shift = 0;
    while (((sm_kmax - kmin) >> shift) >= (unsigned)kBuckets) ++shift;
    for (int p = threadIdx.x; p < V; p += blockDim.x) atomicAdd(&bucket[(skeys[p] - kmin) >> shift], 1u);
    __syncthreads();
    // exclusive prefix over the buckets (two per thread); pivot bucket P: start < nkeep <= start + count
    const unsigned c0 = bucket[2 * threadIdx.x], c1 = bucket[2 * threadIdx.x + 1];
    int total;
    const int ex = block_scan_excl((int)(c0 + c1), scan_smem, &total);
    if (ex < nkeep && ex + (int)c0 >= nkeep) {
      sm_pivot = 2 * threadIdx.x;
      sm_count = ex + (int)c0;
    } else if (ex + (int)c0 < nkeep && ex + (int)(c0 + c1) >= nkeep) {
      sm_pivot = 2 * threadIdx.x + 1;
      sm_count = ex + (int)(c0 + c1);
    }
    __syncthreads();
    const int P = sm_pivot, nsc = sm_count;  // nsc = keys in buckets <= P (>= nkeep)
    if (nsc <= kBucketCap) {  // CTA-uniform
      bucket[2 * threadIdx.x] = (unsigned)ex;  // running write cursors: after the scatter bucket[b] = end of bucket b
      bucket[2 * threadIdx.x + 1] = (unsigned)ex + c0;
      __syncthreads();
      for (int p = threadIdx.x; p < V; p += blockDim.x) {
        const unsigned kv = skeys[p];
        const unsigned bk = (kv - kmin) >> shift;
        if ((int)bk <= P) stage_a[atomicAdd(&bucket[bk], 1u)] = ((unsigned long long)kv << 32) | (unsigned)p;
      }
      __syncthreads();
      for (int i = threadIdx.x; i < nsc; i += blockDim.x) {
        const unsigned long long x = stage_a[i];
        const unsigned bk = ((unsigned)(x >> 32) - kmin) >> shift;
        const int lo = bk ? (int)bucket[bk - 1] : 0, hi = (int)bucket[bk];
        int cnt = 0;
        for (int j = lo; j < hi; ++j) cnt += stage_a[j] < x ? 1 : 0;
        if (lo + cnt < nkeep) ssel[lo + cnt] = x;
      }
      __syncthreads();
      sorted_done = true;
    }
    if (threadIdx.x == 0) sm_count = 0;  // the radix path below counts with it
    __syncthreads();
  }

  if (sorted_done) {
    // head = ssel[0, nkeep) in ascending (key, rank) order
  } else if (nkeep == V) {
    for (int p = threadIdx.x; p < npad; p += blockDim.x)
      sel[p] = p < V ? (((unsigned long long)key_at(p) << 32) | (unsigned)p) : ~0ull;
    __syncthreads();
  } else {
    // MSB-first radix select of the nkeep-th smallest key
    for (int pass = 0; pass < 4; ++pass) {
      const int shift = 24 - 8 * pass;
      const unsigned mask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
      for (int i = threadIdx.x; i < 256; i += blockDim.x) hist256[i] = 0u;
      __syncthreads();
      const unsigned prefix = sm_prefix;
      for (int it = 0; it < niter; ++it) {
        const int p = it * blockDim.x + threadIdx.x;
        const unsigned kv = p < V ? key_at(p) : 0u;
        hist_add(hist256, (kv >> shift) & 0xffu, p < V && (kv & mask) == prefix);
      }
      __syncthreads();
      if (warp == 0) {  // digit search: 8 bins per lane, warp scan
        unsigned h[8];
        int tot = 0;
#pragma unroll
        for (int d = 0; d < 8; ++d) {
          h[d] = hist256[lane * 8 + d];
          tot += (int)h[d];
        }
        const int incl = warp_scan_incl(tot);
        const int excl = incl - tot;
        const int need = sm_need;
        __syncwarp();  // every lane has read sm_need before one lane overwrites it
        if (excl < need && incl >= need) {
          int acc = excl;
#pragma unroll
          for (int d = 0; d < 8; ++d) {
            if (acc + (int)h[d] >= need) {
              sm_need = need - acc;
              sm_prefix = prefix | ((unsigned)(lane * 8 + d) << shift);
              sm_eq_total = (int)h[d];
              break;
            }
            acc += (int)h[d];
          }
        }
      }
      __syncthreads();
    }
      const unsigned pivot = sm_prefix;
    const int need_eq = sm_need;
    const bool ordered_ties = need_eq < sm_eq_total;  // only the lowest-ranked of the pivot-valued keys make it
    if (ordered_ties) {
      for (int it = 0; it < niter; ++it) {
        const int p = it * blockDim.x + threadIdx.x;
        const unsigned m = __ballot_sync(kFullMask, p < V && key_at(p) == pivot);
        if (lane == 0) wtab[it * 32 + warp] = __popc(m);
      }
      __syncthreads();
      for (int base = 0; base < niter * 32; base += blockDim.x) {
        const int i = base + threadIdx.x;
        const int v = i < niter * 32 ? wtab[i] : 0;
        int total;
        const int ex = block_scan_excl(v, scan_smem, &total);
        const int carry = carry_smem;
        if (i < niter * 32) wtab[i] = carry + ex;
        __syncthreads();
        if (threadIdx.x == 0) carry_smem = carry + total;
        __syncthreads();
      }
    }
    for (int it = 0; it < niter; ++it) {  // gather (any order: the 64-bit keys are unique and get sorted next)
      const int p = it * blockDim.x + threadIdx.x;
      const unsigned kv = p < V ? key_at(p) : 0xffffffffu;
      bool take = p < V && kv <= pivot;
      if (ordered_ties) {
        const bool eq = p < V && kv == pivot;
        const unsigned m = __ballot_sync(kFullMask, eq);
        if (eq) take = wtab[it * 32 + warp] + __popc(m & ((1u << lane) - 1u)) < need_eq;
      }
      const int q = warp_append(&sm_count, take);
      if (take) sel[q] = ((unsigned long long)kv << 32) | (unsigned)p;
    }
    __syncthreads();
  }
  const int ssort = npad < kRegSortThreads ? kRegSortThreads : npad;
  if (sorted_done) {
  } else if (ssort <= 4 * kRegSortThreads && ssort <= a.sel_cap) {
    // bitonic sort in the registers of four warps, ssort / 128 keys per thread (sel == ssel here)
    if (threadIdx.x < kRegSortThreads) {
      switch (ssort / kRegSortThreads) {
        case 1: bitonic_sort_regs<1>(ssel, nkeep); break;
        case 2: bitonic_sort_regs<2>(ssel, nkeep); break;
        default: bitonic_sort_regs<4>(ssel, nkeep); break;
      }
    }
    __syncthreads();
  } else {
    for (int p = nkeep + threadIdx.x; p < npad; p += blockDim.x) sel[p] = ~0ull;
    __syncthreads();
    bitonic_sort_u64(sel, npad);  // ends with __syncthreads()
  }

  trace_point(sorted_done ? 10 : 11);
  // head rows: rank p lives in tile t = last tile with tbase[t] <= p, slot p - tbase[t] of that tile
  float *out = a.out + (size_t)b * A * 7;
  unsigned short *row_cls = a.row_cls + (size_t)b * a.cls_stride;
  float4 *row_box = a.row_box + (size_t)b * A;
  int my_p = 0, my_c = -1;  // this thread's head row (per-class head lists: nkeep <= blockDim.x)
  for (int r = threadIdx.x; r < nkeep; r += blockDim.x) {
    const int p = (int)(unsigned)(sel[r] & 0xffffffffull);
    int lo = 0, hi = T - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (tbase[mid] <= p) lo = mid; else hi = mid - 1;
    }
    const float *src = a.slot_rows + ((size_t)b * a.Apad + (size_t)lo * a.tile + (size_t)(p - tbase[lo])) * 7;
    const float s0 = src[0], s1 = src[1], s2 = src[2], s3 = src[3], s4 = src[4], s5 = src[5], s6 = src[6];
    float *o = out + (size_t)r * 7;
    o[0] = s0;
    o[1] = s1;
    o[2] = s2;
    o[3] = s3;
    o[4] = s4;
    o[5] = s5;
    o[6] = s6;
    row_cls[r] = (unsigned short)s0;
    row_box[r] = make_float4(s2, s3, s4, s5);
    if (a.head_list) {
      my_p = p;
      my_c = (int)s0;
    } else if (a.head_rank) {
      a.head_rank[(size_t)b * a.Apad + r] = p | ((int)s0 << 24);  // pass-1 rank | class << 24
    }
  }
  if (a.head_list) {
    // Per-class head lists for the resolve of the pair kernel: a stable counting sort of the <= kHeadCap head rows by
    // class (one row per thread, so thread order == head position order).  Lanes of a warp with the same class find
    // each other with MATCH.ANY; per-(warp, class) counts are scanned over the warps by the warp that owns the class,
    // the class totals over the classes by warp 0.  Each pair CTA then reads only its own ~nkeep/classes entries, in
    // order, instead of filtering and ordering all nkeep head positions.
    static_assert(kSortThreads == 1024 && kV2ClsPad == 32, "one (warp, class) table entry per thread");
    unsigned *wc = bucket;                                   // [32 warps][32 classes], dead after the sort
    int *ctot = reinterpret_cast<int *>(hist256);            // [32] class totals, then [32..63] class offsets
    wc[threadIdx.x] = 0u;
    __syncthreads();
    const bool on = my_c >= 0 && my_c < kV2ClsPad;
    const unsigned peers = __match_any_sync(kFullMask, on ? my_c : -1);
    const int within = __popc(peers & ((1u << lane) - 1u));
    if (on && within == 0) wc[warp * 32 + my_c] = (unsigned)__popc(peers);
    __syncthreads();
    {  // warp `c` scans the counts of class c over the 32 warps (lane = warp index)
      const int v = (int)wc[lane * 32 + warp];
      const int incl = warp_scan_incl(v);
      wc[lane * 32 + warp] = (unsigned)(incl - v);
      if (lane == 31) ctot[warp] = incl;
    }
    __syncthreads();
    if (warp == 0) {
      const int v = ctot[lane];
      const int incl = warp_scan_incl(v);
      ctot[32 + lane] = incl - v;
      int *ho = a.head_off + (size_t)b * (kV2ClsPad + 1);
      ho[lane] = incl - v;
      if (lane == 31) ho[32] = incl;
    }
    __syncthreads();
    if (on) a.head_list[(size_t)b * kHeadCap + ctot[32 + my_c] + (int)wc[warp * 32 + my_c] + within] = make_int2((int)threadIdx.x, my_p);
  }
  // the NMS kernel reads row_cls eight entries at a time: define the entries between V and the next multiple of 8
  if (threadIdx.x < 8 && V + (int)threadIdx.x < ((V + 7) & ~7)) row_cls[V + threadIdx.x] = (unsigned short)0xffffu;
  if (a.tail_copy && nkeep < V && sm_need_rows) {
    // Fork/join pipeline, rare case: a class of this image is too large for the pair kernel's shared-memory mask and
    // will take the NMS on final-order rows, which reads row_cls / row_box of the tail rows [nkeep, V) as well.  (The
    // tail rows of `out` themselves are written by the pair kernel, each class its own.)  One warp per tile, three
    // tiles at a time: all loads of the three runs are issued before the first store.
    const int nwarps = (int)(blockDim.x >> 5);
    constexpr int kU = 3;
    for (int t0 = (int)warp; t0 < T; t0 += kU * nwarps) {
      int base[kU], nt[kU], skip[kU];
      int maxrows = 0;
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int t = t0 + u * nwarps;
        base[u] = nt[u] = skip[u] = 0;
        if (t < T) {
          base[u] = tbase[t];
          nt[u] = (t + 1 < T ? tbase[t + 1] : V) - base[u];
          skip[u] = min(nt[u], max(0, nkeep - base[u]));
          if (skip[u] < nt[u]) maxrows = max(maxrows, nt[u]);
        }
      }
      for (int q = (int)lane; q < maxrows; q += 32) {
        unsigned short cl[kU];
        float4 bx[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u) {
          const int t = t0 + u * nwarps;
          const bool on = q >= skip[u] && q < nt[u];
          cl[u] = on ? a.slot_cls[(size_t)b * a.cls_stride + (size_t)t * a.tile + q] : (unsigned short)0;
          bx[u] = on ? a.slot_box[(size_t)b * a.Apad + (size_t)t * a.tile + q] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < kU; ++u)
          if (q >= skip[u] && q < nt[u]) {
            row_cls[base[u] + q] = cl[u];
            row_box[base[u] + q] = bx[u];
          }
      }
    }
  }
}

struct NmsArgs {
  float *out;
  const int *nms_rows;
  int *cursor;
  const unsigned short *row_cls;
  const float4 *row_box;
  int *seg_list;
  float4 *seg_box;
  unsigned char *seg_dead;
  float *seg_area;
  int A, cls_stride, C;
  float nms_threshold;
  int force_suppress, mask_rows, smem_rows, debug;
};

// IoU >= thr test of multibox_detection.cc:44-51,162 on boxes staged by stage_box().
//  * Disjoint boxes have i = 0, hence iou = 0 < thr (thr > 0 whenever NMS runs).  min(x2) - max(x1) > 0 is four
//    compares per axis pair (a - b > 0 <=> a > b in IEEE arithmetic with gradual underflow); degenerate boxes were
//    given x1 = +inf by stage_box so they fail the test.
//  * For overlapping pairs RN(i/u) >= thr is decided without dividing unless i lies within 2^-20 (relative) of
//    thr*u: rounding is monotonic, so i/u > thr(1+2^-21) implies RN(i/u) >= thr and i/u < thr(1-2^-21) implies
//    RN(i/u) < thr; the products thr_hi*u, thr_lo*u carry < 2^-22 relative error.  Only the band in between
//    (and tiny unions, where those products could underflow) takes the IEEE division.
struct NmsThr {
  float thr, lo, hi, shrink;
};
__device__ __forceinline__ NmsThr make_thr(float thr) {
  NmsThr t;
  t.thr = thr;
  t.lo = fmul(thr, 1.0f - 0x1p-20f);
  t.hi = fmul(thr, 1.0f + 0x1p-20f);
  t.shrink = __fmul_rd(thr, 1.0f - 0x1p-10f);
  return t;
}
__device__ __forceinline__ float4 stage_box(float4 b, float *area) {
  *area = fmul(fsub(b.z, b.x), fsub(b.w, b.y));  // (a[2]-a[0])*(a[3]-a[1]) of CalculateOverlap
  if (!(b.z > b.x && b.w > b.y)) b.x = __int_as_float(0x7f800000);
  return b;
}
// threshold test for a pair already known to overlap
__device__ __forceinline__ bool iou_ge(float4 a, float area_a, float4 b, float area_b, NmsThr t) {
  const float w = fsub(fminf(a.z, b.z), fmaxf(a.x, b.x));
  const float h = fsub(fminf(a.w, b.w), fmaxf(a.y, b.y));
  const float i = fmul(w, h);
  const float u = fsub(fadd(area_a, area_b), i);
  if (u <= 0.f) return false;
  if (u >= 1e-30f) {
    if (i > fmul(t.hi, u)) return true;
    if (i < fmul(t.lo, u)) return false;
  }
  return fdiv(i, u) >= t.thr;
}
__device__ __forceinline__ bool suppresses(float4 a, float area_a, float4 b, float area_b, NmsThr t) {
  if (!(a.z > b.x && b.z > a.x && a.w > b.y && b.w > a.y)) return false;
  return iou_ge(a, area_a, b, area_b, t);
}

constexpr int kNmsQueue = 256;  // per-warp candidate queue entries of the small-segment path
constexpr int kNmsTab = 512;  // ballot-count table entries: 64 iterations x 8 warps = 131072 rows per sweep

// NMS of one (image, class) segment -- or of the whole image with force_suppress -- on rows in FINAL order
// (row_cls / row_box).  Body of det_nms_kernel; det_pair_kernel calls it for segments too large for its mask.
constexpr int kNmsBodySmallBytes = kNmsMaskRows * (16 + 40 + 4 + 4) + (kNmsThreads / 32) * 256 * 2;
constexpr int kNmsBodyLargeBytes = kNmsSmemRows * 16 + kNmsSmemRows + 512 + kNmsSmemRows * 4;
constexpr int kNmsBodyBytes = kNmsBodySmallBytes > kNmsBodyLargeBytes ? kNmsBodySmallBytes : kNmsBodyLargeBytes;

__device__ __forceinline__ void nms_final_order_body(const NmsArgs &a, const int b, const int seg,
                                                     unsigned char *smem_raw /* kNmsBodyBytes, 16-byte aligned */) {
  // shared memory is a union of the two paths:
  //   small (n <= mask_rows <= 320): boxes[320] float4 + mask[320 * 5] u64 + list[320] int + areas[320]  (20 KB)
  //   large:                         boxes[1024] float4 + dead[1024] + 64 words + areas[1024]             (21.5 KB)
  // kept near 22 KB so that 8 CTAs (all 2048 threads) fit on an SM and the usual grid is a single wave.
  constexpr int kOffMask = kNmsMaskRows * 16, kOffList = kOffMask + kNmsMaskRows * 5 * 8, kOffArea = kOffList + kNmsMaskRows * 4;
  constexpr int kOffQueue = kOffArea + kNmsMaskRows * 4, kSmallBytes = kOffQueue + (kNmsThreads / 32) * kNmsQueue * 2;
  constexpr int kOffDead = kNmsSmemRows * 16, kOffWord = kOffDead + kNmsSmemRows, kOffAreaL = kOffWord + 512;
  constexpr int kLargeBytes = kOffAreaL + kNmsSmemRows * 4;
  static_assert(kSmallBytes <= kNmsBodyBytes && kLargeBytes <= kNmsBodyBytes, "caller's stage is too small");
  __shared__ int wtab[kNmsTab];
  __shared__ unsigned long long rowany[8];
  __shared__ unsigned char unit_tab[64];
  __shared__ unsigned sm_deadbits[2];
  __shared__ unsigned long long sm_alive, sm_deadmask;
  __shared__ int scan_smem[kNmsThreads / 32 + 1];
  __shared__ int carry_smem, sm_base;

  const int V = a.nms_rows[b];
  if (V == 0 || (a.debug & 4)) return;
  float *out = a.out + (size_t)b * a.A * 7;
  const float4 *row_box = a.row_box + (size_t)b * a.A;
  const bool identity = a.force_suppress != 0;  // single segment: row q is list entry q
  const NmsThr thr = make_thr(a.nms_threshold);
  const unsigned lane = lane_id(), warp = warp_id(), nwarps = blockDim.x >> 5;
  const int rows_per_iter = blockDim.x * 8;
  const int niter = (V + rows_per_iter - 1) / rows_per_iter;

  // ---------------- ordered member list of this class: ballot-count tables, no atomics ----------------
  int n = V;
  int *slist = reinterpret_cast<int *>(smem_raw + kOffList);  // small path only
  int *glist = nullptr;
  if (!identity) {
    const uint4 *cls8 = reinterpret_cast<const uint4 *>(a.row_cls + (size_t)b * a.cls_stride);
    const unsigned short want = (unsigned short)seg;
    const unsigned want2 = (unsigned)want * 0x10001u;
    auto hits_raw = [&](int it) -> unsigned {  // 8-bit mask of the rows 8*(it*blockDim + tid) .. +7 in this class
      const int g = it * blockDim.x + threadIdx.x;
      const int r0 = g * 8;
      if (r0 >= V) return 0u;
      const uint4 q = __ldg(cls8 + g);
      const unsigned w[4] = {q.x ^ want2, q.y ^ want2, q.z ^ want2, q.w ^ want2};
      unsigned h = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if ((w[k] & 0xffffu) == 0u) h |= 1u << (2 * k);
        if ((w[k] >> 16) == 0u) h |= 2u << (2 * k);
      }
      return V - r0 >= 8 ? h : h & ((1u << (V - r0)) - 1u);
    };
    // the first four iterations (8192 rows) are remembered between the two sweeps, 8 bits each
    unsigned hcache = 0u;
    auto hits_of = [&](int it) -> unsigned { return it < 4 ? (hcache >> (8 * it)) & 0xffu : hits_raw(it); };
#pragma unroll
    for (int it = 0; it < 4; ++it)
      if (it < niter) hcache |= hits_raw(it) << (8 * it);
    if (niter <= kNmsTab / 8) {
      // usual case (V <= 131072 rows): one sweep -- the scan of the per-(iteration, warp) counts yields both the
      // segment size and every warp's write offset
      const int entries = niter * (int)nwarps;
      for (int it = 0; it < niter; ++it) {
        const int c = warp_sum_i32(__popc(hits_of(it)));
        if (lane == 0) wtab[it * nwarps + warp] = c;
      }
      __syncthreads();
      if (entries <= 32) {
        if (warp == 0) {
          const int v = (int)lane < entries ? wtab[lane] : 0;
          const int incl = warp_scan_incl(v);
          if ((int)lane < entries) wtab[lane] = incl - v;
          if (lane == 31) carry_smem = incl;
        }
        __syncthreads();
      } else {
        if (threadIdx.x == 0) carry_smem = 0;
        __syncthreads();
        for (int base = 0; base < entries; base += blockDim.x) {
          const int i = base + threadIdx.x;
          const int v = i < entries ? wtab[i] : 0;
          int total;
          const int ex = block_scan_excl(v, scan_smem, &total);
          const int carry = carry_smem;
          if (i < entries) wtab[i] = carry + ex;
          __syncthreads();
          if (threadIdx.x == 0) carry_smem = carry + total;
          __syncthreads();
        }
      }
      n = carry_smem;
      if (n < 2) return;
      if (n > a.mask_rows) {  // large segment: claim a region of the per-image list buffer
        if (threadIdx.x == 0) sm_base = atomicAdd(&a.cursor[b], n);
        __syncthreads();
        glist = a.seg_list + (size_t)b * a.A + sm_base;
      }
      for (int it = 0; it < niter; ++it) {
        const unsigned h = hits_of(it);
        const int c = __popc(h);
        int pos = wtab[it * nwarps + warp] + warp_scan_incl(c) - c;
        const int r0 = (it * blockDim.x + threadIdx.x) * 8;
        for (unsigned m = h; m; m &= m - 1) {
          const int r = r0 + __ffs(m) - 1;
          if (glist) glist[pos] = r; else slist[pos] = r;
          ++pos;
        }
      }
      __syncthreads();
    } else {
    // sweep 1: counts per (iteration, warp)
    int total_n = 0;
    for (int it0 = 0; it0 < niter; it0 += kNmsTab / 8) {
      const int its = min(niter - it0, kNmsTab / 8);
      for (int it = 0; it < its; ++it) {
        const int c = warp_sum_i32(__popc(hits_of(it0 + it)));
        if (lane == 0) wtab[it * nwarps + warp] = c;
      }
      __syncthreads();
      if (threadIdx.x == 0) carry_smem = 0;
      __syncthreads();
      for (int base = 0; base < its * (int)nwarps; base += blockDim.x) {
        const int i = base + threadIdx.x;
        const int v = i < its * (int)nwarps ? wtab[i] : 0;
        int total;
        block_scan_excl(v, scan_smem, &total);
        __syncthreads();
        if (threadIdx.x == 0) carry_smem += total;
        __syncthreads();
      }
      total_n += carry_smem;
      __syncthreads();
    }
    n = total_n;
    if (n < 2) return;
    if (n > a.mask_rows) {  // large segment: claim a region of the per-image list buffer
      if (threadIdx.x == 0) sm_base = atomicAdd(&a.cursor[b], n);
      __syncthreads();
      glist = a.seg_list + (size_t)b * a.A + sm_base;
    }
    // sweep 2: ordered positions
    int running = 0;
    for (int it0 = 0; it0 < niter; it0 += kNmsTab / 8) {
      const int its = min(niter - it0, kNmsTab / 8);
      for (int it = 0; it < its; ++it) {
        const int c = warp_sum_i32(__popc(hits_of(it0 + it)));
        if (lane == 0) wtab[it * nwarps + warp] = c;
      }
      __syncthreads();
      if (threadIdx.x == 0) carry_smem = running;
      __syncthreads();
      for (int base = 0; base < its * (int)nwarps; base += blockDim.x) {
        const int i = base + threadIdx.x;
        const int v = i < its * (int)nwarps ? wtab[i] : 0;
        int total;
        const int ex = block_scan_excl(v, scan_smem, &total);
        const int carry = carry_smem;
        if (i < its * (int)nwarps) wtab[i] = carry + ex;
        __syncthreads();
        if (threadIdx.x == 0) carry_smem = carry + total;
        __syncthreads();
      }
      for (int it = 0; it < its; ++it) {
        const unsigned h = hits_of(it0 + it);
        const int c = __popc(h);
        int pos = wtab[it * nwarps + warp] + warp_scan_incl(c) - c;
        const int r0 = ((it0 + it) * blockDim.x + threadIdx.x) * 8;
        for (unsigned m = h; m; m &= m - 1) {
          const int r = r0 + __ffs(m) - 1;
          if (glist) glist[pos] = r; else slist[pos] = r;
          ++pos;
        }
      }
      running = carry_smem;
      __syncthreads();
    }
    }
  } else if (n < 2) {
    return;
  }

  if (a.debug & 2) return;
  if (n <= a.mask_rows) {
    // ---------------- small segment: full bit mask in shared memory + word-serial resolve ----------------
    float4 *boxes = reinterpret_cast<float4 *>(smem_raw);
    unsigned long long *mask = reinterpret_cast<unsigned long long *>(smem_raw + kOffMask);
    const int W = (n + 63) >> 6;
    float *areas = reinterpret_cast<float *>(smem_raw + kOffArea);
    unsigned short *queue = reinterpret_cast<unsigned short *>(smem_raw + kOffQueue) + warp * kNmsQueue;
    const int npad = (n + 31) & ~31;
    for (int q = threadIdx.x; q < npad; q += blockDim.x) {
      if (q < n) {
        boxes[q] = stage_box(__ldg(row_box + (identity ? q : slist[q])), &areas[q]);
      } else {
        boxes[q] = make_float4(__int_as_float(0x7f800000), 0.f, 0.f, 0.f);  // overlaps nothing
      }
    }
    for (int q = threadIdx.x; q < n * W; q += blockDim.x) mask[q] = 0ull;
    if (threadIdx.x < 8) rowany[threadIdx.x] = 0ull;  // W <= 5
    const int ngroups = npad >> 5;
    const int nunits = ngroups * (ngroups + 1) / 2;  // <= 55
    if ((int)threadIdx.x < nunits) {  // unit u -> (row group, column group >= row group), row-major over the triangle
      int rg = 0, rem = threadIdx.x;
      while (rem >= ngroups - rg) {
        rem -= ngroups - rg;
        ++rg;
      }
      unit_tab[threadIdx.x] = (unsigned char)((rg << 4) | (rg + rem));
    }
    __syncthreads();
    // mask[i * W + w] bit j: row i suppresses row 64 w + j (> i).  A unit = 32 consecutive rows (one per lane) x 32
    // consecutive columns, dealt round-robin to the warps, in two phases so that no lane idles in a divergent branch:
    //   (A) branch-free overlap test against the 32 column boxes (one broadcast LDS.128 each) -> candidate bits;
    //   (B) the candidates of the whole unit (a few percent of the pairs) are compacted into a per-warp queue and the
    //       exact IoU >= thr test runs on dense lanes, hits are OR-ed into the mask with shared-memory atomics.
    unsigned *mask32 = reinterpret_cast<unsigned *>(mask);
    constexpr int kWarps = kNmsThreads / 32;
    {
      for (int u = warp; u < nunits; u += kWarps) {
        const unsigned rc = unit_tab[u];
        const int rg = (int)(rc >> 4), cg = (int)(rc & 15u);
        const int i = (rg << 5) + lane;
        const bool row_ok = i < n && !(a.debug & 1);
        float4 bi = boxes[i];
        if (!row_ok) bi.x = __int_as_float(0x7f800000);  // fails every overlap test
        // Necessary condition for IoU >= thr: the intersection is at least thr x the width and height of row i's
        // box (inter <= iw * h_i and union >= w_i * h_i), i.e. box j must reach into box i shrunk by thr x (w_i, h_i)
        // on every side.  The shrunk box is rounded outwards and uses thr(1 - 2^-10), far more than the 2^-21 the
        // roundings of the exact test can move the decision; boxes too small for that error analysis (areas near
        // the denormal range) are not shrunk.  Costs the same four compares as a plain overlap test and halves the
        // pairs that reach the exact test.
        const float wi = __fsub_rd(bi.z, bi.x), hi = __fsub_rd(bi.w, bi.y);
        const bool shrink_ok = wi >= 0x1p-40f && hi >= 0x1p-40f;
        const float tw = shrink_ok ? __fmul_rd(thr.shrink, wi) : 0.f, th = shrink_ok ? __fmul_rd(thr.shrink, hi) : 0.f;
        const float sl = __fadd_rd(bi.x, tw), sr = __fsub_ru(bi.z, tw), st = __fadd_rd(bi.y, th), sb = __fsub_ru(bi.w, th);
        const int jb = cg << 5;
        unsigned cand = 0u;
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) {
          const float4 bj = boxes[jb + jj];
          // four chained FSETP + one predicated OR (the compiler's own lowering spends 8-9 instructions on selects)
          asm("{\n\t.reg .pred p;\n\t"
              "setp.gt.f32 p, %1, %2;\n\t"
              "setp.gt.and.f32 p, %3, %4, p;\n\t"
              "setp.gt.and.f32 p, %5, %6, p;\n\t"
              "setp.gt.and.f32 p, %7, %8, p;\n\t"
              "@p or.b32 %0, %0, %9;\n\t}"
              : "+r"(cand)
              : "f"(sr), "f"(bj.x), "f"(bj.z), "f"(sl), "f"(sb), "f"(bj.y), "f"(bj.w), "f"(st), "r"(1u << jj));
        }
        if (cg == rg) cand &= lane == 31 ? 0u : ~0u << (lane + 1);  // j > i
        const int c = __popc(cand);
        const int incl = warp_scan_incl(c);
        const int total = __shfl_sync(kFullMask, incl, 31);
        for (int base = 0; base < total; base += kNmsQueue) {
          int pos = incl - c - base;
          if (total <= kNmsQueue) {
            for (unsigned m = cand; m; m &= m - 1, ++pos) queue[pos] = (unsigned short)((lane << 5) | (__ffs(m) - 1));
          } else {
            for (unsigned m = cand; m; m &= m - 1, ++pos) {
              if (pos >= 0 && pos < kNmsQueue) queue[pos] = (unsigned short)((lane << 5) | (__ffs(m) - 1));
            }
          }
          __syncwarp();
          const int cnt = min(kNmsQueue, total - base);
          for (int q = lane; q < cnt; q += 32) {
            const unsigned e = queue[q];
            const int r = (rg << 5) + (int)(e >> 5), jj = (int)(e & 31u);
            if (iou_ge(boxes[r], areas[r], boxes[jb + jj], areas[jb + jj], thr)) {
              atomicOr(&mask32[r * 2 * W + cg], 1u << jj);
              atomicOr(&rowany[r >> 6], 1ull << (r & 63));
            }
          }
          __syncwarp();
        }
      }
    }
    __syncthreads();
    if (warp == 0) {
      // lane w owns word w of the removed set.  Chunk c: only rows that suppress something (rowany) need the
      // serial treatment; every lane resolves them redundantly from broadcast reads, then the rows of the live
      // suppressors are OR-ed into the later words.
      unsigned long long remv = 0ull;
      for (int c = 0; c < W; ++c) {
        unsigned long long cur = __shfl_sync(kFullMask, remv, c);
        unsigned long long live = 0ull;
        for (unsigned long long cand = rowany[c]; cand; cand &= cand - 1) {
          const int t = __ffsll((long long)cand) - 1;
          if (!((cur >> t) & 1ull)) {
            live |= 1ull << t;
            cur |= mask[((c << 6) + t) * W + c];
          }
        }
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
      // suppressed rows: only the id field is overwritten (multibox_detection.cc:163)
      for (int w = 0; w < W; ++w) {
        const unsigned long long word = __shfl_sync(kFullMask, remv, w);
        for (int half = 0; half < 2; ++half) {
          const int q = (w << 6) + half * 32 + lane;
          if (q < n && ((word >> (half * 32 + lane)) & 1ull)) out[(size_t)(identity ? q : slist[q]) * 7] = -1.f;
        }
      }
    }
    return;
  }

  // ---------------- large segment: 64-row chunks, ballot mask + serial resolve + parallel sweep ----------------
  float4 *boxes;
  unsigned char *dead;
  float *areas;
  unsigned long long *sm_word = reinterpret_cast<unsigned long long *>(smem_raw + kOffWord);
  if (n <= a.smem_rows) {
    boxes = reinterpret_cast<float4 *>(smem_raw);
    dead = smem_raw + kOffDead;
    areas = reinterpret_cast<float *>(smem_raw + kOffAreaL);
  } else {
    // claim scratch in segment order (the list region doubles as the allocator for identity segments)
    if (identity) {
      if (threadIdx.x == 0) sm_base = atomicAdd(&a.cursor[b], n);
      __syncthreads();
    }
    boxes = a.seg_box + (size_t)b * a.A + sm_base;
    dead = a.seg_dead + (size_t)b * a.A + sm_base;
    areas = a.seg_area + (size_t)b * a.A + sm_base;
  }
  for (int q = threadIdx.x; q < n; q += blockDim.x) {
    boxes[q] = stage_box(__ldg(row_box + (identity ? q : glist[q])), &areas[q]);
    dead[q] = 0;
  }
  __syncthreads();
  for (int c0 = 0; c0 < n; c0 += 64) {
    const int m = min(64, n - c0);
    // (1) 64x64 upper-triangular suppression mask of the chunk, one row per warp iteration
    if (warp < 2) {
      const int t = warp * 32 + lane;
      const unsigned bal = __ballot_sync(kFullMask, t < m && dead[c0 + t]);
      if (lane == 0) sm_deadbits[warp] = bal;
    }
    for (int i = warp; i < m; i += nwarps) {
      unsigned lo = 0, hi = 0;
      if (!dead[c0 + i]) {
        const float4 bi = boxes[c0 + i];
        const float ai = areas[c0 + i];
        const int j0 = lane, j1 = lane + 32;
        const bool s0 = j0 > i && j0 < m && suppresses(bi, ai, boxes[c0 + j0], areas[c0 + j0], thr);
        const bool s1 = j1 > i && j1 < m && suppresses(bi, ai, boxes[c0 + j1], areas[c0 + j1], thr);
        lo = __ballot_sync(kFullMask, s0);
        hi = __ballot_sync(kFullMask, s1);
      }
      if (lane == 0) sm_word[i] = ((unsigned long long)hi << 32) | lo;
    }
    __syncthreads();
    // (2) serial greedy resolve inside the chunk on 64-bit words
    if (threadIdx.x == 0) {
      unsigned long long dm = ((unsigned long long)sm_deadbits[1] << 32) | sm_deadbits[0];
      unsigned long long alive = 0;
      for (int t = 0; t < m; ++t)
        if (!((dm >> t) & 1ull)) {
          alive |= 1ull << t;
          dm |= sm_word[t];
        }
      sm_alive = alive;
      sm_deadmask = dm;
    }
    __syncthreads();
    const unsigned long long alive = sm_alive;
    if ((int)threadIdx.x < m) dead[c0 + threadIdx.x] = (unsigned char)((sm_deadmask >> threadIdx.x) & 1ull);
    // (3) sweep every later row against the surviving pivots of this chunk
    for (int j = c0 + 64 + threadIdx.x; j < n; j += blockDim.x) {
      if (dead[j]) continue;
      const float4 bj = boxes[j];
      const float aj = areas[j];
      unsigned long long rem = alive;
      while (rem) {
        const int t = __ffsll((long long)rem) - 1;
        rem &= rem - 1;
        if (suppresses(boxes[c0 + t], areas[c0 + t], bj, aj, thr)) {
          dead[j] = 1;
          break;
        }
      }
    }
    __syncthreads();
  }
  for (int q = threadIdx.x; q < n; q += blockDim.x)
    if (dead[q]) out[(size_t)(identity ? q : glist[q]) * 7] = -1.f;
}

__global__ void __launch_bounds__(kNmsThreads, 5) det_nms_kernel(const __grid_constant__ NmsArgs a) {
  __shared__ __align__(16) unsigned char stage[kNmsBodyBytes];
  nms_final_order_body(a, (int)blockIdx.y, (int)blockIdx.x, stage);
}

// ====================================================================================================
// Fork/join pipeline (the default for the VOC / Cityscapes heads):
//
//   det_stream_bulk_kernel<.., kV2>  as above + a class-grouped copy of every tile's boxes (cbox / crank / tile_cls)
//   det_sort_kernel (grid B)  ||  det_pair_kernel (grid (C-1, B))      -- both depend only on the stream kernel
//
// Which pairs of a class overlap by IoU >= thr does not depend on the row order -- only the greedy resolve does
// (multibox_detection.cc:153-167 walks the rows in final order).  det_pair_kernel therefore builds, per (image,
// class), the member list in pass-1 RANK order straight from the tiles' class-grouped runs (no pass over all V rows)
// and the SYMMETRIC suppression mask of the segment while the sort kernel is still selecting and sorting the
// nms_topk head and moving the tail rows [nkeep, V) to their final positions (all columns but the id, which the
// resolve owns).  The final order of a class is  [its head rows in sorted order] ++ [its members
// of rank >= nkeep in rank order]; a row can appear in both parts (the tail quirk) or in neither (rank < nkeep but
// not among the nms_topk best).  The resolve replays the greedy loop on that sequence with the precomputed mask:
//   * a node (row of the final order) is kept iff no EARLIER adjacent node is kept; evaluated as a fixed point over
//     all nodes in parallel (a node dies as soon as one earlier neighbour is known alive and lives once all of them
//     are known dead; the earliest undecided node always decides; real overlap graphs settle in a few rounds);
//   * head nodes and tail nodes of the same member are distinct nodes, and a self bit (IoU(box, box) >= thr, i.e. the
//     box is not degenerate) lets a kept head row remove its own duplicate in the tail, as the reference does.
// The pair kernel is launched as a PROGRAMMATIC DEPENDENT of the sort kernel (whose CTAs trigger
// griddepcontrol.launch_dependents on entry): it starts once all 32 sort CTAs are resident -- so the big sort CTAs are
// never locked out by 640 small ones -- runs its pair tests beside the sort, executes griddepcontrol.wait (the sort
// grid has completed and its writes are visible) and resolves its segment right there from shared memory.  Only
// segments with more than mask_rows members take nms_final_order_body (chunk sweep on final order) after the wait.
struct PairArgs {
  float *out;
  const int *tile_count;
  const unsigned *tile_cls;
  const float *slot_rows;
  const unsigned short *slot_cls;
  const float4 *slot_box;
  const float4 *cbox;
  const unsigned short *crank;
  unsigned short *row_cls;
  float4 *row_box;
  const int *nms_rows;   // resolve kernel only
  const int *head_rank;
  const int2 *head_list;   // per-class head lists written by the sort kernel (0 < nms_topk <= kHeadCap), or null
  const int *head_off;
  int A, T, tile, Apad, cls_stride, nfg;
  float nms_threshold;
  int nms_topk, mask_rows;
};

// Pass-1 row of a tail member (rank >= nkeep), copied from its tile slot to its final position `rank`; the resolve
// later overwrites the id of suppressed rows only (multibox_detection.cc:163).
__device__ __forceinline__ void copy_tail_row(const PairArgs &a, int b, size_t slot, int rank) {
  const float *src = a.slot_rows + ((size_t)b * a.Apad + slot) * 7;
  const float s0 = src[0], s1 = src[1], s2 = src[2], s3 = src[3], s4 = src[4], s5 = src[5], s6 = src[6];
  float *o = a.out + ((size_t)b * a.A + rank) * 7;
  o[0] = s0;
  o[1] = s1;
  o[2] = s2;
  o[3] = s3;
  o[4] = s4;
  o[5] = s5;
  o[6] = s6;
}


// Exclusive block scan of one 64-bit value per thread (two packed 32-bit counters); `smem`: blockDim.x/32 + 1 words.
__device__ __forceinline__ unsigned long long block_scan_excl_u64(unsigned long long v, unsigned long long *smem,
                                                                  unsigned long long *total) {
  unsigned long long incl = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const unsigned long long o = __shfl_up_sync(kFullMask, incl, d);
    if ((int)lane_id() >= d) incl += o;
  }
  const unsigned w = warp_id(), l = lane_id(), nw = (blockDim.x + 31) >> 5;
  if (l == 31) smem[w] = incl;
  __syncthreads();
  if (w == 0) {
    const unsigned long long s = l < nw ? smem[l] : 0ull;
    unsigned long long si = s;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned long long o = __shfl_up_sync(kFullMask, si, d);
      if ((int)l >= d) si += o;
    }
    if (l < nw) smem[l] = si - s;
    if (l == 31) smem[nw] = si;
  }
  __syncthreads();
  *total = smem[nw];
  return smem[w] + incl - v;
}

__device__ __forceinline__ int lower_bound_i32(const int *v, int n, int x) {  // first index with v[i] >= x
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (v[mid] < x) lo = mid + 1; else hi = mid;
  }
  return lo;
}

constexpr int kResolveCand = 128;  // nodes with an earlier neighbour that get a compact neighbour list
constexpr int kResolveNbr = 8;     // entries of such a list; longer ones fall back to the bit-mask walk

struct ResolveScratch {
  int *hq, *tmpq, *tmpr;      // [kNmsMaskRows] head positions of the class, ordered / as found (+ their ranks)
  unsigned short *hm, *hp;    // [kNmsMaskRows] member of head node k / head node of member m (0xffff: none)
  unsigned short *open;       // [2 * kNmsMaskRows] nodes that have an earlier neighbour
  unsigned char *status;      // [2 * kNmsMaskRows] 0 undecided, 1 kept, 2 suppressed
  unsigned short *nbr;        // [kResolveCand * kResolveNbr] earlier neighbours (node ids) of the listed nodes
  unsigned char *nnbr;        // [kResolveCand] their number (0xff: too many, walk the mask)
  int *counter;               // [2]
};

// Part of the resolve that does not depend on the sort: call (whole CTA) before griddepcontrol.wait.
__device__ __forceinline__ void resolve_prepare(const int n, const ResolveScratch s) {
  for (int q = threadIdx.x; q < n; q += blockDim.x) s.hp[q] = (unsigned short)0xffffu;
  for (int q = threadIdx.x; q < 2 * n; q += blockDim.x) s.status[q] = 1;  // kept unless an earlier neighbour says no
  if (threadIdx.x < 2) s.counter[threadIdx.x] = 0;
  __syncthreads();
}

// Earlier neighbours of node q in final order, as node ids: calls f(id) for each.
template <typename F>
__device__ __forceinline__ void for_each_earlier(const int q, const int nh, const int m0, const int W,
                                                 const unsigned long long *mask, const unsigned long long *rowany,
                                                 const unsigned long long *selfadj, const ResolveScratch &s, F f) {
  const bool tail = q >= nh;
  const int m = tail ? m0 + (q - nh) : (int)s.hm[q];
  if ((rowany[m >> 6] >> (m & 63)) & 1ull) {
    for (int w = 0; w < W; ++w) {
      for (unsigned long long bits = mask[m * W + w]; bits; bits &= bits - 1) {
        const int j = (w << 6) + __ffsll((long long)bits) - 1;
        const int hj = s.hp[j];
        if (hj != 0xffff && (tail || hj < q)) f(hj);          // the head copy of member j precedes this node
        if (tail && j >= m0 && j < m) f(nh + j - m0);         // the tail copy of member j precedes this tail node
      }
    }
  }
  // a tail row's own copy in the sorted head
  if (tail && s.hp[m] != 0xffff && ((selfadj[m >> 6] >> (m & 63)) & 1ull)) f((int)s.hp[m]);
}

// Greedy resolve of one segment (n <= kNmsMaskRows members in rank order, symmetric mask with row stride W) in the
// image's final row order.  Whole CTA, after resolve_prepare and after the sort's results are visible.
__device__ __forceinline__ void resolve_segment(const PairArgs &a, const int b, const int c, const int n, const int V,
                                                const unsigned long long *mask, const int *ranks,
                                                const unsigned long long *rowany, const unsigned long long *selfadj,
                                                const ResolveScratch s) {
  const int nkeep = (a.nms_topk > 0 && a.nms_topk < V) ? a.nms_topk : V;
  const int W = (n + 63) >> 6;
  // head rows of this class (their positions q in the sorted head; every one of them is a member, so <= n)
  int nh;
  if (a.head_list) {
    // the sort kernel (another SM) handed over this class's head rows already grouped and in position order
    const int *ho = a.head_off + (size_t)b * (kV2ClsPad + 1);
    const int off0 = __ldcg(ho + c);
    nh = __ldcg(ho + c + 1) - off0;
    const int2 *hl = a.head_list + (size_t)b * kHeadCap + off0;
    for (int k = threadIdx.x; k < nh; k += blockDim.x) {
      const int2 e = __ldcg(hl + k);  // (head position, pass-1 rank)
      const int m = lower_bound_i32(ranks, n, e.y);
      s.hq[k] = e.x;
      s.hm[k] = (unsigned short)m;
      s.hp[m] = (unsigned short)k;
    }
    DSPMB_STAMP(5);
  } else {
    const int *hinfo = a.head_rank + (size_t)b * a.Apad;  // rank | class << 24, written by the sort kernel (another SM)
    for (int q = threadIdx.x; q < nkeep; q += blockDim.x) {
      const int info = __ldcg(hinfo + q);
      if ((info >> 24) == c) {
        const int k = atomicAdd(&s.counter[0], 1);
        s.tmpq[k] = q;
        s.tmpr[k] = info & 0xffffff;
      }
    }
    __syncthreads();
    DSPMB_STAMP(5);
    nh = s.counter[0];
    for (int k = threadIdx.x; k < nh; k += blockDim.x) {
      const int myq = s.tmpq[k];
      int ord = 0;
      for (int i = 0; i < nh; ++i) ord += s.tmpq[i] < myq ? 1 : 0;
      const int m = lower_bound_i32(ranks, n, s.tmpr[k]);
      s.hq[ord] = myq;
      s.hm[ord] = (unsigned short)m;
      s.hp[m] = (unsigned short)ord;
    }
  }
  const int m0 = lower_bound_i32(ranks, n, nkeep);  // members m0 .. n-1 are the tail rows of the class
  // Nodes in final order: q < nh is head row hm[q]; q >= nh is tail member m0 + (q - nh).  The greedy loop of
  // multibox_detection.cc:153-167 keeps a node iff no EARLIER adjacent node is kept.  Evaluated as a fixed point: a
  // node is suppressed as soon as one earlier neighbour is known kept and kept once all of them are known suppressed;
  // the earliest open node always decides.  Only nodes that HAVE an earlier neighbour take part (the others are kept,
  // which is what resolve_prepare wrote): they are listed once, densely, together with the node ids of those
  // neighbours, and the rounds -- a handful for real overlap graphs -- only chase those short lists.
  const int ns = nh + (n - m0);
  __syncthreads();
  DSPMB_STAMP(6);
  for (int q0 = 0; q0 < ns; q0 += blockDim.x) {
    const int q = q0 + (int)threadIdx.x;
    bool has = false;
    if (q < ns) {
      const bool tail = q >= nh;
      const int m = tail ? m0 + (q - nh) : (int)s.hm[q];
      has = ((rowany[m >> 6] >> (m & 63)) & 1ull) || (tail && s.hp[m] != 0xffff);
    }
    const unsigned bal = __ballot_sync(kFullMask, has);
    if (bal) {
      int base = 0;
      if (lane_id() == 0) base = atomicAdd(&s.counter[1], __popc(bal));
      base = __shfl_sync(kFullMask, base, 0);
      if (has) s.open[base + __popc(bal & ((1u << lane_id()) - 1u))] = (unsigned short)q;
    }
  }
  __syncthreads();
  const int nopen = s.counter[1];
  volatile unsigned char *status = s.status;
  for (int i = threadIdx.x; i < nopen; i += blockDim.x) {  // neighbour lists; a node without any is kept
    const int q = s.open[i];
    int cnt = 0;
    for_each_earlier(q, nh, m0, W, mask, rowany, selfadj, s, [&](int id) {
      if (i < kResolveCand && cnt < kResolveNbr) s.nbr[i * kResolveNbr + cnt] = (unsigned short)id;
      ++cnt;
    });
    if (i < kResolveCand) s.nnbr[i] = cnt <= kResolveNbr ? (unsigned char)cnt : (unsigned char)0xff;
    if (cnt) status[q] = 0;
  }
  __syncthreads();
  while (true) {
    bool open = false;
    for (int i = threadIdx.x; i < nopen; i += blockDim.x) {
      const int q = s.open[i];
      if (status[q]) continue;
      bool any_alive = false, any_open = false;
      const int cnt = i < kResolveCand ? (int)s.nnbr[i] : 0xff;
      if (cnt != 0xff) {
        for (int k = 0; k < cnt; ++k) {
          const int st = status[s.nbr[i * kResolveNbr + k]];
          any_alive |= st == 1;
          any_open |= st == 0;
        }
      } else {
        for_each_earlier(q, nh, m0, W, mask, rowany, selfadj, s, [&](int id) {
          const int st = status[id];
          any_alive |= st == 1;
          any_open |= st == 0;
        });
      }
      if (any_alive) status[q] = 2;
      else if (!any_open) status[q] = 1;
      else open = true;
    }
    if (!__syncthreads_or(open)) break;
  }
  DSPMB_STAMP(7);
  // ids: a suppressed row gets -1 and keeps everything else (multibox_detection.cc:163)
  float *out = a.out + (size_t)b * a.A * 7;
  for (int q = threadIdx.x; q < ns; q += blockDim.x)
    if (s.status[q] == 2) out[(size_t)(q < nh ? s.hq[q] : ranks[m0 + (q - nh)]) * 7] = -1.f;
  DSPMB_STAMP(8);
}

__global__ void __launch_bounds__(kNmsThreads, 6) det_pair_kernel(const __grid_constant__ PairArgs a,
                                                                  const __grid_constant__ NmsArgs na) {
  constexpr int kOffMask = kNmsMaskRows * 16, kOffArea = kOffMask + kNmsMaskRows * 5 * 8;
  constexpr int kOffRank = kOffArea + kNmsMaskRows * 4, kOffQueue = kOffRank + kNmsMaskRows * 4;
  constexpr int kOffCol = kOffQueue + (kNmsThreads / 32) * kNmsQueue * 2, kOffRow = kOffCol + kNmsMaskRows * 4;
  constexpr int kOffSlot = kOffRow + kNmsMaskRows * 4;
  constexpr int kBytes = kOffSlot + kNmsMaskRows * 4;
  constexpr int kWarps = kNmsThreads / 32;
  static_assert(kNmsMaskRows * 16 >= 3 * kNmsMaskRows * 4 + 2 * kNmsMaskRows * 2, "hq / tmpq / tmpr / open alias the box stage");
  static_assert(kWarps * kNmsQueue * 2 >= 2 * kNmsMaskRows * 2 + 2 * kNmsMaskRows, "hm / hp / status alias the queues");
  static_assert(kOffSlot - kOffCol >= kResolveCand * kResolveNbr * 2 + kResolveCand, "neighbour lists alias the packed boxes");
  __shared__ __align__(16) unsigned char smem_raw[kBytes];
  __shared__ int sm_tbase[kV2MaxTiles], sm_mbase[kV2MaxTiles];
  __shared__ unsigned short sm_coff[kV2MaxTiles];
  __shared__ unsigned long long scan_smem[kWarps + 1];
  __shared__ unsigned long long sm_carry;
  __shared__ unsigned long long rowany[8], selfadj[8];
  __shared__ float4 gbb[kNmsMaskRows / 32];
  __shared__ unsigned char unit_tab[64];
  __shared__ int sm_nunits, sm_next, sm_ctr[2];
  TraceScope trace_(2);
  DSPMB_STAMP(0);

  const int b = blockIdx.y, c = blockIdx.x, T = a.T;
  const unsigned lane = lane_id(), warp = warp_id();
  const int *cnt = a.tile_count + (size_t)b * T;
  const unsigned *tcls = a.tile_cls + ((size_t)b * kV2ClsPad + c) * T;
  if (threadIdx.x == 0) {
    sm_carry = 0ull;
    sm_nunits = 0;
    sm_next = 0;
  }
  if (threadIdx.x < 8) {
    rowany[threadIdx.x] = 0ull;
    selfadj[threadIdx.x] = 0ull;
  }
  __syncthreads();
  // rank base of every tile (low word) and member base of this class in every tile (high word), one packed scan
  for (int base = 0; base < T; base += blockDim.x) {
    const int t = base + threadIdx.x;
    const unsigned cw = t < T ? tcls[t] : 0u;
    const unsigned long long v = t < T ? ((unsigned long long)(unsigned)cnt[t] | ((unsigned long long)(cw >> 16) << 32)) : 0ull;
    unsigned long long total;
    const unsigned long long ex = block_scan_excl_u64(v, scan_smem, &total);
    const unsigned long long carry = sm_carry;
    if (t < T) {
      sm_tbase[t] = (int)(unsigned)(carry + ex);
      sm_mbase[t] = (int)((carry + ex) >> 32);
      sm_coff[t] = (unsigned short)(cw & 0xffffu);
    }
    __syncthreads();
    if (threadIdx.x == 0) sm_carry = carry + total;
    __syncthreads();
  }
  const int V = (int)(unsigned)sm_carry, n = (int)(sm_carry >> 32);
  const int nkeep = (a.nms_topk > 0 && a.nms_topk < V) ? a.nms_topk : V;
  static_assert(kBytes >= kNmsBodyBytes, "the final-order path reuses this kernel's stage");
  // member m of the class (rank order) -> tile, slot inside the tile's run, pass-1 rank
  auto locate = [&](int m, size_t *cslot, size_t *slot) -> int {
    int lo = 0, hi = T;  // first tile whose member base exceeds m
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (sm_mbase[mid] <= m) lo = mid + 1; else hi = mid;
    }
    const int t = lo - 1;
    *cslot = (size_t)t * a.tile + sm_coff[t] + (m - sm_mbase[t]);
    const int k = (int)a.crank[(size_t)b * a.cls_stride + *cslot];
    *slot = (size_t)t * a.tile + k;
    return sm_tbase[t] + k;
  };
  if (V < 1 || n < 1 || n > a.mask_rows) {
    if (V >= 1 && n > a.mask_rows) {
      // a segment too large for the shared-memory mask: its tail rows are moved here like everybody's ...
      for (int m = threadIdx.x; m < n; m += blockDim.x) {
        size_t cslot, slot;
        const int rank = locate(m, &cslot, &slot);
        if (rank >= nkeep) copy_tail_row(a, b, slot, rank);
      }
    }
    // every CTA waits for the sort grid before it exits, so that the completion of THIS grid implies the sort's
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (V >= 1 && n > a.mask_rows) {
      // ... then chunk-sweep NMS on the rows in FINAL order, which the sort grid has just completed (row_cls / row_box
      // of every row of the image: the sort kernel saw the same class sizes and copied the tail's as well)
      __syncthreads();
      nms_final_order_body(na, b, c, smem_raw);
    }
    return;
  }

  // ---- members of class c in rank order: member m lives in the tile whose member base is the last one <= m ----
  float4 *boxes = reinterpret_cast<float4 *>(smem_raw);
  unsigned long long *mask = reinterpret_cast<unsigned long long *>(smem_raw + kOffMask);
  float *areas = reinterpret_cast<float *>(smem_raw + kOffArea);
  int *ranks = reinterpret_cast<int *>(smem_raw + kOffRank);
  unsigned short *queue = reinterpret_cast<unsigned short *>(smem_raw + kOffQueue) + warp * kNmsQueue;
  unsigned *colpack = reinterpret_cast<unsigned *>(smem_raw + kOffCol);
  unsigned *rowpack = reinterpret_cast<unsigned *>(smem_raw + kOffRow);
  int *mslot = reinterpret_cast<int *>(smem_raw + kOffSlot);  // slot of member m inside the image's tile slots
  const NmsThr thr = make_thr(a.nms_threshold);
  const int W = (n + 63) >> 6, npad = (n + 31) & ~31;
  for (int m = threadIdx.x; m < npad; m += blockDim.x) {
    unsigned cp = 0x7f7f7f7fu, rp = 0x80808080u;  // padded / degenerate rows: reach nothing, are reached by nothing
    if (m < n) {
      size_t s0, slot;
      const int rank = locate(m, &s0, &slot);
      ranks[m] = rank;
      mslot[m] = (int)slot;
      const float4 bi = stage_box(__ldg(a.cbox + (size_t)b * a.Apad + s0), &areas[m]);
      boxes[m] = bi;
      if (bi.x < __int_as_float(0x7f800000)) {
        // Candidate filter, 4 x 7 bit per box.  Necessary condition for IoU >= thr: box j must reach into box i shrunk
        // by thr x (w_i, h_i) on every side (the intersection is at least thr x the width and height of box i, because
        // inter <= iw * h_i and union >= w_i * h_i).  The shrunk box is rounded outwards with thr(1 - 2^-10), far more
        // than the 2^-21 the roundings of the exact test can move the decision; boxes too small for that error
        // analysis are not shrunk.  The four compares  sr > x1_j, x2_j > sl, sb > y1_j, y2_j > st  are taken on
        // coordinates quantised to 1/127 -- the row side rounded up, the column side down, so no true candidate is lost
        // (19.9 k instead of 16.5 k candidates per image on the benchmark batch) -- as ONE subtraction of packed bytes:
        // (row | 0x80) - col keeps a byte's top bit iff row >= col, and no byte ever borrows from its neighbour.
        const float wi = __fsub_rd(bi.z, bi.x), hi = __fsub_rd(bi.w, bi.y);
        const bool shrink_ok = wi >= 0x1p-40f && hi >= 0x1p-40f;
        const float tw = shrink_ok ? __fmul_rd(thr.shrink, wi) : 0.f, th = shrink_ok ? __fmul_rd(thr.shrink, hi) : 0.f;
        const float sl = __fadd_rd(bi.x, tw), sr = __fsub_ru(bi.z, tw), st = __fadd_rd(bi.y, th), sb = __fsub_ru(bi.w, th);
        auto qd = [](float v) { return (unsigned)min(127, max(0, __float2int_rd(v * 127.f))); };
        auto qu = [](float v) { return (unsigned)min(127, max(0, __float2int_ru(v * 127.f))); };
        cp = qd(bi.x) | (qd(fsub(1.f, bi.z)) << 8) | (qd(bi.y) << 16) | (qd(fsub(1.f, bi.w)) << 24);
        rp = 0x80808080u | qu(sr) | (qu(fsub(1.f, sl)) << 8) | (qu(sb) << 16) | (qu(fsub(1.f, st)) << 24);
      }
    } else {
      boxes[m] = make_float4(__int_as_float(0x7f800000), 0.f, 0.f, 0.f);  // overlaps nothing
    }
    colpack[m] = cp;
    rowpack[m] = rp;
  }
  for (int q = threadIdx.x; q < n * W; q += blockDim.x) mask[q] = 0ull;
  __syncthreads();
  DSPMB_STAMP(2);

  // ---- bounding box of every 32-row group (unit culling) and the self bits ----
  const int ngroups = npad >> 5;
  for (int g = warp; g < ngroups; g += kWarps) {
    const int q = (g << 5) + lane;
    const float4 bx = boxes[q];
    const bool ok = bx.x < __int_as_float(0x7f800000);  // padded / degenerate rows never pass the candidate test
    float x1 = ok ? bx.x : __int_as_float(0x7f800000), y1 = ok ? bx.y : __int_as_float(0x7f800000);
    float x2 = ok ? bx.z : __int_as_float(0xff800000), y2 = ok ? bx.w : __int_as_float(0xff800000);
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
      x1 = fminf(x1, __shfl_xor_sync(kFullMask, x1, m));
      y1 = fminf(y1, __shfl_xor_sync(kFullMask, y1, m));
      x2 = fmaxf(x2, __shfl_xor_sync(kFullMask, x2, m));
      y2 = fmaxf(y2, __shfl_xor_sync(kFullMask, y2, m));
    }
    if (lane == 0) gbb[g] = make_float4(x1, y1, x2, y2);
    // a row appearing both in the sorted head and in the tail suppresses its own copy iff IoU(box, box) >= thr
    const bool self = q < n && suppresses(bx, areas[q], bx, areas[q], thr);
    const unsigned bal = __ballot_sync(kFullMask, self);
    if (lane == 0) reinterpret_cast<unsigned *>(selfadj)[g] = bal;
  }
  __syncthreads();
  {  // units (row group <= column group) whose group boxes touch, in any order
    const int nall = ngroups * (ngroups + 1) / 2;  // <= 55
    if ((int)threadIdx.x < nall) {
      int rg = 0, rem = threadIdx.x;
      while (rem >= ngroups - rg) {
        rem -= ngroups - rg;
        ++rg;
      }
      const int cg = rg + rem;
      const float4 p = gbb[rg], q = gbb[cg];
      if (p.z > q.x && q.z > p.x && p.w > q.y && q.w > p.y) unit_tab[atomicAdd(&sm_nunits, 1)] = (unsigned char)((rg << 4) | cg);
    }
  }
  __syncthreads();
  const int nunits = sm_nunits;
  unsigned *mask32 = reinterpret_cast<unsigned *>(mask);
  unsigned *rowany32 = reinterpret_cast<unsigned *>(rowany);
  if (warp == kWarps - 1) {
    // Tail rows [nkeep, V) keep their pass-1 content at their pass-1 position (multibox_detection.cc:146-151): the
    // class's members of rank >= nkeep are moved from their tile slots to `out` by ONE warp while the others start on
    // the units (it joins them afterwards): a flat loop over (member, column) elements, eight independent loads in
    // flight per lane, so the copy costs a handful of memory round trips beside the pair tests instead of before them.
    // The resolve later overwrites the id of suppressed rows only.
    const int mt = lower_bound_i32(ranks, n, nkeep);
    const int total = (n - mt) * 7;
    const float *srows = a.slot_rows + (size_t)b * a.Apad * 7;
    float *orows = a.out + (size_t)b * a.A * 7;
    for (int base = 0; base < total; base += 256) {
      float v[8];
      int dst[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int idx = base + k * 32 + (int)lane;
        const int mm = (idx * 9363) >> 16, col = idx - mm * 7;  // idx / 7 for idx < 13107
        dst[k] = -1;
        if (idx < total) {
          v[k] = srows[(size_t)mslot[mt + mm] * 7 + col];
          dst[k] = ranks[mt + mm] * 7 + col;
        }
      }
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (dst[k] >= 0) orows[dst[k]] = v[k];
    }
  }
  while (true) {  // the warps draw units from a shared counter (diagonal units and culled neighbourhoods differ in cost)
    int u = 0;
    if (lane == 0) u = atomicAdd(&sm_next, 1);
    u = __shfl_sync(kFullMask, u, 0);
    if (u >= nunits) break;
    const unsigned rc = unit_tab[u];
    const int rg = (int)(rc >> 4), cg = (int)(rc & 15u);
    const unsigned rp = rowpack[(rg << 5) + lane];
    const int jb = cg << 5;
    unsigned cand = 0u;
#pragma unroll
    for (int j4 = 0; j4 < 8; ++j4) {
      const uint4 cp = *reinterpret_cast<const uint4 *>(colpack + jb + 4 * j4);  // one broadcast LDS.128 = 4 columns
      if ((~(rp - cp.x) & 0x80808080u) == 0u) cand |= 1u << (4 * j4);
      if ((~(rp - cp.y) & 0x80808080u) == 0u) cand |= 2u << (4 * j4);
      if ((~(rp - cp.z) & 0x80808080u) == 0u) cand |= 4u << (4 * j4);
      if ((~(rp - cp.w) & 0x80808080u) == 0u) cand |= 8u << (4 * j4);
    }
    if (cg == rg) cand &= lane == 31 ? 0u : ~0u << (lane + 1);  // each unordered pair once
    const int cnum = __popc(cand);
    const int incl = warp_scan_incl(cnum);
    const int total = __shfl_sync(kFullMask, incl, 31);
    for (int base = 0; base < total; base += kNmsQueue) {
      int pos = incl - cnum - base;
      if (total <= kNmsQueue) {
        for (unsigned m = cand; m; m &= m - 1, ++pos) queue[pos] = (unsigned short)((lane << 5) | (__ffs(m) - 1));
      } else {
        for (unsigned m = cand; m; m &= m - 1, ++pos)
          if (pos >= 0 && pos < kNmsQueue) queue[pos] = (unsigned short)((lane << 5) | (__ffs(m) - 1));
      }
      __syncwarp();
      const int qn = min(kNmsQueue, total - base);
      for (int q = lane; q < qn; q += 32) {
        const unsigned e = queue[q];
        const int r = (rg << 5) + (int)(e >> 5), jj = (int)(e & 31u), j = jb + jj;
        if (iou_ge(boxes[r], areas[r], boxes[j], areas[j], thr)) {
          atomicOr(&mask32[r * 2 * W + cg], 1u << jj);        // row r, column j  (j >> 5 == cg)
          atomicOr(&mask32[j * 2 * W + rg], 1u << (r & 31));  // and the transpose
          atomicOr(&rowany32[rg], 1u << (r & 31));
          atomicOr(&rowany32[cg], 1u << jj);
        }
      }
      __syncwarp();
    }
  }
  __syncthreads();
  ResolveScratch rs;
  rs.hq = reinterpret_cast<int *>(smem_raw);  // the box stage and the queues are dead now
  rs.tmpq = rs.hq + kNmsMaskRows;
  rs.tmpr = rs.tmpq + kNmsMaskRows;
  rs.open = reinterpret_cast<unsigned short *>(rs.tmpr + kNmsMaskRows);
  rs.hm = reinterpret_cast<unsigned short *>(smem_raw + kOffQueue);
  rs.hp = rs.hm + kNmsMaskRows;
  rs.status = reinterpret_cast<unsigned char *>(rs.hp + kNmsMaskRows);
  rs.nbr = reinterpret_cast<unsigned short *>(smem_raw + kOffCol);  // the packed filter boxes are dead as well
  rs.nnbr = reinterpret_cast<unsigned char *>(rs.nbr + kResolveCand * kResolveNbr);
  rs.counter = sm_ctr;
  resolve_prepare(n, rs);
  DSPMB_STAMP(3);
  // the sort grid has completed and flushed: head_rank and the head rows of `out` are final
  asm volatile("griddepcontrol.wait;" ::: "memory");
  DSPMB_STAMP(4);
  resolve_segment(a, b, c, n, V, mask, ranks, rowany, selfadj, rs);
}

// Ordered compaction of the surviving rows (id >= 0) of every image into (B, K, 7), padded with -1, plus the
// per-image count: the `det[:, 0] >= 0` filter every consumer of the op applies on the host
// (detect/multitask_detector.py:268-271, multi_solver.py:419-432), done on the device so that only K rows per
// image have to cross NVLink / PCIe.
__global__ void __launch_bounds__(256) det_compact_kernel(const float *__restrict__ out, const int *__restrict__ valid,
                                                          int A, int K, float *__restrict__ dst, int *__restrict__ counts) {
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
    const int keep = (r < V && src[(size_t)r * 7] >= 0.f) ? 1 : 0;
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

}  // namespace
}  // namespace dspmb

using namespace dspmb;

extern "C" int dspmb_detection_compact_f32(const float *out, const int32_t *valid_count, int B, int A, int K,
                                           float *dst, int32_t *counts, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  DSPMB_REQUIRE(B >= 0 && A > 0 && K > 0, "detection_compact: bad shape B=%d A=%d K=%d", B, A, K);
  DSPMB_REQUIRE(out && dst && counts, "detection_compact: NULL tensor");
  if (B == 0) return DSPMB_OK;
  {
    ProfileScope _p(kSlotDetCompact, stream);
    det_compact_kernel<<<B, 256, 0, stream>>>(out, valid_count, A, K, dst, counts);
  }
  DSPMB_CUDA_TRY(cudaGetLastError());
  return DSPMB_OK;
}

extern "C" int dspmb_debug_trace(unsigned long long *device_buffer) {
  DSPMB_CUDA_TRY(cudaMemcpyToSymbol(g_trace, &device_buffer, sizeof(device_buffer)));
  return DSPMB_OK;
}

extern "C" int dspmb_debug_stamps(unsigned long long *device_buffer) {
  DSPMB_CUDA_TRY(cudaMemcpyToSymbol(g_stamps, &device_buffer, sizeof(device_buffer)));
  return DSPMB_OK;
}

// The batch is processed as up to kMaxSplit groups of consecutive images (DSPMB_TUNE_DET_SPLIT): the select / sort /
// pair tests / resolve of group g only depend on the stream kernel of group g, so inside the library's CUDA graph they
// run on a parallel branch while the stream kernel of group g + 1 -- which is HBM-bound and leaves most issue slots
// idle -- is reading its class tensor.  Every group owns its own slice of the workspace.
constexpr int kMaxSplit = 4;
static int split_groups(int B, int want) {
  int g = want < 1 ? 1 : (want > kMaxSplit ? kMaxSplit : want);
  while (g > 1 && B / g < 4) --g;  // at least four images per group
  return g;
}
static size_t split_workspace_bytes(int B, int A, int C, int groups) {
  size_t total = 0;
  for (int g = 0; g < groups; ++g) {
    const int b0 = (int)((long long)B * g / groups), b1 = (int)((long long)B * (g + 1) / groups);
    total += align_up(carve(nullptr, b1 - b0, A, C).bytes, 256);
  }
  return total;
}

extern "C" size_t dspmb_detection_workspace_bytes(int B, int A, int C) {
  if (B <= 0 || A <= 0 || C <= 0) return 0;
  size_t need = 0;
  for (int g = 1; g <= kMaxSplit; ++g) {
    const size_t n = split_workspace_bytes(B, A, C, split_groups(B, g));
    need = n > need ? n : need;
  }
  return need;
}

// Tensor map of cls_prob viewed as a 2-D fp32 tensor (inner dimension: A anchors, outer: B*C class rows) with a box of
// {256 anchors, nfg rows}.  cuTensorMapEncodeTiled is resolved through the runtime (no link-time libcuda dependency).
static int make_cls_tensor_map(CUtensorMap *tmap, const float *cls_prob, int B, int C, int A, int nfg) {
  typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                               const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn encode = nullptr;
  if (!encode) {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) {
      cudaGetLastError();
      return DSPMB_ERR_CUDA;
    }
    encode = (EncodeFn)fn;
  }
  const cuuint64_t dims[2] = {(cuuint64_t)A, (cuuint64_t)B * C};
  const cuuint64_t strides[1] = {(cuuint64_t)A * sizeof(float)};
  const cuuint32_t box[2] = {256u, (cuuint32_t)nfg};
  const cuuint32_t estr[2] = {1u, 1u};
  const CUresult r = encode(tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(cls_prob), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? DSPMB_OK : DSPMB_ERR_CUDA;
}

static int detection_run(const HeadsArgs *heads, const float *cls_prob, const float *loc_pred, const float *anchors,
                         float *out, int B, int A, int C, float threshold, int clip, const float *variances,
                         float nms_threshold, int force_suppress, int nms_topk, int32_t *valid_count_out, void *workspace,
                         size_t workspace_bytes, void *stream_);

extern "C" int dspmb_detection_f32(const float *cls_prob, const float *loc_pred, const float *anchors, float *out,
                                   int B, int A, int C, float threshold, int clip, const float *variances,
                                   float nms_threshold, int force_suppress, int nms_topk, int32_t *valid_count_out,
                                   void *workspace, size_t workspace_bytes, void *stream_) {
  DSPMB_REQUIRE(cls_prob && loc_pred, "MultiBoxDetection: NULL tensor");
  return detection_run(nullptr, cls_prob, loc_pred, anchors, out, B, A, C, threshold, clip, variances, nms_threshold,
                       force_suppress, nms_topk, valid_count_out, workspace, workspace_bytes, stream_);
}

// Tile table of the head-fed stream kernel: whole cells per tile, scale by scale.
static int heads_describe(const float *const *cls_heads, const float *const *loc_heads, const int *head_hw,
                          const int *head_na, int nscales, int A, HeadsArgs *h) {
  DSPMB_REQUIRE(nscales >= 1 && nscales <= kMaxScales, "MultiBoxDetection (heads): 1..%d scales supported (got %d)", kMaxScales,
                nscales);
  DSPMB_REQUIRE(cls_heads && loc_heads && head_hw && head_na, "MultiBoxDetection (heads): NULL table");
  memset(h, 0, sizeof(*h));
  h->nscales = nscales;
  int tiles = 0, anchors = 0;
  for (int k = 0; k < nscales; ++k) {
    const int hw = head_hw[2 * k] * head_hw[2 * k + 1], na = head_na[k];
    DSPMB_REQUIRE(hw > 0 && na > 0 && na <= kHeadTile, "MultiBoxDetection (heads): bad scale %d (H*W=%d, anchors/cell=%d)", k, hw, na);
    DSPMB_REQUIRE(cls_heads[k] && loc_heads[k], "MultiBoxDetection (heads): NULL head %d", k);
    h->cls[k] = cls_heads[k];
    h->loc[k] = loc_heads[k];
    h->hw[k] = hw;
    h->na[k] = na;
    h->cpt[k] = kHeadTile / na;
    h->tile_off[k] = tiles;
    h->anchor_off[k] = anchors;
    tiles += ceil_div(hw, h->cpt[k]);
    anchors += hw * na;
  }
  h->tile_off[nscales] = tiles;
  DSPMB_REQUIRE(anchors == A, "MultiBoxDetection (heads): the heads hold %d anchors, the anchor tensor %d", anchors, A);
  return DSPMB_OK;
}

extern "C" size_t dspmb_detection_heads_workspace_bytes(int B, int A, int C, const int *head_hw, const int *head_na,
                                                        int nscales) {
  if (B <= 0 || A <= 0 || C <= 0 || !head_hw || !head_na || nscales < 1 || nscales > kMaxScales) return 0;
  int tiles = 0;
  for (int k = 0; k < nscales; ++k) {
    if (head_na[k] <= 0 || head_na[k] > kHeadTile) return 0;
    tiles += ceil_div(head_hw[2 * k] * head_hw[2 * k + 1], kHeadTile / head_na[k]);
  }
  return carve(nullptr, B, A, C, tiles, (size_t)tiles * kHeadTile).bytes;
}

extern "C" int dspmb_detection_heads_f32(const float *const *cls_heads, const float *const *loc_heads, const int *head_hw,
                                         const int *head_na, int nscales, const float *anchors, float *out, int B, int A,
                                         int C, float threshold, int clip, const float *variances, float nms_threshold,
                                         int force_suppress, int nms_topk, int32_t *valid_count_out, void *workspace,
                                         size_t workspace_bytes, void *stream_) {
  HeadsArgs h;
  const int rc = heads_describe(cls_heads, loc_heads, head_hw, head_na, nscales, A, &h);
  if (rc != DSPMB_OK) return rc;
  DSPMB_REQUIRE(C == 21 || C == 9, "MultiBoxDetection (heads): built for 21 (VOC) and 9 (Cityscapes) classes, got %d", C);
  return detection_run(&h, nullptr, nullptr, anchors, out, B, A, C, threshold, clip, variances, nms_threshold,
                       force_suppress, nms_topk, valid_count_out, workspace, workspace_bytes, stream_);
}

static int detection_run(const HeadsArgs *heads, const float *cls_prob, const float *loc_pred, const float *anchors,
                         float *out, int B, int A, int C, float threshold, int clip, const float *variances,
                         float nms_threshold, int force_suppress, int nms_topk, int32_t *valid_count_out, void *workspace,
                         size_t workspace_bytes, void *stream_) {
  // Shape CHECKs of MultiBoxDetectionProp::InferShape (multibox_detection-inl.h:149-171).
  DSPMB_REQUIRE(B >= 0 && A > 0 && C > 0, "MultiBoxDetection: bad shape B=%d A=%d C=%d", B, A, C);
  DSPMB_REQUIRE(anchors && out && variances, "MultiBoxDetection: NULL tensor");
  DSPMB_REQUIRE(B <= 65535 && C <= 65535, "MultiBoxDetection: batch / classes > 65535 not supported in one call");
  DSPMB_REQUIRE(A < (1 << 24), "MultiBoxDetection: more than 2^24 anchors are not supported");
  DSPMB_REQUIRE(((uintptr_t)anchors & 15) == 0, "MultiBoxDetection: anchors must be 16-byte aligned");
  if (B == 0) return DSPMB_OK;
  const int Th = heads ? heads->tile_off[heads->nscales] : 0;
  const size_t need = heads ? carve(nullptr, B, A, C, Th, (size_t)Th * kHeadTile).bytes : dspmb_detection_workspace_bytes(B, A, C);
  if (!workspace || workspace_bytes < need || ((uintptr_t)workspace & 255)) {
    set_error("MultiBoxDetection: workspace must be 256-byte aligned and >= %zu bytes (got %zu)", need, workspace_bytes);
    return DSPMB_ERR_WORKSPACE;
  }
  const int nms_debug = getenv("DSPMB_NMS_DEBUG") ? atoi(getenv("DSPMB_NMS_DEBUG")) : 0;

  struct {
    const void *p[6];
    int i[7];
    float f[6];
    HeadsArgs h;
  } key;
  memset(&key, 0, sizeof(key));
  if (heads) key.h = *heads;
  key.p[0] = cls_prob, key.p[1] = loc_pred, key.p[2] = anchors, key.p[3] = out, key.p[4] = valid_count_out, key.p[5] = workspace;
  key.i[0] = B, key.i[1] = A, key.i[2] = C, key.i[3] = clip, key.i[4] = force_suppress, key.i[5] = nms_topk, key.i[6] = nms_debug;
  key.f[0] = threshold, key.f[1] = nms_threshold;
  for (int k = 0; k < 4; ++k) key.f[2 + k] = variances[k];

  const float *const cls_prob_all = cls_prob, *const loc_pred_all = loc_pred;
  float *const out_all = out;
  int32_t *const valid_all = valid_count_out;
  const int B_all = B;
  return graph_cached_launch(&key, sizeof(key), (cudaStream_t)stream_, [&](const LaunchCtx &ctx) -> int {
  const bool vec4_all = (A % 4 == 0) && (((uintptr_t)cls_prob_all | (uintptr_t)out_all | (uintptr_t)loc_pred_all) & 15) == 0;
  const int variant_all = tuning(DSPMB_TUNE_DET_STREAM_VARIANT);
  const bool v2_all = tuning(DSPMB_TUNE_DET_PIPELINE) != 0 && vec4_all && (C == 21 || C == 9) && variant_all == 2 &&
                      !force_suppress && nms_threshold > 0.f && nms_threshold <= 1.f && C - 1 <= kV2ClsPad &&
                      ceil_div(A, 256) <= kV2MaxTiles;
  const int groups = (v2_all && !heads) ? split_groups(B_all, tuning(DSPMB_TUNE_DET_SPLIT)) : 1;
  size_t ws_off = 0;
  for (int grp = 0; grp < groups; ++grp) {
  const int b0 = (int)((long long)B_all * grp / groups), B = (int)((long long)B_all * (grp + 1) / groups) - b0;
  const float *cls_prob = heads ? nullptr : cls_prob_all + (size_t)b0 * C * A;
  const float *loc_pred = heads ? nullptr : loc_pred_all + (size_t)b0 * A * 5;
  float *out = out_all + (size_t)b0 * A * 7;
  int32_t *valid_count_out = valid_all ? valid_all + b0 : nullptr;
  DetWorkspace w = carve((char *)workspace + ws_off, B, A, C, Th, (size_t)Th * kHeadTile);
  ws_off += align_up(w.bytes, 256);
  // the stream kernels of all groups run back to back on the main branch; everything after a group's stream kernel
  // goes to a side branch (alternating between the two) when the call is being captured into the library's graph
  cudaStream_t stream = ctx.stream;
  const bool vec4 = (A % 4 == 0) && (((uintptr_t)cls_prob | (uintptr_t)out | (uintptr_t)loc_pred) & 15) == 0;
  // TMA-fed persistent kernel whenever the ring fits (2..4 stages); plain kernels otherwise
  const size_t stage_bytes = ((size_t)(C - 1) * kPipeTile + (size_t)kPipeTile * 5) * sizeof(float);
  int stages = 0, ctas_per_sm = 2;
  const int variant = tuning(DSPMB_TUNE_DET_STREAM_VARIANT);  // 0 generic, 1 TMA ring, else register-resident
  if (vec4 && C > 1 && variant == 1) {
    stages = (int)(100 * 1024 / stage_bytes);
    if (stages < 2) {
      ctas_per_sm = 1;
      stages = (int)(200 * 1024 / stage_bytes);
    }
    if (stages > 4) stages = 4;
    if (stages < 2) stages = 0;
  }
  const bool bulk_ok = vec4 && (C == 21 || C == 9) && ((uintptr_t)cls_prob & 15) == 0;
  const int bulk_threads = variant == 3 ? 256 : 128, bulk_vec = variant == 5 ? 4 : 2;
  const bool bulk_variant = bulk_ok && (variant == 2 || variant == 3 || variant == 5);
  const bool reg_variant = vec4 && !bulk_variant && variant > 2 && (C == 21 || C == 9);
  const int tile = heads ? kHeadTile
                         : (stages ? kPipeTile
                                   : (bulk_variant ? bulk_threads * bulk_vec
                                                   : (reg_variant ? kRegThreads * 4 : kStreamThreads * (vec4 ? 4 : 1))));
  const int T = heads ? Th : ceil_div(A, tile);
  const int Adef = ((A + 3) & ~3) + 4 * kStreamThreads;
  const int Apad = heads && Th * kHeadTile > Adef ? Th * kHeadTile : Adef;
  const bool nms_on = nms_threshold > 0.f && nms_threshold <= 1.f && (C > 1 || force_suppress);
  // fork/join pipeline: per-class segments, the default (or the head-fed) stream kernel, tile tables that fit in shared memory
  const bool v2 = tuning(DSPMB_TUNE_DET_PIPELINE) != 0 && (heads || (bulk_variant && variant == 2 && !stages)) &&
                  !force_suppress && nms_on && T <= kV2MaxTiles && C - 1 <= kV2ClsPad;

  StreamArgs sa;
  sa.cls_prob = cls_prob;
  sa.loc_pred = loc_pred;
  sa.anchors = anchors;
  sa.out = out;
  sa.tile_count = w.tile_count;
  sa.slot_rows = w.slot_rows;
  sa.slot_keys = w.slot_keys;
  sa.slot_cls = w.slot_cls;
  sa.slot_box = w.slot_box;
  sa.cbox = w.cbox;
  sa.crank = w.crank;
  sa.tile_cls = w.tile_cls;
  sa.A = A;
  sa.C = C;
  sa.T = T;
  sa.Apad = Apad;
  sa.cls_stride = (Apad + 7) & ~7;
  sa.threshold = threshold;
  sa.clip = clip;
  sa.vx = variances[0];
  sa.vy = variances[1];
  sa.vw = variances[2];
  sa.vh = variances[3];
  sa.fma_build = libm_fma_mode();
  sa.prefetch = tuning(DSPMB_TUNE_DET_PREFETCH);
  const int phases = tuning(DSPMB_TUNE_PHASES);
  if (!(phases & 1)) {
    // stream phase skipped (per-phase timing: the workspace still holds the previous call's records)
  } else if (heads) {
    dim3 gridh(T, B);
    ProfileScope _p(kSlotDetStream, stream);
#define DSPMB_LAUNCH_HEADS(CC, V2)                                                                    \
  do {                                                                                                \
    constexpr size_t kBytes = sizeof(HeadSmem<CC>);                                                   \
    DSPMB_ENSURE_DYN_SMEM((det_stream_heads_kernel<CC, V2>), kBytes);                                 \
    det_stream_heads_kernel<CC, V2><<<gridh, kHeadThreads, kBytes, stream>>>(sa, *heads);             \
  } while (0)
    if (C == 21 && v2) DSPMB_LAUNCH_HEADS(21, true);
    else if (C == 21) DSPMB_LAUNCH_HEADS(21, false);
    else if (v2) DSPMB_LAUNCH_HEADS(9, true);
    else DSPMB_LAUNCH_HEADS(9, false);
#undef DSPMB_LAUNCH_HEADS
  } else if (stages) {
    PipeArgs pa;
    pa.s = sa;
    pa.num_tiles = B * T;
    pa.stages = stages;
    pa.stage_floats = (int)(stage_bytes / sizeof(float));
    DSPMB_ENSURE_DYN_SMEM(det_stream_tma_kernel, 200 * 1024);
    int grid = kNumSMs * ctas_per_sm;
    if (grid > pa.num_tiles) grid = pa.num_tiles;
    ProfileScope _p(kSlotDetStream, stream);
    det_stream_tma_kernel<<<grid, kPipeThreads, stages * stage_bytes, stream>>>(pa);
  } else {
    dim3 grid1(T, B);
    ProfileScope _p(kSlotDetStream, stream);
    if (bulk_variant) {
#define DSPMB_LAUNCH_BULK(NFG, TH, VEC, V2)                                                                       \
  do {                                                                                                            \
    constexpr size_t kBytes = sizeof(BulkSmem<NFG, TH * VEC>);                                                    \
    DSPMB_ENSURE_DYN_SMEM((det_stream_bulk_kernel<NFG, TH, VEC, V2>), kBytes);                                    \
    det_stream_bulk_kernel<NFG, TH, VEC, V2><<<grid1, TH, kBytes, stream>>>(sa, tmap);                            \
  } while (0)
#define DSPMB_LAUNCH_BULK_LEAN(NFG)                                                                               \
  do {                                                                                                            \
    constexpr size_t kBytes = sizeof(BulkSmem<NFG, 256, true>);                                                   \
    DSPMB_ENSURE_DYN_SMEM((det_stream_bulk_kernel<NFG, 128, 2, true, true>), kBytes);                             \
    det_stream_bulk_kernel<NFG, 128, 2, true, true><<<grid1, 128, kBytes, stream>>>(sa, tmap);                    \
  } while (0)
#define DSPMB_LAUNCH_BULK_TENSOR(NFG)                                                                             \
  do {                                                                                                            \
    constexpr size_t kBytes = sizeof(BulkSmem<NFG, 256, true>);                                                   \
    DSPMB_ENSURE_DYN_SMEM((det_stream_bulk_kernel<NFG, 128, 2, true, true, true>), kBytes);                       \
    det_stream_bulk_kernel<NFG, 128, 2, true, true, true><<<grid1, 128, kBytes, stream>>>(sa, tmap);              \
  } while (0)
      const int lean = tuning(DSPMB_TUNE_DET_LEAN);
      // 2: class tile through ONE cp.async.bulk.tensor copy (needs the tensor map; rows of A*4 bytes must be 16-byte
      // multiples, which vec4 guarantees); 1: NFG 1-D bulk copies
      CUtensorMap tmap;
      memset(&tmap, 0, sizeof(tmap));
      bool tensor = false;
      if (v2 && lean == 2) tensor = make_cls_tensor_map(&tmap, cls_prob, B, C, A, C - 1) == DSPMB_OK;
      if (C == 21 && v2 && tensor) DSPMB_LAUNCH_BULK_TENSOR(20);
      else if (v2 && tensor) DSPMB_LAUNCH_BULK_TENSOR(8);
      else if (C == 21 && v2 && lean) DSPMB_LAUNCH_BULK_LEAN(20);
      else if (v2 && lean) DSPMB_LAUNCH_BULK_LEAN(8);
      else if (C == 21 && v2) DSPMB_LAUNCH_BULK(20, 128, 2, true);
      else if (v2) DSPMB_LAUNCH_BULK(8, 128, 2, true);
      else if (C == 21 && variant == 2) DSPMB_LAUNCH_BULK(20, 128, 2, false);
      else if (C == 21 && variant == 3) DSPMB_LAUNCH_BULK(20, 256, 2, false);
      else if (C == 21) DSPMB_LAUNCH_BULK(20, 128, 4, false);
      else if (variant == 2) DSPMB_LAUNCH_BULK(8, 128, 2, false);
      else if (variant == 3) DSPMB_LAUNCH_BULK(8, 256, 2, false);
      else DSPMB_LAUNCH_BULK(8, 128, 4, false);
#undef DSPMB_LAUNCH_BULK
#undef DSPMB_LAUNCH_BULK_LEAN
#undef DSPMB_LAUNCH_BULK_TENSOR
    }
    else if (reg_variant && C == 21)
      det_stream_reg_kernel<20, kRegThreads><<<grid1, kRegThreads, 0, stream>>>(sa);
    else if (reg_variant && C == 9)
      det_stream_reg_kernel<8, kRegThreads><<<grid1, kRegThreads, 0, stream>>>(sa);
    else if (vec4)
      det_stream_kernel<4><<<grid1, kStreamThreads, 0, stream>>>(sa);
    else
      det_stream_kernel<1><<<grid1, kStreamThreads, 0, stream>>>(sa);
  }
  if (phases & 1) ++ctx.launches;
  DSPMB_CUDA_TRY(cudaGetLastError());
  if (groups > 1 && ctx.forked()) {
    const int rc = ctx.fork_side(grp % LaunchCtx::kSides);
    if (rc != DSPMB_OK) return rc;
    stream = ctx.side[grp % LaunchCtx::kSides];
  }

  SortArgs so;
  so.out = out;
  so.tile_count = w.tile_count;
  so.slot_rows = w.slot_rows;
  so.slot_keys = w.slot_keys;
  so.slot_cls = w.slot_cls;
  so.slot_box = w.slot_box;
  so.keys = w.keys;
  so.tile_base = w.tile_base;
  so.valid = w.valid;
  so.nms_rows = w.nms_rows;
  so.cursor = w.cursor;
  so.row_cls = w.row_cls;
  so.row_box = w.row_box;
  so.sort_keys = w.sort_keys;
  const bool head_lists = v2 && nms_topk > 0 && nms_topk <= kHeadCap;
  so.head_rank = v2 ? w.head_rank : nullptr;
  so.head_list = head_lists ? w.head_list : nullptr;
  so.head_off = w.head_off;
  so.tile_cls = w.tile_cls;
  so.nfg = C - 1;
  so.mask_rows = tuning(DSPMB_TUNE_NMS_MASK_ROWS);
  so.tail_copy = v2 ? 1 : 0;
  so.valid_count_out = valid_count_out;
  so.header = w.header;
  so.T = T;
  so.tile = tile;
  so.rank_parts = v2 ? 0 : (ceil_div(T, 6) < 16 ? ceil_div(T, 6) : 16);  // v2: the pair kernel moves the tail rows
  so.debug = nms_debug;
  so.A = A;
  so.Apad = Apad;
  so.cls_stride = (Apad + 7) & ~7;
  so.npad_max = next_pow2(A);
  so.niter_max = ceil_div(A, kSortThreads);
  so.nms_threshold = nms_threshold;
  so.force_suppress = force_suppress;
  so.nms_topk = nms_topk;
  const bool keys_in_smem = A <= kKeySmemMax;
  // sort keys in shared memory: enough for the top-k head, or for everything when no top-k limit applies
  const int smem_keys = tuning(DSPMB_TUNE_SORT_SMEM_KEYS) < 2 ? 2 : tuning(DSPMB_TUNE_SORT_SMEM_KEYS);
  const int want0 = (nms_topk > 0 && nms_topk < A) ? next_pow2(nms_topk) : so.npad_max;
  const int want = want0 < kRegSortThreads ? kRegSortThreads : want0;  // the register sort pads to 128 keys
  so.sel_cap = want < smem_keys ? want : smem_keys;
  const size_t smem2 = sizeof(unsigned long long) * so.sel_cap + (keys_in_smem ? sizeof(unsigned) * (size_t)((A + 3) & ~3) : 0) +
                       sizeof(int) * 32 * (size_t)so.niter_max;
  DSPMB_REQUIRE(smem2 <= 200 * 1024, "MultiBoxDetection: too many anchors for the sort kernel (A=%d)", A);
  DSPMB_ENSURE_DYN_SMEM(det_sort_kernel<true>, 200 * 1024);  // + 21.3 KB static <= 227 KB per CTA
  DSPMB_ENSURE_DYN_SMEM(det_sort_kernel<false>, 200 * 1024);
  NmsArgs na;
  na.out = out;
  na.nms_rows = w.nms_rows;
  na.cursor = w.cursor;
  na.row_cls = w.row_cls;
  na.row_box = w.row_box;
  na.seg_list = w.seg_list;
  na.seg_box = w.seg_box;
  na.seg_dead = w.seg_dead;
  na.seg_area = w.seg_area;
  na.A = A;
  na.cls_stride = (Apad + 7) & ~7;
  na.C = C;
  na.nms_threshold = nms_threshold;
  na.force_suppress = force_suppress;
  na.mask_rows = tuning(DSPMB_TUNE_NMS_MASK_ROWS);
  na.smem_rows = tuning(DSPMB_TUNE_NMS_SMEM_ROWS);
  na.debug = nms_debug;
  if (phases & 2) {
    ProfileScope _p(kSlotDetSort, stream);
    cudaLaunchConfig_t scfg = {};
    scfg.gridDim = dim3(B, 1 + so.rank_parts);
    scfg.blockDim = dim3(kSortThreads);
    scfg.dynamicSmemBytes = smem2;
    scfg.stream = stream;
    cudaLaunchAttribute sattr[1];
    sattr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    // only behind a stream kernel launched on this very stream (one image group) that has a tail to hide behind
    sattr[0].val.programmaticStreamSerializationAllowed =
        (tuning(DSPMB_TUNE_DET_SORT_PDL) && groups == 1 && (phases & 1) && (long long)T * B > kNumSMs) ? 1 : 0;
    scfg.attrs = sattr;
    scfg.numAttrs = 1;
    if (keys_in_smem)
      DSPMB_CUDA_TRY(cudaLaunchKernelEx(&scfg, det_sort_kernel<true>, so));
    else
      DSPMB_CUDA_TRY(cudaLaunchKernelEx(&scfg, det_sort_kernel<false>, so));
    ++ctx.launches;
  }
  DSPMB_CUDA_TRY(cudaGetLastError());
  if (v2) {
    if (phases & 4) {
      PairArgs pa;
      pa.out = out;
      pa.tile_count = w.tile_count;
      pa.tile_cls = w.tile_cls;
      pa.slot_rows = w.slot_rows;
      pa.slot_cls = w.slot_cls;
      pa.slot_box = w.slot_box;
      pa.cbox = w.cbox;
      pa.crank = w.crank;
      pa.row_cls = w.row_cls;
      pa.row_box = w.row_box;
      pa.nms_rows = w.nms_rows;
      pa.head_rank = w.head_rank;
      pa.head_list = head_lists ? w.head_list : nullptr;
      pa.head_off = w.head_off;
      pa.A = A;
      pa.T = T;
      pa.tile = tile;
      pa.Apad = Apad;
      pa.cls_stride = (Apad + 7) & ~7;
      pa.nfg = C - 1;
      pa.nms_threshold = nms_threshold;
      pa.nms_topk = nms_topk;
      pa.mask_rows = na.mask_rows;
      // programmatic dependent of the sort kernel, same stream: starts when the sort CTAs are resident
      ProfileScope _p(kSlotDetPair, stream);
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(C - 1, B);
      cfg.blockDim = dim3(kNmsThreads);
      cfg.dynamicSmemBytes = 0;
      cfg.stream = stream;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      DSPMB_CUDA_TRY(cudaLaunchKernelEx(&cfg, det_pair_kernel, pa, na));
      ++ctx.launches;
    }
  } else if ((phases & 4) && nms_on) {
    ProfileScope _p(kSlotDetNms, stream);
    det_nms_kernel<<<dim3(force_suppress ? 1 : C - 1, B), kNmsThreads, 0, stream>>>(na);
    ++ctx.launches;
  }
  DSPMB_CUDA_TRY(cudaGetLastError());
  }  // groups
  if (groups > 1 && ctx.forked()) return ctx.join();
  return DSPMB_OK;
  });
}

// ====================================================================================================
// Fused compaction + all-gather over NVLink peer memory (the one exchange step of the path).
//
// Every rank owns a gather buffer (cudaMalloc'd, shared with its peers through CUDA IPC):
//   2 slots of  [world][B][K*7] float rows | [world][B] int counts | [world][B][SW] int stats | [world] u64 flags
//   then        [2][world] u64 acks | 2 int CTA counters | int error
// submit(step):  one CTA per local image compacts the surviving rows (id >= 0, row order, at most K, padded with -1)
//   into shared memory once and stores the block -- and the image's SW target statistics -- into section `rank` of
//   EVERY rank's buffer with 128-bit stores; the remote ones travel over NVLink/NVSwitch while other CTAs are still
//   compacting, no NCCL launch, no host round trip.  The last CTA to finish (system-scope fences around a counter)
//   stores the sequence number step + 1 into flag[rank] of every peer.
// Flow control: slot = step & 1 is reused every two steps.  A reader that has finished with generation g of a slot
//   (dspmb_detection_gather_ack) stores g + 1 into ack[slot][reader] of every WRITER; the gather kernel of
//   generation g + 2 does not touch a peer's slot before its own copy of all acks has reached g + 1.  Sequence
//   numbers (not cumulative counters) make both waits exact, so a rank that runs ahead can neither overwrite an
//   unread slot nor satisfy a wait with newer data.  Waits are bounded; a timeout latches the error word, which
//   dspmb_gather_error reads back.
namespace dspmb {
namespace {

constexpr int kMaxPeers = 16;
constexpr long long kGatherSpins = 20000000ll;  // x 200 ns: a few seconds, then the error word is set

struct GatherLayout {
  size_t rows, counts, stats, flags, slot_bytes, ack_off, done_off, err_off, total;
};
inline GatherLayout gather_layout(int B, int K, int SW, int world) {
  GatherLayout g;
  g.rows = align_up((size_t)world * B * K * 7 * sizeof(float), 256);
  g.counts = align_up((size_t)world * B * sizeof(int), 256);
  g.stats = align_up((size_t)world * B * SW * sizeof(int), 256);
  g.flags = align_up((size_t)world * sizeof(unsigned long long), 256);
  g.slot_bytes = g.rows + g.counts + g.stats + g.flags;
  g.ack_off = 2 * g.slot_bytes;
  g.done_off = g.ack_off + align_up((size_t)2 * world * sizeof(unsigned long long), 256);
  g.err_off = g.done_off + 256;
  g.total = g.err_off + 256;
  return g;
}

struct GatherArgs {
  const float *out;
  const int *valid;
  const int *stats;
  unsigned char *peer[kMaxPeers];
  int B, A, K, SW, rank, world, slot;
  unsigned long long seq;
  size_t slot_off, counts_off, stats_off, flags_off, ack_off, done_off, err_off;  // byte offsets inside a buffer
};

__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys_u64(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// Bounded wait until *p >= want; returns false (and latches the error word) on a timeout.
__device__ __forceinline__ bool spin_until_ge(const unsigned long long *p, unsigned long long want, int *err) {
  for (long long spins = 0; spins < kGatherSpins; ++spins) {
    if (ld_acquire_sys_u64(p) >= want) return true;
    __nanosleep(200);
  }
  atomicExch(err, 1);
  return false;
}

__global__ void __launch_bounds__(256) det_gather_kernel(const __grid_constant__ GatherArgs g) {
  extern __shared__ __align__(16) float stage[];  // K * 7 floats
  __shared__ int scan_smem[256 / 32 + 1];
  __shared__ int carry_smem, sm_last;
  const int b = blockIdx.x;
  unsigned char *local = g.peer[g.rank];
  // every reader has released the generation that used this slot two steps ago
  if (g.seq >= 3ull && (int)threadIdx.x < g.world)
    spin_until_ge(reinterpret_cast<const unsigned long long *>(local + g.ack_off) + (size_t)g.slot * g.world + threadIdx.x,
                  g.seq - 2ull, reinterpret_cast<int *>(local + g.err_off));
  if (threadIdx.x == 0) carry_smem = 0;
  __syncthreads();
  const int K = g.K;
  int n = 0;
  if (K > 0) {
    const float *src = g.out + (size_t)b * g.A * 7;
    const int V = g.valid ? min(g.valid[b], g.A) : g.A;
    for (int base = 0; base < V; base += blockDim.x) {
      const int r = base + threadIdx.x;
      const int keep = (r < V && src[(size_t)r * 7] >= 0.f) ? 1 : 0;
      int total;
      const int ex = block_scan_excl(keep, scan_smem, &total);
      const int carry = carry_smem;
      const int pos = carry + ex;
      if (keep && pos < K) {
#pragma unroll
        for (int c = 0; c < 7; ++c) stage[pos * 7 + c] = src[(size_t)r * 7 + c];
      }
      __syncthreads();
      if (threadIdx.x == 0) carry_smem = carry + total;
      __syncthreads();
      if (carry_smem >= K) break;
    }
    n = min(carry_smem, K);
    for (int q = n * 7 + threadIdx.x; q < K * 7; q += blockDim.x) stage[q] = -1.f;
    __syncthreads();
  }
  const size_t row_off = g.slot_off + ((size_t)g.rank * g.B + b) * K * 7 * sizeof(float);
  const size_t cnt_off = g.counts_off + ((size_t)g.rank * g.B + b) * sizeof(int);
  const size_t st_off = g.stats_off + ((size_t)g.rank * g.B + b) * g.SW * sizeof(int);
  const int nvec = (K * 7) >> 2;  // K * 7 * 4 bytes is a multiple of 16 when K % 4 == 0 (checked on the host)
  for (int p = 0; p < g.world; ++p) {
    const int peer = (g.rank + p) % g.world;  // start with the local copy, spread the remote targets
    float4 *dst = reinterpret_cast<float4 *>(g.peer[peer] + row_off);
    for (int q = threadIdx.x; q < nvec; q += blockDim.x) dst[q] = reinterpret_cast<const float4 *>(stage)[q];
    if (threadIdx.x == 0 && K > 0) *reinterpret_cast<int *>(g.peer[peer] + cnt_off) = n;
    if (g.stats && (int)threadIdx.x < g.SW)
      reinterpret_cast<int *>(g.peer[peer] + st_off)[threadIdx.x] = g.stats[(size_t)b * g.SW + threadIdx.x];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {  // the last CTA of this launch publishes the sequence number (threadFenceReduction pattern)
    int *done = reinterpret_cast<int *>(local + g.done_off) + g.slot;
    const int prev = atomicAdd(done, 1);
    sm_last = prev == (int)gridDim.x - 1;
    if (sm_last) *done = 0;
  }
  __syncthreads();
  if (sm_last && (int)threadIdx.x < g.world) {
    __threadfence_system();
    st_release_sys_u64(reinterpret_cast<unsigned long long *>(g.peer[threadIdx.x] + g.flags_off) + g.rank, g.seq);
  }
}

// Returns once flag[r] >= seq for every writer r of this rank's slot (bounded).
__global__ void det_gather_wait_kernel(const unsigned long long *flags, int world, unsigned long long seq, int *err) {
  if ((int)threadIdx.x < world) spin_until_ge(flags + threadIdx.x, seq, err);
}

// Tells every writer that this rank has finished reading generation `seq` of the slot.
__global__ void det_gather_ack_kernel(const __grid_constant__ GatherArgs g) {
  if ((int)threadIdx.x < g.world)
    st_release_sys_u64(reinterpret_cast<unsigned long long *>(g.peer[threadIdx.x] + g.ack_off) + (size_t)g.slot * g.world + g.rank,
                       g.seq);
}

int fill_gather_args(GatherArgs &g, int B, int A, int K, int SW, int rank, int world, void *const *peer_bases, int slot,
                     long long seq) {
  DSPMB_REQUIRE(B > 0 && K >= 0 && (K % 4) == 0 && SW >= 0 && SW <= 32,
                "detection_gather: need B > 0, K a multiple of 4, 0 <= stats_width <= 32");
  DSPMB_REQUIRE(world >= 1 && world <= kMaxPeers && rank >= 0 && rank < world, "detection_gather: bad rank/world");
  DSPMB_REQUIRE(peer_bases && (slot == 0 || slot == 1) && seq >= 1, "detection_gather: bad argument");
  for (int p = 0; p < world; ++p) {
    DSPMB_REQUIRE(peer_bases[p] != nullptr, "detection_gather: peer %d has no buffer", p);
    g.peer[p] = (unsigned char *)peer_bases[p];
  }
  const GatherLayout l = gather_layout(B, K, SW, world);
  g.B = B;
  g.A = A;
  g.K = K;
  g.SW = SW;
  g.rank = rank;
  g.world = world;
  g.slot = slot;
  g.seq = (unsigned long long)seq;
  g.slot_off = (size_t)slot * l.slot_bytes;
  g.counts_off = g.slot_off + l.rows;
  g.stats_off = g.counts_off + l.counts;
  g.flags_off = g.stats_off + l.stats;
  g.ack_off = l.ack_off;
  g.done_off = l.done_off;
  g.err_off = l.err_off;
  return DSPMB_OK;
}

}  // namespace
}  // namespace dspmb

extern "C" size_t dspmb_gather_buffer_bytes(int B, int K, int stats_width, int world) {
  return gather_layout(B, K, stats_width, world).total;
}

extern "C" int dspmb_p2p_alloc(size_t bytes, void **dev_ptr, unsigned char *handle_out) {
  DSPMB_REQUIRE(dev_ptr && handle_out && bytes > 0, "p2p_alloc: bad argument");
  DSPMB_CUDA_TRY(cudaMalloc(dev_ptr, bytes));
  DSPMB_CUDA_TRY(cudaMemset(*dev_ptr, 0, bytes));
  cudaIpcMemHandle_t h;
  DSPMB_CUDA_TRY(cudaIpcGetMemHandle(&h, *dev_ptr));
  memcpy(handle_out, &h, sizeof(h));
  return DSPMB_OK;
}

extern "C" int dspmb_p2p_open(const unsigned char *handle, void **dev_ptr) {
  DSPMB_REQUIRE(handle && dev_ptr, "p2p_open: bad argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  DSPMB_CUDA_TRY(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return DSPMB_OK;
}

extern "C" int dspmb_p2p_close(void *dev_ptr) {
  DSPMB_CUDA_TRY(cudaIpcCloseMemHandle(dev_ptr));
  return DSPMB_OK;
}

extern "C" int dspmb_p2p_free(void *dev_ptr) {
  DSPMB_CUDA_TRY(cudaFree(dev_ptr));
  return DSPMB_OK;
}

extern "C" int dspmb_detection_gather_f32(const float *out, const int32_t *valid_count, const int32_t *stats, int B, int A,
                                          int K, int stats_width, int rank, int world, void *const *peer_bases, int slot,
                                          long long seq, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GatherArgs g;
  const int rc = fill_gather_args(g, B, A, K, stats_width, rank, world, peer_bases, slot, seq);
  if (rc != DSPMB_OK) return rc;
  DSPMB_REQUIRE(K == 0 || (out && A > 0), "detection_gather: rows requested without an output tensor");
  DSPMB_REQUIRE(stats_width == 0 || stats, "detection_gather: stats_width > 0 without a statistics tensor");
  DSPMB_REQUIRE((size_t)K * 7 * sizeof(float) <= 160 * 1024, "detection_gather: K too large");
  g.out = out;
  g.valid = valid_count;
  g.stats = stats_width ? stats : nullptr;
  const size_t smem = (size_t)K * 7 * sizeof(float);
  DSPMB_ENSURE_DYN_SMEM(det_gather_kernel, 160 * 1024);
  {
    ProfileScope _p(kSlotDetCompact, stream);
    det_gather_kernel<<<B, 256, smem, stream>>>(g);
  }
  DSPMB_CUDA_TRY(cudaGetLastError());
  return DSPMB_OK;
}

extern "C" int dspmb_detection_gather_wait(void *local_base, int B, int K, int stats_width, int world, int slot,
                                           long long seq, void *stream_) {
  DSPMB_REQUIRE(local_base && (slot == 0 || slot == 1) && seq >= 1 && world >= 1 && world <= kMaxPeers,
                "detection_gather_wait: bad argument");
  const GatherLayout l = gather_layout(B, K, stats_width, world);
  unsigned char *base = (unsigned char *)local_base;
  const unsigned long long *flags =
      (const unsigned long long *)(base + (size_t)slot * l.slot_bytes + l.rows + l.counts + l.stats);
  det_gather_wait_kernel<<<1, 32, 0, (cudaStream_t)stream_>>>(flags, world, (unsigned long long)seq, (int *)(base + l.err_off));
  DSPMB_CUDA_TRY(cudaGetLastError());
  return DSPMB_OK;
}

extern "C" int dspmb_detection_gather_ack(int B, int K, int stats_width, int rank, int world, void *const *peer_bases,
                                          int slot, long long seq, void *stream_) {
  GatherArgs g;
  const int rc = fill_gather_args(g, B, 1, K, stats_width, rank, world, peer_bases, slot, seq);
  if (rc != DSPMB_OK) return rc;
  g.out = nullptr;
  g.valid = nullptr;
  g.stats = nullptr;
  det_gather_ack_kernel<<<1, 32, 0, (cudaStream_t)stream_>>>(g);
  DSPMB_CUDA_TRY(cudaGetLastError());
  return DSPMB_OK;
}

// ---- the whole per-step exchange of a rank in ONE host call ----
// submit() used to be ~10 Python-level calls per step (two event records, two stream waits, up to three kernel launches,
// each through ctypes / torch): at 50 us per step that host work was as long as the step itself.  The context owns the
// side stream and the events; dspmb_gather_submit enqueues, in order: compute stream waits for the gather kernel that
// read `out` two steps ago; (side) bounded wait + acknowledgement of the generation the slot still holds when nobody
// read it; side stream waits for the compute stream's work so far; gather kernel; `done` event.
struct GatherCtx {
  cudaStream_t side;
  cudaEvent_t ready[2], done[2];
};

extern "C" void *dspmb_gather_ctx_create(void) {
  GatherCtx *c = new GatherCtx();
  if (cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking) != cudaSuccess) {
    delete c;
    set_error("gather_ctx_create: cudaStreamCreateWithFlags failed");
    return nullptr;
  }
  for (int i = 0; i < 2; ++i) {
    cudaEventCreateWithFlags(&c->ready[i], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&c->done[i], cudaEventDisableTiming);
  }
  return c;
}

extern "C" void *dspmb_gather_ctx_side_stream(void *ctx) { return ctx ? (void *)((GatherCtx *)ctx)->side : nullptr; }

extern "C" int dspmb_gather_ctx_destroy(void *ctx) {
  if (!ctx) return DSPMB_OK;
  GatherCtx *c = (GatherCtx *)ctx;
  for (int i = 0; i < 2; ++i) {
    cudaEventDestroy(c->ready[i]);
    cudaEventDestroy(c->done[i]);
  }
  cudaStreamDestroy(c->side);
  delete c;
  return DSPMB_OK;
}

extern "C" int dspmb_gather_submit(void *ctx, const float *out, const int32_t *valid_count, const int32_t *stats, int B,
                                   int A, int K, int stats_width, int rank, int world, void *const *peer_bases, int slot,
                                   long long seq, long long prev_seq, int release_prev, void *compute_stream_) {
  DSPMB_REQUIRE(ctx && (slot == 0 || slot == 1), "gather_submit: bad argument");
  GatherCtx *c = (GatherCtx *)ctx;
  cudaStream_t compute = (cudaStream_t)compute_stream_;
  if (prev_seq > 0) DSPMB_CUDA_TRY(cudaStreamWaitEvent(compute, c->done[slot], 0));
  if (prev_seq > 0 && release_prev) {
    int rc = dspmb_detection_gather_wait(peer_bases[rank], B, K, stats_width, world, slot, prev_seq, c->side);
    if (rc != DSPMB_OK) return rc;
    rc = dspmb_detection_gather_ack(B, K, stats_width, rank, world, peer_bases, slot, prev_seq, c->side);
    if (rc != DSPMB_OK) return rc;
  }
  DSPMB_CUDA_TRY(cudaEventRecord(c->ready[slot], compute));
  DSPMB_CUDA_TRY(cudaStreamWaitEvent(c->side, c->ready[slot], 0));
  const int rc = dspmb_detection_gather_f32(out, valid_count, stats, B, A, K, stats_width, rank, world, peer_bases, slot, seq, c->side);
  if (rc != DSPMB_OK) return rc;
  DSPMB_CUDA_TRY(cudaEventRecord(c->done[slot], c->side));
  return DSPMB_OK;
}

extern "C" int dspmb_detection_gather_read(const void *local_base, int B, int K, int stats_width, int world, int slot,
                                           float *rows_out, int32_t *counts_out, int32_t *stats_out, void *stream_) {
  DSPMB_REQUIRE(local_base && (slot == 0 || slot == 1), "detection_gather_read: bad argument");
  const GatherLayout l = gather_layout(B, K, stats_width, world);
  const unsigned char *base = (const unsigned char *)local_base + (size_t)slot * l.slot_bytes;
  cudaStream_t stream = (cudaStream_t)stream_;
  if (rows_out && K > 0)
    DSPMB_CUDA_TRY(cudaMemcpyAsync(rows_out, base, (size_t)world * B * K * 7 * sizeof(float), cudaMemcpyDeviceToDevice, stream));
  if (counts_out && K > 0)
    DSPMB_CUDA_TRY(cudaMemcpyAsync(counts_out, base + l.rows, (size_t)world * B * sizeof(int), cudaMemcpyDeviceToDevice, stream));
  if (stats_out && stats_width > 0)
    DSPMB_CUDA_TRY(cudaMemcpyAsync(stats_out, base + l.rows + l.counts, (size_t)world * B * stats_width * sizeof(int),
                                   cudaMemcpyDeviceToDevice, stream));
  return DSPMB_OK;
}

extern "C" int dspmb_gather_error(const void *local_base, int B, int K, int stats_width, int world) {
  DSPMB_REQUIRE(local_base, "gather_error: bad argument");
  const GatherLayout l = gather_layout(B, K, stats_width, world);
  int err = 0;
  DSPMB_CUDA_TRY(cudaMemcpy(&err, (const unsigned char *)local_base + l.err_off, sizeof(int), cudaMemcpyDeviceToHost));
  if (err) set_error("detection_gather: a peer did not arrive / acknowledge within the spin bound");
  return err;
}
