// d2pc_device.cuh -- shared by the sm_100a translation units: per-frame state block, workspace
// layout, launch geometry and small warp/block primitives.
#ifndef D2PC_DEVICE_CUH_
#define D2PC_DEVICE_CUH_

#include <cuda_runtime.h>
#include <stdint.h>

#include "d2pc_math.h"

namespace d2pc {

// ------------------------------------------------------------------------------------------
// tunables
// ------------------------------------------------------------------------------------------
constexpr int kSampleSize = 8192;      // level-1 sample per frame (stratified), power of two
constexpr int kSample2Size = 4096;     // level-2 sample over a bracket's candidates
constexpr int kSortCap = 16384;        // frames with at most this many pixels skip sampling
constexpr int kSelBits = 12;           // bucket histogram levels of the exact selection: 4096 bins
constexpr int kScanThreads = 256;
constexpr int kScanPerThread = 32;
constexpr int kScanTile = kScanThreads * kScanPerThread;  // 4096 pixels per CTA
constexpr int kSelThreads = 1024;      // sample kernel
constexpr int kSelectThreads = 512;    // stand-alone select kernel
constexpr int kEmitThreads = 256;
constexpr int kEmitPerThread = 4;
constexpr int kEmitTile = kEmitThreads * kEmitPerThread;  // 1024 output points per CTA
constexpr int kFbTargets = 4;          // simultaneous ranks in the fallback radix select
constexpr uint32_t kCandDivisor = 16;  // candidate capacity per bracket = max(n/16, ...)

// ------------------------------------------------------------------------------------------
// geometry derived from D2pcConfig (host computes once, passed by value to kernels)
// ------------------------------------------------------------------------------------------
struct Geom {
  int32_t H, W, C, h, w, step;
  int32_t nu, nv;        // output grid: ceil(W/step) x ceil(H/step)
  uint32_t P;            // H*W   pixels of the (virtually resized) map = n of the percentiles
  uint32_t N;            // nu*nv output rows per frame (capacity of the output slot)
  uint32_t D;            // h*w   elements of one depth frame
  int32_t native;        // (h, w) == (H, W)
  double scale_x, scale_y;  // (double)w / W, (double)h / H   (bilinear coordinate scale)
};

inline Geom make_geom(const D2pcConfig &c) {
  Geom g;
  g.H = c.img_h; g.W = c.img_w; g.C = c.img_c; g.h = c.dep_h; g.w = c.dep_w; g.step = c.step;
  g.nu = (c.img_w + c.step - 1) / c.step;
  g.nv = (c.img_h + c.step - 1) / c.step;
  g.P = (uint32_t)c.img_h * (uint32_t)c.img_w;
  g.N = (uint32_t)g.nu * (uint32_t)g.nv;
  g.D = (uint32_t)c.dep_h * (uint32_t)c.dep_w;
  g.native = (c.dep_h == c.img_h && c.dep_w == c.img_w) ? 1 : 0;
  g.scale_x = (double)c.dep_w / (double)c.img_w;
  g.scale_y = (double)c.dep_h / (double)c.img_h;
  return g;
}

// ------------------------------------------------------------------------------------------
// per-frame state (lives in the caller's workspace)
// ------------------------------------------------------------------------------------------
struct __align__(256) FrameState {
  // level-1 brackets (inclusive key bounds) for the 2% / 98% order statistics
  uint32_t brL[2], brU[2];
  uint32_t sample_ok;     // 0: sample saw a non-finite value
  // streaming pass results (atomically accumulated by scan CTAs)
  uint32_t below[2];      // [0]: finite values < brL[0] (counted by the scan); [1]: derived by select
  uint32_t above1;        // finite values > brU[1] (counted by the scan)
  uint32_t eqL[2];        // finite keys == brL
  uint32_t inside[2];     // finite keys strictly inside (brL, brU); appended to the candidate list
  uint32_t eqU[2];        // finite keys == brU (brU != brL)
  uint32_t n_nonfinite, n_nan;
  // values the scan deferred to the two raw queues; reserved as ONE 64-bit atomic (low word queue 0, high word
  // queue 1): same-line atomics serialise in L2 at ~16 ns each, and a frame has hundreds of scan tiles
  alignas(8) uint32_t nqueue[2];
  uint32_t min_key, max_key;  // of the repaired map (fallback path only)
  // exact selection
  uint32_t sel_key[4];    // keys at ranks lo2, hi2, lo98, hi98
  uint32_t sel_fail;
  uint32_t sel_done;      // bracket CTAs finished (0..2)
  // fallback radix select
  uint32_t fb_prefix[kFbTargets];
  uint32_t fb_rank[kFbTargets];
  uint32_t fb_active;     // number of live targets in the current stage
  uint32_t fb_any_nan;    // repaired map still holds NaN (median itself NaN)
  // emit: depth-space mask of the frame (ax-1): mode 0 = exact z compare, 1 = [mask_lo, mask_hi] on raw depth
  int32_t mask_mode;
  float mask_lo, mask_hi;
  uint32_t emit_count;
  uint32_t bounds_min[3], bounds_max[3];  // ordered keys of kept x, y, z
  // What the emit consumes, in ONE 128-byte line: the persistent path kernel takes a snapshot of it with a single
  // warp-wide load (status is written last, after a fence: a snapshot that shows READY shows the final norm).
  alignas(128) NormParams norm;
  int32_t status;         // D2PC_FRAME_*
};
static_assert(sizeof(NormParams) + 4 <= 128, "norm + status must share one 128-byte line");
constexpr int kStatusWord = (int)(sizeof(NormParams) / 4);  // index of `status` in the line's 32 words

// Per-frame scratch of the cooperative selection inside the persistent path kernel (several CTAs per bracket):
// global bucket histograms, classification counts, the located buckets and their members.  Zeroed per step by
// the sample kernel.
constexpr uint32_t kSelBins = 4096, kSelListCap = 1024;
struct __align__(256) SelShared {
  uint32_t counts[2][8];     // below, eqL, inside, eqU, non-finite, NaN
  uint32_t scan_done;        // scan tiles finished (the selection waits for all of them); on this line, not on
                             // the FrameState line the scan's reservations hammer
  uint32_t a_done[2], c_done[2];
  uint32_t bin_ready[2];     // 0 pending, 1 members wanted, 2 bracket finished without a collect
  uint32_t bin[2][2], base[2][2], need[2][2], want[2][2], shared_bin[2], key[2][2];
  uint32_t mcount[2][2], mmin[2][2], mmax[2][2];   // mmin holds ~min (zero-initialised scratch)
  uint32_t pad_[7];           // header = 64 words: members / hist stay 16-byte aligned
  uint32_t members[2][2][kSelListCap];
  uint32_t hist[2][kSelBins];
};

// One bilinear tap of the IPP model (d2pc_math.h axis_tap), 8 bytes, precomputed per call for
// every destination column / row so the per-pixel kernels do no float64 coordinate math.
struct __align__(8) TapEntry {
  int32_t i0c;  // source index of tap 0; bit 31 set = clamped (output copies tap 0)
  float t;      // float32 weight of tap 1
};

// scheduler words of the persistent path kernel: kSchedQueues work counters, one per 128-byte line, then the abort
// flag on its own line; the per-frame trace follows
constexpr int kSchedQueues = 32;
constexpr size_t kSchedBytes = 128 * (kSchedQueues + 1);

struct WsLayout {
  size_t state_off, cand_off, tile_off, fbhist_off, tap_off, sched_off, sel_off, resized_off, total;
  uint32_t cand_cap;      // keys per (frame, bracket)
  uint32_t emit_tiles;    // emit CTAs per frame
};

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

inline WsLayout make_layout(const D2pcConfig &c) {
  Geom g = make_geom(c);
  WsLayout L;
  uint32_t cap = g.P / kCandDivisor + 1;
  uint32_t small = g.P < (uint32_t)kSortCap ? g.P : (uint32_t)kSortCap;
  if (cap < small) cap = small;
  cap = (cap + 3u) & ~3u;
  L.cand_cap = cap;
  L.emit_tiles = (g.N + kEmitTile - 1) / kEmitTile;
  size_t off = 0;
  L.state_off = off;  off = align_up(off + sizeof(FrameState) * (size_t)c.batch, 256);
  L.cand_off = off;   off = align_up(off + (size_t)c.batch * 2 * cap * sizeof(uint32_t), 256);
  L.tile_off = off;   off = align_up(off + (size_t)c.batch * L.emit_tiles * sizeof(unsigned long long), 256);
  L.fbhist_off = off; off = align_up(off + (size_t)c.batch * kFbTargets * 256 * sizeof(uint32_t), 256);
  L.tap_off = off;    off = align_up(off + ((size_t)c.img_w + (size_t)c.img_h) * sizeof(TapEntry), 256);
  // work counter and abort flag of the persistent path kernel, then 8 timestamps per frame (its trace)
  L.sched_off = off;  off = align_up(off + kSchedBytes + (size_t)c.batch * 64, 256);
  L.sel_off = off;    off = align_up(off + (size_t)c.batch * sizeof(SelShared), 256);
  // resized depth only: the scan materialises the (H x W) map once; every later kernel reads it
  L.resized_off = off;
  if (!g.native) off = align_up(off + (size_t)c.batch * g.P * sizeof(float), 256);
  L.total = off;
  return L;
}

// everything a kernel needs to find its frame's data
struct KParams {
  Geom g;
  int32_t batch;
  const float *depth;       // [batch, h, w]
  FrameState *state;        // [batch]
  uint32_t *cand;           // [batch][2][cand_cap] per-bracket raw queues of deferred values (float bits)
  unsigned long long *tile_state;  // [batch][emit_tiles]
  uint32_t *fb_hist;        // [batch][kFbTargets][256]
  const TapEntry *xtab, *ytab;  // [W], [H] bilinear taps (resized depth only)
  SelShared *sel;           // [batch] cooperative-selection scratch (persistent path kernel)
  uint32_t *sched;          // scheduler words of the persistent path kernel (kSchedBytes), then its trace
  float *resized;           // [batch][P] materialised resized map (resized depth only), else nullptr
  uint32_t cand_cap, emit_tiles;
  int32_t force_fallback;
  int32_t hints;            // kHint* bits: L2 eviction priorities of the sub-batch pipeline (0 = none)
};

// L2 eviction-priority hints (d2pc_path.cu): keep what the emit will read again, let everything that is
// touched once leave first
constexpr int kHintScanKeep = 1;       // scan: depth loads evict-last (the emit reads the map again)
constexpr int kHintEmitDepthFirst = 2; // emit: depth loads evict-first (last use) -- native depth only
constexpr int kHintStreamFirst = 4;    // emit: colour loads and row stores evict-first
constexpr int kHintResizedKeep = 8;    // scan: stores of the materialised resized map evict-last
constexpr int kHintPipeline = 15;

inline KParams make_kparams(const D2pcConfig &c, const float *d_depth, void *ws) {
  KParams k;
  WsLayout L = make_layout(c);
  char *base = (char *)ws;
  k.g = make_geom(c);
  k.batch = c.batch;
  k.depth = d_depth;
  k.state = (FrameState *)(base + L.state_off);
  k.cand = (uint32_t *)(base + L.cand_off);
  k.tile_state = (unsigned long long *)(base + L.tile_off);
  k.fb_hist = (uint32_t *)(base + L.fbhist_off);
  k.xtab = (const TapEntry *)(base + L.tap_off);
  k.ytab = k.xtab + c.img_w;
  k.sched = (uint32_t *)(base + L.sched_off);
  k.sel = (SelShared *)(base + L.sel_off);
  k.resized = k.g.native ? nullptr : (float *)(base + L.resized_off);
  k.cand_cap = L.cand_cap;
  k.emit_tiles = L.emit_tiles;
  k.force_fallback = c.force_fallback;
  k.hints = 0;
  return k;
}

// View of the per-pixel (H x W) depth map for the kernels that run after the scan: the input itself
// when it already has the image size, otherwise the map the scan materialised.
inline KParams per_pixel_view(const KParams &kp) {
  KParams v = kp;
  if (!kp.g.native) {
    v.depth = kp.resized;
    v.g.D = kp.g.P;
    v.g.h = kp.g.H;
    v.g.w = kp.g.W;
    v.g.native = 1;
  }
  return v;
}

// frames [b0, b0 + nb) of a call's parameter block.  ring_slot >= 0: the slice's materialised resized maps
// live at frames [ring_slot, ring_slot + nb) of the resized area (sub-batch pipeline: the area is reused as a
// small ring so the maps stay in L2 between the scan that writes them and the emit that reads them).
inline KParams slice_kparams(const KParams &kp, int b0, int nb, int ring_slot) {
  KParams s = kp;
  s.batch = nb;
  s.depth = kp.depth + (size_t)b0 * kp.g.D;
  s.state = kp.state + b0;
  s.cand = kp.cand + (size_t)b0 * 2 * kp.cand_cap;
  s.tile_state = kp.tile_state + (size_t)b0 * kp.emit_tiles;
  s.fb_hist = kp.fb_hist + (size_t)b0 * kFbTargets * 256;
  s.sel = kp.sel + b0;
  if (kp.resized) s.resized = kp.resized + (size_t)(ring_slot >= 0 ? ring_slot : b0) * kp.g.P;
  return s;
}

struct EmitArgs {
  const uint8_t *bgr;
  float *xyz, *rgb;
  uint32_t *count;
  PixelConsts pc;
  int32_t use_z, drop_nf, want_bounds;
  float z_min, z_max;
  uint32_t frame0;   // output slot of the parameter block's frame 0 (sub-batch slices keep xyz / rgb unsliced)
};

int validate_config(const D2pcConfig *cfg);   // d2pc_api.cu
// d2pc_stats.cu
int check_workspace(const D2pcConfig *cfg, const void *ws, size_t ws_bytes);
int stats_prepare();
int taps_launch(const KParams &kp, cudaStream_t st);
constexpr int kStatsSample = 1, kStatsScan = 2, kStatsSelect = 4, kStatsScanSelect = kStatsScan | kStatsSelect;
int stats_launch(KParams kp, cudaStream_t st, int phases);
int status_launch(const KParams &kp, int32_t *d_status, int32_t *d_any, cudaStream_t st);
// d2pc_emit.cu
EmitArgs make_emit_args(const D2pcConfig &cfg, const uint8_t *d_bgr, float *d_xyz, float *d_rgb, uint32_t *d_count);
EmitArgs slice_emit_args(const EmitArgs &ea, const Geom &g, int b0);
int emit_init_launch(const KParams &kp, cudaStream_t st);
int bounds_export_launch(const KParams &kp, float *d_bounds, cudaStream_t st);
int emit_validate(const D2pcConfig *cfg, const float *d_depth, const uint8_t *d_bgr, void *d_workspace,
                  size_t workspace_bytes, float *d_xyz, float *d_rgb, uint32_t *d_count, float *d_bounds);
int emit_launch(const D2pcConfig &cfg, const KParams &kp, const EmitArgs &ea, float *d_bounds, cudaStream_t st,
                int32_t smooth_k, const double *h_kernel, void *d_scratch);
int record_cuda_error(cudaError_t e);          // d2pc_api.cu: returns D2PC_ERR_CUDA, stores text

#define D2PC_CHECK_LAUNCH()                                  \
  do {                                                       \
    cudaError_t e__ = cudaGetLastError();                    \
    if (e__ != cudaSuccess) return record_cuda_error(e__);   \
  } while (0)

#if defined(__CUDACC__)
// ------------------------------------------------------------------------------------------
// device primitives
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 ldg_stream_f4(const float *p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
// loads / stores carrying an explicit L2 eviction policy (createpolicy descriptor)
__device__ __forceinline__ uint64_t l2_policy(bool first, bool last) {
  uint64_t pf, pl, pn;
  asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pf));
  asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pl));
  asm("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pn));
  return first ? pf : (last ? pl : pn);
}
__device__ __forceinline__ float4 ldg_f4_pol(const float *p, uint64_t pol) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p), "l"(pol));
  return r;
}
__device__ __forceinline__ uint32_t ldg_u32_pol(const void *p, uint64_t pol) {
  uint32_t r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(r) : "l"(p), "l"(pol));
  return r;
}
__device__ __forceinline__ void stg_f4_pol(float *p, float4 v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;"
               :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
}
// shared -> global bulk copy (TMA unit) with an L2 policy; caller commits / waits the bulk group
__device__ __forceinline__ void bulk_store_pol(void *gdst, uint32_t smem_addr, uint32_t bytes, uint64_t pol) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;"
               :: "l"(gdst), "r"(smem_addr), "r"(bytes), "l"(pol) : "memory");
}
__device__ __forceinline__ uint32_t ldg_stream_u32(const void *p) {
  uint32_t r;
  asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_f1_pol(float *p, float v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.f32 [%0], %1, %2;" :: "l"(p), "f"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ void stg_stream_f4(float *p, float4 v) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void stg_stream_f1(float *p, float v) {
  asm volatile("st.global.cs.f32 [%0], %1;" :: "l"(p), "f"(v) : "memory");
}

__device__ __forceinline__ uint32_t warp_sum(uint32_t v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ uint32_t warp_min(uint32_t v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ uint32_t warp_max(uint32_t v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Bitonic sort of n (power of two) uint32 keys in shared memory by the whole CTA, ascending.
__device__ __forceinline__ void block_bitonic_sort(uint32_t *s, uint32_t n) {
  const uint32_t half = n >> 1;
  for (uint32_t k = 2; k <= n; k <<= 1) {
    for (uint32_t j = k >> 1; j > 0; j >>= 1) {
      for (uint32_t i = threadIdx.x; i < half; i += blockDim.x) {
        uint32_t a = ((i & ~(j - 1)) << 1) | (i & (j - 1));
        uint32_t b = a | j;
        uint32_t x = s[a], y = s[b];
        bool up = (a & k) == 0;
        if ((x > y) == up) { s[a] = y; s[b] = x; }
      }
      __syncthreads();
    }
  }
}

// ------------------------------------------------------------------------------------------
// Exact order statistics without sorting: multi-level bucket histogram over the key range.
// For every target t, out[t] = the rank[t]-th smallest key (0-based) among the keys the CTA
// visits through `for_each(f)` (f is called once per key owned by the calling thread; the
// functor must visit the same keys on every call).  Level 1 splits the common range [lo0, hi0]
// into 2^NBLOG equal power-of-two buckets and locates each target's bucket by a block scan;
// the next level does the same inside that bucket, until the bucket width is one key.  At most
// ceil(32 / NBLOG) passes; keys spread over the range give conflict-free shared atomics.
//   s_hist: T << NBLOG words;  s_res: 2*T + 40 words.  All threads must call; every thread
//   gets the results.  A rank outside the population returns ok = false.
// ------------------------------------------------------------------------------------------
template <int T, int NBLOG, typename ForEach>
__device__ __forceinline__ bool block_hist_select(ForEach for_each, uint32_t lo0, uint32_t hi0,
                                                  const uint32_t (&rank)[T], uint32_t (&out)[T],
                                                  uint32_t *s_hist, uint32_t *s_res) {
  constexpr uint32_t NB = 1u << NBLOG;
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int lane = tid & 31, warp = tid >> 5, nwarp = nthr >> 5;
  uint32_t lo[T], span[T], base[T];
  bool done[T];
#pragma unroll
  for (int t = 0; t < T; ++t) { lo[t] = lo0; span[t] = hi0 - lo0; base[t] = 0; done[t] = false; out[t] = 0; }
  bool shared_range = true, ok = true;
  uint32_t *s_warp = s_res + 2 * T;  // [<= 32] warp totals
  for (int level = 0; level < 6; ++level) {
    bool all_done = true;
#pragma unroll
    for (int t = 0; t < T; ++t) all_done = all_done && done[t];
    if (all_done) break;
    int shift[T];
#pragma unroll
    for (int t = 0; t < T; ++t) {
      const int bits = span[t] ? 32 - __clz(span[t]) : 0;
      shift[t] = bits > NBLOG ? bits - NBLOG : 0;
    }
    const int nh = shared_range ? 1 : T;
    for (uint32_t i = tid; i < (uint32_t)nh * NB; i += nthr) s_hist[i] = 0u;
    if (tid < 2 * T) s_res[tid] = 0xFFFFFFFFu;
    __syncthreads();
    if (shared_range) {
      for_each([&](uint32_t key) {
        const uint32_t d = key - lo[0];
        if (d <= span[0]) atomicAdd(&s_hist[d >> shift[0]], 1u);
      });
    } else {
      for_each([&](uint32_t key) {
#pragma unroll
        for (int t = 0; t < T; ++t) {
          const uint32_t d = key - lo[t];
          if (!done[t] && d <= span[t]) atomicAdd(&s_hist[t * NB + (d >> shift[t])], 1u);
        }
      });
    }
    __syncthreads();
    // locate each target's bucket: block-wide exclusive scan of every live histogram
    for (int h = 0; h < nh; ++h) {
      if (!shared_range && done[h]) continue;  // uniform
      const uint32_t *hist = s_hist + (size_t)h * NB;
      constexpr int PER = 4;  // consecutive bins per thread per chunk
      uint32_t carry = 0;
      for (uint32_t chunk = 0; chunk < NB; chunk += (uint32_t)nthr * PER) {
        const uint32_t b0 = chunk + (uint32_t)tid * PER;
        uint32_t c[PER], sum = 0;
#pragma unroll
        for (int i = 0; i < PER; ++i) { c[i] = (b0 + i < NB) ? hist[b0 + i] : 0u; sum += c[i]; }
        uint32_t incl = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          uint32_t y = __shfl_up_sync(0xffffffffu, incl, d);
          if (lane >= d) incl += y;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        uint32_t woff = 0, total = 0;
        for (int w = 0; w < nwarp; ++w) { const uint32_t x = s_warp[w]; if (w < warp) woff += x; total += x; }
        uint32_t pref = carry + woff + incl - sum;
#pragma unroll
        for (int i = 0; i < PER; ++i) {
#pragma unroll
          for (int t = 0; t < T; ++t) {
            const bool mine = shared_range ? !done[t] : (t == h);
            const uint32_t r = rank[t] - base[t];
            if (mine && c[i] && pref <= r && r < pref + c[i]) { s_res[2 * t] = b0 + i; s_res[2 * t + 1] = pref; }
          }
          pref += c[i];
        }
        carry += total;
        __syncthreads();
      }
    }
    __syncthreads();
#pragma unroll
    for (int t = 0; t < T; ++t) {
      if (done[t]) continue;
      const uint32_t bin = s_res[2 * t], pref = s_res[2 * t + 1];
      if (bin == 0xFFFFFFFFu) { ok = false; done[t] = true; continue; }
      base[t] += pref;
      lo[t] += bin << shift[t];
      span[t] = shift[t] ? ((1u << shift[t]) - 1u) : 0u;
      if (shift[t] == 0) { done[t] = true; out[t] = lo[t]; }
    }
    shared_range = false;
    __syncthreads();
  }
  return ok;
}

// ------------------------------------------------------------------------------------------
// Two-pass exact selection (the common case of block_hist_select at a fraction of its instruction
// count): one bucket histogram over [lo0, lo0 + span] (kFastBins buckets of 2^shift keys), a block
// scan that locates each wanted rank's bucket, one more visit of the keys that collects the
// members of those buckets into short shared lists, and a rank pick inside the list.  The pieces
// are separate so that a caller can fold its own per-key work into the first visit.
// ------------------------------------------------------------------------------------------
constexpr int kFastBinsLog = 12;
constexpr uint32_t kFastBins = 1u << kFastBinsLog;
constexpr uint32_t kFastListCap = 1024;   // members kept per wanted bucket

__device__ __forceinline__ int fast_shift(uint32_t span) {
  const int bits = span ? 32 - __clz(span) : 0;
  return bits > kFastBinsLog ? bits - kFastBinsLog : 0;
}

// Locate, for every wanted rank, its bucket in a kFastBins-bucket histogram and the number of keys
// in the buckets before it.  All threads call (blockDim.x * PER == kFastBins, PER a multiple of 4);
// every thread gets the results.  bin[t] = 0xFFFFFFFF: rank outside the population.
//   s_warp: 33 words, s_res: 2 * T words
template <int T, bool GLOBAL = false>
__device__ __forceinline__ void block_locate(const uint32_t *s_hist, const uint32_t (&rank)[T], const bool (&want)[T],
                                             uint32_t (&bin)[T], uint32_t (&base)[T], uint32_t *s_warp,
                                             uint32_t *s_res) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
  const int per = (int)(kFastBins / blockDim.x);  // 4 for 1024 threads
  if (tid < 2 * T) s_res[tid] = 0xFFFFFFFFu;
  const uint32_t b0 = (uint32_t)tid * (uint32_t)per;
  uint32_t sum = 0;
  for (int i = 0; i < per; i += 4) {
    const uint4 c = GLOBAL ? __ldcg(reinterpret_cast<const uint4 *>(s_hist + b0 + i))
                           : *reinterpret_cast<const uint4 *>(s_hist + b0 + i);
    sum += c.x + c.y + c.z + c.w;
  }
  uint32_t incl = sum;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t y = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += y;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    const uint32_t x = lane < nwarp ? s_warp[lane] : 0u;
    uint32_t in2 = x;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, in2, d);
      if (lane >= d) in2 += y;
    }
    __syncwarp();
    s_warp[lane] = in2 - x;  // exclusive prefix of the warp totals
  }
  __syncthreads();
  uint32_t pref = s_warp[warp] + incl - sum;
  bool mine = false;
#pragma unroll
  for (int t = 0; t < T; ++t) mine = mine || (want[t] && sum && pref <= rank[t] && rank[t] < pref + sum);
  if (mine) {
    for (int i = 0; i < per; ++i) {
      const uint32_t c = GLOBAL ? __ldcg(s_hist + b0 + i) : s_hist[b0 + i];
#pragma unroll
      for (int t = 0; t < T; ++t)
        if (want[t] && c && pref <= rank[t] && rank[t] < pref + c) { s_res[2 * t] = b0 + i; s_res[2 * t + 1] = pref; }
      pref += c;
    }
  }
  __syncthreads();
#pragma unroll
  for (int t = 0; t < T; ++t) { bin[t] = s_res[2 * t]; base[t] = s_res[2 * t + 1]; }
}

// r-th smallest (0-based) of the n keys in a shared list; all threads call, result
// broadcast through *s_out (caller synchronises before reading).
__device__ __forceinline__ void block_pick(const uint32_t *list, uint32_t n, uint32_t r, uint32_t *s_out) {
  for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
    const uint32_t k = list[i];
    uint32_t smaller = 0, equal = 0;
    for (uint32_t j = 0; j < n; ++j) {
      const uint32_t x = list[j];
      smaller += x < k ? 1u : 0u;
      equal += x == k ? 1u : 0u;
    }
    if (smaller <= r && r < smaller + equal) *s_out = k;
  }
}

__device__ __forceinline__ uint32_t hash_u32(uint32_t x) {  // lowbias32
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

// a1 from the tap tables: same arithmetic as d2pc_math.h bilinear_sample (horizontal lerp on two
// source rows, then vertical; clamped taps copy; corner blocks turn +-inf into NaN).
__device__ __forceinline__ float bilinear_taps(const float *src, int32_t src_w, TapEntry tx, TapEntry ty) {
  const int32_t x0 = tx.i0c & 0x7FFFFFFF, y0 = ty.i0c & 0x7FFFFFFF;
  const bool cx = tx.i0c < 0, cy = ty.i0c < 0;
  const float *row0 = src + (size_t)y0 * src_w;
  const float a0 = __ldg(row0 + x0);
  float r0 = a0;
  if (!cx) r0 = fmaf(__ldg(row0 + x0 + 1) - a0, tx.t, a0);
  if (cy) return (cx && is_inf_f32(r0)) ? nan_f32() : r0;
  const float *row1 = row0 + src_w;
  const float a1 = __ldg(row1 + x0);
  float r1 = a1;
  if (!cx) r1 = fmaf(__ldg(row1 + x0 + 1) - a1, tx.t, a1);
  return fmaf(r1 - r0, ty.t, r0);
}

// Raw value of pixel p (row-major index into the H x W grid) of one frame's (virtually resized)
// depth map.  NATIVE: the depth map already has the image size.
template <bool NATIVE>
__device__ __forceinline__ float depth_at(const float *frame, const KParams &kp, uint32_t p) {
  if (NATIVE) {
    return __ldg(frame + p);
  } else {
    const uint32_t v = p / (uint32_t)kp.g.W;
    const uint32_t u = p - v * (uint32_t)kp.g.W;
    return bilinear_taps(frame, kp.g.w, kp.xtab[u], kp.ytab[v]);
  }
}
#endif  // __CUDACC__

}  // namespace d2pc
#endif  // D2PC_DEVICE_CUH_
