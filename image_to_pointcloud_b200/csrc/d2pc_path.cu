// d2pc_path.cu -- the whole hot path (statistics + emission) of a batch as a software pipeline over
// sub-batches of a few frames, so that every depth map is read from HBM once.
//
// A frame's normalisation needs its exact percentiles before the first point can be emitted, so the depth
// map is visited twice: by the scan (statistics) and by the emit.  Run batch-wide (d2pc_stats_enqueue then
// d2pc_emit_enqueue) the second visit comes from HBM again, and a resized depth map makes one more round
// trip.  Here the batch is cut into sub-batches of `sub_batch` frames whose depth maps (or materialised
// resized maps) fit in the 126 MB L2 together with the next sub-batch's:
//
//   auxiliary stream (high priority):  stats(0) stats(1)      stats(2)      stats(3) ...
//   caller's stream:                            emit(0)       emit(1)       emit(2)  ...
//                                                 ^ stats(k+1) runs while emit(k) streams its output;
//                                                   stats(k+1) may start once emit(k-1) has finished
//
// The scan loads depth with an L2 evict-last policy, the emit reads it back (L2 hit) with evict-first and
// sends colours in and rows out with evict-first hints, so the streaming traffic does not push the waiting
// depth maps out.  The latency-bound sample / select launches of sub-batch k+1 hide under the
// bandwidth-bound emit of sub-batch k.  A resized depth map is materialised into a ring of
// (lookahead + 1) sub-batch slots at the start of the workspace's resized area: it is rewritten while still
// dirty in L2 and never reaches HBM.
//
// The fork / join is expressed with events on two streams; with D2PC_PATH_GRAPH the whole step is captured
// once into a CUDA graph (on streams the handle owns -- the caller's stream may be the legacy default
// stream, which cannot capture) and replayed with one launch while the arguments repeat.
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "d2pc_emit_dev.cuh"
#include "d2pc_stats_dev.cuh"

using namespace d2pc;

// ------------------------------------------------------------------------------------------
// The persistent path kernel: one launch per step.  Every CTA claims work items from a global
// counter, in an order in which an item only ever depends on items with a smaller index:
//
//   group g (g = 0 .. F + D - 1):   select(frame g-1, bracket 0), select(frame g-1, bracket 1),
//                                   then scan tiles of frame g interleaved 1 : R with emit tiles of frame g-D
//
// so the scan of frame g (depth HBM -> L2, evict-last) runs D frames ahead of the emit that reads the map
// again (L2 hit), and a frame's exact selection has D - 1 frame times to finish.  Waiting is a spin of one
// thread on a flag in the frame's state block; the claim order makes it deadlock-free as long as all CTAs
// are resident (grid = SMs x occupancy).  A spin that times out raises the abort flag: every CTA stops
// claiming, the unfinished frames stay PENDING and the host re-runs them through the two-phase calls.
// ------------------------------------------------------------------------------------------
namespace d2pc {

__global__ void sched_reset_kernel(KParams kp) {
  if (threadIdx.x <= kSchedQueues) kp.sched[32 * threadIdx.x] = 0u;
}
int sched_reset_launch(const KParams &kp, cudaStream_t st) {
  sched_reset_kernel<<<1, 64, 0, st>>>(kp);
  D2PC_CHECK_LAUNCH();
  return D2PC_OK;
}

struct SchedParams {
  int32_t F, D;
  uint32_t ns, ne, R;        // scan / emit tiles per frame, emit tiles per scan tile in the interleave
  uint32_t K;                // CTAs that share one bracket's selection
  uint32_t skip;             // measurement aid: 1 skip selections, 2 skip scan tiles, 4 skip emit tiles
  uint32_t nq;               // work queues (item i belongs to queue i % nq; a CTA serves queue blockIdx.x % nq)
  uint32_t group_items, total_items;
  int32_t vec_ok;
};

constexpr int kPathMinBlocks = 6;
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// per-frame trace (8 x u64 after the scheduler words): 0 first scan tile, 1 last scan tile done, 2/3 selection of
// bracket 0 start/end, 4/5 bracket 1, 6 first emit tile start, 7 last emit tile end
__device__ __forceinline__ unsigned long long *trace_slot(const KParams &kp, uint32_t f, int k) {
  return reinterpret_cast<unsigned long long *>(reinterpret_cast<unsigned char *>(kp.sched) + kSchedBytes) + (size_t)f * 8 + k;
}

constexpr int kPathScanPT = 16;   // pixels per thread of a scan tile inside the path kernel (4096-pixel tiles: 16 KB)

// One work item per CTA; the item is the CTA's index.  CTAs are dispatched in index order, so when a CTA runs,
// everything it may have to wait for (smaller indices) is running or done; the peers of a cooperative selection
// (the next few indices) follow as slots free up.
template <bool BOUNDS, bool TRACE>
__global__ void __launch_bounds__(kEmitThreads, kPathMinBlocks) path_kernel(KParams kp, EmitArgs ea, FastArgs fa,
                                                                           SchedParams sp) {
  extern __shared__ __align__(16) unsigned char s_dyn[];
  __shared__ EmitSmall s_emit;
  __shared__ SelPartSmall s_sel;
  __shared__ uint32_t s_go;
  __shared__ NormParams s_norm;
  const int tid = threadIdx.x;
  uint32_t *abort_flag = kp.sched + 32u * kSchedQueues;
  const uint32_t item = blockIdx.x;
  const uint32_t g = item / sp.group_items, i = item - g * sp.group_items;
  if (i < 2u * sp.K) {  // ---- part k of the exact selection of bracket br of frame g - 1
    if (g < 1u || g > (uint32_t)sp.F || (sp.skip & 1u)) return;
    const int f = (int)g - 1;
    const uint32_t br = i / sp.K, k = i - br * sp.K;
    if (tid == 0) s_go = spin_until(&kp.sel[f].scan_done, abort_flag, [&](uint32_t v) { return v >= sp.ns; }) ? 1u : 0u;
    __syncthreads();
    if (!s_go) return;
    if (TRACE && tid == 0 && k == 0u) *trace_slot(kp, f, 2 + 2 * (int)br) = global_ns();
    select_part(kp, f, (int)br, k, sp.K, reinterpret_cast<float *>(s_dyn), kSliceCap, s_sel,
                [&](const uint32_t *p, auto pred) { return spin_until(p, abort_flag, pred); });
    if (TRACE) {
      __syncthreads();
      if (tid == 0) atomicMax(trace_slot(kp, f, 3 + 2 * (int)br), global_ns());
    }
    return;
  }
  const uint32_t j = i - 2u * sp.K, q = j / (sp.R + 1u), r = j - q * (sp.R + 1u);
  if (r == 0u) {  // ---- scan tile q of frame g
    if (g >= (uint32_t)sp.F || (sp.skip & 2u)) return;
    if (TRACE && tid == 0) atomicMin(trace_slot(kp, g, 0), global_ns());
    scan_tile<kPathScanPT>(kp, (int)g, q, sp.vec_ok, *reinterpret_cast<ScanTileSmemT<kPathScanPT> *>(s_dyn));
    __syncthreads();
    if (tid == 0) {
      __threadfence();
      atomicAdd(&kp.sel[g].scan_done, 1u);
      if (TRACE) atomicMax(trace_slot(kp, g, 1), global_ns());
    }
    return;
  }
  // ---- emit tile e of frame g - D
  const uint32_t e = q * sp.R + (r - 1u);
  if (g < (uint32_t)sp.D || e >= sp.ne || (sp.skip & 4u)) return;
  const uint32_t f = g - (uint32_t)sp.D;
  if (f >= (uint32_t)sp.F) return;
  if (TRACE && tid == 0) atomicMin(trace_slot(kp, f, 6), global_ns());
  // the tile's loads are in flight when the frame's line (norm | status) is looked at
  auto norm_fn = [&]() -> const NormParams * {
    const uint32_t *line = reinterpret_cast<const uint32_t *>(&kp.state[f].norm);  // 32 words: norm | status
    uint32_t go = 0u;
    for (int attempt = 0; attempt < 2 && go == 0u; ++attempt) {
      if (tid < 32) {  // one warp-wide load = one snapshot of the line
        const uint32_t w = ld_relaxed_u32(line + tid);
        const uint32_t st = __shfl_sync(0xffffffffu, w, kStatusWord);
        if (st == (uint32_t)D2PC_FRAME_READY && tid < kStatusWord) reinterpret_cast<uint32_t *>(&s_norm)[tid] = w;
        if (tid == 0) s_go = st == (uint32_t)D2PC_FRAME_READY ? 1u : (st == (uint32_t)D2PC_FRAME_PENDING ? 0u : 3u);
      }
      __syncthreads();
      go = s_go;
      if (go == 0u) {  // still pending: thread 0 waits for the status word, then the line is read again
        __syncthreads();
        if (tid == 0) s_go = spin_until(line + kStatusWord, abort_flag, [&](uint32_t x) { return x != (uint32_t)D2PC_FRAME_PENDING; }) ? 0u : 2u;
        __syncthreads();
        go = s_go;
        __syncthreads();
      }
    }
    return go == 1u ? &s_norm : nullptr;  // 3: the frame needs the exact fallback (the host re-runs it); 2 / 0: aborted
  };
  emit_fast_tile<1, false, BOUNDS>(kp, ea, fa, f, e, reinterpret_cast<float *>(s_dyn), s_emit, norm_fn);
  if (TRACE) {
    __syncthreads();
    if (tid == 0) atomicMax(trace_slot(kp, f, 7), global_ns());
  }
}

}  // namespace d2pc

constexpr int kMaxAux = 8, kMaxEmit = 4;

struct D2pcPath {
  int device;
  cudaStream_t s_cap;              // origin of a graph capture
  cudaStream_t s_aux[kMaxAux];     // statistics (high priority)
  cudaStream_t s_emit[kMaxEmit];   // emits when more than one emit stream is used
  std::vector<cudaEvent_t> events;
  // cached graph
  cudaGraphExec_t exec;
  std::vector<unsigned char> key;
};

namespace {

struct PathArgs {
  D2pcConfig cfg;
  const float *d_depth;
  const uint8_t *d_bgr;
  void *d_workspace;
  size_t workspace_bytes;
  float *d_xyz, *d_rgb;
  uint32_t *d_count;
  float *d_bounds;
  int32_t *d_status, *d_any_fallback;
  int32_t sub_batch, lookahead, flags;
};

int ensure_events(D2pcPath *p, size_t n) {
  while (p->events.size() < n) {
    cudaEvent_t e;
    cudaError_t err = cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    if (err != cudaSuccess) return record_cuda_error(err);
    p->events.push_back(e);
  }
  return D2PC_OK;
}

#define PATH_CUDA(x)                                         \
  do {                                                       \
    cudaError_t e__ = (x);                                   \
    if (e__ != cudaSuccess) return record_cuda_error(e__);   \
  } while (0)

int path_kernel_supported(const D2pcConfig &cfg, const KParams &kp, const EmitArgs &ea) {
  const bool mask = cfg.use_z_range || cfg.drop_nonfinite;
  return kp.g.native && cfg.step == 1 && !mask && cfg.img_c == 3 && (cfg.img_w % 4) == 0 &&
         (((uintptr_t)kp.depth & 15u) == 0u) && (((uintptr_t)ea.bgr & 3u) == 0u) && (kp.g.N & 3u) == 0u &&
         ((unsigned long long)kp.g.N * (unsigned long long)kp.g.nu < (1ull << 40)) && kp.g.P > (uint32_t)kSortCap;
}

size_t path_kernel_smem() {
  size_t m = sizeof(ScanTileSmemT<kPathScanPT>);
  if (kEmitStageBytes > m) m = kEmitStageBytes;
  return m;  // a selection slice uses what is there (kSliceCap floats fit)
}

static int env_int(const char *name, int dflt, int lo, int hi);

// sample (all frames) -> persistent kernel -> status, all on one stream
int path_ordered(const PathArgs &a, cudaStream_t st) {
  const D2pcConfig &cfg = a.cfg;
  KParams kp = make_kparams(cfg, a.d_depth, a.d_workspace);
  kp.hints = (a.flags & D2PC_PATH_NO_L2_HINTS) ? 0 : (kHintScanKeep | kHintEmitDepthFirst | kHintStreamFirst);
  if (const char *h = getenv("D2PC_HINTS")) kp.hints = atoi(h);
  const EmitArgs ea = make_emit_args(cfg, a.d_bgr, a.d_xyz, a.d_rgb, a.d_count);
  int rc;
  const int skip = env_int("D2PC_PATH_SKIP", 0, 0, 15);
  if (skip) {
    if ((rc = sched_reset_launch(kp, st)) != D2PC_OK) return rc;
  } else if ((rc = stats_launch(kp, st, kStatsSample)) != D2PC_OK) return rc;
  if (cfg.want_bounds) {
    if ((rc = emit_init_launch(kp, st)) != D2PC_OK) return rc;
  }
  FastArgs fa;
  fa.tiles_per_frame = kp.emit_tiles;
  fa.total_tiles = kp.emit_tiles * (uint32_t)cfg.batch;
  fa.batch = (uint32_t)cfg.batch;
  fa.magic_w = ((1ull << 40) + (unsigned long long)kp.g.nu - 1ull) / (unsigned long long)kp.g.nu;
  fa.pc_simple = consts_simple(ea.pc) ? 1 : 0;
  SchedParams sp;
  sp.F = cfg.batch;
  sp.D = env_int("D2PC_PATH_D", a.lookahead > 0 ? a.lookahead : 3, 2, 64);
  sp.ns = (kp.g.P + kPathScanPT * kScanThreads - 1) / (kPathScanPT * kScanThreads);
  sp.ne = kp.emit_tiles;
  sp.R = (sp.ne + sp.ns - 1) / sp.ns;
  {
    const uint32_t k = (kp.cand_cap + kSliceCap - 1u) / kSliceCap;
    sp.K = (uint32_t)env_int("D2PC_PATH_K", (int)(k < 8u ? 8u : (k > 64u ? 64u : k)), 1, 64);
  }
  sp.skip = (uint32_t)skip;
  sp.nq = (uint32_t)env_int("D2PC_PATH_NQ", 16, 1, kSchedQueues);
  sp.group_items = 2u * sp.K + sp.ns * (sp.R + 1u);
  const unsigned long long total = (unsigned long long)(sp.F + sp.D) * sp.group_items;
  if (total >= 0xFFFF0000ull) return D2PC_ERR_UNSUPPORTED;
  sp.total_items = (uint32_t)total;
  sp.vec_ok = ((kp.g.P & 3u) == 0u) && (((uintptr_t)kp.depth & 15u) == 0u);
  const size_t smem = path_kernel_smem();
  const bool trace = env_int("D2PC_PATH_TRACE", 0, 0, 1) != 0;
  void (*kern)(KParams, EmitArgs, FastArgs, SchedParams) =
      cfg.want_bounds ? (trace ? path_kernel<true, true> : path_kernel<true, false>)
                      : (trace ? path_kernel<false, true> : path_kernel<false, false>);
  PATH_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<sp.total_items, kEmitThreads, smem, st>>>(kp, ea, fa, sp);
  D2PC_CHECK_LAUNCH();
  if (cfg.want_bounds) {
    if ((rc = bounds_export_launch(kp, a.d_bounds, st)) != D2PC_OK) return rc;
  }
  if (a.d_status || a.d_any_fallback) {
    if ((rc = status_launch(kp, a.d_status, a.d_any_fallback, st)) != D2PC_OK) return rc;
  }
  return D2PC_OK;
}

static int env_int(const char *name, int dflt, int lo, int hi) {
  const char *v = getenv(name);
  int x = v ? atoi(v) : dflt;
  return x < lo ? lo : (x > hi ? hi : x);
}

// Enqueue the pipelined step.  Sub-batch k: statistics on auxiliary stream k % n_aux, emit on emit stream
// k % n_emit (the caller's stream when n_emit == 1); the statistics of sub-batch k wait for the emit of
// sub-batch k - window, so at most `window` sub-batches of depth wait in L2 (and a ring slot is free again
// before it is rewritten).
int path_body(D2pcPath *p, const PathArgs &a, cudaStream_t s_main) {
  const D2pcConfig &cfg = a.cfg;
  KParams kp = make_kparams(cfg, a.d_depth, a.d_workspace);
  const bool generic = !kp.g.native && resize_is_generic(kp.g.h, kp.g.w);
  const int B = cfg.batch, S = generic ? B : a.sub_batch;
  const int n_sub = (B + S - 1) / S;
  const bool overlap = !(a.flags & D2PC_PATH_NO_OVERLAP) && n_sub > 1;
  const int window = (a.lookahead > 0 ? a.lookahead : 1) + 1;                       // sub-batches in flight (scanned, not yet emitted)
  const int n_aux = overlap ? env_int("D2PC_PATH_AUX", window < kMaxAux ? window : kMaxAux, 1, kMaxAux) : 1;
  const int n_emit = overlap ? env_int("D2PC_PATH_EMIT", 2, 1, kMaxEmit) : 1;
  const bool sample_first = n_sub > 1 && env_int("D2PC_PATH_SAMPLE_FIRST", 1, 0, 1) != 0;
  const bool ring = kp.resized != nullptr && n_sub > 1;
  // one stage: nothing is read twice from L2, so everything the emit touches is marked evict-first (its depth
  // and colour loads are last uses, its rows are never re-read): measured 1.527 against 1.538 ms per 128 frames
  kp.hints = (a.flags & D2PC_PATH_NO_L2_HINTS) ? 0 : (n_sub > 1 ? kHintPipeline : (kHintEmitDepthFirst | kHintStreamFirst));
  // a ring slot is rewritten while its lines are still dirty in L2: do not demote them after the emit's read
  if (ring) kp.hints &= ~kHintEmitDepthFirst;
  if (const char *h = getenv("D2PC_HINTS")) kp.hints = atoi(h);  // measurement aid
  const EmitArgs ea = make_emit_args(cfg, a.d_bgr, a.d_xyz, a.d_rgb, a.d_count);
  int rc;
  if ((rc = ensure_events(p, 1 + 3 * (size_t)n_sub)) != D2PC_OK) return rc;
  cudaEvent_t ev_fork = p->events[0];
  cudaEvent_t *ev_stats = &p->events[1], *ev_emit = &p->events[1 + n_sub], *ev_scan = &p->events[1 + 2 * n_sub];
  // Staggered chain (measurement aid, D2PC_PATH_STAGGER=1): scan(k + 1) starts when scan(k) has finished, so the
  // latency-bound selection of sub-batch k runs under the scan of sub-batch k + 1 and the emits behind it instead
  // of all sub-batches moving in lockstep.  Measured slower than one stage (128 x 1080p: 1.65 ms staggered in
  // halves against 1.57 ms): kernels that share the GPU each slow down by what the overlap would have saved.
  const bool stagger = overlap && n_aux > 1 && env_int("D2PC_PATH_STAGGER", 0, 0, 1) != 0;
  if ((rc = taps_launch(kp, s_main)) != D2PC_OK) return rc;
  if (sample_first) {  // the sample depends on the input only: all frames at once, off the per-stage chain
    if ((rc = stats_launch(kp, s_main, kStatsSample)) != D2PC_OK) return rc;
  }
  if (overlap) {
    PATH_CUDA(cudaEventRecord(ev_fork, s_main));
    for (int i = 0; i < n_aux; ++i) PATH_CUDA(cudaStreamWaitEvent(p->s_aux[i], ev_fork, 0));
    if (n_emit > 1)
      for (int i = 0; i < n_emit; ++i) PATH_CUDA(cudaStreamWaitEvent(p->s_emit[i], ev_fork, 0));
  }
  for (int k = 0; k < n_sub; ++k) {
    const int b0 = k * S, nb = (b0 + S <= B) ? S : B - b0;
    const KParams ks = slice_kparams(kp, b0, nb, ring ? (k % window) * S : -1);
    cudaStream_t s_stats = overlap ? p->s_aux[k % n_aux] : s_main;
    cudaStream_t s_em = (overlap && n_emit > 1) ? p->s_emit[k % n_emit] : s_main;
    if (overlap && k >= window) PATH_CUDA(cudaStreamWaitEvent(s_stats, ev_emit[k - window], 0));
    if (stagger) {
      if (!sample_first && (rc = stats_launch(ks, s_stats, kStatsSample)) != D2PC_OK) return rc;
      if (k > 0) PATH_CUDA(cudaStreamWaitEvent(s_stats, ev_scan[k - 1], 0));
      if ((rc = stats_launch(ks, s_stats, kStatsScan)) != D2PC_OK) return rc;
      PATH_CUDA(cudaEventRecord(ev_scan[k], s_stats));
      if ((rc = stats_launch(ks, s_stats, kStatsSelect)) != D2PC_OK) return rc;
    } else if ((rc = stats_launch(ks, s_stats, sample_first ? kStatsScanSelect : (kStatsSample | kStatsScanSelect))) != D2PC_OK)
      return rc;
    if (overlap) {
      PATH_CUDA(cudaEventRecord(ev_stats[k], s_stats));
      PATH_CUDA(cudaStreamWaitEvent(s_em, ev_stats[k], 0));
    }
    if ((rc = emit_launch(cfg, ks, slice_emit_args(ea, kp.g, b0), a.d_bounds ? a.d_bounds + 6 * (size_t)b0 : nullptr,
                          s_em, 0, nullptr, nullptr)) != D2PC_OK)
      return rc;
    if (overlap) PATH_CUDA(cudaEventRecord(ev_emit[k], s_em));
  }
  if (overlap && n_emit > 1) {  // join: the last emit of every emit stream (the statistics streams joined them)
    for (int k = n_sub - 1; k >= 0 && k >= n_sub - n_emit; --k) PATH_CUDA(cudaStreamWaitEvent(s_main, ev_emit[k], 0));
  }
  if (a.d_status || a.d_any_fallback) {
    if ((rc = status_launch(kp, a.d_status, a.d_any_fallback, s_main)) != D2PC_OK) return rc;
  }
  return D2PC_OK;
}

}  // namespace

extern "C" int d2pc_path_create(D2pcPath **path) {
  if (!path) return D2PC_ERR_INVALID_ARGUMENT;
  *path = nullptr;
  D2pcPath *p = new D2pcPath();
  p->exec = nullptr;
  p->s_cap = nullptr;
  for (int i = 0; i < kMaxAux; ++i) p->s_aux[i] = nullptr;
  for (int i = 0; i < kMaxEmit; ++i) p->s_emit[i] = nullptr;
  cudaError_t e = cudaGetDevice(&p->device);
  int lo = 0, hi = 0;
  if (e == cudaSuccess) e = cudaDeviceGetStreamPriorityRange(&lo, &hi);  // hi = numerically lowest = highest priority
  if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&p->s_cap, cudaStreamNonBlocking, lo);
  for (int i = 0; i < kMaxAux && e == cudaSuccess; ++i)
    e = cudaStreamCreateWithPriority(&p->s_aux[i], cudaStreamNonBlocking, getenv("D2PC_PATH_NOPRIO") ? lo : hi);
  for (int i = 0; i < kMaxEmit && e == cudaSuccess; ++i)
    e = cudaStreamCreateWithPriority(&p->s_emit[i], cudaStreamNonBlocking, lo);
  if (e != cudaSuccess) {
    d2pc_path_destroy(p);
    return record_cuda_error(e);
  }
  int rc = stats_prepare();
  if (rc) { d2pc_path_destroy(p); return rc; }
  *path = p;
  return D2PC_OK;
}

extern "C" void d2pc_path_destroy(D2pcPath *p) {
  if (!p) return;
  if (p->exec) cudaGraphExecDestroy(p->exec);
  for (cudaEvent_t e : p->events) cudaEventDestroy(e);
  if (p->s_cap) cudaStreamDestroy(p->s_cap);
  for (int i = 0; i < kMaxAux; ++i)
    if (p->s_aux[i]) cudaStreamDestroy(p->s_aux[i]);
  for (int i = 0; i < kMaxEmit; ++i)
    if (p->s_emit[i]) cudaStreamDestroy(p->s_emit[i]);
  delete p;
}

extern "C" int d2pc_path_enqueue(D2pcPath *p, const D2pcConfig *cfg, const float *d_depth, const uint8_t *d_bgr,
                                 void *d_workspace, size_t workspace_bytes, float *d_xyz, float *d_rgb,
                                 uint32_t *d_count, float *d_bounds, int32_t *d_status, int32_t *d_any_fallback,
                                 int32_t sub_batch, int32_t lookahead, int32_t flags, void *stream) {
  if (!p) return D2PC_ERR_INVALID_ARGUMENT;
  int rc = check_workspace(cfg, d_workspace, workspace_bytes);
  if (rc) return rc;
  rc = emit_validate(cfg, d_depth, d_bgr, d_workspace, workspace_bytes, d_xyz, d_rgb, d_count, d_bounds);
  if (rc) return rc;
  int dev = -1;
  PATH_CUDA(cudaGetDevice(&dev));
  if (dev != p->device) return D2PC_ERR_INVALID_ARGUMENT;  // the handle's streams belong to its device
  if (sub_batch <= 0 || sub_batch > cfg->batch) sub_batch = cfg->batch;
  if (lookahead < 0) lookahead = 0;
  PathArgs a;
  memset(&a, 0, sizeof(a));
  a.cfg = *cfg;
  a.d_depth = d_depth; a.d_bgr = d_bgr; a.d_workspace = d_workspace; a.workspace_bytes = workspace_bytes;
  a.d_xyz = d_xyz; a.d_rgb = d_rgb; a.d_count = d_count; a.d_bounds = d_bounds;
  a.d_status = d_status; a.d_any_fallback = d_any_fallback;
  a.sub_batch = sub_batch; a.lookahead = lookahead; a.flags = flags;
  cudaStream_t st = (cudaStream_t)stream;
  {
    const KParams kp0 = make_kparams(a.cfg, a.d_depth, a.d_workspace);
    const EmitArgs ea0 = make_emit_args(a.cfg, a.d_bgr, a.d_xyz, a.d_rgb, a.d_count);
    if ((flags & D2PC_PATH_ORDERED) && path_kernel_supported(a.cfg, kp0, ea0)) return path_ordered(a, st);
  }
  if (!(flags & D2PC_PATH_GRAPH)) return path_body(p, a, st);

  const unsigned char *kb = reinterpret_cast<const unsigned char *>(&a);
  if (!p->exec || p->key.size() != sizeof(a) || memcmp(p->key.data(), kb, sizeof(a)) != 0) {
    if (p->exec) { cudaGraphExecDestroy(p->exec); p->exec = nullptr; }
    p->key.clear();
    cudaGraph_t graph = nullptr;
    PATH_CUDA(cudaStreamBeginCapture(p->s_cap, cudaStreamCaptureModeThreadLocal));
    rc = path_body(p, a, p->s_cap);
    cudaError_t e = cudaStreamEndCapture(p->s_cap, &graph);
    if (rc != D2PC_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (e != cudaSuccess) return record_cuda_error(e);
    e = cudaGraphInstantiate(&p->exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) { p->exec = nullptr; return record_cuda_error(e); }
    p->key.assign(kb, kb + sizeof(a));
  }
  PATH_CUDA(cudaGraphLaunch(p->exec, st));
  return D2PC_OK;
}

// measurement aid: where the persistent kernel leaves its per-frame trace (8 x uint64 ns per frame) in the workspace
extern "C" int d2pc_path_trace_offset(const D2pcConfig *cfg, size_t *offset) {
  int rc = validate_config(cfg);
  if (rc) return rc;
  if (!offset) return D2PC_ERR_INVALID_ARGUMENT;
  *offset = make_layout(*cfg).sched_off + kSchedBytes;
  return D2PC_OK;
}
