// d2pc_stats.cu -- reference steps a1..a5 (backend/app.py:186-206) on the device:
// exact 2nd/98th-percentile order statistics of every frame's (virtually resized) depth map,
// non-finite repair (np.nanmedian) and the frame's normalisation parameters.
//
// Fast path (3-4 launches per batch, depth map read from HBM exactly once):
//   sample_kernel  1 CTA/frame   stratified sample (registers) -> exact sample order statistics
//                                (bucket histogram + member pick) -> key brackets [L, U]
//                                that contain the wanted ranks with ~6 sigma margin
//   scan_*_kernel  streaming     per pixel, 4 instructions: values strictly between the two brackets
//                                (about 94%) need nothing; everything else, and every NaN, is deferred
//                                to the frame's two raw queues (one reservation per CTA).  Native
//                                depth, large grids: several tiles per CTA with the next tile's loads
//                                in flight during a tile's epilogue (scan_native_multi_kernel);
//                                resized depth: the interpolation is fused (scan_resized_tiled_kernel)
//   select_kernel  2 CTA/frame   classify the queued values against the bracket (below / equal to a
//                                bound / strictly inside), resolve the wanted ranks, exact selection
//                                inside the bracket (bucket histogram over the L2-resident queue, then
//                                the members of the wanted buckets); the last CTA of a frame evaluates
//                                NumPy's _lerp in float64 and writes the parameter block.  Small
//                                batches and 4K frames: K CTAs per bracket in two launches
//                                (select_a_kernel, select_c_kernel); one or two frames: scan and
//                                selection in one index-ordered launch (stats_ordered_kernel)
// Frames the fast path cannot finish *exactly* (non-finite values, bracket miss, queue
// overflow, collapsed percentiles) are only marked; d2pc_stats_fallback_enqueue runs the
// input-agnostic exact path (8-bit radix select, nanmedian repair) for those.
#include <stdlib.h>

#include "d2pc_stats_dev.cuh"

namespace d2pc {

__device__ __forceinline__ int32_t bracket_margin(double q, int S) {
  double sd = sqrt((double)S * q * (1.0 - q));
  return (int32_t)ceil(6.0 * sd) + 4;
}

// ------------------------------------------------------------------------------------------
// taps_kernel: a1 coordinate math once per call (float64, exactly d2pc_math.h axis_tap)
// ------------------------------------------------------------------------------------------
__global__ void taps_kernel(KParams kp) {
  TapEntry *xt = const_cast<TapEntry *>(kp.xtab), *yt = const_cast<TapEntry *>(kp.ytab);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < kp.g.W) {
    AxisTap a = axis_tap(i, kp.g.scale_x, kp.g.w);
    xt[i].i0c = a.i0 | (a.clamped ? (int32_t)0x80000000 : 0);
    xt[i].t = a.t;
  } else if (i < kp.g.W + kp.g.H) {
    AxisTap a = axis_tap(i - kp.g.W, kp.g.scale_y, kp.g.h);
    yt[i - kp.g.W].i0c = a.i0 | (a.clamped ? (int32_t)0x80000000 : 0);
    yt[i - kp.g.W].t = a.t;
  }
}

// ------------------------------------------------------------------------------------------
// resize_generic_kernel: a1 for sources with a 1-pixel side (d2pc_math.h generic_sample), materialised
// before the statistics; everything after it runs on the (H x W) map like a native-size input.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) resize_generic_kernel(KParams kp) {
  const uint32_t b = blockIdx.y;
  const float *src = kp.depth + (size_t)b * kp.g.D;
  float *dst = kp.resized + (size_t)b * kp.g.P;
  for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < kp.g.P; p += gridDim.x * blockDim.x) {
    const uint32_t v = p / (uint32_t)kp.g.W, u = p - v * (uint32_t)kp.g.W;
    dst[p] = generic_sample(src, kp.g.w, generic_tap((int32_t)u, kp.g.scale_x, kp.g.w, 1),
                            generic_tap((int32_t)v, kp.g.scale_y, kp.g.h, 0));
  }
}

// ------------------------------------------------------------------------------------------
// sample_kernel
// ------------------------------------------------------------------------------------------
template <bool NATIVE>
__global__ void __launch_bounds__(kSelThreads, 1) sample_kernel(KParams kp) {
  extern __shared__ uint32_t s_hist[];  // 4 << kSelBits words
  __shared__ uint32_t s_bad, s_min, s_max;
  __shared__ uint32_t s_res[2 * 4 + 40];
  __shared__ uint32_t s_warp[33], s_cnt[4], s_lmin[4], s_lmax[4], s_out[4];
  const int b = blockIdx.x;
  FrameState *fs = kp.state + b;
  const float *frame = kp.depth + (size_t)b * kp.g.D;
  const uint32_t n = kp.g.P;
  const int tid = threadIdx.x;
  if (tid == 0) {
    s_bad = 0; s_min = 0xFFFFFFFFu; s_max = 0u;
    for (int i = 0; i < 2; ++i) {
      fs->below[i] = 0; fs->eqL[i] = 0; fs->inside[i] = 0; fs->eqU[i] = 0;
    }
    fs->above1 = 0;
    fs->n_nonfinite = 0; fs->n_nan = 0; fs->nqueue[0] = 0; fs->nqueue[1] = 0;
    fs->min_key = 0xFFFFFFFFu; fs->max_key = 0u;
    for (int i = 0; i < 4; ++i) fs->sel_key[i] = 0;
    fs->sel_fail = 0; fs->sel_done = 0;
    if (b == 0)
      for (int i = 0; i <= kSchedQueues; ++i) kp.sched[32 * i] = 0u;
    {  // trace of the persistent path kernel: min-slots start at ~0, max-slots at 0
      unsigned long long *tr = reinterpret_cast<unsigned long long *>(reinterpret_cast<unsigned char *>(kp.sched) + kSchedBytes) + (size_t)b * 8;
      for (int i = 0; i < 8; ++i) tr[i] = (i == 0 || i == 6) ? ~0ull : 0ull;
    }
    fs->status = D2PC_FRAME_PENDING;
    fs->fb_active = 0; fs->fb_any_nan = 0;
    fs->norm.has_nonfinite = 0;
    fs->norm.median = 0.0f;
    fs->norm.simple = 0;
  }
  {  // scratch of the cooperative selection (persistent path kernel)
    uint32_t *z = reinterpret_cast<uint32_t *>(kp.sel + b);
    for (uint32_t i = tid; i < sizeof(SelShared) / 4; i += kSelThreads) z[i] = 0u;
  }
  if (n <= (uint32_t)kSortCap) {  // small frame: every finite key is a candidate
    if (tid == 0) {
      fs->brL[0] = fs->brL[1] = 0u;
      fs->brU[0] = fs->brU[1] = 0xFFFFFFFFu;
      fs->sample_ok = 1;
    }
    return;
  }
  // stratified sample, kSampleSize / blockDim keys per thread, kept in registers
  constexpr int E = kSampleSize / kSelThreads;
  const uint32_t S = kSampleSize;
  uint32_t k[E];
  bool bad = false;
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const uint32_t j = (uint32_t)e * kSelThreads + (uint32_t)tid;
    uint32_t start = (uint32_t)(((unsigned long long)j * n) / S);
    uint32_t end = (uint32_t)(((unsigned long long)(j + 1) * n) / S);
    uint32_t idx = start + hash_u32(j * 0x9E3779B9u + (uint32_t)b * 0x85EBCA6Bu + 12345u) % (end - start);
    float v;
    if (NATIVE) {  // one scattered 4-byte sample per load: ask L2 for a 64-byte fill instead of the default 128 (LDG.LTC64B)
      asm volatile("ld.global.nc.L2::64B.f32 %0, [%1];" : "=f"(v) : "l"(frame + idx));
    } else {
      v = depth_at<NATIVE>(frame, kp, idx);
    }
    k[e] = 0xFFFFFFFFu;
    if (is_finite_f32(v)) k[e] = float_to_key(v); else bad = true;
  }
  __syncthreads();
  if (bad) atomicOr(&s_bad, 1u);
  {  // key range of the sample (makes the first histogram level adaptive)
    uint32_t mn = 0xFFFFFFFFu, mx = 0u;
#pragma unroll
    for (int e = 0; e < E; ++e) { mn = min(mn, k[e]); mx = max(mx, k[e]); }
    mn = warp_min(mn); mx = warp_max(mx);
    if ((tid & 31) == 0) { atomicMin(&s_min, mn); atomicMax(&s_max, mx); }
  }
  __syncthreads();
  // sample ranks that bracket the wanted order statistics with a ~6 sigma margin
  uint32_t rank[4], out[4];
  bool open_lo[2], open_hi[2];
#pragma unroll
  for (int br = 0; br < 2; ++br) {
    const double q = br ? D2PC_Q98 : D2PC_Q02;
    RankPair rp = percentile_ranks(n, q);
    long long i_lo = (long long)(((unsigned long long)rp.lo * S) / n);
    long long i_hi = (long long)(((unsigned long long)rp.hi * S) / n) + 1;
    int32_t m = bracket_margin(q, (int)S);
    long long iL = i_lo - m, iU = i_hi + m;
    open_lo[br] = iL < 0;
    open_hi[br] = iU >= (long long)S;
    rank[2 * br + 0] = open_lo[br] ? 0u : (uint32_t)iL;
    rank[2 * br + 1] = open_hi[br] ? S - 1 : (uint32_t)iU;
  }
  // exact sample order statistics: two-pass selection (bucket histogram -> the wanted buckets' members)
  {
    uint32_t *s_list = s_hist + kFastBins;  // [4][kFastListCap]
    const uint32_t lo0 = s_min, span = s_max - s_min;
    const int shift = fast_shift(span);
    for (uint32_t i = tid; i < kFastBins; i += kSelThreads) s_hist[i] = 0u;
    if (tid < 4) { s_cnt[tid] = 0u; s_lmin[tid] = 0xFFFFFFFFu; s_lmax[tid] = 0u; s_out[tid] = 0u; }
    __syncthreads();
#pragma unroll
    for (int e = 0; e < E; ++e) atomicAdd(&s_hist[(k[e] - lo0) >> shift], 1u);
    __syncthreads();
    uint32_t bin[4], base[4];
    const bool want[4] = {true, true, true, true};
    block_locate<4>(s_hist, rank, want, bin, base, s_warp, s_res);
    bool slow = false;
#pragma unroll
    for (int t = 0; t < 4; ++t) slow = slow || bin[t] == 0xFFFFFFFFu;  // cannot happen (ranks < S)
    if (!slow && shift == 0) {
#pragma unroll
      for (int t = 0; t < 4; ++t) out[t] = lo0 + bin[t];
    } else if (!slow) {
      int owner[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        owner[t] = t;
#pragma unroll
        for (int u = t - 1; u >= 0; --u)
          if (bin[u] == bin[t]) owner[t] = u;
      }
#pragma unroll
      for (int e = 0; e < E; ++e) {
        const uint32_t bk = (k[e] - lo0) >> shift;
#pragma unroll
        for (int t = 0; t < 4; ++t)
          if (owner[t] == t && bk == bin[t]) {
            const uint32_t idx = atomicAdd(&s_cnt[t], 1u);
            if (idx < kFastListCap) s_list[t * kFastListCap + idx] = k[e];
            atomicMin(&s_lmin[t], k[e]);
            atomicMax(&s_lmax[t], k[e]);
          }
      }
      __syncthreads();
#pragma unroll
      for (int t = 0; t < 4; ++t)
        if (owner[t] == t && s_cnt[t] > kFastListCap && s_lmin[t] != s_lmax[t]) slow = true;
      if (!slow) {
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int o = owner[t];
          if (s_lmin[o] == s_lmax[o]) { if (tid == 0) s_out[t] = s_lmin[o]; }
          else block_pick(s_list + o * kFastListCap, s_cnt[o], rank[t] - base[t], &s_out[t]);
        }
        __syncthreads();
#pragma unroll
        for (int t = 0; t < 4; ++t) out[t] = s_out[t];
      }
    }
    if (slow) {  // uniform: heavy ties inside one bucket -> the general multi-level selection
      __syncthreads();
      block_hist_select<4, kSelBits>([&](auto f) {
#pragma unroll
        for (int e = 0; e < E; ++e) f(k[e]);
      }, s_min, s_max, rank, out, s_hist, s_res);
    }
  }
  if (tid < 2) {
    fs->brL[tid] = open_lo[tid] ? 0u : out[2 * tid + 0];
    fs->brU[tid] = open_hi[tid] ? 0xFFFFFFFFu : out[2 * tid + 1];
  }
  if (tid == 0) fs->sample_ok = s_bad ? 0u : 1u;
}

// ------------------------------------------------------------------------------------------
// scan_kernel
// ------------------------------------------------------------------------------------------
// One streaming pass, 4 instructions per pixel on the common path.  With the brackets [L0, U0] (around the
// 2% ranks) and [L1, U1] (around the 98% ranks), about 94% of the pixels lie strictly between U0 and L1 and
// need nothing at all: they are below bracket 1 and above bracket 0, which the frame's totals account for.
// Every other value (about 6%: below or inside bracket 0, inside or above bracket 1, or non-finite -- every
// ordered compare with NaN is false) is appended to the thread's private queue column in shared memory and
// classified in the epilogue: counted (below L0, above U1, non-finite) or forwarded to the bracket's raw
// queue in global memory.  Compares are float compares (-0.0 == +0.0), here and in the selection,
// consistently.  min/max are not tracked: a frame whose percentiles collapse (p98 <= p2) goes to the exact
// fallback, which computes them.
template <int QD>
struct ScanSharedT {
  float pqueue[QD][kScanThreads];  // per-thread deferred values, slot-major
  ScanFlushSmem flush;
};
using ScanShared = ScanSharedT<kScanPerThread>;

// Resized depth, up-scaling geometry: one CTA owns a kRzRows x kRzCols tile of the (H x W) map.
// The horizontal lerp of every source row the tile needs is computed once into shared memory
// (a source row feeds ~1/scale_y destination rows), then each thread finishes 4 consecutive
// columns of a row with one vertical lerp from two LDS.128, scans them and writes the resized map
// with one 16 B store.  Same arithmetic as bilinear_taps (horizontal fmaf, then vertical fmaf;
// clamped taps copy; corner blocks turn +-inf into NaN).
constexpr int kRzCols = 128, kRzRows = 32, kRzSrcRows = 40;
constexpr int kRzPerThread = kRzCols * kRzRows / kScanThreads;  // 16

__global__ void __launch_bounds__(kScanThreads) scan_resized_tiled_kernel(KParams kp) {
  __shared__ ScanSharedT<kRzPerThread> sh;
  __shared__ __align__(16) float s_h[kRzSrcRows][kRzCols];
  const int b = blockIdx.z, tid = threadIdx.x;
  FrameState *fs = kp.state + b;
  const float *frame = kp.depth + (size_t)b * kp.g.D;
  const int32_t W = kp.g.W, H = kp.g.H, sw = kp.g.w;
  float Lf[2], Uf[2];
  bracket_floats(fs, Lf, Uf);
  const uint32_t q0 = (uint32_t)__cvta_generic_to_shared(&sh.pqueue[0][tid]);
  uint32_t qaddr = q0;
  const int32_t u0 = blockIdx.x * kRzCols, v0 = blockIdx.y * kRzRows;
  const int32_t rows = min(kRzRows, H - v0);
  const int32_t ys = kp.ytab[v0].i0c & 0x7FFFFFFF;
  const TapEntry tl = kp.ytab[v0 + rows - 1];
  const int32_t nsr = (tl.i0c & 0x7FFFFFFF) + (tl.i0c < 0 ? 0 : 1) - ys + 1;  // <= kRzSrcRows (host checked)
  {  // phase 1: horizontally resized source rows of this tile
    const int32_t col = tid & (kRzCols - 1), u = u0 + col;
    if (u < W) {
      const TapEntry tx = kp.xtab[u];
      const int32_t x0 = tx.i0c & 0x7FFFFFFF;
      const bool cx = tx.i0c < 0;
      for (int32_t r = tid / kRzCols; r < nsr; r += kScanThreads / kRzCols) {
        const float *row = frame + (size_t)(ys + r) * sw;
        const float a = __ldg(row + x0);
        float h = a;
        if (!cx) h = fmaf(__ldg(row + x0 + 1) - a, tx.t, a);
        s_h[r][col] = h;
      }
    }
  }
  __syncthreads();
  {  // phase 2: vertical lerp, scan, materialise
    const int32_t c4 = (tid & 31) * 4, uu = u0 + c4, warp = tid >> 5;
    const uint64_t pol_keep = l2_policy(false, (kp.hints & kHintResizedKeep) != 0);
    if (uu < W) {  // W % 4 == 0: the four columns are inside together
      bool cxk[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) cxk[k] = kp.xtab[uu + k].i0c < 0;
#pragma unroll 1
      for (int32_t rr = warp; rr < rows; rr += kScanThreads / 32) {
        const int32_t v = v0 + rr;
        const TapEntry ty = kp.ytab[v];
        const int32_t y0 = (ty.i0c & 0x7FFFFFFF) - ys;
        const float4 r0 = *reinterpret_cast<const float4 *>(&s_h[y0][c4]);
        float o[4] = {r0.x, r0.y, r0.z, r0.w};
        if (ty.i0c < 0) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (cxk[k] && is_inf_f32(o[k])) o[k] = nan_f32();
        } else {
          const float4 r1 = *reinterpret_cast<const float4 *>(&s_h[y0 + 1][c4]);
          o[0] = fmaf(r1.x - r0.x, ty.t, r0.x);
          o[1] = fmaf(r1.y - r0.y, ty.t, r0.y);
          o[2] = fmaf(r1.z - r0.z, ty.t, r0.z);
          o[3] = fmaf(r1.w - r0.w, ty.t, r0.w);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) scan_value(o[k], Uf[0], Lf[1], qaddr);
        float *dst = kp.resized + (size_t)b * kp.g.P + (size_t)v * W + uu;
        stg_f4_pol(dst, make_float4(o[0], o[1], o[2], o[3]), pol_keep);
      }
    }
  }
  scan_flush(&sh.pqueue[0][0], (qaddr - q0) / (uint32_t)(kScanThreads * 4), Uf[0], Lf[1], fs, kp, b, sh.flush);
}

// native-size (or already materialised) depth: one kTilePx tile per CTA (small grids)
template <int PT>
__global__ void __launch_bounds__(kScanThreads) scan_native_kernel(KParams kp, int vec_ok) {
  extern __shared__ __align__(16) unsigned char s_dyn[];
  scan_tile<PT>(kp, blockIdx.y, blockIdx.x, vec_ok, *reinterpret_cast<ScanTileSmemT<PT> *>(s_dyn));
}

// The same for grids of many tiles: TPC consecutive kTilePx tiles of one frame per CTA.  A tile's
// epilogue (two barriers and the round trip of the queue reservation) keeps a CTA's slot busy with nothing in
// flight; the next tile's loads are therefore issued into the registers the current tile has just consumed,
// BEFORE its epilogue, so the epilogue's latency hides under them.
template <int PT, int TPC, int MINB>
__global__ void __launch_bounds__(kScanThreads, MINB) scan_native_multi_kernel(KParams kp, int vec_ok) {
  extern __shared__ __align__(16) unsigned char s_dyn[];
  ScanTileSmemT<PT> &sm = *reinterpret_cast<ScanTileSmemT<PT> *>(s_dyn);
  constexpr int kPx = PT * kScanThreads;
  const int tid = threadIdx.x, b = blockIdx.y;
  FrameState *fs = kp.state + b;
  const float *frame = kp.depth + (size_t)b * kp.g.D;
  const uint32_t n = kp.g.P;
  const uint32_t ntiles = (n + (uint32_t)kPx - 1u) / (uint32_t)kPx;
  const uint32_t t0 = blockIdx.x * (uint32_t)TPC;
  const float U0 = bracket_hi_float(fs->brU[0]), L1 = bracket_lo_float(fs->brL[1]);
  const uint32_t q0 = (uint32_t)__cvta_generic_to_shared(&sm.pq[0][tid]);
  const uint64_t pol = l2_policy((kp.hints & kHintScanKeep) == 0, (kp.hints & kHintScanKeep) != 0);   // one stage: the map streams through (evict-first)
  auto full = [&](uint32_t tile) { return vec_ok && tile < ntiles && (tile + 1u) * (uint32_t)kPx <= n; };
  float4 r[PT / 4];
  if (full(t0)) {
    const float *src = frame + t0 * (uint32_t)kPx + 4u * (uint32_t)tid;
#pragma unroll
    for (int j = 0; j < PT / 4; ++j) r[j] = ldg_f4_pol(src + (size_t)j * (4 * kScanThreads), pol);
  }
#pragma unroll 1
  for (uint32_t i = 0; i < (uint32_t)TPC; ++i) {
    const uint32_t tile = t0 + i;
    if (tile >= ntiles) break;  // uniform
    uint32_t qaddr = q0;
    if (full(tile)) {
#pragma unroll
      for (int j = 0; j < PT / 4; ++j) {
        scan_value(r[j].x, U0, L1, qaddr);
        scan_value(r[j].y, U0, L1, qaddr);
        scan_value(r[j].z, U0, L1, qaddr);
        scan_value(r[j].w, U0, L1, qaddr);
      }
    } else {  // last tile of a frame / unaligned frames
      const uint32_t tile_base = tile * (uint32_t)kPx;
#pragma unroll 1
      for (int j = 0; j < PT / 4; ++j) {
        const uint32_t p = tile_base + 4u * (uint32_t)(j * kScanThreads + tid);
        for (uint32_t k = 0; k < 4u; ++k)
          if (p + k < n) scan_value(__ldg(frame + p + k), U0, L1, qaddr);
      }
    }
    if (i + 1u < (uint32_t)TPC && full(tile + 1u)) {  // the next tile's loads fly during this tile's epilogue
      const float *src = frame + (tile + 1u) * (uint32_t)kPx + 4u * (uint32_t)tid;
#pragma unroll
      for (int j = 0; j < PT / 4; ++j) r[j] = ldg_f4_pol(src + (size_t)j * (4 * kScanThreads), pol);
    }
    scan_flush(&sm.pq[0][0], (qaddr - q0) / (uint32_t)(kScanThreads * 4), U0, L1, fs, kp, b, sm.flush);
    __syncthreads();  // every thread has read its columns (and the flush scratch) before the next tile's pushes
  }
}

// resized depth, geometries the tiled kernel does not cover: direct gather of the four taps per pixel
__global__ void __launch_bounds__(kScanThreads) scan_gather_kernel(KParams kp, int vec_ok) {
  __shared__ ScanShared sh;
  const int b = blockIdx.y;
  const int tid = threadIdx.x;
  FrameState *fs = kp.state + b;
  const float *frame = kp.depth + (size_t)b * kp.g.D;
  const uint32_t n = kp.g.P;
  float Lf[2], Uf[2];
  bracket_floats(fs, Lf, Uf);
  const uint32_t q0 = (uint32_t)__cvta_generic_to_shared(&sh.pqueue[0][tid]);
  uint32_t qaddr = q0;
  const uint32_t tile_base = blockIdx.x * (uint32_t)kScanTile;
  const uint32_t W = (uint32_t)kp.g.W;
#pragma unroll 2
  for (int j = 0; j < kScanPerThread / 4; ++j) {
    uint32_t p = tile_base + 4u * (uint32_t)(j * kScanThreads + tid);
    if (p >= n) continue;
    uint32_t v = p / W;
    uint32_t u = p - v * W;
    TapEntry ty = kp.ytab[v];
    float val[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
    for (uint32_t k = 0; k < 4u; ++k) {
      if (p + k < n) {
        if (u >= W) {
          u -= W;
          v += 1;
          ty = kp.ytab[v];
        }
        val[k] = bilinear_taps(frame, kp.g.w, kp.xtab[u], ty);
        scan_value(val[k], Uf[0], Lf[1], qaddr);
        u += 1;
      }
    }
    // materialise the resized map for the kernels that follow (emit, mask count, fallback)
    float *dst = kp.resized + (size_t)b * n + p;
    if (vec_ok && p + 3u < n) *reinterpret_cast<float4 *>(dst) = make_float4(val[0], val[1], val[2], val[3]);
    else
      for (uint32_t k = 0; k < 4u; ++k)
        if (p + k < n) dst[k] = val[k];
  }
  scan_flush(&sh.pqueue[0][0], (qaddr - q0) / (uint32_t)(kScanThreads * 4), Uf[0], Lf[1], fs, kp, b, sh.flush);
}

// ------------------------------------------------------------------------------------------
// select_kernel
// ------------------------------------------------------------------------------------------
// One CTA per (frame, bracket): shared-memory bucket histogram + member pick (select_bracket).  Measured against K
// cooperating CTAs per bracket with a global histogram as two launches (nothing waits): 91 us against 53 us per 64
// frames -- the ~62 000 L2 atomics per bracket cost more than the one CTA's two walks over the queue.
__global__ void __launch_bounds__(kSelectThreads, 2) select_kernel(KParams kp) {
  extern __shared__ uint32_t s_sel[];  // kSelectSmemWords
  __shared__ SelectSmall ss;
  select_bracket(kp, blockIdx.y, blockIdx.x, s_sel, ss);
}

// The same selection by K CTAs per (frame, bracket) as two launches (nothing waits): step A classifies the slices
// (shared-memory bucket histograms, merged into the frame's global one) and the last CTA of a bracket locates the
// wanted buckets; step C collects their members and the last CTA picks the exact keys.  For large frames / few
// frames, where one CTA per bracket walking a quarter of a million queued values twice is the whole stage
// (4K x 16: 166 us).
constexpr int kCoopThreads = 512;
constexpr uint32_t kCoopSmemWords = 2u << 11;   // bucket histogram in step A; lists / general selection in step C
__global__ void __launch_bounds__(kCoopThreads) select_a_kernel(KParams kp, uint32_t K) {
  extern __shared__ uint32_t s_sel[];
  __shared__ SelPartSmall ss;
  select_phase_a<true>(kp, blockIdx.z, blockIdx.y, blockIdx.x, K, reinterpret_cast<float *>(s_sel), 0u, ss);
}
__global__ void __launch_bounds__(kCoopThreads) select_c_kernel(KParams kp, uint32_t K) {
  extern __shared__ uint32_t s_sel[];
  __shared__ SelPartSmall ss;
  const int b = blockIdx.z, br = blockIdx.y;
  if (*reinterpret_cast<volatile uint32_t *>(&kp.sel[b].bin_ready[br]) != 1u) return;  // finished in step A
  select_phase_c(kp, b, br, blockIdx.x, K, reinterpret_cast<float *>(s_sel), 0u, false, ss);
}

// ------------------------------------------------------------------------------------------
// stats_ordered_kernel: scan and exact selection of a whole batch in ONE launch
// ------------------------------------------------------------------------------------------
// One work item per CTA, the item is the CTA's index; CTAs are dispatched in index order.  Group g holds the
// 2 K parts of frame g-1's cooperative selection, then the scan tiles of frame g: a frame's selection (a chain
// of latencies, ~20 us) runs under the scans of the frames behind it instead of after the whole batch's scan as
// a second, latency-bound launch, and a single frame needs two launches (sample, this) instead of three.  A
// selection part waits for its frame's scan tiles (smaller indices: dispatched before it) and for its peers
// (the next few indices); waits are bounded, a time-out raises the abort flag and leaves the frames PENDING,
// which the status kernel turns into NEEDS_FALLBACK.
constexpr uint32_t kStatsSliceCap = (uint32_t)(sizeof(ScanTileSmem) / sizeof(float));   // 8192 + a few words
__global__ void __launch_bounds__(kScanThreads, 5) stats_ordered_kernel(KParams kp, int vec_ok, uint32_t ns, uint32_t K) {
  extern __shared__ __align__(16) unsigned char s_dyn[];
  __shared__ SelPartSmall s_sel;
  __shared__ uint32_t s_go;
  const int tid = threadIdx.x;
  const uint32_t group_items = 2u * K + ns;
  const uint32_t g = blockIdx.x / group_items, i = blockIdx.x - g * group_items;
  uint32_t *abort_flag = sched_abort_flag(kp);
  if (i < 2u * K) {
    if (g < 1u) return;
    const int f = (int)g - 1;
    const uint32_t br = i / K, k = i - br * K;
    if (tid == 0) s_go = spin_until(&kp.sel[f].scan_done, abort_flag, [&](uint32_t v) { return v >= ns; }) ? 1u : 0u;
    __syncthreads();
    if (!s_go) return;
    select_part(kp, f, (int)br, k, K, reinterpret_cast<float *>(s_dyn), kStatsSliceCap, s_sel,
                [&](const uint32_t *p, auto pred) { return spin_until(p, abort_flag, pred); });
    return;
  }
  if (g >= (uint32_t)kp.batch) return;
  scan_tile<kTilePerThread>(kp, (int)g, i - 2u * K, vec_ok, *reinterpret_cast<ScanTileSmem *>(s_dyn));
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    atomicAdd(&kp.sel[g].scan_done, 1u);
  }
}

// ------------------------------------------------------------------------------------------
// fallback: exact 8-bit radix select over the whole map for flagged frames
// ------------------------------------------------------------------------------------------
enum { kStageMedian = 0, kStagePercentile = 1 };

// The fallback does not trust anything the fast path counted: it recounts NaN / non-finite
// values of the (virtually resized) map itself.
__global__ void fb_reset_kernel(KParams kp) {
  FrameState *fs = kp.state + blockIdx.x;
  if (threadIdx.x == 0 && fs->status == D2PC_FRAME_NEEDS_FALLBACK) {
    fs->n_nonfinite = 0; fs->n_nan = 0;
    fs->min_key = 0xFFFFFFFFu; fs->max_key = 0u;
  }
}

template <bool NATIVE>
__global__ void __launch_bounds__(kScanThreads) fb_count_kernel(KParams kp) {
  const int b = blockIdx.y, tid = threadIdx.x;
  FrameState *fs = kp.state + b;
  if (fs->status != D2PC_FRAME_NEEDS_FALLBACK) return;
  const float *frame = kp.depth + (size_t)b * kp.g.D;
  const uint32_t n = kp.g.P;
  const uint32_t tile_base = blockIdx.x * (uint32_t)kScanTile;
  uint32_t nf = 0, nan = 0;
  for (int j = 0; j < kScanPerThread; ++j) {
    uint32_t p = tile_base + (uint32_t)(j * kScanThreads + tid);
    if (p >= n) break;
    float v = depth_at<NATIVE>(frame, kp, p);
    if (!is_finite_f32(v)) { nf++; if (is_nan_f32(v)) nan++; }
  }
  nf = warp_sum(nf);
  nan = warp_sum(nan);
  if ((tid & 31) == 0) {
    if (nf) atomicAdd(&fs->n_nonfinite, nf);
    if (nan) atomicAdd(&fs->n_nan, nan);
  }
}

__global__ void fb_begin_kernel(KParams kp) {
  const int b = blockIdx.x;
  FrameState *fs = kp.state + b;
  uint32_t *hist = kp.fb_hist + (size_t)b * kFbTargets * 256;
  for (int i = threadIdx.x; i < kFbTargets * 256; i += blockDim.x) hist[i] = 0;
  if (threadIdx.x != 0) return;
  fs->fb_active = 0;
  fs->fb_any_nan = 0;
  if (fs->status != D2PC_FRAME_NEEDS_FALLBACK) return;
  if (fs->n_nonfinite == 0) return;  // no repair needed: median stage idle
  const uint32_t m = kp.g.P - fs->n_nan;  // np.nanmedian drops NaN, keeps +-inf
  if (m == 0) return;
  for (int t = 0; t < kFbTargets; ++t) fs->fb_prefix[t] = 0;
  if (m & 1u) {
    fs->fb_rank[0] = m / 2;
    fs->fb_active = 1;
  } else {
    fs->fb_rank[0] = m / 2 - 1;
    fs->fb_rank[1] = m / 2;
    fs->fb_active = 2;
  }
}

template <bool NATIVE>
__global__ void __launch_bounds__(kScanThreads) fb_hist_kernel(KParams kp, int stage, int pass) {
  __shared__ uint32_t sh[kFbTargets][256];
  const int b = blockIdx.y, tid = threadIdx.x;
  FrameState *fs = kp.state + b;
  if (fs->status != D2PC_FRAME_NEEDS_FALLBACK) return;
  const int T = (int)fs->fb_active;
  if (T == 0) return;
  const float *frame = kp.depth + (size_t)b * kp.g.D;
  const uint32_t n = kp.g.P;
  for (int i = tid; i < kFbTargets * 256; i += kScanThreads) (&sh[0][0])[i] = 0;
  uint32_t prefix[kFbTargets];
  for (int t = 0; t < kFbTargets; ++t) prefix[t] = fs->fb_prefix[t];
  const float med = fs->norm.median;
  __syncthreads();
  const int shift = 24 - 8 * pass;
  const uint32_t tile_base = blockIdx.x * (uint32_t)kScanTile;
  uint32_t kmin = 0xFFFFFFFFu, kmax = 0u;
  for (int j = 0; j < kScanPerThread; ++j) {
    uint32_t p = tile_base + (uint32_t)(j * kScanThreads + tid);
    if (p >= n) break;
    float v = depth_at<NATIVE>(frame, kp, p);
    if (stage == kStageMedian) {
      if (is_nan_f32(v)) continue;
    } else {
      if (!is_finite_f32(v)) v = med;  // np.where(finite, d, med)
    }
    uint32_t key = float_to_key(v);
    kmin = min(kmin, key);
    kmax = max(kmax, key);
    for (int t = 0; t < T; ++t) {
      bool match = (pass == 0) || ((key >> (shift + 8)) == (prefix[t] >> (shift + 8)));
      if (match) atomicAdd(&sh[t][(key >> shift) & 255u], 1u);
    }
  }
  if (stage == kStagePercentile && pass == 0) {  // d.min() / d.max() of the repaired map
    kmin = warp_min(kmin);
    kmax = warp_max(kmax);
    if ((tid & 31) == 0) {
      if (kmin != 0xFFFFFFFFu) atomicMin(&fs->min_key, kmin);
      if (kmax != 0u) atomicMax(&fs->max_key, kmax);
    }
  }
  __syncthreads();
  uint32_t *hist = kp.fb_hist + (size_t)b * kFbTargets * 256;
  for (int i = tid; i < T * 256; i += kScanThreads) {
    uint32_t c = (&sh[0][0])[i];
    if (c) atomicAdd(&hist[i], c);
  }
}

__global__ void fb_pick_kernel(KParams kp, int pass) {
  const int b = blockIdx.x;
  FrameState *fs = kp.state + b;
  uint32_t *hist = kp.fb_hist + (size_t)b * kFbTargets * 256;
  const bool live = fs->status == D2PC_FRAME_NEEDS_FALLBACK && fs->fb_active > 0;
  if (live && threadIdx.x < fs->fb_active) {
    const int t = threadIdx.x;
    const int shift = 24 - 8 * pass;
    uint32_t rank = fs->fb_rank[t], cum = 0;
    int d = 0;
    for (; d < 256; ++d) {
      uint32_t c = hist[t * 256 + d];
      if (rank < cum + c) break;
      cum += c;
    }
    if (d > 255) d = 255;  // cannot happen for a consistent rank; stay in range
    fs->fb_prefix[t] |= (uint32_t)d << shift;
    fs->fb_rank[t] = rank - cum;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kFbTargets * 256; i += blockDim.x) hist[i] = 0;
}

// after the median stage: store the repair value, then arm the percentile stage
__global__ void fb_median_kernel(KParams kp) {
  const int b = blockIdx.x;
  FrameState *fs = kp.state + b;
  if (threadIdx.x != 0 || fs->status != D2PC_FRAME_NEEDS_FALLBACK) return;
  const uint32_t n = kp.g.P;
  float med = 0.0f;
  int has_nf = 0;
  if (fs->n_nonfinite != 0) {
    has_nf = 1;
    const uint32_t m = n - fs->n_nan;
    if (m == 0) {
      med = nan_f32();
    } else {
      const int T = (int)fs->fb_active;
      med = median_from_ranks(key_to_float(fs->fb_prefix[0]), key_to_float(fs->fb_prefix[T - 1]), m);
    }
  }
  fs->norm.median = med;
  fs->norm.has_nonfinite = has_nf;
  const bool any_nan = has_nf && is_nan_f32(med);
  fs->fb_any_nan = any_nan ? 1u : 0u;
  if (any_nan) {
    fs->fb_active = 0;
    return;
  }
  RankPair r2 = percentile_ranks(n, D2PC_Q02), r98 = percentile_ranks(n, D2PC_Q98);
  fs->fb_rank[0] = r2.lo; fs->fb_rank[1] = r2.hi; fs->fb_rank[2] = r98.lo; fs->fb_rank[3] = r98.hi;
  for (int t = 0; t < kFbTargets; ++t) fs->fb_prefix[t] = 0;
  fs->fb_active = 4;
}

__global__ void fb_finalise_kernel(KParams kp) {
  const int b = blockIdx.x;
  FrameState *fs = kp.state + b;
  if (threadIdx.x != 0 || fs->status != D2PC_FRAME_NEEDS_FALLBACK) return;
  const uint32_t n = kp.g.P;
  NormParams np_;
  const float med = fs->norm.median;
  const int has_nf = fs->norm.has_nonfinite;
  if (fs->fb_any_nan) {
    finalise_norm(0.0, 0.0, 0.0f, 0.0f, true, &np_);
  } else {
    RankPair r2 = percentile_ranks(n, D2PC_Q02), r98 = percentile_ranks(n, D2PC_Q98);
    double p2 = lerp_percentile(key_to_float(fs->fb_prefix[0]), key_to_float(fs->fb_prefix[1]), r2.gamma);
    double p98 = lerp_percentile(key_to_float(fs->fb_prefix[2]), key_to_float(fs->fb_prefix[3]), r98.gamma);
    const uint32_t kmin = fs->min_key, kmax = fs->max_key;  // of the repaired map (stage P, pass 0)
    finalise_norm(p2, p98, key_to_float(kmin), key_to_float(kmax), false, &np_);
  }
  np_.median = med;
  np_.has_nonfinite = has_nf;
  finish_norm(&np_);
  fs->norm = np_;
  __threadfence();
  fs->status = D2PC_FRAME_READY;
}

// ------------------------------------------------------------------------------------------
// status / parameter export
// ------------------------------------------------------------------------------------------
__global__ void status_kernel(KParams kp, int32_t *d_status, int32_t *d_any) {
  __shared__ int s_any;
  if (threadIdx.x == 0) s_any = 0;
  __syncthreads();
  for (int b = threadIdx.x; b < kp.batch; b += blockDim.x) {
    int32_t s = kp.state[b].status;
    if (s == D2PC_FRAME_PENDING) {  // an index-ordered kernel gave up waiting (abort flag): the exact path takes over
      s = D2PC_FRAME_NEEDS_FALLBACK;
      kp.state[b].status = s;
    }
    if (d_status) d_status[b] = s;
    if (s == D2PC_FRAME_NEEDS_FALLBACK) s_any = 1;
  }
  __syncthreads();
  if (threadIdx.x == 0 && d_any) *d_any = s_any;
}

__global__ void params_kernel(KParams kp, D2pcFrameParams *out) {
  for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < kp.batch; b += gridDim.x * blockDim.x) {
    const FrameState &fs = kp.state[b];
    D2pcFrameParams o;
    o.p2 = fs.norm.p2; o.p98 = fs.norm.p98; o.den = fs.norm.den; o.inv_den = fs.norm.inv_den;
    o.lo32 = fs.norm.lo32; o.hi32 = fs.norm.hi32; o.den32 = fs.norm.den32; o.median = fs.norm.median;
    o.branch = fs.norm.branch; o.status = fs.status;
    o.n_nonfinite = fs.n_nonfinite; o.n_nan = fs.n_nan;
    o.n_cand[0] = fs.nqueue[0]; o.n_cand[1] = fs.nqueue[1];
    o.reserved[0] = fs.sel_fail; o.reserved[1] = fs.sample_ok;
    out[b] = o;
  }
}

}  // namespace d2pc

using namespace d2pc;

static int check_ws(const D2pcConfig *cfg, const void *ws, size_t ws_bytes) {
  int rc = validate_config(cfg);
  if (rc) return rc;
  if (!ws) return D2PC_ERR_INVALID_ARGUMENT;
  if (ws_bytes < make_layout(*cfg).total) return D2PC_ERR_WORKSPACE_TOO_SMALL;
  if (((uintptr_t)ws & 255u) != 0) return D2PC_ERR_INVALID_ARGUMENT;
  return D2PC_OK;
}

// sample -> scan -> select for the nb frames of a (sliced) parameter block.  The tap tables of a resized
// depth are per call, not per slice: the caller runs taps_launch once.
namespace d2pc {

int taps_launch(const KParams &kp, cudaStream_t st) {
  if (kp.g.native || resize_is_generic(kp.g.h, kp.g.w)) return D2PC_OK;
  taps_kernel<<<(kp.g.W + kp.g.H + 255) / 256, 256, 0, st>>>(kp);
  D2PC_CHECK_LAUNCH();
  return D2PC_OK;
}

int stats_prepare() {
  const size_t sample_smem = (size_t)(4u << kSelBits) * sizeof(uint32_t);  // 64 KB (general selection)
  cudaError_t e = cudaFuncSetAttribute(sample_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sample_smem);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kSelectSmemWords * sizeof(uint32_t)));
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(sample_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sample_smem);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(scan_native_kernel<kTilePerThread>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ScanTileSmem));
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(scan_native_multi_kernel<kTilePerThread, 2, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ScanTileSmem));
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(scan_native_multi_kernel<kTilePerThread, 4, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ScanTileSmem));
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(stats_ordered_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ScanTileSmem));
  if (e != cudaSuccess) return record_cuda_error(e);
  return D2PC_OK;
}

int stats_launch(KParams kp, cudaStream_t st, int phases) {
  const int nb = kp.batch;
  if (!kp.g.native && resize_is_generic(kp.g.h, kp.g.w)) {
    const uint32_t blocks = (kp.g.P + 255u) / 256u;
    if (phases & kStatsSample) {
      resize_generic_kernel<<<dim3(blocks < 148u * 8u ? blocks : 148u * 8u, nb), 256, 0, st>>>(kp);
      D2PC_CHECK_LAUNCH();
    }
    kp = per_pixel_view(kp);   // the statistics read the materialised map
  }
  const size_t sample_smem = (size_t)(4u << kSelBits) * sizeof(uint32_t);
  const int vec_ok = ((kp.g.P & 3u) == 0u) && (((uintptr_t)kp.depth & 15u) == 0u);
  dim3 scan_grid((kp.g.P + kScanTile - 1) / kScanTile, nb);
  if (phases & kStatsSample) {
    if (kp.g.native) sample_kernel<true><<<nb, kSelThreads, sample_smem, st>>>(kp);
    else sample_kernel<false><<<nb, kSelThreads, sample_smem, st>>>(kp);
    D2PC_CHECK_LAUNCH();
  }
  if (!(phases & kStatsScanSelect)) return D2PC_OK;
  // One or two frames (the drop-in call): scan + cooperative selection in one index-ordered launch (49 us instead
  // of 66 us for a 1080p frame).  Batches: the selection's waiting CTAs would hold slots the scan needs (measured
  // 0.45 ms against 0.32 ms per 128 frames), so scan and selection stay two launches there.
  const char *mode = getenv("D2PC_STATS_ORDERED");  // measurement aid: "0" never, "1" always
  const bool ordered = (phases & kStatsScanSelect) == kStatsScanSelect && (mode ? atoi(mode) != 0 : nb <= 2);
  if (kp.g.native && kp.g.P > (uint32_t)kSortCap && ordered) {
    const uint32_t ns = (kp.g.P + kTilePx - 1) / kTilePx;
    uint32_t K = (kp.cand_cap + kStatsSliceCap - 1u) / kStatsSliceCap;
    K = K < 4u ? 4u : (K > 64u ? 64u : K);
    const unsigned long long total = (unsigned long long)(nb + 1) * (2ull * K + ns);
    if (total < 0x7FFFFFFFull) {
      stats_ordered_kernel<<<(unsigned)total, kScanThreads, sizeof(ScanTileSmem), st>>>(kp, vec_ok, ns, K);
      D2PC_CHECK_LAUNCH();
      return D2PC_OK;
    }
  }
  if (!(phases & kStatsScan)) {
  } else if (kp.g.native) {
    // Several tiles per CTA (the next tile's loads fly during a tile's epilogue) once the grid still fills the
    // GPU's 4 x 148 slots at least twice: 128 x 1080p 0.316 -> 0.286 ms, 16 x 4K 0.189 -> 0.175 ms per statistics pass
    const uint32_t nt = (kp.g.P + kTilePx - 1) / kTilePx;
    const char *te = getenv("D2PC_SCAN_TPC");  // measurement aid
    const unsigned long long tiles = (unsigned long long)nt * (unsigned long long)nb;
    const int tpc = te ? atoi(te) : (tiles >= 4ull * 1184ull ? 4 : (tiles >= 2ull * 1184ull ? 2 : 1));
    if (tpc >= 4) scan_native_multi_kernel<kTilePerThread, 4, 4><<<dim3((nt + 3) / 4, nb), kScanThreads, sizeof(ScanTileSmem), st>>>(kp, vec_ok);
    else if (tpc >= 2) scan_native_multi_kernel<kTilePerThread, 2, 4><<<dim3((nt + 1) / 2, nb), kScanThreads, sizeof(ScanTileSmem), st>>>(kp, vec_ok);
    else scan_native_kernel<kTilePerThread><<<dim3(nt, nb), kScanThreads, sizeof(ScanTileSmem), st>>>(kp, vec_ok);
  } else {
    // tiled kernel: needs 16 B-aligned rows of the resized map and every tile's source rows in shared memory
    bool tiled = (kp.g.W & 3) == 0;
    for (int32_t v0 = 0; tiled && v0 < kp.g.H; v0 += kRzRows) {
      const int32_t v1 = (v0 + kRzRows < kp.g.H ? v0 + kRzRows : kp.g.H) - 1;
      const AxisTap t0 = axis_tap(v0, kp.g.scale_y, kp.g.h), t1 = axis_tap(v1, kp.g.scale_y, kp.g.h);
      if (t1.i1 - t0.i0 + 1 > kRzSrcRows) tiled = false;
    }
    if (tiled) {
      dim3 tg((kp.g.W + kRzCols - 1) / kRzCols, (kp.g.H + kRzRows - 1) / kRzRows, nb);
      scan_resized_tiled_kernel<<<tg, kScanThreads, 0, st>>>(kp);
    } else {
      scan_gather_kernel<<<scan_grid, kScanThreads, 0, st>>>(kp, (kp.g.P & 3u) == 0u ? 1 : 0);
    }
  }
  D2PC_CHECK_LAUNCH();
  if (phases & kStatsSelect) {
    // One CTA per bracket is a chain of latencies (1080p: 52 us, 4K: 166 us) that only pays when the batch
    // fills the GPU with such CTAs (2 nb >= ~150).  Smaller batches and 4K frames split each bracket over K
    // CTAs, two launches (measured, statistics of 16 frames: 1080p 0.105 -> 0.085 ms, 4K 0.300 -> 0.193 ms;
    // 128 x 1080p would lose: 0.316 -> 0.36 ms).
    const char *cm = getenv("D2PC_SELECT_COOP");  // measurement aid: "0" never, K > 0 always with K parts
    uint32_t K = 296u / (2u * (uint32_t)nb);
    if (K > 16u) K = 16u;
    if (kp.g.P >= 4u * 1024u * 1024u && K < 4u) K = 4u;
    if (K < 4u) K = 0u;
    if (cm) K = (uint32_t)atoi(cm);
    if (K > 64u) K = 64u;
    if (K >= 2u && kp.g.P > (uint32_t)kSortCap) {
      select_a_kernel<<<dim3(K, 2, nb), kCoopThreads, kCoopSmemWords * sizeof(uint32_t), st>>>(kp, K);
      D2PC_CHECK_LAUNCH();
      select_c_kernel<<<dim3(K, 2, nb), kCoopThreads, kCoopSmemWords * sizeof(uint32_t), st>>>(kp, K);
    } else {
      select_kernel<<<dim3(2, nb), kSelectThreads, (size_t)kSelectSmemWords * sizeof(uint32_t), st>>>(kp);
    }
    D2PC_CHECK_LAUNCH();
  }
  return D2PC_OK;
}

int status_launch(const KParams &kp, int32_t *d_status, int32_t *d_any, cudaStream_t st) {
  status_kernel<<<1, 256, 0, st>>>(kp, d_status, d_any);
  D2PC_CHECK_LAUNCH();
  return D2PC_OK;
}

int check_workspace(const D2pcConfig *cfg, const void *ws, size_t ws_bytes) { return check_ws(cfg, ws, ws_bytes); }

}  // namespace d2pc

extern "C" int d2pc_stats_enqueue(const D2pcConfig *cfg, const float *d_depth, void *d_workspace,
                                  size_t workspace_bytes, void *stream) {
  int rc = check_ws(cfg, d_workspace, workspace_bytes);
  if (rc) return rc;
  if (!d_depth) return D2PC_ERR_INVALID_ARGUMENT;
  cudaStream_t st = (cudaStream_t)stream;
  KParams kp = make_kparams(*cfg, d_depth, d_workspace);
  if ((rc = stats_prepare()) != D2PC_OK) return rc;
  if ((rc = taps_launch(kp, st)) != D2PC_OK) return rc;
  return stats_launch(kp, st, kStatsSample | kStatsScanSelect);
}

extern "C" int d2pc_stats_fallback_enqueue(const D2pcConfig *cfg, const float *d_depth,
                                           void *d_workspace, size_t workspace_bytes, void *stream) {
  int rc = check_ws(cfg, d_workspace, workspace_bytes);
  if (rc) return rc;
  if (!d_depth) return D2PC_ERR_INVALID_ARGUMENT;
  cudaStream_t st = (cudaStream_t)stream;
  // the scan has already materialised a resized map, so the fallback always reads a per-pixel map
  KParams kp = per_pixel_view(make_kparams(*cfg, d_depth, d_workspace));
  dim3 grid((kp.g.P + kScanTile - 1) / kScanTile, cfg->batch);
  fb_reset_kernel<<<cfg->batch, 32, 0, st>>>(kp);
  D2PC_CHECK_LAUNCH();
  fb_count_kernel<true><<<grid, kScanThreads, 0, st>>>(kp);
  D2PC_CHECK_LAUNCH();
  fb_begin_kernel<<<cfg->batch, 256, 0, st>>>(kp);
  D2PC_CHECK_LAUNCH();
  for (int stage = 0; stage < 2; ++stage) {
    for (int pass = 0; pass < 4; ++pass) {
      fb_hist_kernel<true><<<grid, kScanThreads, 0, st>>>(kp, stage, pass);
      D2PC_CHECK_LAUNCH();
      fb_pick_kernel<<<cfg->batch, 256, 0, st>>>(kp, pass);
      D2PC_CHECK_LAUNCH();
    }
    if (stage == 0) fb_median_kernel<<<cfg->batch, 32, 0, st>>>(kp);
    else fb_finalise_kernel<<<cfg->batch, 32, 0, st>>>(kp);
    D2PC_CHECK_LAUNCH();
  }
  return D2PC_OK;
}

extern "C" int d2pc_frame_status(const D2pcConfig *cfg, const void *d_workspace, int32_t *d_status,
                                 int32_t *d_any_fallback, void *stream) {
  int rc = validate_config(cfg);
  if (rc) return rc;
  if (!d_workspace) return D2PC_ERR_INVALID_ARGUMENT;
  KParams kp = make_kparams(*cfg, nullptr, (void *)d_workspace);
  return status_launch(kp, d_status, d_any_fallback, (cudaStream_t)stream);
}

extern "C" int d2pc_frame_params(const D2pcConfig *cfg, const void *d_workspace,
                                 D2pcFrameParams *d_params, void *stream) {
  int rc = validate_config(cfg);
  if (rc) return rc;
  if (!d_workspace || !d_params) return D2PC_ERR_INVALID_ARGUMENT;
  KParams kp = make_kparams(*cfg, nullptr, (void *)d_workspace);
  params_kernel<<<(cfg->batch + 127) / 128, 128, 0, (cudaStream_t)stream>>>(kp, d_params);
  D2PC_CHECK_LAUNCH();
  return D2PC_OK;
}
