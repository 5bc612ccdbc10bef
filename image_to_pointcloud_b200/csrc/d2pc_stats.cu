// d2pc_stats.cu -- reference steps a1..a5 (backend/app.py:186-206) on the device:
// exact 2nd/98th-percentile order statistics of every frame's (virtually resized) depth map,
// non-finite repair (np.nanmedian) and the frame's normalisation parameters.
//
// Fast path (3 launches per batch, depth map read from HBM exactly once):
//   sample_kernel  1 CTA/frame   stratified sample (registers) -> exact sample order statistics
//                                by a multi-level bucket histogram -> key brackets [L, U]
//                                that contain the wanted ranks with ~6 sigma margin
//   scan_kernel    streaming     per pixel, branch-free: count values below each bracket and
//                                defer the few values inside a bracket (about 2% each) or
//                                non-finite to the frame's raw queue.  Pure compares, no
//                                histogram; the only atomics are per CTA.
//   select_kernel  2 CTA/frame   classify the queued values against the bracket (equal to a
//                                bound / strictly inside), resolve the wanted ranks, exact
//                                selection inside the bracket (multi-level bucket histogram over
//                                the L2-resident queue); the last CTA of a frame evaluates
//                                NumPy's _lerp in float64 and writes the parameter block.
// Frames the fast path cannot finish *exactly* (non-finite values, bracket miss, queue
// overflow, collapsed percentiles) are only marked; d2pc_stats_fallback_enqueue runs the
// input-agnostic exact path (8-bit radix select, nanmedian repair) for those.
#include "d2pc_device.cuh"

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
__global__ void __launch_bounds__(kSelThreads) sample_kernel(KParams kp) {
  extern __shared__ uint32_t s_hist[];  // 4 << kSelBits words
  __shared__ uint32_t s_bad, s_min, s_max;
  __shared__ uint32_t s_res[2 * 4 + 40];
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
    fs->n_nonfinite = 0; fs->n_nan = 0; fs->nqueue[0] = 0; fs->nqueue[1] = 0;
    fs->min_key = 0xFFFFFFFFu; fs->max_key = 0u;
    for (int i = 0; i < 4; ++i) fs->sel_key[i] = 0;
    fs->sel_fail = 0; fs->sel_done = 0;
    fs->status = D2PC_FRAME_PENDING;
    fs->fb_active = 0; fs->fb_any_nan = 0;
    fs->norm.has_nonfinite = 0;
    fs->norm.median = 0.0f;
    fs->norm.simple = 0;
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
    float v = depth_at<NATIVE>(frame, kp, idx);
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
  block_hist_select<4, kSelBits>([&](auto f) {
#pragma unroll
    for (int e = 0; e < E; ++e) f(k[e]);
  }, s_min, s_max, rank, out, s_hist, s_res);
  if (tid < 2) {
    fs->brL[tid] = open_lo[tid] ? 0u : out[2 * tid + 0];
    fs->brU[tid] = open_hi[tid] ? 0xFFFFFFFFu : out[2 * tid + 1];
  }
  if (tid == 0) fs->sample_ok = s_bad ? 0u : 1u;
}

// ------------------------------------------------------------------------------------------
// scan_kernel
// ------------------------------------------------------------------------------------------
// One streaming pass, ~10 instructions per pixel on the common path: float compares against the
// bracket bounds (as floats; -0.0 == +0.0 here and in the rare path, consistently), per-thread
// "below" counters, and a rare path (about 4% of pixels: inside a bracket, or non-finite) that
// appends the key to the CTA's staging list or bumps an equal-to-bound / non-finite counter.
// min/max are not tracked: a frame whose percentiles collapse (p98 <= p2) goes to the exact
// fallback, which computes them.
template <int QD>
struct ScanSharedT {
  float pqueue[QD][kScanThreads];  // per-thread deferred values, slot-major
  uint32_t qtotal[2], gbase[2];
  uint32_t red[2][kScanThreads / 32];
};
using ScanShared = ScanSharedT<kScanPerThread>;

// Common path, 11 predicated instructions, no branch, no atomic: count "below" for both brackets
// and, when the value is inside a bracket or non-finite (about 4% of pixels), store it in the
// thread's private queue column in shared memory (qaddr = shared-space byte address of the next
// free slot).  NaN: every ordered compare is false; setp.gtu (unordered greater) is true.
__device__ __forceinline__ void scan_value(float v, const float Lf[2], const float Uf[2], uint32_t &b0,
                                           uint32_t &b1, uint32_t &qaddr) {
  asm volatile(
      "{\n\t"
      ".reg .pred lt0, lt1, in0, any;\n\t"
      ".reg .f32 av;\n\t"
      "setp.lt.f32 lt0, %3, %4;\n\t"
      "setp.lt.f32 lt1, %3, %5;\n\t"
      "@lt0 add.u32 %0, %0, 1;\n\t"
      "@lt1 add.u32 %1, %1, 1;\n\t"
      "setp.le.and.f32 in0, %3, %6, !lt0;\n\t"
      "setp.le.and.f32 any, %3, %7, !lt1;\n\t"
      "or.pred any, any, in0;\n\t"
      "abs.f32 av, %3;\n\t"
      "setp.gtu.or.f32 any, av, 0f7F7FFFFF, any;\n\t"
      "@any st.shared.f32 [%2], %3;\n\t"
      "@any add.u32 %2, %2, %8;\n\t"
      "}"
      : "+r"(b0), "+r"(b1), "+r"(qaddr)
      : "f"(v), "f"(Lf[0]), "f"(Lf[1]), "f"(Uf[0]), "f"(Uf[1]), "n"(kScanThreads * 4)
      : "memory");
}

__device__ __forceinline__ void bracket_floats(const FrameState *fs, float Lf[2], float Uf[2]) {
#pragma unroll
  for (int br = 0; br < 2; ++br) {
    const uint32_t L = fs->brL[br], U = fs->brU[br];
    Lf[br] = (L == 0u) ? -__int_as_float(0x7F800000) : key_to_float(L);           // open below
    Uf[br] = (U == 0xFFFFFFFFu) ? __int_as_float(0x7F800000) : key_to_float(U);   // open above
  }
}

// CTA totals -> 4 global atomics; deferred values -> the two brackets' raw queues
// (queue 0: v <= U0 or non-finite; queue 1: v >= L1; a value inside both goes to both)
template <int QD>
__device__ __forceinline__ void scan_epilogue(ScanSharedT<QD> &sh, FrameState *fs, const KParams &kp, int b, uint32_t b0,
                                              uint32_t b1, uint32_t qaddr, uint32_t q0, const float Lf[2],
                                              const float Uf[2]) {
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const uint32_t nq = (qaddr - q0) / (uint32_t)(kScanThreads * 4);
  uint32_t n0 = 0, n1 = 0;
  for (uint32_t j = 0; j < nq; ++j) {
    const float v = sh.pqueue[j][tid];
    n0 += !(v > Uf[0]) ? 1u : 0u;   // also true for NaN
    n1 += (v >= Lf[1]) ? 1u : 0u;   // +inf lands here too; harmless (it is in queue 0 as well)
  }
  const uint32_t pos0 = n0 ? atomicAdd(&sh.qtotal[0], n0) : 0u;
  const uint32_t pos1 = n1 ? atomicAdd(&sh.qtotal[1], n1) : 0u;
  b0 = warp_sum(b0);
  b1 = warp_sum(b1);
  if (lane == 0) { sh.red[0][warp] = b0; sh.red[1][warp] = b1; }
  __syncthreads();
  if (tid < 2) {
    uint32_t r = 0;
    for (int w = 0; w < kScanThreads / 32; ++w) r += sh.red[tid][w];
    if (r) atomicAdd(&fs->below[tid], r);
  } else if (tid < 4) {
    const uint32_t t = sh.qtotal[tid - 2];
    sh.gbase[tid - 2] = t ? atomicAdd(&fs->nqueue[tid - 2], t) : 0u;
  }
  __syncthreads();
  if (nq) {
    float *gq0 = reinterpret_cast<float *>(kp.cand) + (size_t)b * 2 * kp.cand_cap;
    float *gq1 = gq0 + kp.cand_cap;
    uint32_t g0 = sh.gbase[0] + pos0, g1 = sh.gbase[1] + pos1;
    for (uint32_t j = 0; j < nq; ++j) {
      const float v = sh.pqueue[j][tid];
      if (!(v > Uf[0])) { if (g0 < kp.cand_cap) gq0[g0] = v; ++g0; }
      if (v >= Lf[1]) { if (g1 < kp.cand_cap) gq1[g1] = v; ++g1; }
    }
  }
}

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
  if (tid < 2) sh.qtotal[tid] = 0;
  uint32_t b0 = 0, b1 = 0;
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
        for (int k = 0; k < 4; ++k) scan_value(o[k], Lf, Uf, b0, b1, qaddr);
        float *dst = kp.resized + (size_t)b * kp.g.P + (size_t)v * W + uu;
        *reinterpret_cast<float4 *>(dst) = make_float4(o[0], o[1], o[2], o[3]);
      }
    }
  }
  scan_epilogue(sh, fs, kp, b, b0, b1, qaddr, q0, Lf, Uf);
}

template <bool NATIVE>
__global__ void __launch_bounds__(kScanThreads) scan_kernel(KParams kp, int vec_ok) {
  __shared__ ScanShared sh;
  const int b = blockIdx.y;
  const int tid = threadIdx.x;
  FrameState *fs = kp.state + b;
  const float *frame = kp.depth + (size_t)b * kp.g.D;
  const uint32_t n = kp.g.P;
  float Lf[2], Uf[2];
  bracket_floats(fs, Lf, Uf);
  if (tid < 2) sh.qtotal[tid] = 0;
  __syncthreads();
  uint32_t b0 = 0, b1 = 0;
  const uint32_t q0 = (uint32_t)__cvta_generic_to_shared(&sh.pqueue[0][tid]);
  uint32_t qaddr = q0;
  const uint32_t tile_base = blockIdx.x * (uint32_t)kScanTile;

  if (NATIVE) {
    if (vec_ok && tile_base + (uint32_t)kScanTile <= n) {
      // full tile: groups of 16 pixels per thread; the next group's four 16 B loads are issued
      // before the current group is compared.  The rolled loop keeps ptxas from hoisting all 160
      // compares ahead of their uses (which spills predicates into bit-masks).
      constexpr int NG = kScanPerThread / 16;
      const float *src = frame + tile_base + 4u * (uint32_t)tid;
      float4 r[4], nx[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) r[j] = ldg_stream_f4(src + (size_t)j * (4 * kScanThreads));
#pragma unroll 1
      for (int g = 0; g < NG; ++g) {
        if (g + 1 < NG) {
#pragma unroll
          for (int j = 0; j < 4; ++j) nx[j] = ldg_stream_f4(src + (size_t)((g + 1) * 4 + j) * (4 * kScanThreads));
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          scan_value(r[j].x, Lf, Uf, b0, b1, qaddr);
          scan_value(r[j].y, Lf, Uf, b0, b1, qaddr);
          scan_value(r[j].z, Lf, Uf, b0, b1, qaddr);
          scan_value(r[j].w, Lf, Uf, b0, b1, qaddr);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) r[j] = nx[j];
      }
    } else {  // last tile of a frame / unaligned frames
#pragma unroll 1
      for (int j = 0; j < kScanPerThread / 4; ++j) {
        const uint32_t p = tile_base + 4u * (uint32_t)(j * kScanThreads + tid);
        for (uint32_t k = 0; k < 4u; ++k)
          if (p + k < n) scan_value(__ldg(frame + p + k), Lf, Uf, b0, b1, qaddr);
      }
    }
  } else {
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
          scan_value(val[k], Lf, Uf, b0, b1, qaddr);
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
  }

  scan_epilogue(sh, fs, kp, b, b0, b1, qaddr, q0, Lf, Uf);
}

// ------------------------------------------------------------------------------------------
// select_kernel
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t next_pow2(uint32_t x) {
  uint32_t p = 1;
  while (p < x) p <<= 1;
  return p;
}

__device__ void finalise_fast(FrameState *fs, uint32_t n) {
  volatile FrameState *vfs = fs;
  if (vfs->sel_fail) {
    vfs->status = D2PC_FRAME_NEEDS_FALLBACK;
    return;
  }
  RankPair r2 = percentile_ranks(n, D2PC_Q02), r98 = percentile_ranks(n, D2PC_Q98);
  double p2 = lerp_percentile(key_to_float(vfs->sel_key[0]), key_to_float(vfs->sel_key[1]), r2.gamma);
  double p98 = lerp_percentile(key_to_float(vfs->sel_key[2]), key_to_float(vfs->sel_key[3]), r98.gamma);
  if (!(p98 > p2)) {  // degenerate percentiles need min/max: the exact path computes them
    vfs->status = D2PC_FRAME_NEEDS_FALLBACK;
    return;
  }
  NormParams np_;
  finalise_norm(p2, p98, 0.0f, 0.0f, false, &np_);
  np_.median = 0.0f;
  np_.has_nonfinite = 0;
  finish_norm(&np_);
  fs->norm = np_;
  __threadfence();
  vfs->status = D2PC_FRAME_READY;
}

// visit this thread's share of a raw queue, 8 independent loads in flight per thread
template <typename F>
__device__ __forceinline__ void for_each_queued(const float *q, uint32_t nq, F f) {
  const uint32_t stride = blockDim.x;
  uint32_t i = threadIdx.x;
  for (; i + 7u * stride < nq; i += 8u * stride) {
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = __ldg(q + i + (uint32_t)k * stride);
#pragma unroll
    for (int k = 0; k < 8; ++k) f(v[k]);
  }
  for (; i < nq; i += stride) f(__ldg(q + i));
}

__global__ void __launch_bounds__(kSelThreads, 2) select_kernel(KParams kp) {
  extern __shared__ uint32_t s_hist[];  // 2 << kSelBits words
  __shared__ uint32_t s_res[2 * 2 + 40];
  __shared__ uint32_t s_c[5];  // eqL, inside, eqU, non-finite, NaN
  const int br = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  FrameState *fs = kp.state + b;
  const uint32_t n = kp.g.P;
  const uint32_t qcap = kp.cand_cap;
  const float *q = reinterpret_cast<const float *>(kp.cand) + ((size_t)b * 2 + br) * qcap;
  const uint32_t nqueue = fs->nqueue[br];
  const uint32_t nq = min(nqueue, qcap);
  const uint32_t L = fs->brL[br], U = fs->brU[br];
  float Lf2[2], Uf2[2];
  bracket_floats(fs, Lf2, Uf2);
  const float Lf = Lf2[br], Uf = Uf2[br];
  if (tid < 5) s_c[tid] = 0;
  __syncthreads();
  {  // pass 0: classify the queued values against this bracket
    uint32_t c_eqL = 0, c_in = 0, c_eqU = 0, c_nf = 0, c_nan = 0;
    for_each_queued(q, nq, [&](float v) {
      if (!(fabsf(v) < __int_as_float(0x7F800000))) { c_nf++; if (v != v) c_nan++; }
      else if (v >= Lf && v <= Uf) {
        if (v == Lf) c_eqL++;
        else if (v < Uf) c_in++;
        else c_eqU++;
      }
    });
    c_eqL = warp_sum(c_eqL); c_in = warp_sum(c_in); c_eqU = warp_sum(c_eqU);
    c_nf = warp_sum(c_nf); c_nan = warp_sum(c_nan);
    if ((tid & 31) == 0) {
      if (c_eqL) atomicAdd(&s_c[0], c_eqL);
      if (c_in) atomicAdd(&s_c[1], c_in);
      if (c_eqU) atomicAdd(&s_c[2], c_eqU);
      if (c_nf) atomicAdd(&s_c[3], c_nf);
      if (c_nan) atomicAdd(&s_c[4], c_nan);
    }
  }
  __syncthreads();
  const uint32_t below = fs->below[br], eqL = s_c[0], nin = s_c[1], eqU = s_c[2];
  // a non-finite value anywhere in the frame shows up in queue 0; both CTAs must agree to fail,
  // which the shared sel_fail flag takes care of
  bool fail = (s_c[3] != 0u) || (nqueue > qcap) || (fs->sample_ok == 0u) || (kp.force_fallback != 0);
  RankPair rp = percentile_ranks(n, br ? D2PC_Q98 : D2PC_Q02);
  long long need[2] = {-1, -1};
  uint32_t key[2] = {0u, 0u};
  for (int t = 0; t < 2; ++t) {
    uint32_t r = t ? rp.hi : rp.lo;
    if (r < below) { fail = true; continue; }
    uint32_t r1 = r - below;
    if (r1 < eqL) { key[t] = L; continue; }
    uint32_t r2 = r1 - eqL;
    if (r2 < nin) { need[t] = r2; continue; }
    uint32_t r3 = r2 - nin;
    if (r3 < eqU) { key[t] = U; continue; }
    fail = true;
  }
  const bool any_need = need[0] >= 0 || need[1] >= 0;

  if (!fail && any_need) {  // uniform over the CTA
    // strictly inside (L, U) in float order; -0.0 == +0.0 there, so when a bound is a zero the
    // other zero's key can sit one step outside the key interval: widen the key range by one
    const uint32_t lo0 = L > 0u ? L - 1u : 0u;
    const uint32_t hi0 = U < 0xFFFFFFFFu ? U + 1u : U;
    uint32_t rk[2], ok_[2];
    for (int t = 0; t < 2; ++t) rk[t] = (uint32_t)(need[t] >= 0 ? need[t] : need[1 - t]);
    const bool ok = block_hist_select<2, kSelBits>([&](auto f) {
      for_each_queued(q, nq, [&](float v) {
        if (v > Lf && v < Uf) f(float_to_key(v));
      });
    }, lo0, hi0, rk, ok_, s_hist, s_res);
    if (!ok) fail = true;
    for (int t = 0; t < 2; ++t)
      if (need[t] >= 0) key[t] = ok_[t];
  }

  if (tid == 0) {
    fs->sel_key[2 * br + 0] = key[0];
    fs->sel_key[2 * br + 1] = key[1];
    fs->eqL[br] = eqL; fs->inside[br] = nin; fs->eqU[br] = eqU;
    if (br == 0) { fs->n_nonfinite = s_c[3]; fs->n_nan = s_c[4]; }  // non-finite values live in queue 0
    if (fail) atomicOr(&fs->sel_fail, 1u);
    __threadfence();
    uint32_t ticket = atomicAdd(&fs->sel_done, 1u);
    if (ticket == 1u) {
      __threadfence();
      finalise_fast(fs, n);
    }
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

extern "C" int d2pc_stats_enqueue(const D2pcConfig *cfg, const float *d_depth, void *d_workspace,
                                  size_t workspace_bytes, void *stream) {
  int rc = check_ws(cfg, d_workspace, workspace_bytes);
  if (rc) return rc;
  if (!d_depth) return D2PC_ERR_INVALID_ARGUMENT;
  cudaStream_t st = (cudaStream_t)stream;
  KParams kp = make_kparams(*cfg, d_depth, d_workspace);
  if (!kp.g.native && resize_is_generic(kp.g.h, kp.g.w)) {
    const uint32_t blocks = (kp.g.P + 255u) / 256u;
    resize_generic_kernel<<<dim3(blocks < 148u * 8u ? blocks : 148u * 8u, cfg->batch), 256, 0, st>>>(kp);
    D2PC_CHECK_LAUNCH();
    kp = per_pixel_view(kp);   // the statistics read the materialised map
    d_depth = kp.depth;
  }
  const size_t sample_smem = (size_t)(4u << kSelBits) * sizeof(uint32_t);  // 64 KB
  const size_t select_smem = (size_t)(2u << kSelBits) * sizeof(uint32_t);  // 32 KB
  {
    cudaError_t e = kp.g.native
        ? cudaFuncSetAttribute(sample_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sample_smem)
        : cudaFuncSetAttribute(sample_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sample_smem);
    if (e != cudaSuccess) return record_cuda_error(e);
  }
  const int vec_ok = ((kp.g.P & 3u) == 0u) && (((uintptr_t)d_depth & 15u) == 0u);
  dim3 scan_grid((kp.g.P + kScanTile - 1) / kScanTile, cfg->batch);
  const size_t scan_smem_bytes = 0;
  if (!kp.g.native) {
    taps_kernel<<<(kp.g.W + kp.g.H + 255) / 256, 256, 0, st>>>(kp);
    D2PC_CHECK_LAUNCH();
  }
  if (kp.g.native) {
    sample_kernel<true><<<cfg->batch, kSelThreads, sample_smem, st>>>(kp);
    D2PC_CHECK_LAUNCH();
    scan_kernel<true><<<scan_grid, kScanThreads, scan_smem_bytes, st>>>(kp, vec_ok);
  } else {
    sample_kernel<false><<<cfg->batch, kSelThreads, sample_smem, st>>>(kp);
    D2PC_CHECK_LAUNCH();
    // tiled kernel: needs 16 B-aligned rows of the resized map and every tile's source rows in shared memory
    bool tiled = (kp.g.W & 3) == 0;
    for (int32_t v0 = 0; tiled && v0 < kp.g.H; v0 += kRzRows) {
      const int32_t v1 = (v0 + kRzRows < kp.g.H ? v0 + kRzRows : kp.g.H) - 1;
      const AxisTap t0 = axis_tap(v0, kp.g.scale_y, kp.g.h), t1 = axis_tap(v1, kp.g.scale_y, kp.g.h);
      if (t1.i1 - t0.i0 + 1 > kRzSrcRows) tiled = false;
    }
    if (tiled) {
      dim3 tg((kp.g.W + kRzCols - 1) / kRzCols, (kp.g.H + kRzRows - 1) / kRzRows, cfg->batch);
      scan_resized_tiled_kernel<<<tg, kScanThreads, 0, st>>>(kp);
    } else {
      scan_kernel<false><<<scan_grid, kScanThreads, scan_smem_bytes, st>>>(kp, (kp.g.P & 3u) == 0u ? 1 : 0);
    }
  }
  D2PC_CHECK_LAUNCH();
  select_kernel<<<dim3(2, cfg->batch), kSelThreads, select_smem, st>>>(kp);
  D2PC_CHECK_LAUNCH();
  return D2PC_OK;
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
  status_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(kp, d_status, d_any_fallback);
  D2PC_CHECK_LAUNCH();
  return D2PC_OK;
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
