// d2pc_emit.cu -- reference steps a4..a11 (backend/app.py:200-246) plus the ax-1 extension,
// fused into one pass per output point:
//   (resize) -> non-finite repair -> clip/normalise -> invert -> scale -> back-projection ->
//   BGR->RGB gather -> (depth-range mask -> ordered compaction) -> AoS float32 packing.
//
// Memory plan (HBM-bound by design: 4 B depth + 3 B BGR in, 24 B out per point):
//   * emit_fast_kernel<STEP, MASK, BOUNDS>: 3-channel image, any density (STEP 1 / 2 / 4), with or without the
//     depth-range mask.  Each thread owns 4 consecutive output rows of one image row: vector loads of the sampled
//     sectors, FP64 chain in registers, 12-byte AoS records staged through shared memory and handed to the TMA unit
//     as one bulk copy per array and CTA (cp.async.bulk.global.shared::cta).  MASK: the tile's first output row
//     comes from mask_count_kernel + mask_offsets_kernel, the rows are compacted inside the tile.
//   * emit_generic_kernel: BGRA / grey images, widths that do not fit the vector tiling, smooth=True.  Same
//     staging; the compacted rows of a tile go to `prefix` rows found by a decoupled look-back over the frame's
//     tiles.  Either way the output keeps raster order (the reference's preview stride points[::stride],
//     app.py:498-500, depends on it).
#include <stdlib.h>

#include "d2pc_emit_dev.cuh"

namespace d2pc {


__global__ void __launch_bounds__(32) mask_prepare_kernel(KParams kp, EmitArgs ea, int pc_simple) {
  FrameState *fs = kp.state + blockIdx.x;
  if (fs->status != D2PC_FRAME_READY) return;
  const bool by_depth = ea.use_z && fs->norm.simple && pc_simple;  // uniform over the warp
  float lo_f = 1.0f, hi_f = 0.0f;  // empty interval
  if (by_depth) {
    const NormParams sn = fs->norm;
    const uint32_t kmin = float_to_key(-3.402823466e38f), kmax = float_to_key(3.402823466e38f);
    const bool dec = ea.pc.invert != 0;
    // z32(d) is monotone in d (see mask_interval): the kept set is the key interval [a, b]
    //   a = first key whose z passes the bound that turns TRUE as d grows
    //   b = last key whose z passes the bound that turns FALSE as d grows = first failing key - 1
    auto rising = [&](uint32_t k) {
      const float z = simple_z32(key_to_float(k), sn, ea.pc);
      return dec ? (z <= ea.z_max) : (z >= ea.z_min);
    };
    auto fallen = [&](uint32_t k) {
      const float z = simple_z32(key_to_float(k), sn, ea.pc);
      return !(dec ? (z >= ea.z_min) : (z <= ea.z_max));
    };
    if (rising(kmax) && !fallen(kmin)) {
      const uint32_t a = warp_first_true(kmin, kmax, rising);
      // first key in [kmin, kmax] that has fallen, or kmax + 1 if none
      const uint32_t f = fallen(kmax) ? warp_first_true(kmin, kmax, fallen) : kmax + 1u;
      if (f > a) { lo_f = key_to_float(a); hi_f = key_to_float(f - 1u); }
    }
  }
  if (threadIdx.x == 0) {
    fs->mask_mode = by_depth ? 1 : 0;
    fs->mask_lo = lo_f;
    fs->mask_hi = hi_f;
  }
}

// kept points per emit tile (same tiling and predicate as emit_fast_kernel<MASK>): reads the
// per-pixel depth map only (4 B/px).  One warp per 1024-pixel tile: 8 independent 16 B loads per
// lane, a warp reduction, no shared memory, no barrier.  Mode-0 frames evaluate the exact chain.
constexpr int kCountWarps = 8;
template <int STEP>
__global__ void __launch_bounds__(kCountWarps * 32, 6) mask_count_kernel(KParams kp, EmitArgs ea, uint32_t tiles_per_frame,
                                                                     uint32_t total_tiles, unsigned long long magic_w) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t t = blockIdx.x * (uint32_t)kCountWarps + (uint32_t)warp;
  if (t >= total_tiles) return;
  const uint32_t b = t / tiles_per_frame, tile = t - b * tiles_per_frame;
  const FrameState *fs = kp.state + b;
  const uint32_t N = kp.g.N, W = (uint32_t)kp.g.W, NU = (uint32_t)kp.g.nu;
  const uint32_t tile_base = tile * (uint32_t)kEmitTile;
  const float *frame = kp.depth + (size_t)b * kp.g.P;
  const int32_t status = fs->status;   // looked at after the tile's loads have been issued
  // the four sampled depths of row group j (same pixels as emit_fast_tile)
  auto load_group = [&](int j, bool *ok) -> float4 {
    const uint32_t p0 = tile_base + 4u * (uint32_t)(j * 32 + lane);
    *ok = p0 < N;  // N % 4 == 0 on this path
    if (!*ok) return make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    if (STEP == 1) return ldg_stream_f4(frame + p0);
    const uint32_t jv = (uint32_t)(((unsigned long long)p0 * magic_w) >> 40), ju = p0 - jv * NU;
    const float *src = frame + (size_t)(jv * (uint32_t)STEP) * W + ju * (uint32_t)STEP;
    if (STEP == 2) {
      const float4 da = ldg_stream_f4(src), db = ldg_stream_f4(src + 4);
      return make_float4(da.x, da.z, db.x, db.z);
    }
    return make_float4(__ldg(src), __ldg(src + 4), __ldg(src + 8), __ldg(src + 12));
  };
  float4 r[8];
  bool ok[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) r[j] = load_group(j, &ok[j]);
  uint32_t cnt = 0;
  const MaskParams mp = load_mask(fs);
  if (status != D2PC_FRAME_READY) return;   // uniform per warp
  const bool by_depth = !ea.use_z || mp.mode == 1;
  if (by_depth) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (ok[j]) {
        cnt += mask_keep(r[j].x, 0.0f, mp, ea) ? 1u : 0u;
        cnt += mask_keep(r[j].y, 0.0f, mp, ea) ? 1u : 0u;
        cnt += mask_keep(r[j].z, 0.0f, mp, ea) ? 1u : 0u;
        cnt += mask_keep(r[j].w, 0.0f, mp, ea) ? 1u : 0u;
      }
  } else {  // rare: a rolled loop that loads again (L1 / L2 hits) -- indexing r[] dynamically would put it in local memory
    const NormParams np_ = fs->norm;
#pragma unroll 1
    for (int j = 0; j < 8; ++j) {
      bool okj;
      const float4 q = load_group(j, &okj);
      if (!okj) continue;
      const float raw[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {  // same expression as back_project's z
        const double n = normalised_depth(raw[k], np_, ea.pc.invert);
        cnt += mask_keep(raw[k], (float)(n * ea.pc.scale), mp, ea) ? 1u : 0u;
      }
    }
  }
  cnt = warp_sum(cnt);
  if (lane == 0) kp.tile_state[(size_t)b * tiles_per_frame + tile] = (unsigned long long)cnt;
}

// per frame: exclusive scan of the tile counts -> high word of tile_state; total -> count[b]
__global__ void __launch_bounds__(1024) mask_offsets_kernel(KParams kp, uint32_t *count) {
  __shared__ uint32_t s_warp[32];
  __shared__ uint32_t s_carry;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (kp.state[b].status != D2PC_FRAME_READY) return;
  unsigned long long *ts = kp.tile_state + (size_t)b * kp.emit_tiles;
  if (tid == 0) s_carry = 0;
  __syncthreads();
  for (uint32_t base = 0; base < kp.emit_tiles; base += 1024u) {
    const uint32_t i = base + (uint32_t)tid;
    const uint32_t c = i < kp.emit_tiles ? (uint32_t)ts[i] : 0u;
    uint32_t incl = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      uint32_t y = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += y;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t woff = 0, total = 0;
    for (int w = 0; w < 32; ++w) { const uint32_t x = s_warp[w]; if (w < warp) woff += x; total += x; }
    const uint32_t excl = s_carry + woff + incl - c;
    if (i < kp.emit_tiles) ts[i] = ((unsigned long long)excl << 32) | (unsigned long long)c;
    __syncthreads();
    if (tid == 0) s_carry += total;
    __syncthreads();
  }
  if (tid == 0) count[b] = s_carry;
}

template <int STEP, bool MASK, int MIN_BLOCKS>
static int launch_emit_fast(const KParams &kp, const EmitArgs &ea, const FastArgs &fa, cudaStream_t st) {
  const size_t smem = kEmitStageBytes;
  if (ea.want_bounds) emit_fast_kernel<STEP, MASK, true, 5><<<fa.total_tiles, kEmitThreads, smem, st>>>(kp, ea, fa);
  else emit_fast_kernel<STEP, MASK, false, MIN_BLOCKS><<<fa.total_tiles, kEmitThreads, smem, st>>>(kp, ea, fa);
  return D2PC_OK;
}

// ------------------------------------------------------------------------------------------
// generic path (any stride, BGRA / grey images, widths that are not a multiple of 4)
// ------------------------------------------------------------------------------------------
// a6: smoothing scratch (two float64 maps per frame) and the separable filter's column pass
struct SmoothArgs {
  const double *rows;  // [batch][P] row-filtered normalised map
  int32_t ksize;
  double k[D2PC_MAX_SMOOTH_KSIZE];
};

__device__ __forceinline__ int32_t reflect101(int32_t i, int32_t n) {
  if (n == 1) return 0;
  while (i < 0 || i >= n) {
    if (i < 0) i = -i;
    if (i >= n) i = 2 * (n - 1) - i;
  }
  return i;
}

// column pass of cv2's float64 separable filter: kc*x0 + sum_j k(c+j) * (x(+j) + x(-j)), no FMA
__device__ __forceinline__ double smooth_column(const SmoothArgs &sa, const double *rows, int32_t u, int32_t v,
                                                int32_t W, int32_t H) {
  const int32_t r = sa.ksize >> 1;
  double s = sa.k[r] * rows[(size_t)v * W + u];
  for (int32_t j = 1; j <= r; ++j) {
    const double a = rows[(size_t)reflect101(v + j, H) * W + u];
    const double b = rows[(size_t)reflect101(v - j, H) * W + u];
    s = s + sa.k[r + j] * (a + b);
  }
  return s;
}

// normalised (and inverted) map of every pixel as float64: the input of the blur (app.py:205-212)
__global__ void __launch_bounds__(256) smooth_norm_kernel(KParams kp, int32_t invert, double *out) {
  const int b = blockIdx.y;
  const FrameState *fs = kp.state + b;
  if (fs->status != D2PC_FRAME_READY) return;
  const NormParams np_ = fs->norm;
  const float *frame = kp.depth + (size_t)b * kp.g.D;
  for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < kp.g.P; p += gridDim.x * blockDim.x)
    out[(size_t)b * kp.g.P + p] = normalised_depth(frame[p], np_, invert);
}

// row pass: vector body (x < W - W%4) accumulates with FMA, the scalar tail with multiply + add
__global__ void __launch_bounds__(256) smooth_rows_kernel(KParams kp, SmoothArgs sa, const double *in, double *out) {
  const int b = blockIdx.y;
  if (kp.state[b].status != D2PC_FRAME_READY) return;
  const int32_t W = kp.g.W, r = sa.ksize >> 1, wb = W - (W & 3);
  const double *src = in + (size_t)b * kp.g.P;
  for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < kp.g.P; p += gridDim.x * blockDim.x) {
    const int32_t v = (int32_t)(p / (uint32_t)W), u = (int32_t)(p - (uint32_t)v * (uint32_t)W);
    const double *row = src + (size_t)v * W;
    double s = sa.k[0] * row[reflect101(u - r, W)];
    if (u < wb) {
      for (int32_t j = 1; j < sa.ksize; ++j) s = fma(sa.k[j], row[reflect101(u + j - r, W)], s);
    } else {
      for (int32_t j = 1; j < sa.ksize; ++j) s = s + sa.k[j] * row[reflect101(u + j - r, W)];
    }
    out[(size_t)b * kp.g.P + p] = s;
  }
}

template <bool NATIVE, bool MASK, bool SMOOTH>
__global__ void __launch_bounds__(kEmitThreads) emit_generic_kernel(KParams kp, EmitArgs ea, SmoothArgs sa) {
  __shared__ __align__(16) float s_xyz[kEmitTile * 3 + 4];
  __shared__ __align__(16) float s_rgb[kEmitTile * 3 + 4];
  __shared__ uint32_t s_warp[kEmitThreads / 32];
  __shared__ uint32_t s_prefix;
  __shared__ uint32_t s_b[6][kEmitThreads / 32];
  const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tile = blockIdx.x;
  FrameState *fs = kp.state + b;
  if (fs->status != D2PC_FRAME_READY) return;
  const NormParams np_ = fs->norm;
  const MaskParams mp = load_mask(fs);
  const Geom &g = kp.g;
  const float *frame = kp.depth + (size_t)b * g.D;
  const uint32_t tile_base = (uint32_t)tile * (uint32_t)kEmitTile;
  const uint32_t i0 = tile_base + (uint32_t)kEmitPerThread * (uint32_t)tid;
  float o[12], col[12];
  bool keep[4];
  uint32_t mn[3] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu}, mx[3] = {0u, 0u, 0u};
  uint32_t my_cnt = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const uint32_t i = i0 + k;
    keep[k] = false;
    if (i >= g.N) continue;
    const uint32_t jv = i / (uint32_t)g.nu, ju = i - jv * (uint32_t)g.nu;
    const uint32_t u = ju * (uint32_t)g.step, v = jv * (uint32_t)g.step;
    const uint32_t p = v * (uint32_t)g.W + u;
    const float raw = depth_at<NATIVE>(frame, kp, p);
    const double n = SMOOTH ? smooth_column(sa, sa.rows + (size_t)b * g.P, (int32_t)u, (int32_t)v, g.W, g.H)
                            : normalised_depth(raw, np_, ea.pc.invert);
    back_project(n, (int32_t)u, (int32_t)v, ea.pc, &o[3 * k], &o[3 * k + 1], &o[3 * k + 2]);
    if (g.C >= 3) {
      const uint8_t *cp = ea.bgr + ((size_t)b * g.P + p) * (size_t)g.C;
      col[3 * k + 0] = (float)cp[2];
      col[3 * k + 1] = (float)cp[1];
      col[3 * k + 2] = (float)cp[0];
    } else {
      col[3 * k + 0] = col[3 * k + 1] = col[3 * k + 2] = 128.0f;
    }
    bool kk = true;
    if (MASK) kk = mask_keep(raw, o[3 * k + 2], mp, ea);
    keep[k] = kk;
    if (kk) {
      my_cnt++;
      if (ea.want_bounds) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          uint32_t key = float_to_key(o[3 * k + c]);
          mn[c] = min(mn[c], key); mx[c] = max(mx[c], key);
        }
      }
    }
  }

  // local offsets: exclusive scan of my_cnt over the CTA
  uint32_t incl = my_cnt;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t y = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += y;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  uint32_t warp_off = 0, total = 0;
#pragma unroll
  for (int w = 0; w < kEmitThreads / 32; ++w) {
    uint32_t c = s_warp[w];
    if (w < warp) warp_off += c;
    total += c;
  }
  uint32_t local = warp_off + incl - my_cnt;

  // destination row of this tile
  uint32_t dest_row;
  if (MASK) {
    volatile unsigned long long *ts = kp.tile_state + (size_t)b * kp.emit_tiles;
    if (warp == 0) {
      if (lane == 0 && tile > 0) ts[tile] = (1ull << 32) | (unsigned long long)total;
      uint32_t ex = (tile == 0) ? 0u : lookback_exclusive(ts, tile);
      if (lane == 0) {
        ts[tile] = (2ull << 32) | (unsigned long long)(ex + total);
        s_prefix = ex;
        if (tile == (int)kp.emit_tiles - 1) ea.count[b] = ex + total;
      }
    }
    __syncthreads();
    dest_row = s_prefix;
  } else {
    dest_row = tile_base;
    if (tile == 0 && tid == 0) ea.count[b] = g.N;
  }

  const size_t g0 = ((size_t)((uint32_t)b + ea.frame0) * g.N + dest_row) * 3;
  const uint32_t s_off = (uint32_t)(g0 & 3);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (!keep[k]) continue;
    const uint32_t w = s_off + 3u * local;
    s_xyz[w] = o[3 * k]; s_xyz[w + 1] = o[3 * k + 1]; s_xyz[w + 2] = o[3 * k + 2];
    s_rgb[w] = col[3 * k]; s_rgb[w + 1] = col[3 * k + 1]; s_rgb[w + 2] = col[3 * k + 2];
    local++;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  const uint64_t pol_o = l2_policy((kp.hints & kHintStreamFirst) != 0, false);
  copy_out(s_xyz, s_off, total * 3u, ea.xyz, g0, pol_o);
  copy_out(s_rgb, s_off, total * 3u, ea.rgb, g0, pol_o);
  if (ea.want_bounds) reduce_bounds(fs, mn, mx, s_b);
}

__global__ void emit_init_kernel(KParams kp, int clear_tiles) {
  const int b = blockIdx.x;
  FrameState *fs = kp.state + b;
  if (threadIdx.x == 0) {
    fs->emit_count = 0;
    for (int c = 0; c < 3; ++c) { fs->bounds_min[c] = 0xFFFFFFFFu; fs->bounds_max[c] = 0u; }
  }
  if (clear_tiles) {
    unsigned long long *ts = kp.tile_state + (size_t)b * kp.emit_tiles;
    for (uint32_t i = threadIdx.x; i < kp.emit_tiles; i += blockDim.x) ts[i] = 0ull;
  }
}

// bounds keys -> floats; an empty frame (no kept point) reports NaN bounds
__global__ void bounds_export_kernel(KParams kp, float *d_bounds) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= kp.batch * 6) return;
  const int b = i / 6, c = i % 6;
  const FrameState &fs = kp.state[b];
  const bool empty = fs.bounds_min[0] == 0xFFFFFFFFu && fs.bounds_max[0] == 0u;
  uint32_t key = c < 3 ? fs.bounds_min[c] : fs.bounds_max[c - 3];
  d_bounds[i] = empty ? nan_f32() : key_to_float(key);
}

}  // namespace d2pc

using namespace d2pc;

namespace d2pc {

EmitArgs make_emit_args(const D2pcConfig &cfg, const uint8_t *d_bgr, float *d_xyz, float *d_rgb, uint32_t *d_count) {
  EmitArgs ea;
  ea.bgr = d_bgr; ea.xyz = d_xyz; ea.rgb = d_rgb; ea.count = d_count;
  ea.pc.scale = cfg.depth_scale; ea.pc.cx = cfg.cx; ea.pc.cy = cfg.cy; ea.pc.f = cfg.f;
  ea.pc.inv_f = 1.0 / cfg.f; ea.pc.invert = cfg.invert;
  ea.use_z = cfg.use_z_range; ea.drop_nf = cfg.drop_nonfinite; ea.want_bounds = cfg.want_bounds;
  ea.z_min = cfg.z_min; ea.z_max = cfg.z_max;
  ea.frame0 = 0;
  return ea;
}

// frames [b0, b0 + nb) of the output side of a call (the parameter block is sliced by slice_kparams)
EmitArgs slice_emit_args(const EmitArgs &ea, const Geom &g, int b0) {
  EmitArgs s = ea;
  if (ea.bgr) s.bgr = ea.bgr + (size_t)b0 * g.P * (size_t)g.C;
  s.frame0 = ea.frame0 + (uint32_t)b0;  // xyz / rgb keep their 16-byte aligned base: rows are addressed through frame0
  s.count = ea.count + b0;
  return s;
}

// Emission of the kp.batch frames of a (sliced) parameter block; kp is the block as make_kparams /
// slice_kparams produced it (the per-pixel view is applied here).
int emit_launch(const D2pcConfig &cfg, const KParams &kp_in, const EmitArgs &ea, float *d_bounds, cudaStream_t st,
                int32_t smooth_k, const double *h_kernel, void *d_scratch) {
  const bool smooth = smooth_k > 0;
  // emit always reads a per-pixel map: the input, or the resized map the scan materialised
  const KParams kp = per_pixel_view(kp_in);
  const int nb = kp.batch;
  const bool mask = cfg.use_z_range || cfg.drop_nonfinite;
  dim3 grid(kp.emit_tiles, nb);
  SmoothArgs sa;
  sa.rows = nullptr;
  sa.ksize = 0;
  if (smooth) {
    double *n64 = (double *)d_scratch, *rows = n64 + (size_t)nb * kp.g.P;
    sa.rows = rows;
    sa.ksize = smooth_k;
    for (int i = 0; i < smooth_k; ++i) sa.k[i] = h_kernel[i];
    dim3 sg(148 * 4, nb);
    smooth_norm_kernel<<<sg, 256, 0, st>>>(kp, cfg.invert, n64);
    D2PC_CHECK_LAUNCH();
    smooth_rows_kernel<<<sg, 256, 0, st>>>(kp, sa, n64, rows);
    D2PC_CHECK_LAUNCH();
  }
  // fast path: 3-channel image, 4 consecutive output rows per thread in one image row, vector loads.
  // stride 1: W % 4 == 0; stride 2: W % 8 == 0; stride 4: W % 16 == 0
  const bool fast = !smooth && cfg.img_c == 3 && (cfg.img_w % (4 * cfg.step)) == 0 &&
                    (((uintptr_t)kp.depth & 15u) == 0u) && (((uintptr_t)ea.bgr & 3u) == 0u) && (kp.g.N & 3u) == 0u &&
                    ((unsigned long long)kp.g.N * (unsigned long long)kp.g.nu < (1ull << 40));
  if (mask || cfg.want_bounds) {
    emit_init_kernel<<<nb, 256, 0, st>>>(kp, (mask && !fast) ? 1 : 0);
    D2PC_CHECK_LAUNCH();
  }
  if (mask) {
    // a smoothed z is not a function of the pixel's own depth: exact z compare (mode 0) in that case
    mask_prepare_kernel<<<nb, 32, 0, st>>>(kp, ea, (!smooth && consts_simple(ea.pc)) ? 1 : 0);
    D2PC_CHECK_LAUNCH();
  }
  if (fast) {
    FastArgs fa;
    fa.tiles_per_frame = kp.emit_tiles;
    fa.total_tiles = kp.emit_tiles * (uint32_t)nb;
    fa.batch = (uint32_t)nb;
    fa.magic_w = ((1ull << 40) + (unsigned long long)kp.g.nu - 1ull) / (unsigned long long)kp.g.nu;
    fa.pc_simple = consts_simple(ea.pc) ? 1 : 0;
    if (mask) {
      const uint32_t total_tiles = fa.total_tiles, cg = (total_tiles + kCountWarps - 1) / kCountWarps;
      if (cfg.step == 1) mask_count_kernel<1><<<cg, kCountWarps * 32, 0, st>>>(kp, ea, kp.emit_tiles, total_tiles, fa.magic_w);
      else if (cfg.step == 2) mask_count_kernel<2><<<cg, kCountWarps * 32, 0, st>>>(kp, ea, kp.emit_tiles, total_tiles, fa.magic_w);
      else mask_count_kernel<4><<<cg, kCountWarps * 32, 0, st>>>(kp, ea, kp.emit_tiles, total_tiles, fa.magic_w);
      D2PC_CHECK_LAUNCH();
      mask_offsets_kernel<<<nb, 1024, 0, st>>>(kp, ea.count);
      D2PC_CHECK_LAUNCH();
    }
    int rcl;
    if (mask) rcl = cfg.step == 1 ? launch_emit_fast<1, true, 5>(kp, ea, fa, st)
                  : cfg.step == 2 ? launch_emit_fast<2, true, 5>(kp, ea, fa, st)
                                  : launch_emit_fast<4, true, 5>(kp, ea, fa, st);
    else if (cfg.step == 1) rcl = launch_emit_fast<1, false, 6>(kp, ea, fa, st);
    else if (cfg.step == 2) rcl = launch_emit_fast<2, false, 6>(kp, ea, fa, st);
    else rcl = launch_emit_fast<4, false, 6>(kp, ea, fa, st);
    if (rcl) return rcl;
  } else if (smooth) {
    if (mask) emit_generic_kernel<true, true, true><<<grid, kEmitThreads, 0, st>>>(kp, ea, sa);
    else emit_generic_kernel<true, false, true><<<grid, kEmitThreads, 0, st>>>(kp, ea, sa);
  } else {
    if (mask) emit_generic_kernel<true, true, false><<<grid, kEmitThreads, 0, st>>>(kp, ea, sa);
    else emit_generic_kernel<true, false, false><<<grid, kEmitThreads, 0, st>>>(kp, ea, sa);
  }
  D2PC_CHECK_LAUNCH();
  if (cfg.want_bounds) {
    bounds_export_kernel<<<(nb * 6 + 127) / 128, 128, 0, st>>>(kp, d_bounds);
    D2PC_CHECK_LAUNCH();
  }
  return D2PC_OK;
}

int emit_init_launch(const KParams &kp, cudaStream_t st) {
  emit_init_kernel<<<kp.batch, 256, 0, st>>>(kp, 0);
  D2PC_CHECK_LAUNCH();
  return D2PC_OK;
}

int bounds_export_launch(const KParams &kp, float *d_bounds, cudaStream_t st) {
  bounds_export_kernel<<<(kp.batch * 6 + 127) / 128, 128, 0, st>>>(kp, d_bounds);
  D2PC_CHECK_LAUNCH();
  return D2PC_OK;
}

int emit_validate(const D2pcConfig *cfg, const float *d_depth, const uint8_t *d_bgr, void *d_workspace,
                  size_t workspace_bytes, float *d_xyz, float *d_rgb, uint32_t *d_count, float *d_bounds) {
  int rc = validate_config(cfg);
  if (rc) return rc;
  if (!d_workspace || !d_depth || !d_xyz || !d_rgb || !d_count) return D2PC_ERR_INVALID_ARGUMENT;
  if (cfg->img_c >= 3 && !d_bgr) return D2PC_ERR_INVALID_ARGUMENT;
  if (cfg->want_bounds && !d_bounds) return D2PC_ERR_INVALID_ARGUMENT;
  if (workspace_bytes < make_layout(*cfg).total) return D2PC_ERR_WORKSPACE_TOO_SMALL;
  if ((((uintptr_t)d_xyz | (uintptr_t)d_rgb) & 15u) != 0) return D2PC_ERR_INVALID_ARGUMENT;
  return D2PC_OK;
}

}  // namespace d2pc

static int emit_impl(const D2pcConfig *cfg, const float *d_depth, const uint8_t *d_bgr, void *d_workspace,
                     size_t workspace_bytes, float *d_xyz, float *d_rgb, uint32_t *d_count, float *d_bounds,
                     void *stream, int32_t smooth_k, const double *h_kernel, void *d_scratch, size_t scratch_bytes) {
  int rc = emit_validate(cfg, d_depth, d_bgr, d_workspace, workspace_bytes, d_xyz, d_rgb, d_count, d_bounds);
  if (rc) return rc;
  if (smooth_k > 0) {
    if ((smooth_k & 1) == 0 || smooth_k < 3 || smooth_k > D2PC_MAX_SMOOTH_KSIZE) return D2PC_ERR_UNSUPPORTED;
    if (!h_kernel || !d_scratch) return D2PC_ERR_INVALID_ARGUMENT;
    const size_t need = 2 * (size_t)cfg->batch * (size_t)cfg->img_h * (size_t)cfg->img_w * sizeof(double);
    if (scratch_bytes < need) return D2PC_ERR_WORKSPACE_TOO_SMALL;
    if (((uintptr_t)d_scratch & 15u) != 0) return D2PC_ERR_INVALID_ARGUMENT;
  }
  const KParams kp = make_kparams(*cfg, d_depth, d_workspace);
  const EmitArgs ea = make_emit_args(*cfg, d_bgr, d_xyz, d_rgb, d_count);
  return emit_launch(*cfg, kp, ea, d_bounds, (cudaStream_t)stream, smooth_k, h_kernel, d_scratch);
}

// ------------------------------------------------------------------------------------------
// f2 depth preview: normalise -> (d * 255).astype(uint8) -> colour map   (app.py:127-153)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) preview_kernel(KParams kp, int32_t invert, const uint8_t *lut, uint8_t *out) {
  __shared__ uint8_t s_lut[256 * 3];
  for (int i = threadIdx.x; i < 256 * 3; i += blockDim.x) s_lut[i] = lut[i];
  __syncthreads();
  const int b = blockIdx.y;
  const FrameState *fs = kp.state + b;
  if (fs->status != D2PC_FRAME_READY) return;
  const NormParams np_ = fs->norm;
  const float *frame = kp.depth + (size_t)b * kp.g.D;
  uint8_t *dst = out + (size_t)b * kp.g.P * 3;
  for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < kp.g.P; p += gridDim.x * blockDim.x) {
    const double n = normalised_depth(frame[p], np_, invert);
    // float64 map in the percentile branch, float32 otherwise; astype(uint8) truncates
    const uint32_t idx = (np_.branch == D2PC_BRANCH_PCT) ? (uint32_t)(int32_t)(n * 255.0)
                                                        : (uint32_t)(int32_t)((float)n * 255.0f);
    const uint32_t j = (idx & 255u) * 3u;
    dst[3 * (size_t)p + 0] = s_lut[j];
    dst[3 * (size_t)p + 1] = s_lut[j + 1];
    dst[3 * (size_t)p + 2] = s_lut[j + 2];
  }
}

extern "C" int d2pc_preview_enqueue(const D2pcConfig *cfg, const float *d_depth, void *d_workspace,
                                    size_t workspace_bytes, const uint8_t *d_lut_bgr, uint8_t *d_out_bgr,
                                    void *stream) {
  int rc = validate_config(cfg);
  if (rc) return rc;
  if (!d_depth || !d_workspace || !d_lut_bgr || !d_out_bgr) return D2PC_ERR_INVALID_ARGUMENT;
  if (cfg->dep_h != cfg->img_h || cfg->dep_w != cfg->img_w) return D2PC_ERR_INVALID_ARGUMENT;
  if (workspace_bytes < make_layout(*cfg).total) return D2PC_ERR_WORKSPACE_TOO_SMALL;
  KParams kp = make_kparams(*cfg, d_depth, d_workspace);
  preview_kernel<<<dim3(148 * 2, cfg->batch), 256, 0, (cudaStream_t)stream>>>(kp, cfg->invert, d_lut_bgr, d_out_bgr);
  D2PC_CHECK_LAUNCH();
  return D2PC_OK;
}

extern "C" int d2pc_emit_enqueue(const D2pcConfig *cfg, const float *d_depth, const uint8_t *d_bgr,
                                 void *d_workspace, size_t workspace_bytes, float *d_xyz,
                                 float *d_rgb, uint32_t *d_count, float *d_bounds, void *stream) {
  return emit_impl(cfg, d_depth, d_bgr, d_workspace, workspace_bytes, d_xyz, d_rgb, d_count, d_bounds, stream, 0,
                   nullptr, nullptr, 0);
}

extern "C" int d2pc_smooth_scratch_bytes(const D2pcConfig *cfg, size_t *bytes) {
  int rc = validate_config(cfg);
  if (rc) return rc;
  if (!bytes) return D2PC_ERR_INVALID_ARGUMENT;
  *bytes = 2 * (size_t)cfg->batch * (size_t)cfg->img_h * (size_t)cfg->img_w * sizeof(double);
  return D2PC_OK;
}

extern "C" int d2pc_emit_smooth_enqueue(const D2pcConfig *cfg, const float *d_depth, const uint8_t *d_bgr,
                                        void *d_workspace, size_t workspace_bytes, int32_t ksize,
                                        const double *h_kernel, void *d_scratch, size_t scratch_bytes,
                                        float *d_xyz, float *d_rgb, uint32_t *d_count, float *d_bounds,
                                        void *stream) {
  if (ksize <= 0) return D2PC_ERR_INVALID_ARGUMENT;
  return emit_impl(cfg, d_depth, d_bgr, d_workspace, workspace_bytes, d_xyz, d_rgb, d_count, d_bounds, stream, ksize,
                   h_kernel, d_scratch, scratch_bytes);
}
