// d2pc_emit_dev.cuh -- device-side pieces of the fused emission (reference steps a4..a11, backend/app.py:200-246)
// shared by the stand-alone emit kernels (d2pc_emit.cu) and the persistent path kernel (d2pc_path.cu).
#ifndef D2PC_EMIT_DEV_CUH_
#define D2PC_EMIT_DEV_CUH_

#include "d2pc_device.cuh"

namespace d2pc {

__device__ __forceinline__ void stage_f4(float *s, float a, float b, float c, float d) {
  *reinterpret_cast<float4 *>(s) = make_float4(a, b, c, d);
}

// per-CTA reduction of kept-point bounds -> 6 global atomics on ordered keys
__device__ __forceinline__ void reduce_bounds(FrameState *fs, uint32_t mn[3], uint32_t mx[3],
                                              uint32_t (*s_b)[kEmitThreads / 32]) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    uint32_t a = warp_min(mn[c]), b = warp_max(mx[c]);
    if (lane == 0) { s_b[c][warp] = a; s_b[3 + c][warp] = b; }
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    const int c = threadIdx.x;
    uint32_t r = s_b[c][0];
    for (int w = 1; w < kEmitThreads / 32; ++w) r = c < 3 ? min(r, s_b[c][w]) : max(r, s_b[c][w]);
    if (c < 3) { if (r != 0xFFFFFFFFu) atomicMin(&fs->bounds_min[c], r); }
    else       { if (r != 0u) atomicMax(&fs->bounds_max[c - 3], r); }
  }
}

// Copy n_f floats from shared staging (whose word 0 corresponds to global float index
// g0 - s_off, i.e. staging is offset so that src and dst share 16 B alignment) to global.
// The 16-byte aligned middle goes out as one TMA bulk copy issued by thread 0; the (at most three)
// floats before and after it are stored by single threads.  The caller has executed
// fence.proxy.async.shared::cta and a __syncthreads() after the last staging write.
__device__ __forceinline__ void copy_out(const float *s, uint32_t s_off, uint32_t n_f, float *gbase,
                                         size_t g0, uint64_t pol) {
  float *galigned = gbase + (g0 - s_off);
  const uint32_t span = s_off + n_f;
  const uint32_t w_first = s_off ? 4u : 0u;      // first float of the aligned middle
  const uint32_t w_last = span & ~3u;            // one past its last float
  const uint32_t tid = threadIdx.x;
  if (w_last > w_first) {
    if (tid == 0) {
      const uint32_t sa = (uint32_t)__cvta_generic_to_shared(s + w_first);
      bulk_store_pol(galigned + w_first, sa, (w_last - w_first) * 4u, pol);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    if (tid >= 32 && tid < 36) {        // head: floats [s_off, min(w_first, span))
      const uint32_t w = s_off + (tid - 32);
      if (w < w_first && w < span) stg_stream_f1(galigned + w, s[w]);
    } else if (tid >= 64 && tid < 68) { // tail: floats [w_last, span)
      const uint32_t w = w_last + (tid - 64);
      if (w < span) stg_stream_f1(galigned + w, s[w]);
    }
  } else {  // fewer than one aligned chunk: a handful of floats
    if (tid < 8) {
      const uint32_t w = s_off + tid;
      if (w < span) stg_stream_f1(galigned + w, s[w]);
    }
  }
}

// ------------------------------------------------------------------------------------------
// fast path: step 1, native depth, 3-channel image, W % 4 == 0, no mask
// ------------------------------------------------------------------------------------------
// u8 -> f32 without the conversion pipe: byte k of w into the mantissa of 2^23, minus 2^23.
__device__ __forceinline__ float byte_to_float(uint32_t w, uint32_t selector) {
  return __uint_as_float(__byte_perm(w, 0x4B000000u, selector)) - 8388608.0f;
}
#define D2PC_B0 0x7440u
#define D2PC_B1 0x7441u
#define D2PC_B2 0x7442u
#define D2PC_B3 0x7443u

// tile_state word: (flag << 32) | value ; flag 0 = not ready, 1 = tile aggregate, 2 = inclusive
__device__ __forceinline__ uint32_t lookback_exclusive(volatile unsigned long long *ts, int tile) {
  const int lane = threadIdx.x & 31;
  uint32_t exclusive = 0;
  int base_idx = tile - 1;
  while (true) {
    const int idx = base_idx - lane;
    unsigned long long st = (idx >= 0) ? ts[idx] : ((2ull << 32) | 0ull);
    const uint32_t flag = (uint32_t)(st >> 32);
    const unsigned inval = __ballot_sync(0xffffffffu, flag == 0u);
    const unsigned incl = __ballot_sync(0xffffffffu, flag == 2u);
    const int first_incl = incl ? (__ffs(incl) - 1) : 32;
    const unsigned need = first_incl >= 31 ? 0xffffffffu : ((2u << first_incl) - 1u);
    if (inval & need) continue;  // a needed predecessor has not published yet: re-read
    uint32_t val = (lane <= first_incl) ? (uint32_t)st : 0u;
    exclusive += warp_sum(val);
    if (first_incl < 32) break;
    base_idx -= 32;
  }
  return exclusive;
}

// ax-1 predicate.  Frames whose z is provably monotone in the raw depth (mask_mode 1) compare the
// raw value against the frame's depth-space interval; all other frames compare the emitted z.
struct MaskParams {
  int32_t mode;
  float lo, hi;
};
__device__ __forceinline__ MaskParams load_mask(const FrameState *fs) {
  MaskParams m;
  m.mode = fs->mask_mode; m.lo = fs->mask_lo; m.hi = fs->mask_hi;
  return m;
}
__device__ __forceinline__ bool mask_keep_v(float raw, float z32, const MaskParams &m, int32_t use_z, int32_t drop_nf,
                                            float z_min, float z_max) {
  bool kk = true;
  if (use_z) {
    if (m.mode == 1) kk = (raw >= m.lo) && (raw <= m.hi);
    else kk = (z32 >= z_min) && (z32 <= z_max);
  }
  if (drop_nf && !is_finite_f32(raw)) kk = false;
  return kk;
}
__device__ __forceinline__ bool mask_keep(float raw, float z32, const MaskParams &m, const EmitArgs &ea) {
  return mask_keep_v(raw, z32, m, ea.use_z, ea.drop_nf, ea.z_min, ea.z_max);
}

// mask_interval (d2pc_math.h) by one warp: 32 probes of the ordered key space per round instead of
// one, so the two searches take ~7 rounds each.  The predicate is monotone (false -> true).
template <typename Pred>
__device__ __forceinline__ uint32_t warp_first_true(uint32_t lo, uint32_t hi, Pred pred) {
  // invariant: pred is false for keys < lo and true for hi
  const int lane = threadIdx.x & 31;
  while (lo < hi) {
    const unsigned long long span = (unsigned long long)(hi - lo);
    const uint32_t k = lo + (uint32_t)((span * (unsigned long long)(lane + 1)) / 33ull);  // < hi
    const bool p = pred(k);
    const unsigned m = __ballot_sync(0xffffffffu, p);
    if (m == 0u) {
      lo = __shfl_sync(0xffffffffu, k, 31) + 1u;
    } else {
      const int j = __ffs(m) - 1;
      hi = __shfl_sync(0xffffffffu, k, j);
      if (j > 0) lo = __shfl_sync(0xffffffffu, k, j - 1) + 1u;
    }
  }
  return lo;
}

// One tile of 1024 consecutive output rows per CTA, 4 consecutive rows per thread (they share an
// image row).  STEP 1: one 16 B depth load + 12 colour bytes; STEP 2 / 4 (density medium / low):
// the same tiling over the strided output grid, loading only the sectors that hold sampled
// pixels.  The depth map is always per-pixel here (the scan materialises a resized map).  MASK (any stride):
// depth-range / non-finite mask with ordered compaction; the tile's first output row comes from the per-tile
// kept counts that mask_count_kernel + mask_offsets_kernel computed over the same sampled pixels.
// High occupancy matters more than per-thread ILP here (measured: 6 CTAs/SM beat 3-5 and beat a
// persistent register-prefetching variant), hence MIN_BLOCKS.
struct FastArgs {
  uint32_t tiles_per_frame, total_tiles, batch;
  unsigned long long magic_w;  // ceil(2^40 / W): p / W == (p * magic_w) >> 40 for p * W < 2^40
  int32_t pc_simple;
};

// FrameState fields the emit reads.  COHERENT: inside the persistent path kernel they were written by another
// CTA of the same launch, and this SM's L1 may hold an older copy of the line: read through L2.
template <bool COHERENT>
__device__ __forceinline__ NormParams load_norm(const FrameState *fs) {
  if (!COHERENT) return fs->norm;
  NormParams np_;
  static_assert(sizeof(NormParams) % 8 == 0, "NormParams is copied as 8-byte words");
  const unsigned long long *src = reinterpret_cast<const unsigned long long *>(&fs->norm);
  unsigned long long *dst = reinterpret_cast<unsigned long long *>(&np_);
#pragma unroll
  for (int i = 0; i < (int)(sizeof(NormParams) / 8); ++i) dst[i] = __ldcg(src + i);
  return np_;
}

constexpr size_t kEmitStageBytes = 2 * ((size_t)kEmitTile * 3 + 4) * sizeof(float);
struct EmitSmall {  // static shared scratch of one emit tile
  uint32_t warp[kEmitThreads / 32];
  uint32_t b[6][kEmitThreads / 32];
  uint32_t dest;   // MASK: the tile's first output row
};

// norm_fn(): called by every thread of the CTA after the tile's loads have been issued; returns a pointer to the
// frame's NormParams (global memory for the stand-alone kernels; a snapshot in shared memory inside the path kernel,
// which may first have to wait for the frame's selection), or nullptr to give the tile up.  Returns false then.
template <int STEP, bool MASK, bool BOUNDS, typename NormFn>
__device__ __forceinline__ bool emit_fast_tile(const KParams &kp, const EmitArgs &ea, const FastArgs &fa, uint32_t b,
                                               uint32_t tile, float *s_stage, EmitSmall &es, NormFn norm_fn) {
  uint32_t *s_warp = es.warp;
  uint32_t (*s_b)[kEmitThreads / 32] = es.b;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // P = rows of the output grid (nu x nv); for STEP 1 the grid is the image itself
  const uint32_t P = kp.g.N, W = (uint32_t)kp.g.W, NU = (uint32_t)kp.g.nu;
  FrameState *fs = kp.state + b;
  const uint32_t tile_base = tile * (uint32_t)kEmitTile;
  const uint32_t p0 = tile_base + 4u * (uint32_t)tid;   // first of this thread's 4 output rows
  float *s_xyz = s_stage;
  float *s_rgb = s_stage + (kEmitTile * 3 + 4);
  float o[12];
  uint32_t c0 = 0, c1 = 0, c2 = 0;
  bool keep[4] = {false, false, false, false};
  uint32_t my_cnt = 0;
  uint32_t mn[3] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu}, mx[3] = {0u, 0u, 0u};
  float raw[4] = {0.0f, 0.0f, 0.0f, 0.0f};
  uint32_t u = 0, v = 0;
  if (p0 < P) {
    // (jv, ju) in the output grid; magic_w divides by nu (== W for STEP 1)
    const uint32_t jv = (uint32_t)(((unsigned long long)p0 * fa.magic_w) >> 40), ju = p0 - jv * NU;
    v = jv * (uint32_t)STEP; u = ju * (uint32_t)STEP;
    const size_t pix = (size_t)b * kp.g.P + (size_t)v * W + u;  // first source pixel (same row for all 4)
    const uint8_t *cp = ea.bgr + pix * 3;
    const uint64_t pol_d = l2_policy((kp.hints & kHintEmitDepthFirst) != 0, false);
    const uint64_t pol_c = l2_policy((kp.hints & kHintStreamFirst) != 0, false);
    if (STEP == 1) {
      const float4 d4 = ldg_f4_pol(kp.depth + pix, pol_d);
      raw[0] = d4.x; raw[1] = d4.y; raw[2] = d4.z; raw[3] = d4.w;
      c0 = ldg_u32_pol(cp, pol_c); c1 = ldg_u32_pol(cp + 4, pol_c); c2 = ldg_u32_pol(cp + 8, pol_c);
    } else if (STEP == 2) {
      // pixels u, u+2, u+4, u+6: two 16 B depth loads, the 24 colour bytes of 8 pixels
      const float4 da = ldg_f4_pol(kp.depth + pix, pol_d), db = ldg_f4_pol(kp.depth + pix + 4, pol_d);
      raw[0] = da.x; raw[1] = da.z; raw[2] = db.x; raw[3] = db.z;
      const uint32_t w0 = ldg_u32_pol(cp, pol_c), w1 = ldg_u32_pol(cp + 4, pol_c), w2 = ldg_u32_pol(cp + 8, pol_c);
      const uint32_t w3 = ldg_u32_pol(cp + 12, pol_c), w4 = ldg_u32_pol(cp + 16, pol_c), w5 = ldg_u32_pol(cp + 20, pol_c);
      // repack as the STEP 1 layout: c0 = B0 G0 R0 B1, c1 = G1 R1 B2 G2, c2 = R2 B3 G3 R3
      const uint32_t px1 = __byte_perm(w1, w2, 0x0432);  // bytes 6,7,8   -> B1 G1 R1 .
      const uint32_t px3 = __byte_perm(w4, w5, 0x0432);  // bytes 18,19,20 -> B3 G3 R3 .
      c0 = __byte_perm(w0, px1, 0x4210);                 // B0 G0 R0 B1
      c1 = __byte_perm(px1, w3, 0x5421);                 // G1 R1 B2 G2
      c2 = __byte_perm(w3, px3, 0x6542);                 // R2 B3 G3 R3
    } else {
      // pixels u, u+4, u+8, u+12: one 4 B load each (12-byte colour pitch keeps u32 loads aligned)
      raw[0] = __ldg(kp.depth + pix); raw[1] = __ldg(kp.depth + pix + 4);
      raw[2] = __ldg(kp.depth + pix + 8); raw[3] = __ldg(kp.depth + pix + 12);
      const uint32_t q0 = ldg_u32_pol(cp, pol_c), q1 = ldg_u32_pol(cp + 12, pol_c), q2 = ldg_u32_pol(cp + 24, pol_c),
                     q3 = ldg_u32_pol(cp + 36, pol_c);
      c0 = __byte_perm(q0, q1, 0x4210);   // B0 G0 R0 B1
      c1 = __byte_perm(q1, q2, 0x5421);   // G1 R1 B2 G2
      c2 = __byte_perm(q2, q3, 0x6542);   // R2 B3 G3 R3
    }
    if (!MASK) {
      // colours first: once staged, their registers are free for the float64 chain below
      float *sr = s_rgb + 12 * tid;
      // bytes (little endian): c0 = B0 G0 R0 B1, c1 = G1 R1 B2 G2, c2 = R2 B3 G3 R3
      stage_f4(sr, byte_to_float(c0, D2PC_B2), byte_to_float(c0, D2PC_B1), byte_to_float(c0, D2PC_B0),
               byte_to_float(c1, D2PC_B1));
      stage_f4(sr + 4, byte_to_float(c1, D2PC_B0), byte_to_float(c0, D2PC_B3), byte_to_float(c2, D2PC_B0),
               byte_to_float(c1, D2PC_B3));
      stage_f4(sr + 8, byte_to_float(c1, D2PC_B2), byte_to_float(c2, D2PC_B3), byte_to_float(c2, D2PC_B2),
               byte_to_float(c2, D2PC_B1));
    }
  }
  // MASK: the tile's first output row (written by mask_offsets_kernel before this launch) is fetched by one thread
  // while the pixel loads are in flight and parked in shared memory: fetched where it is used (after the CTA scan)
  // every thread waited a full L2 round trip for it, and fetched early by every thread it cost two live registers
  if (MASK && tid == 0) es.dest = (uint32_t)(__ldg(kp.tile_state + (size_t)b * kp.emit_tiles + tile) >> 32);
  const NormParams *np_src = norm_fn();
  if (np_src == nullptr) return false;  // uniform
  if (p0 < P) {
    const NormParams np_ = *np_src;
    const MaskParams mp = load_mask(fs);
    if (np_.simple && fa.pc_simple) {  // uniform per frame
      const double ux0 = (double)(int32_t)u - ea.pc.cx;
      const double vy = (double)(int32_t)v - ea.pc.cy;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        simple_point(raw[k], ux0 + (double)(k * STEP), vy, np_, ea.pc, &o[3 * k], &o[3 * k + 1], &o[3 * k + 2]);
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const double n = normalised_depth(raw[k], np_, ea.pc.invert);
        back_project(n, (int32_t)u + k * STEP, (int32_t)v, ea.pc, &o[3 * k], &o[3 * k + 1], &o[3 * k + 2]);
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      bool kk = true;
      if (MASK) kk = mask_keep(raw[k], o[3 * k + 2], mp, ea);  // same predicate as mask_count_kernel
      keep[k] = kk;
      my_cnt += kk ? 1u : 0u;
      if (BOUNDS && kk) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          uint32_t key = float_to_key(o[3 * k + c]);
          mn[c] = min(mn[c], key); mx[c] = max(mx[c], key);
        }
      }
    }
  }
  if (!MASK) {
    if (p0 < P) {
      float *sx = s_xyz + 12 * tid;
      stage_f4(sx, o[0], o[1], o[2], o[3]);
      stage_f4(sx + 4, o[4], o[5], o[6], o[7]);
      stage_f4(sx + 8, o[8], o[9], o[10], o[11]);
    }
    // The tile's rows are contiguous in both outputs (rows * 12 bytes each, a multiple of 16 at 16-byte
    // aligned addresses): one elected thread hands each staged array to the TMA unit as a bulk
    // shared -> global copy, the other threads are done.
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // staged rows visible to the async proxy
    __syncthreads();
    const uint32_t rows = min((uint32_t)kEmitTile, P - tile_base);
    const size_t g0 = ((size_t)(b + ea.frame0) * kp.g.N + tile_base) * 3;
    if (tid == 0) {
      const uint32_t bytes = rows * 12u;  // rows % 4 == 0 here
      const uint32_t sx = (uint32_t)__cvta_generic_to_shared(s_xyz), sr = (uint32_t)__cvta_generic_to_shared(s_rgb);
      const uint64_t pol_o = l2_policy((kp.hints & kHintStreamFirst) != 0, false);
      bulk_store_pol(ea.xyz + g0, sx, bytes, pol_o);
      bulk_store_pol(ea.rgb + g0, sr, bytes, pol_o);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the staging may be released once it has been read
      if (tile == 0) ea.count[b] = kp.g.N;
    }
  } else {
    // CTA exclusive scan of the per-thread keep counts
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
    // destination row: exclusive prefix of the kept counts, computed before this launch by
    // mask_count_kernel + mask_offsets_kernel (no inter-CTA dependency inside emit)
    const uint32_t dest_row = es.dest;   // (the barrier of the scan above ordered thread 0's store)
    const size_t g0 = ((size_t)(b + ea.frame0) * kp.g.N + dest_row) * 3;
    const uint32_t s_off = (uint32_t)(g0 & 3);
    const float col[12] = {
        byte_to_float(c0, D2PC_B2), byte_to_float(c0, D2PC_B1), byte_to_float(c0, D2PC_B0),
        byte_to_float(c1, D2PC_B1), byte_to_float(c1, D2PC_B0), byte_to_float(c0, D2PC_B3),
        byte_to_float(c2, D2PC_B0), byte_to_float(c1, D2PC_B3), byte_to_float(c1, D2PC_B2),
        byte_to_float(c2, D2PC_B3), byte_to_float(c2, D2PC_B2), byte_to_float(c2, D2PC_B1)};
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
  }
  if (BOUNDS) reduce_bounds(fs, mn, mx, s_b);
  return true;
}

template <int STEP, bool MASK, bool BOUNDS, int MIN_BLOCKS>
__global__ void __launch_bounds__(kEmitThreads, MIN_BLOCKS) emit_fast_kernel(KParams kp, EmitArgs ea, FastArgs fa) {
  extern __shared__ __align__(16) float s_stage[];  // xyz [3072 + 4] | rgb [3072 + 4]
  __shared__ EmitSmall es;
  const uint32_t t = blockIdx.x;
  const uint32_t b = t / fa.tiles_per_frame, tile = t - b * fa.tiles_per_frame;
  if (kp.state[b].status != D2PC_FRAME_READY) return;  // uniform per CTA (and per frame: no tile of it publishes)
  emit_fast_tile<STEP, MASK, BOUNDS>(kp, ea, fa, b, tile, s_stage, es, [&]() { return &kp.state[b].norm; });
}


}  // namespace d2pc
#endif  // D2PC_EMIT_DEV_CUH_
