// d2pc_stats_dev.cuh -- device-side pieces of the exact percentile statistics (reference steps a3/a4,
// backend/app.py:197-204) that run both as stand-alone kernels (d2pc_stats.cu) and as work items of the
// persistent path kernel (d2pc_path.cu): one scan tile, one bracket's exact selection, the frame's
// normalisation parameters.
#ifndef D2PC_STATS_DEV_CUH_
#define D2PC_STATS_DEV_CUH_

#include "d2pc_device.cuh"

namespace d2pc {

constexpr int kTilePerThread = 32;
constexpr int kTilePx = kScanThreads * kTilePerThread;  // 8192 pixels per scan tile

// ------------------------------------------------------------------------------------------
// scan: 4 instructions per pixel on the common path.  With the brackets [L0, U0] (around the 2% ranks) and
// [L1, U1] (around the 98% ranks), about 94% of the pixels lie strictly between U0 and L1 and need nothing
// at all: they are above bracket 0 and below bracket 1, which the frame's totals account for.  Every other
// value (about 6%, and every NaN: ordered compares with NaN are false) is appended to the thread's private
// queue column in shared memory; afterwards the low side (v <= U0, or NaN) goes to the frame's raw queue 0
// and the high side (v >= L1) to raw queue 1 in global memory -- one reservation per CTA and queue.  The
// selection classifies the queued values against the bracket (below / equal / inside / above / non-finite).
// Compares are float compares (-0.0 == +0.0), here and in the selection, consistently.  min/max are not
// tracked: a frame whose percentiles collapse (p98 <= p2) goes to the exact fallback, which computes them.
// ------------------------------------------------------------------------------------------
// qaddr = shared-space byte address of the thread's next free queue slot
__device__ __forceinline__ void scan_value(float v, float U0, float L1, uint32_t &qaddr) {
  asm volatile(
      "{\n\t"
      ".reg .pred mid;\n\t"
      "setp.gt.f32 mid, %1, %2;\n\t"
      "setp.lt.and.f32 mid, %1, %3, mid;\n\t"
      "@!mid st.shared.f32 [%0], %1;\n\t"
      "@!mid add.u32 %0, %0, %4;\n\t"
      "}"
      : "+r"(qaddr)
      : "f"(v), "f"(U0), "f"(L1), "n"(kScanThreads * 4)
      : "memory");
}

__device__ __forceinline__ float bracket_lo_float(uint32_t L) {
  return (L == 0u) ? -__int_as_float(0x7F800000) : key_to_float(L);           // open below
}
__device__ __forceinline__ float bracket_hi_float(uint32_t U) {
  return (U == 0xFFFFFFFFu) ? __int_as_float(0x7F800000) : key_to_float(U);   // open above
}
__device__ __forceinline__ void bracket_floats(const FrameState *fs, float Lf[2], float Uf[2]) {
#pragma unroll
  for (int br = 0; br < 2; ++br) {
    Lf[br] = bracket_lo_float(fs->brL[br]);
    Uf[br] = bracket_hi_float(fs->brU[br]);
  }
}

// Deferred values of a CTA -> the frame's two raw queues.  pq = the CTA's slot-major queue columns
// (pq[j * kScanThreads + tid] = thread tid's j-th value), nq = this thread's count.  Two barriers.
struct ScanFlushSmem {
  uint32_t wtot[kScanThreads / 32];
  uint32_t gbase[2];
};
__device__ __forceinline__ void scan_flush(const float *pq, uint32_t nq, float U0, float L1, FrameState *fs,
                                           const KParams &kp, int b, ScanFlushSmem &sm) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint32_t packed = 0;  // low half: values for queue 0, high half: values for queue 1
  for (uint32_t j = 0; j < nq; ++j) {
    const float v = pq[j * kScanThreads + tid];
    packed += (!(v > U0) ? 1u : 0u) + ((v >= L1) ? 0x10000u : 0u);
  }
  uint32_t incl = packed;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t y = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += y;
  }
  if (lane == 31) sm.wtot[warp] = incl;
  __syncthreads();
  uint32_t woff = 0, total = 0;
#pragma unroll
  for (int w = 0; w < kScanThreads / 32; ++w) {
    const uint32_t x = sm.wtot[w];
    if (w < warp) woff += x;
    total += x;
  }
  if (total == 0u) return;  // uniform
  if (tid == 0) {  // one reservation per CTA for both queues (see FrameState::nqueue)
    const unsigned long long add = (unsigned long long)(total & 0xFFFFu) | ((unsigned long long)(total >> 16) << 32);
    const unsigned long long old = atomicAdd(reinterpret_cast<unsigned long long *>(&fs->nqueue[0]), add);
    sm.gbase[0] = (uint32_t)old;
    sm.gbase[1] = (uint32_t)(old >> 32);
  }
  __syncthreads();
  if (nq == 0u) return;
  float *gq0 = reinterpret_cast<float *>(kp.cand) + (size_t)b * 2 * kp.cand_cap;
  const uint32_t excl = woff + incl - packed;
  // offsets into the frame's queue pair (queue 1 starts cand_cap floats after queue 0)
  uint32_t o0 = sm.gbase[0] + (excl & 0xFFFFu), o1 = sm.gbase[1] + (excl >> 16);
  // the queues are what the selection reads next: evict-last, so that the depth map streaming through L2 (1 GB per
  // 128 frames) does not push them out to DRAM before the selection runs
  const uint64_t pol_keep = l2_policy(false, true);
  if (o0 + (packed & 0xFFFFu) <= kp.cand_cap && o1 + (packed >> 16) <= kp.cand_cap) {
    o1 += kp.cand_cap;
    for (uint32_t j = 0; j < nq; ++j) {
      const float v = pq[j * kScanThreads + tid];
      if (!(v > U0)) stg_f1_pol(gq0 + o0++, v, pol_keep);
      if (v >= L1) stg_f1_pol(gq0 + o1++, v, pol_keep);
    }
  } else {  // a queue overflows (the frame will take the fallback): drop what does not fit
    for (uint32_t j = 0; j < nq; ++j) {
      const float v = pq[j * kScanThreads + tid];
      if (!(v > U0)) { if (o0 < kp.cand_cap) gq0[o0] = v; ++o0; }
      if (v >= L1) { if (o1 < kp.cand_cap) gq0[kp.cand_cap + o1] = v; ++o1; }
    }
  }
}

template <int PT>
struct ScanTileSmemT {
  float pq[PT][kScanThreads];  // per-thread deferred values, slot-major (conflict-free)
  ScanFlushSmem flush;
};
using ScanTileSmem = ScanTileSmemT<kTilePerThread>;

// One tile of PT * 256 consecutive pixels of frame b's per-pixel depth map (256 threads, PT pixels each).
// The caller synchronises the CTA before reusing the shared memory and before signalling completion.
template <int PT>
__device__ __forceinline__ void scan_tile(const KParams &kp, int b, uint32_t tile, int vec_ok, ScanTileSmemT<PT> &sm) {
  constexpr int kTilePerThread = PT;
  constexpr int kTilePx = PT * kScanThreads;
  const int tid = threadIdx.x;
  FrameState *fs = kp.state + b;
  const float *frame = kp.depth + (size_t)b * kp.g.D;
  const uint32_t n = kp.g.P;
  const float U0 = bracket_hi_float(fs->brU[0]), L1 = bracket_lo_float(fs->brL[1]);
  const uint32_t q0 = (uint32_t)__cvta_generic_to_shared(&sm.pq[0][tid]);
  uint32_t qaddr = q0;
  const uint32_t tile_base = tile * (uint32_t)kTilePx;
  if (vec_ok && tile_base + (uint32_t)kTilePx <= n) {
    const float *src = frame + tile_base + 4u * (uint32_t)tid;
    const uint64_t pol = l2_policy((kp.hints & kHintScanKeep) == 0, (kp.hints & kHintScanKeep) != 0);   // one stage: the map streams through (evict-first)
    float4 r[kTilePerThread / 4];
#pragma unroll
    for (int j = 0; j < kTilePerThread / 4; ++j) r[j] = ldg_f4_pol(src + (size_t)j * (4 * kScanThreads), pol);
#pragma unroll
    for (int j = 0; j < kTilePerThread / 4; ++j) {
      scan_value(r[j].x, U0, L1, qaddr);
      scan_value(r[j].y, U0, L1, qaddr);
      scan_value(r[j].z, U0, L1, qaddr);
      scan_value(r[j].w, U0, L1, qaddr);
    }
  } else {  // last tile of a frame / unaligned frames
#pragma unroll 1
    for (int j = 0; j < kTilePerThread / 4; ++j) {
      const uint32_t p = tile_base + 4u * (uint32_t)(j * kScanThreads + tid);
      for (uint32_t k = 0; k < 4u; ++k)
        if (p + k < n) scan_value(__ldg(frame + p + k), U0, L1, qaddr);
    }
  }
  const uint32_t nq = (qaddr - q0) / (uint32_t)(kScanThreads * 4);
  scan_flush(&sm.pq[0][0], nq, U0, L1, fs, kp, b, sm.flush);
}

// ------------------------------------------------------------------------------------------
// exact selection inside one bracket + the frame's parameters
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void finalise_fast(FrameState *fs, uint32_t n) {
  volatile FrameState *vfs = fs;
  if (vfs->sel_fail) {
    __threadfence();
    vfs->status = D2PC_FRAME_NEEDS_FALLBACK;
    return;
  }
  RankPair r2 = percentile_ranks(n, D2PC_Q02), r98 = percentile_ranks(n, D2PC_Q98);
  double p2 = lerp_percentile(key_to_float(vfs->sel_key[0]), key_to_float(vfs->sel_key[1]), r2.gamma);
  double p98 = lerp_percentile(key_to_float(vfs->sel_key[2]), key_to_float(vfs->sel_key[3]), r98.gamma);
  if (!(p98 > p2)) {  // degenerate percentiles need min/max: the exact path computes them
    __threadfence();
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

// visit this thread's share of a raw queue, U independent loads in flight per thread.  L2 loads (.cg): inside
// the persistent path kernel the queue was written by other CTAs of the same launch.
template <int U, typename F>
__device__ __forceinline__ void for_each_queued(const float *q, uint32_t nq, F f) {
  const uint32_t stride = blockDim.x;
  uint32_t i = threadIdx.x;
  for (; i + (uint32_t)(U - 1) * stride < nq; i += (uint32_t)U * stride) {
    float v[U];
#pragma unroll
    for (int k = 0; k < U; ++k) v[k] = __ldcg(q + i + (uint32_t)k * stride);
#pragma unroll
    for (int k = 0; k < U; ++k) f(v[k]);
  }
  for (; i < nq; i += stride) f(__ldcg(q + i));
}

constexpr int kSelectSmemWords = (int)kFastBins + 2 * (int)kFastListCap;  // 24 KB; general path: 2 << 11 words

struct SelectSmall {  // static shared scratch of one selection
  uint32_t res[2 * 2 + 40];
  uint32_t c[6];  // eqL, inside, eqU, below, non-finite, NaN
  uint32_t warp[33], cnt[2], lmin[2], lmax[2], out[2];
};

// One bracket of one frame by the calling CTA (blockDim.x = 256 ... 1024, a divisor of kFastBins / 4).
// First visit of the bracket's queue: classify against the bounds (equal to L / strictly inside / equal to
// U) and build the bucket histogram of the inside keys in the same loop; the wanted ranks are then resolved
// against below / eqL / inside / eqU, located in the histogram, and a second visit collects the (few) members
// of the wanted buckets for the exact pick.  Heavy ties inside one bucket take the general multi-level
// selection.  The CTA that finishes a frame's second bracket writes the frame's parameters and status.
//   s_hist: kSelectSmemWords words of shared memory.  All threads call; ends with the status published.
__device__ __forceinline__ void select_bracket(const KParams &kp, int b, int br, uint32_t *s_hist, SelectSmall &ss) {
  const int tid = threadIdx.x, nthr = blockDim.x;
  FrameState *fs = kp.state + b;
  const uint32_t n = kp.g.P;
  const uint32_t qcap = kp.cand_cap;
  const float *q = reinterpret_cast<const float *>(kp.cand) + ((size_t)b * 2 + br) * qcap;
  volatile FrameState *vfs = fs;
  const uint32_t nqueue = vfs->nqueue[br];
  const uint32_t nq = min(nqueue, qcap);
  const uint32_t L = fs->brL[br], U = fs->brU[br];
  float Lf2[2], Uf2[2];
  bracket_floats(fs, Lf2, Uf2);
  const float Lf = Lf2[br], Uf = Uf2[br];
  // strictly inside (L, U) in float order; -0.0 == +0.0 there, so when a bound is a zero the
  // other zero's key can sit one step outside the key interval: widen the key range by one
  const uint32_t lo0 = L > 0u ? L - 1u : 0u;
  const uint32_t hi0 = U < 0xFFFFFFFFu ? U + 1u : U;
  const uint32_t span = hi0 - lo0;
  const int shift = fast_shift(span);
  const uint32_t list_cap = kFastListCap;
  uint32_t *s_list = s_hist + kFastBins;  // [2][kFastListCap]
  for (uint32_t i = tid; i < kFastBins; i += nthr) s_hist[i] = 0u;
  if (tid < 6) ss.c[tid] = 0;
  if (tid < 2) { ss.cnt[tid] = 0u; ss.lmin[tid] = 0xFFFFFFFFu; ss.lmax[tid] = 0u; ss.out[tid] = 0u; }
  __syncthreads();
  {  // visit 1: queue 0 holds every value <= U0 (and NaN), queue 1 every value >= L1
    uint32_t c_eqL = 0, c_in = 0, c_eqU = 0, c_below = 0, c_nf = 0;  // c_nf: non-finite | NaN << 16
    for_each_queued<8>(q, nq, [&](float v) {
      if (!(fabsf(v) < __int_as_float(0x7F800000))) c_nf += 1u + ((v != v) ? 0x10000u : 0u);
      else if (v < Lf) c_below++;
      else if (v == Lf) c_eqL++;
      else if (v < Uf) {
        c_in++;
        const uint32_t d = float_to_key(v) - lo0;
        if (d <= span) atomicAdd(&s_hist[d >> shift], 1u);
      } else if (v == Uf) c_eqU++;
      // v > Uf: only when the brackets overlap (tiny frames); above this bracket, not needed
    });
    c_eqL = warp_sum(c_eqL); c_in = warp_sum(c_in); c_eqU = warp_sum(c_eqU); c_below = warp_sum(c_below);
    const uint32_t w_nf = warp_sum(c_nf & 0xFFFFu), w_nan = warp_sum(c_nf >> 16);
    if ((tid & 31) == 0) {
      if (c_eqL) atomicAdd(&ss.c[0], c_eqL);
      if (c_in) atomicAdd(&ss.c[1], c_in);
      if (c_eqU) atomicAdd(&ss.c[2], c_eqU);
      if (c_below) atomicAdd(&ss.c[3], c_below);
      if (w_nf) atomicAdd(&ss.c[4], w_nf);
      if (w_nan) atomicAdd(&ss.c[5], w_nan);
    }
  }
  __syncthreads();
  const uint32_t eqL = ss.c[0], nin = ss.c[1], eqU = ss.c[2], n_nf = ss.c[4];
  // finite values below the bracket: bracket 0 sees them in its queue; for bracket 1 they are everything that
  // is not in queue 1 (exact when the frame has no non-finite value; otherwise the frame fails anyway)
  const uint32_t below = br == 0 ? ss.c[3] : n - nqueue;
  bool fail = (n_nf != 0u) || (nqueue > qcap) || (fs->sample_ok == 0u) || (kp.force_fallback != 0);
  RankPair rp = percentile_ranks(n, br ? D2PC_Q98 : D2PC_Q02);
  bool want[2] = {false, false};
  uint32_t need[2] = {0u, 0u};
  uint32_t key[2] = {0u, 0u};
  for (int t = 0; t < 2; ++t) {
    uint32_t r = t ? rp.hi : rp.lo;
    if (r < below) { fail = true; continue; }
    uint32_t r1 = r - below;
    if (r1 < eqL) { key[t] = L; continue; }
    uint32_t r2 = r1 - eqL;
    if (r2 < nin) { need[t] = r2; want[t] = true; continue; }
    uint32_t r3 = r2 - nin;
    if (r3 < eqU) { key[t] = U; continue; }
    fail = true;
  }

  if (!fail && (want[0] || want[1])) {  // uniform over the CTA
    uint32_t bin[2], base[2];
    block_locate<2>(s_hist, need, want, bin, base, ss.warp, ss.res);
    bool slow = (want[0] && bin[0] == 0xFFFFFFFFu) || (want[1] && bin[1] == 0xFFFFFFFFu);
    if (!slow && shift == 0) {
      for (int t = 0; t < 2; ++t)
        if (want[t]) key[t] = lo0 + bin[t];
    } else if (!slow) {
      const bool shared_bin = want[0] && want[1] && bin[0] == bin[1];
      const bool own[2] = {want[0], want[1] && !shared_bin};
      for_each_queued<8>(q, nq, [&](float v) {
        if (v > Lf && v < Uf) {
          const uint32_t k = float_to_key(v);
          const uint32_t bk = (k - lo0) >> shift;
#pragma unroll
          for (int t = 0; t < 2; ++t)
            if (own[t] && bk == bin[t]) {
              const uint32_t idx = atomicAdd(&ss.cnt[t], 1u);
              if (idx < list_cap) s_list[t * kFastListCap + idx] = k;
              atomicMin(&ss.lmin[t], k);
              atomicMax(&ss.lmax[t], k);
            }
        }
      });
      __syncthreads();
      for (int t = 0; t < 2; ++t)
        if (own[t] && ss.cnt[t] > list_cap && ss.lmin[t] != ss.lmax[t]) slow = true;
      if (!slow) {
        for (int t = 0; t < 2; ++t) {
          if (!want[t]) continue;
          const int o = (t == 1 && shared_bin) ? 0 : t;
          if (ss.lmin[o] == ss.lmax[o]) { if (tid == 0) ss.out[t] = ss.lmin[o]; }
          else block_pick(s_list + o * kFastListCap, ss.cnt[o], need[t] - base[t], &ss.out[t]);
        }
        __syncthreads();
        for (int t = 0; t < 2; ++t)
          if (want[t]) key[t] = ss.out[t];
      }
    }
    if (slow) {  // uniform
      __syncthreads();
      uint32_t rk[2], ok_[2];
      for (int t = 0; t < 2; ++t) rk[t] = want[t] ? need[t] : need[1 - t];
      const bool ok = block_hist_select<2, 11>([&](auto f) {
        for_each_queued<8>(q, nq, [&](float v) {
          if (v > Lf && v < Uf) f(float_to_key(v));
        });
      }, lo0, hi0, rk, ok_, s_hist, ss.res);
      if (!ok) fail = true;
      for (int t = 0; t < 2; ++t)
        if (want[t]) key[t] = ok_[t];
    }
  }

  if (tid == 0) {
    fs->sel_key[2 * br + 0] = key[0];
    fs->sel_key[2 * br + 1] = key[1];
    fs->eqL[br] = eqL; fs->inside[br] = nin; fs->eqU[br] = eqU;
    fs->below[br] = below;
    if (n_nf) { atomicAdd(&fs->n_nonfinite, n_nf); atomicAdd(&fs->n_nan, ss.c[5]); }
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
// bounded waits between CTAs of one launch (index-ordered kernels: a CTA only waits for CTAs with a smaller
// index, which were dispatched before it, or for the few peers of a cooperative selection right behind it)
// ------------------------------------------------------------------------------------------
constexpr uint32_t kSpinLimit = 1u << 19;   // x ~2 us: about a second

__device__ __forceinline__ uint32_t ld_relaxed_u32(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// thread 0 spins until pred(value at p) holds; false: timed out (abort flag raised)
template <typename Pred>
__device__ __forceinline__ bool spin_until(const uint32_t *p, uint32_t *abort_flag, Pred pred) {
  uint32_t spins = 0, ns = 128;
  while (!pred(ld_relaxed_u32(p))) {
    __nanosleep(ns);                 // back off: hundreds of CTAs may poll the same line
    if (ns < 2048u) ns <<= 1;
    if (++spins > kSpinLimit || ((spins & 15u) == 0u && ld_relaxed_u32(abort_flag))) {
      atomicExch(abort_flag, 1u);
      return false;
    }
  }
  (void)ld_acquire_u32(p);
  return true;
}

__device__ __forceinline__ uint32_t *sched_abort_flag(const KParams &kp) { return kp.sched + 32u * kSchedQueues; }

// ------------------------------------------------------------------------------------------
// The same selection, cooperatively: K CTAs of the persistent path kernel share one bracket.  Each takes a
// slice of the bracket's queue, keeps it in shared memory, classifies it and adds the inside keys to the
// frame's global bucket histogram (L2 atomics).  The CTA that finishes last resolves the wanted ranks and
// locates their buckets; the others wait for that (a short spin: they were all claimed within a few
// microseconds of each other), then every CTA looks through its cached slice for members of the wanted
// buckets and appends them to a global list; the last one again picks the exact keys and, as the second
// bracket of the frame to finish, writes the frame's parameters.  ~10 us per frame instead of ~100 us for one
// CTA walking 62 000 queued values twice.
// ------------------------------------------------------------------------------------------
constexpr uint32_t kSliceCap = 6144;  // floats of dynamic shared memory a slice may use (callers pass their capacity)

struct SelPartSmall {
  uint32_t c[6];
  uint32_t res[2 * 2 + 40];
  uint32_t warp[33];
  uint32_t flag_a, flag_spin, flag_c, out[2];
};

// writes the bracket's result and, as the frame's second bracket, the frame's parameters (thread 0)
__device__ __forceinline__ void select_finish(FrameState *fs, SelShared *sh, int br, uint32_t n, const uint32_t key[2],
                                              bool fail) {
  volatile uint32_t *cn = sh->counts[br];
  fs->sel_key[2 * br + 0] = key[0];
  fs->sel_key[2 * br + 1] = key[1];
  fs->eqL[br] = cn[1]; fs->inside[br] = cn[2]; fs->eqU[br] = cn[3];
  if (cn[4]) { atomicAdd(&fs->n_nonfinite, cn[4]); atomicAdd(&fs->n_nan, cn[5]); }
  if (fail) atomicOr(&fs->sel_fail, 1u);
  __threadfence();
  const uint32_t ticket = atomicAdd(&fs->sel_done, 1u);
  if (ticket == 1u) {
    __threadfence();
    finalise_fast(fs, n);
  }
}

// The cooperative selection in two steps.  Step A (select_phase_a): classify the slice, add to the global histogram;
// the part that finishes last resolves the ranks and publishes the wanted buckets (or finishes the bracket if no
// collect is needed).  Step C (select_phase_c): collect the members of the wanted buckets; the part that finishes
// last picks the exact keys and finishes the bracket.  Inside an index-ordered kernel the two steps run in one CTA
// with a bounded wait between them (select_part, slice cached in shared memory); as two launches (select_a_kernel /
// select_c_kernel) nothing waits at all.
//   s_slice: slice_cap floats of shared memory (>= 2 << 11 words, which the general selection needs).
struct SelCommon {
  FrameState *fs;
  SelShared *sh;
  const float *q;
  uint32_t n, qcap, nqueue, nq, L, U, lo0, hi0, span, s_lo, m;
  float Lf, Uf;
  int shift;
  bool cached;
};
__device__ __forceinline__ SelCommon select_common(const KParams &kp, int b, int br, uint32_t k, uint32_t K, uint32_t slice_cap) {
  SelCommon c;
  c.fs = kp.state + b;
  c.sh = kp.sel + b;
  volatile FrameState *vfs = c.fs;
  c.n = kp.g.P; c.qcap = kp.cand_cap;
  c.q = reinterpret_cast<const float *>(kp.cand) + ((size_t)b * 2 + br) * c.qcap;
  c.nqueue = vfs->nqueue[br];
  c.nq = min(c.nqueue, c.qcap);
  c.L = c.fs->brL[br]; c.U = c.fs->brU[br];
  c.Lf = bracket_lo_float(c.L); c.Uf = bracket_hi_float(c.U);
  c.lo0 = c.L > 0u ? c.L - 1u : 0u;
  c.hi0 = c.U < 0xFFFFFFFFu ? c.U + 1u : c.U;
  c.span = c.hi0 - c.lo0;
  c.shift = fast_shift(c.span);
  const uint32_t chunk = (c.nq + K - 1u) / K;
  c.s_lo = min(c.nq, k * chunk);
  c.m = min(c.nq, c.s_lo + chunk) - c.s_lo;
  c.cached = chunk <= slice_cap;
  return c;
}

// returns 0: not the last part of step A; 1: last, buckets published (collect needed); 2: last, bracket finished
// LOCAL_HIST (stand-alone two-launch selection, slice not cached): the slice's bucket histogram is built in
// shared memory (s_slice, kSelBins words) and only its non-zero buckets are added to the frame's global one --
// a few thousand spread-out L2 reductions per CTA instead of one contended L2 atomic per inside value.
template <bool LOCAL_HIST = false>
__device__ __forceinline__ int select_phase_a(const KParams &kp, int b, int br, uint32_t k, uint32_t K, float *s_slice,
                                              uint32_t slice_cap, SelPartSmall &ss) {
  const int tid = threadIdx.x, nthr = blockDim.x;
  const SelCommon cm = select_common(kp, b, br, k, K, slice_cap);
  FrameState *fs = cm.fs;
  SelShared *sh = cm.sh;
  const float *q = cm.q;
  const uint32_t n = cm.n, qcap = cm.qcap, nqueue = cm.nqueue, L = cm.L, U = cm.U, lo0 = cm.lo0, span = cm.span,
                 s_lo = cm.s_lo, m = cm.m;
  const float Lf = cm.Lf, Uf = cm.Uf;
  const int shift = cm.shift;
  const bool cached = cm.cached && !LOCAL_HIST;
  if (tid < 6) ss.c[tid] = 0u;
  if (LOCAL_HIST)
    for (uint32_t i = tid; i < kSelBins; i += nthr) reinterpret_cast<uint32_t *>(s_slice)[i] = 0u;
  __syncthreads();
  {  // phase A: classify the slice, histogram of the inside keys
    uint32_t c_eqL = 0, c_in = 0, c_eqU = 0, c_below = 0, c_nf = 0;
    uint32_t *hist = LOCAL_HIST ? reinterpret_cast<uint32_t *>(s_slice) : sh->hist[br];
    auto visit = [&](float v) {
      if (!(fabsf(v) < __int_as_float(0x7F800000))) c_nf += 1u + ((v != v) ? 0x10000u : 0u);
      else if (v < Lf) c_below++;
      else if (v == Lf) c_eqL++;
      else if (v < Uf) {
        c_in++;
        const uint32_t d = float_to_key(v) - lo0;
        if (d <= span) atomicAdd(&hist[d >> shift], 1u);
      } else if (v == Uf) c_eqU++;
    };
    uint32_t i = tid;
    for (; i + 7u * nthr < m; i += 8u * nthr) {  // eight independent L2 loads in flight per thread
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = __ldcg(q + s_lo + i + (uint32_t)u * nthr);
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (cached) s_slice[i + (uint32_t)u * nthr] = v[u];
        visit(v[u]);
      }
    }
    for (; i < m; i += nthr) {
      const float v = __ldcg(q + s_lo + i);
      if (cached) s_slice[i] = v;
      visit(v);
    }
    c_eqL = warp_sum(c_eqL); c_in = warp_sum(c_in); c_eqU = warp_sum(c_eqU); c_below = warp_sum(c_below);
    const uint32_t w_nf = warp_sum(c_nf & 0xFFFFu), w_nan = warp_sum(c_nf >> 16);
    if ((tid & 31) == 0) {
      if (c_below) atomicAdd(&ss.c[0], c_below);
      if (c_eqL) atomicAdd(&ss.c[1], c_eqL);
      if (c_in) atomicAdd(&ss.c[2], c_in);
      if (c_eqU) atomicAdd(&ss.c[3], c_eqU);
      if (w_nf) atomicAdd(&ss.c[4], w_nf);
      if (w_nan) atomicAdd(&ss.c[5], w_nan);
    }
  }
  __syncthreads();
  if (LOCAL_HIST) {
    const uint32_t *lh = reinterpret_cast<const uint32_t *>(s_slice);
    for (uint32_t i = tid; i < kSelBins; i += nthr) {
      const uint32_t h = lh[i];
      if (h) atomicAdd(&sh->hist[br][i], h);
    }
    __threadfence();
    __syncthreads();   // the ticket below (thread 0) follows every thread's reductions
  }
  if (tid == 0) {
    for (int c = 0; c < 6; ++c)
      if (ss.c[c]) atomicAdd(&sh->counts[br][c], ss.c[c]);
    __threadfence();
    ss.flag_a = atomicAdd(&sh->a_done[br], 1u);
    __threadfence();
  }
  __syncthreads();
  const bool last_a = ss.flag_a == K - 1u;
  if (last_a) {  // uniform: resolve the ranks, locate their buckets, publish
    volatile uint32_t *cn = sh->counts[br];
    const uint32_t eqL = cn[1], nin = cn[2], eqU = cn[3], n_nf = cn[4];
    const uint32_t below = br == 0 ? cn[0] : n - nqueue;
    bool fail = (n_nf != 0u) || (nqueue > qcap) || (fs->sample_ok == 0u) || (kp.force_fallback != 0);
    const RankPair rp = percentile_ranks(n, br ? D2PC_Q98 : D2PC_Q02);
    bool want[2] = {false, false};
    uint32_t need[2] = {0u, 0u}, key[2] = {0u, 0u};
    for (int t = 0; t < 2; ++t) {
      const uint32_t r = t ? rp.hi : rp.lo;
      if (r < below) { fail = true; continue; }
      const uint32_t r1 = r - below;
      if (r1 < eqL) { key[t] = L; continue; }
      const uint32_t r2 = r1 - eqL;
      if (r2 < nin) { need[t] = r2; want[t] = true; continue; }
      const uint32_t r3 = r2 - nin;
      if (r3 < eqU) { key[t] = U; continue; }
      fail = true;
    }
    bool collect = false;
    uint32_t bin[2] = {0u, 0u}, base[2] = {0u, 0u};
    if (!fail && (want[0] || want[1])) {
      block_locate<2, true>(sh->hist[br], need, want, bin, base, ss.warp, ss.res);
      if ((want[0] && bin[0] == 0xFFFFFFFFu) || (want[1] && bin[1] == 0xFFFFFFFFu)) fail = true;  // cannot happen
      else if (shift == 0) {
        for (int t = 0; t < 2; ++t)
          if (want[t]) key[t] = lo0 + bin[t];
      } else collect = true;
    }
    if (tid == 0) {
      if (collect) {
        for (int t = 0; t < 2; ++t) {
          sh->bin[br][t] = bin[t]; sh->base[br][t] = base[t]; sh->need[br][t] = need[t];
          sh->want[br][t] = want[t] ? 1u : 0u;
          sh->key[br][t] = key[t];
        }
        sh->shared_bin[br] = (want[0] && want[1] && bin[0] == bin[1]) ? 1u : 0u;
        __threadfence();
        *reinterpret_cast<volatile uint32_t *>(&sh->bin_ready[br]) = 1u;
      } else {
        select_finish(fs, sh, br, n, key, fail);
        __threadfence();
        *reinterpret_cast<volatile uint32_t *>(&sh->bin_ready[br]) = 2u;
      }
    }
    return collect ? 1 : 2;
  }
  return 0;
}

__device__ __forceinline__ void select_phase_c(const KParams &kp, int b, int br, uint32_t k, uint32_t K,
                                               float *s_slice, uint32_t slice_cap, bool slice_is_cached, SelPartSmall &ss) {
  const int tid = threadIdx.x, nthr = blockDim.x;
  const SelCommon cm = select_common(kp, b, br, k, K, slice_cap);
  FrameState *fs = cm.fs;
  SelShared *sh = cm.sh;
  const float *q = cm.q;
  const uint32_t n = cm.n, nq = cm.nq, lo0 = cm.lo0, hi0 = cm.hi0, s_lo = cm.s_lo, m = cm.m;
  const float Lf = cm.Lf, Uf = cm.Uf;
  const int shift = cm.shift;
  const bool cached = cm.cached && slice_is_cached;
  // phase C: members of the wanted buckets in this slice -> the frame's global lists
  {
    volatile SelShared *vsh = sh;
    const bool shared_bin = vsh->shared_bin[br] != 0u;
    const bool own[2] = {vsh->want[br][0] != 0u, vsh->want[br][1] != 0u && !shared_bin};
    const uint32_t bin0 = vsh->bin[br][0], bin1 = vsh->bin[br][1];
    auto collect = [&](float v) {
      if (v > Lf && v < Uf) {
        const uint32_t key = float_to_key(v);
        const uint32_t bk = (key - lo0) >> shift;
#pragma unroll
        for (int t = 0; t < 2; ++t)
          if (own[t] && bk == (t ? bin1 : bin0)) {
            const uint32_t idx = atomicAdd(&sh->mcount[br][t], 1u);
            if (idx < kSelListCap) sh->members[br][t][idx] = key;
            atomicMax(&sh->mmin[br][t], ~key);   // min through max: the scratch is zero-initialised
            atomicMax(&sh->mmax[br][t], key);
          }
      }
    };
    uint32_t i = tid;
    if (!cached) {  // the slice comes from L2 again: eight independent loads in flight per thread
      for (; i + 7u * nthr < m; i += 8u * nthr) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = __ldcg(q + s_lo + i + (uint32_t)u * nthr);
#pragma unroll
        for (int u = 0; u < 8; ++u) collect(v[u]);
      }
    }
    for (; i < m; i += nthr) collect(cached ? s_slice[i] : __ldcg(q + s_lo + i));
  }
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    ss.flag_c = atomicAdd(&sh->c_done[br], 1u);
    __threadfence();
  }
  __syncthreads();
  if (ss.flag_c != K - 1u) return;
  // last of phase C: exact pick inside the buckets
  {
    volatile SelShared *vsh = sh;
    const bool shared_bin = vsh->shared_bin[br] != 0u;
    bool want[2] = {vsh->want[br][0] != 0u, vsh->want[br][1] != 0u};
    uint32_t key[2] = {vsh->key[br][0], vsh->key[br][1]};
    bool fail = false, slow = false;
    uint32_t *s_list = reinterpret_cast<uint32_t *>(s_slice);  // [2][kSelListCap]
    for (int t = 0; t < 2; ++t) {
      if (!want[t] || (t == 1 && shared_bin)) continue;
      const uint32_t cnt = vsh->mcount[br][t];
      if (cnt > kSelListCap && ~vsh->mmin[br][t] != vsh->mmax[br][t]) slow = true;
      for (uint32_t i = tid; i < min(cnt, kSelListCap); i += nthr) s_list[t * kSelListCap + i] = __ldcg(&sh->members[br][t][i]);
    }
    if (tid < 2) ss.out[tid] = 0u;
    __syncthreads();
    if (!slow) {
      for (int t = 0; t < 2; ++t) {
        if (!want[t]) continue;
        const int o = (t == 1 && shared_bin) ? 0 : t;
        const uint32_t mn = ~vsh->mmin[br][o], mx = vsh->mmax[br][o];
        if (mn == mx) { if (tid == 0) ss.out[t] = mn; }
        else block_pick(s_list + o * kSelListCap, min(vsh->mcount[br][o], kSelListCap), vsh->need[br][t] - vsh->base[br][t], &ss.out[t]);
      }
      __syncthreads();
      for (int t = 0; t < 2; ++t)
        if (want[t]) key[t] = ss.out[t];
    } else {  // heavy ties inside one bucket: general multi-level selection over the whole queue
      uint32_t rk[2], ok_[2];
      for (int t = 0; t < 2; ++t) rk[t] = want[t] ? vsh->need[br][t] : vsh->need[br][1 - t];
      const bool ok = block_hist_select<2, 11>([&](auto f) {
        for_each_queued<8>(q, nq, [&](float v) {
          if (v > Lf && v < Uf) f(float_to_key(v));
        });
      }, lo0, hi0, rk, ok_, reinterpret_cast<uint32_t *>(s_slice), ss.res);
      if (!ok) fail = true;
      for (int t = 0; t < 2; ++t)
        if (want[t]) key[t] = ok_[t];
    }
    if (tid == 0) select_finish(fs, sh, br, n, key, fail);
  }
}

// both steps in one CTA (index-ordered kernels).  spin(ptr, pred) -> bool is the caller's bounded wait (false: abort).
template <typename Spin>
__device__ __forceinline__ bool select_part(const KParams &kp, int b, int br, uint32_t k, uint32_t K, float *s_slice,
                                            uint32_t slice_cap, SelPartSmall &ss, Spin spin) {
  const int st = select_phase_a<false>(kp, b, br, k, K, s_slice, slice_cap, ss);   // uniform over the CTA
  if (st == 2) return true;
  SelShared *sh = kp.sel + b;
  if (st == 0) {
    if (threadIdx.x == 0) ss.flag_spin = spin(&sh->bin_ready[br], [](uint32_t v) { return v != 0u; }) ? 1u : 0u;
    __syncthreads();
    if (!ss.flag_spin) return false;
    if (*reinterpret_cast<volatile uint32_t *>(&sh->bin_ready[br]) == 2u) return true;
  }
  __syncthreads();
  select_phase_c(kp, b, br, k, K, s_slice, slice_cap, true, ss);
  return true;
}

}  // namespace d2pc
#endif  // D2PC_STATS_DEV_CUH_
