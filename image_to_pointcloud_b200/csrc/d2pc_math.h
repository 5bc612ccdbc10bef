// d2pc_math.h -- per-pixel and per-frame arithmetic of the depth -> point-cloud stage, written
// once for host and device so the same code can be checked on a CPU (tests/test_host_math.py
// builds it with g++ as libd2pc_hostmath.so) and then run inside the sm_100a kernels.
//
// Every function cites the reference step it restates (reference backend/app.py:174-250) or the
// third-party arithmetic that step calls.  Bit-exactness rules:
//   * no floating-point contraction anywhere: nvcc is invoked with --fmad=false and g++ with
//     -ffp-contract=off; the only fused operations are the explicit fma()/fmaf() calls below;
//   * the per-pixel chain runs in float64 exactly as NumPy >= 2 runs it (NEP 50 promotes the
//     clip/normalise/invert chain to float64 because np.percentile returns float64 scalars);
//   * divisions by per-frame constants use a reciprocal + two FMA correction steps that return
//     the correctly rounded quotient (see div_by_const), never an approximation.
#ifndef D2PC_MATH_H_
#define D2PC_MATH_H_

#include <math.h>
#include <stdint.h>
#include <string.h>

#include "../../include/d2pc.h"

#if defined(__CUDACC__)
#define D2PC_HD __host__ __device__ __forceinline__
#else
#define D2PC_HD static inline
#endif

namespace d2pc {

// ---------------------------------------------------------------------------------------------
// bit casts and the order-preserving float32 <-> uint32 key
// ---------------------------------------------------------------------------------------------
D2PC_HD uint32_t f32_bits(float f) {
#if defined(__CUDA_ARCH__)
  return __float_as_uint(f);
#else
  uint32_t u;
  memcpy(&u, &f, 4);
  return u;
#endif
}
D2PC_HD float bits_f32(uint32_t u) {
#if defined(__CUDA_ARCH__)
  return __uint_as_float(u);
#else
  float f;
  memcpy(&f, &u, 4);
  return f;
#endif
}
// Monotone map: a < b (as floats, -0 < +0 by convention)  <=>  key(a) < key(b).
// -inf -> 0x007FFFFF, +inf -> 0xFF800000; NaNs fall outside [key(-inf), key(+inf)].
D2PC_HD uint32_t float_to_key(float f) {
  uint32_t b = f32_bits(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
D2PC_HD float key_to_float(uint32_t k) {
  return bits_f32((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k);
}
#define D2PC_KEY_NEG_INF 0x007FFFFFu
#define D2PC_KEY_POS_INF 0xFF800000u

D2PC_HD bool is_finite_f32(float f) { return (f32_bits(f) & 0x7F800000u) != 0x7F800000u; }
D2PC_HD bool is_nan_f32(float f) { return (f32_bits(f) & 0x7FFFFFFFu) > 0x7F800000u; }
D2PC_HD bool is_inf_f32(float f) { return (f32_bits(f) & 0x7FFFFFFFu) == 0x7F800000u; }
D2PC_HD float nan_f32() { return bits_f32(0x7FC00000u); }

// ---------------------------------------------------------------------------------------------
// a1  bilinear resize taps -- cv2.resize(INTER_LINEAR) float32, IPP path (call site app.py:188)
// ---------------------------------------------------------------------------------------------
struct AxisTap {
  int32_t i0, i1;  // source indices (i1 == i0 when clamped)
  float t;         // weight of tap 1, float32 cast of the float64 fraction
  int32_t clamped; // coordinate fell outside [0, n_src-1): output copies tap 0
};

// scale = (double)n_src / (double)n_dst (computed once by the caller, one IEEE division).
D2PC_HD AxisTap axis_tap(int32_t d, double scale, int32_t n_src) {
  AxisTap a;
  double f = ((double)d + 0.5) * scale - 0.5;  // two roundings; contraction is disabled
  double fl = floor(f);
  double t = f - fl;
  int32_t i0 = (int32_t)fl;
  a.clamped = 0;
  if (i0 < 0) {
    i0 = 0;
    t = 0.0;
    a.clamped = 1;
  }
  if (i0 >= n_src - 1) {
    i0 = n_src - 1;
    t = 0.0;
    a.clamped = 1;
  }
  a.i0 = i0;
  a.i1 = a.clamped ? i0 : i0 + 1;
  a.t = (float)t;
  return a;
}

// Horizontal pass on one source row: r = fmaf(S[x1] - S[x0], tx, S[x0]); clamped -> copy.
D2PC_HD float lerp_tap(float a, float b, float t, int32_t clamped) {
  if (clamped) return a;
  float diff = b - a;  // rounded float32 subtraction
  return fmaf(diff, t, a);
}

// Full 2-D sample at destination (xtap, ytap) from a source image with row pitch src_w.
D2PC_HD float bilinear_sample(const float *src, int32_t src_w, const AxisTap &tx,
                              const AxisTap &ty) {
  const float *row0 = src + (size_t)ty.i0 * src_w;
  float r0 = lerp_tap(row0[tx.i0], row0[tx.i1], tx.t, tx.clamped);
  float out;
  if (ty.clamped) {
    out = r0;
    // corner blocks (both clamped): IPP runs arithmetic there; +-inf comes out as NaN
    if (tx.clamped && is_inf_f32(out)) out = nan_f32();
  } else {
    const float *row1 = src + (size_t)ty.i1 * src_w;
    float r1 = lerp_tap(row1[tx.i0], row1[tx.i1], tx.t, tx.clamped);
    out = lerp_tap(r0, r1, ty.t, 0);
  }
  return out;
}

// a1 for sources with a 1-pixel side.  OpenCV does not hand those to IPP: its own two-pass code runs
// (imgproc resize.cpp, HResizeLinear / VResizeLinear) with float32 coordinates and weights, products
// and sums rounded separately (no FMA), column weights clamped and a plain copy from the first column
// with s + 1 >= w on, row weights NOT clamped (only the row indices are).
struct GenericTap {
  int32_t i0, i1;
  float w0, w1;
  int32_t copy;
};
D2PC_HD GenericTap generic_tap(int32_t d, double scale, int32_t n_src, int32_t is_column) {
  GenericTap g;
  float f = (float)(((double)d + 0.5) * scale - 0.5);
  float fl = floorf(f);
  int32_t s = (int32_t)fl;
  float t = f - fl;
  g.copy = 0;
  if (is_column) {
    if (s < 0) { s = 0; t = 0.0f; }
    if (s + 1 >= n_src) { g.copy = 1; s = s < n_src - 1 ? s : n_src - 1; t = 0.0f; }
    g.i0 = s;
    g.i1 = s + 1 < n_src ? s + 1 : n_src - 1;
  } else {
    g.i0 = s < 0 ? 0 : (s > n_src - 1 ? n_src - 1 : s);
    g.i1 = s + 1 < 0 ? 0 : (s + 1 > n_src - 1 ? n_src - 1 : s + 1);
  }
  g.w0 = 1.0f - t;
  g.w1 = t;
  return g;
}
D2PC_HD float generic_blend(float a, float b, float w0, float w1) {
#ifdef __CUDA_ARCH__
  return __fadd_rn(__fmul_rn(a, w0), __fmul_rn(b, w1));
#else
  volatile float p = a * w0, q = b * w1;  // two rounded products, one rounded sum
  return p + q;
#endif
}
D2PC_HD float generic_sample(const float *src, int32_t src_w, const GenericTap &tx, const GenericTap &ty) {
  const float *row0 = src + (size_t)ty.i0 * src_w, *row1 = src + (size_t)ty.i1 * src_w;
  const float r0 = tx.copy ? row0[tx.i0] : generic_blend(row0[tx.i0], row0[tx.i1], tx.w0, tx.w1);
  const float r1 = tx.copy ? row1[tx.i0] : generic_blend(row1[tx.i0], row1[tx.i1], tx.w0, tx.w1);
  return generic_blend(r0, r1, ty.w0, ty.w1);
}
D2PC_HD bool resize_is_generic(int32_t src_h, int32_t src_w) { return src_h == 1 || src_w == 1; }

// ---------------------------------------------------------------------------------------------
// a3  np.percentile(d, [2, 98]), method "linear" (numpy 2.3.5 _quantile/_lerp; app.py:197)
// ---------------------------------------------------------------------------------------------
struct RankPair {
  uint32_t lo, hi;  // 0-based ranks of the two neighbours in sorted order
  double gamma;     // interpolation weight
};
D2PC_HD RankPair percentile_ranks(uint32_t n, double q) {
  RankPair r;
  double vi = (double)(n - 1) * q;
  if (vi >= (double)(n - 1)) {  // _get_indexes: both neighbours = last element, gamma = vi+1
    r.lo = r.hi = n - 1;
    r.gamma = vi + 1.0;
  } else {
    double fl = floor(vi);
    r.lo = (uint32_t)fl;
    r.hi = r.lo + 1;
    r.gamma = vi - fl;
  }
  return r;
}
// _lerp(a, b, t): diff in float32, products/sums in float64, second form when t >= 0.5.
D2PC_HD double lerp_percentile(float a, float b, double g) {
  float diff = b - a;
  double r = (double)a + (double)diff * g;
  if (g >= 0.5) r = (double)b - (double)diff * (1.0 - g);
  return r;
}
#define D2PC_Q02 (2.0 / 100.0)
#define D2PC_Q98 (98.0 / 100.0)

// a2  np.nanmedian of float32 values given the middle order statistic(s) (app.py:195)
D2PC_HD float median_from_ranks(float lo, float hi, uint32_t count) {
  if (count == 0) return nan_f32();
  if (count & 1u) return hi;  // caller passes lo == hi == s[count/2]
  float s = lo + hi;          // float32 sum, then float32 halving (np.mean of two float32)
  return s / 2.0f;
}

// ---------------------------------------------------------------------------------------------
// a3/a4 per-frame normalisation parameters (app.py:197-204)
// ---------------------------------------------------------------------------------------------
struct NormParams {  // what the emit kernel needs, one per frame
  double p2, p98, den, inv_den;
  float lo32, hi32, den32, median;
  int32_t branch;
  int32_t has_nonfinite;
  // guard-free fast path (see finish_norm): valid iff simple != 0
  float clip_lo_f, clip_hi_f;  // for float32 d:  d < p2 <=> d < clip_lo_f ;  d > p98 <=> d > clip_hi_f
  int32_t simple;
  int32_t pad_;
};

// p2/p98: interpolated percentiles (float64); vmin/vmax: min/max of the repaired float32 map.
// any_nan: the repaired map still contains NaN (median itself was NaN) -> np.percentile is NaN.
D2PC_HD void finalise_norm(double p2, double p98, float vmin, float vmax, bool any_nan,
                           NormParams *o) {
  int32_t branch = D2PC_BRANCH_PCT;
  if (any_nan) {
    p2 = p98 = (double)nan_f32();
    // `p98 <= p2` is False for NaN, so the min/max fallback is not taken; `p98 > p2` is False
    branch = D2PC_BRANCH_ZEROS;
  } else if (p98 <= p2) {
    p2 = (double)vmin;  // float(d.min()), float(d.max())
    p98 = (double)vmax;
    branch = (p98 > p2) ? D2PC_BRANCH_MINMAX : D2PC_BRANCH_ZEROS;
  } else if (!(p98 > p2)) {
    branch = D2PC_BRANCH_ZEROS;
  }
  o->p2 = p2;
  o->p98 = p98;
  o->den = (p98 - p2) + 1e-6;
  o->inv_den = 1.0 / o->den;
  o->lo32 = (float)p2;
  o->hi32 = (float)p98;
  o->den32 = (float)o->den;  // np.float32(p98 - p2 + 1e-6): python float -> float32 once
  o->branch = branch;
}

// ---------------------------------------------------------------------------------------------
// correctly rounded a / b for a divisor b whose correctly rounded reciprocal y = RN(1/b) is known
// ---------------------------------------------------------------------------------------------
// q0 = RN(a*y) is within 1.5 ulp of a/b; one FMA residual step makes it faithful; Markstein's
// theorem (final-step lemma: y = RN(1/b), q faithful, r = a - b*q exact => RN(q + r*y) = RN(a/b))
// makes the second step correctly rounded.  Valid for finite a, normal b, quotient and residuals
// in the normal range (always true here: |a| in [2^-160, 2^130], b >= 1e-6); a == 0 -> 0;
// non-finite or out-of-range cases fall back to the IEEE division.
D2PC_HD double div_by_const(double a, double b, double y) {
  double q0 = a * y;
  double r0 = fma(-b, q0, a);
  double q1 = fma(r0, y, q0);
  double r1 = fma(-b, q1, a);
  double q2 = fma(r1, y, q1);
  // guard: anything exotic (inf/NaN operands, overflow, deep underflow) -> plain division
  double aq = fabs(q2);
  if (!(aq < 1e300) || aq < 1e-290) return a / b;  // also a == +-0: IEEE signed zero
  return q2;
}

// Same quotient without the guard, for callers that have established (per frame / per call) that
// a is finite, |a| is 0 or within [1e-200, 1e200], and b, y are finite and normal: then no
// intermediate over/underflows and the two-step correction is exact.  The sign is patched from
// the operands so that a == -0.0 (or +0.0 with b < 0) yields the IEEE signed zero.
D2PC_HD double div_by_const_fast(double a, double b, double y) {
  double q0 = a * y;
  double r0 = fma(-b, q0, a);
  double q1 = fma(r0, y, q0);
  double r1 = fma(-b, q1, a);
  return fma(r1, y, q1);
}
D2PC_HD double patch_quotient_sign(double q, double a, double b) {
#if defined(__CUDA_ARCH__)
  int hi = __double2hiint(q), lo = __double2loint(q);
  int s = (__double2hiint(a) ^ __double2hiint(b)) & 0x80000000;
  return __hiloint2double((hi & 0x7FFFFFFF) | s, lo);
#else
  uint64_t uq, ua, ub;
  memcpy(&uq, &q, 8); memcpy(&ua, &a, 8); memcpy(&ub, &b, 8);
  uq = (uq & 0x7FFFFFFFFFFFFFFFull) | ((ua ^ ub) & 0x8000000000000000ull);
  memcpy(&q, &uq, 8);
  return q;
#endif
}

D2PC_HD bool finite_mid_range(double x) { double a = fabs(x); return a == 0.0 || (a > 1e-100 && a < 1e100); }

// ---------------------------------------------------------------------------------------------
// a2 + a4 + a5 + a9  one pixel: raw (resized) depth -> xyz (app.py:193-206, 231-237)
// ---------------------------------------------------------------------------------------------
struct PixelConsts {  // per-call constants
  double scale;       // float(depth_scale)
  double cx, cy, f, inv_f;
  int32_t invert;
};

// normalised depth as the double the loop sees: float(d[v,u])
D2PC_HD double normalised_depth(float raw, const NormParams &np_, int32_t invert) {
  float d = raw;
  if (np_.has_nonfinite && !is_finite_f32(d)) d = np_.median;  // np.where(finite, d, med)
  double n;
  if (np_.branch == D2PC_BRANCH_PCT) {
    double c = (double)d;
    c = (c < np_.p2) ? np_.p2 : c;    // np.clip = minimum(maximum(d, p2), p98); d is not NaN here
    c = (c > np_.p98) ? np_.p98 : c;
    n = div_by_const(c - np_.p2, np_.den, np_.inv_den);
    if (invert) n = 1.0 - n;
  } else if (np_.branch == D2PC_BRANCH_MINMAX) {
    float c = d;
    c = (c < np_.lo32) ? np_.lo32 : c;
    c = (c > np_.hi32) ? np_.hi32 : c;
    float m = (c - np_.lo32) / np_.den32;  // float32 chain (python floats are weak scalars)
    if (invert) m = 1.0f - m;
    n = (double)m;
  } else {
    float m = 0.0f;
    if (invert) m = 1.0f - m;
    n = (double)m;
  }
  return n;
}

// u, v are pixel coordinates; out[0..2] = float32(x, y, z)
D2PC_HD void back_project(double n, int32_t u, int32_t v, const PixelConsts &pc, float *x, float *y,
                          float *z) {
  double zd = n * pc.scale;                 // z = float(d[v,u]) * float(depth_scale)
  double zz = (zd != 0.0) ? zd : 1e-6;      // (z if z != 0.0 else 1e-6)
  double ax = ((double)u - pc.cx) * zz;     // (u - cx) * zz      (rounded)
  double ay = ((double)v - pc.cy) * zz;
  *x = (float)div_by_const(ax, pc.f, pc.inv_f);  // ... / f   (rounded), then np.float32
  *y = (float)div_by_const(ay, pc.f, pc.inv_f);
  *z = (float)zd;
}

// ---------------------------------------------------------------------------------------------
// "simple frame" fast path: percentile branch, finite mid-range parameters, no non-finite pixel,
// and (per call) depth_scale > 0, f > 0.  Straight-line code without guards or sign patches;
// bit-identical to normalised_depth()/back_project() on such frames:
//   * every quotient is 0 or lies in [1e-200, 1e200], so div_by_const_fast is exact;
//   * numerators are >= +0 or have a non-zero magnitude, so no signed-zero case differs, except
//     a = c - p2 = -0.0 (d == -0.0 with p2 == +0.0), which is patched.
// Call after has_nonfinite has been set.
// ---------------------------------------------------------------------------------------------
D2PC_HD void finish_norm(NormParams *o) {
  o->simple = 0;
  o->clip_lo_f = o->clip_hi_f = 0.0f;
  o->pad_ = 0;
  if (o->branch != D2PC_BRANCH_PCT || o->has_nonfinite) return;
  if (!finite_mid_range(o->p2) || !finite_mid_range(o->p98) || !finite_mid_range(o->den) || o->den == 0.0) return;
  float lo = (float)o->p2;   // round to nearest, then move to the float32 ceiling of p2
  if ((double)lo < o->p2) lo = bits_f32(lo >= 0.0f ? f32_bits(lo) + 1u : f32_bits(lo) - 1u);
  float hi = (float)o->p98;  // float32 floor of p98
  if ((double)hi > o->p98) hi = bits_f32(hi > 0.0f ? f32_bits(hi) - 1u : f32_bits(hi) + 1u);
  if (!is_finite_f32(lo) || !is_finite_f32(hi)) return;
  o->clip_lo_f = lo;
  o->clip_hi_f = hi;
  o->simple = 1;
}
D2PC_HD bool consts_simple(const PixelConsts &pc) {
  return pc.scale > 1e-100 && pc.scale < 1e100 && pc.f > 1e-100 && pc.f < 1e100 && fabs(pc.cx) < 1e100 &&
         fabs(pc.cy) < 1e100;
}

// float32 z the simple path emits for raw depth d (first half of simple_point)
D2PC_HD float simple_z32(float d, const NormParams &sn, const PixelConsts &pc) {
  double c = (double)d;
  c = (d < sn.clip_lo_f) ? sn.p2 : c;
  c = (d > sn.clip_hi_f) ? sn.p98 : c;
  double a = c - sn.p2;
  double n = patch_quotient_sign(div_by_const_fast(a, sn.den, sn.inv_den), a, sn.den);
  if (pc.invert) n = 1.0 - n;
  return (float)(n * pc.scale);
}

// ax-1 in depth space.  On a simple frame z32(d) is a composition of weakly monotone, correctly
// rounded steps (clip, subtract p2, divide by den > 0, optional 1 - x, multiply by scale > 0, round
// to float32): non-decreasing in d without invert, non-increasing with it.  Hence
// {d : z_min <= z32(d) <= z_max} is an interval [lo, hi] of float32 values (empty if lo > hi),
// found by bisection over the ordered float32 keys with the exact z32().  The mask then costs two
// float compares per pixel, and a tile's kept count is known before any point is computed.
D2PC_HD void mask_interval(const NormParams &sn, const PixelConsts &pc, float z_min, float z_max,
                           float *lo_out, float *hi_out) {
  const uint32_t kmin = float_to_key(-3.402823466e38f), kmax = float_to_key(3.402823466e38f);
  const bool dec = pc.invert != 0;
  // lower end: smallest key whose z satisfies the bound that gets TRUE as d grows
  //   increasing z: z >= z_min ;  decreasing z: z <= z_max
  uint32_t a = kmin, b = kmax;  // search the first key in [kmin, kmax] with pred true; kmax+1 if none
  {
    uint32_t lo = kmin, hi = kmax;
    bool any = false;
    // invariant: pred(false) for keys < lo ; if any: pred(true) for keys >= hi
    const float zt = simple_z32(key_to_float(kmax), sn, pc);
    any = dec ? (zt <= z_max) : (zt >= z_min);
    if (!any) { *lo_out = 1.0f; *hi_out = 0.0f; return; }
    while (lo < hi) {
      const uint32_t mid = lo + ((hi - lo) >> 1);
      const float z = simple_z32(key_to_float(mid), sn, pc);
      const bool p = dec ? (z <= z_max) : (z >= z_min);
      if (p) hi = mid; else lo = mid + 1u;
    }
    a = lo;
  }
  {  // upper end: largest key whose z satisfies the bound that gets FALSE as d grows
    const float z0 = simple_z32(key_to_float(kmin), sn, pc);
    const bool any = dec ? (z0 >= z_min) : (z0 <= z_max);
    if (!any) { *lo_out = 1.0f; *hi_out = 0.0f; return; }
    uint32_t lo = kmin, hi = kmax;  // find last key with pred true
    while (lo < hi) {
      const uint32_t mid = lo + ((hi - lo + 1u) >> 1);
      const float z = simple_z32(key_to_float(mid), sn, pc);
      const bool p = dec ? (z >= z_min) : (z <= z_max);
      if (p) lo = mid; else hi = mid - 1u;
    }
    b = lo;
  }
  if (a > b) { *lo_out = 1.0f; *hi_out = 0.0f; return; }
  *lo_out = key_to_float(a);
  *hi_out = key_to_float(b);
}

// ux = (double)u - cx, vy = (double)v - cy (both exact)
D2PC_HD void simple_point(float d, double ux, double vy, const NormParams &sn, const PixelConsts &pc,
                          float *x, float *y, float *z) {
  double c = (double)d;
  c = (d < sn.clip_lo_f) ? sn.p2 : c;
  c = (d > sn.clip_hi_f) ? sn.p98 : c;
  double a = c - sn.p2;  // >= 0, but -0.0 when d == -0.0 and p2 == +0.0: keep the IEEE sign
  double n = patch_quotient_sign(div_by_const_fast(a, sn.den, sn.inv_den), a, sn.den);
  if (pc.invert) n = 1.0 - n;
  double zd = n * pc.scale;
  double zz = (zd != 0.0) ? zd : 1e-6;
  *x = (float)div_by_const_fast(ux * zz, pc.f, pc.inv_f);
  *y = (float)div_by_const_fast(vy * zz, pc.f, pc.inv_f);
  *z = (float)zd;
}

}  // namespace d2pc
#endif  // D2PC_MATH_H_
