// TEST HARNESS (not product, not oracle): compiles the product's arithmetic header
// image_to_pointcloud_b200/csrc/d2pc_math.h for the host, so that the exact per-pixel / per-frame
// math the sm_100a kernels run can be checked against the NumPy oracle and the golden vectors
// on a machine without a GPU.  Order statistics come from std::sort here (the GPU selection
// kernels are checked on the GPU); everything else is the shared header.
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <random>
#include <vector>

#include "../../image_to_pointcloud_b200/csrc/d2pc_math.h"

using namespace d2pc;

static int g_use_simple = 0;
static int g_last_simple = 0;

extern "C" {

void hm_set_simple(int on) { g_use_simple = on; }
int hm_last_simple(void) { return g_last_simple; }

int hm_resize(const float *src, int h, int w, float *dst, int H, int W) {
  const double sx = (double)w / (double)W, sy = (double)h / (double)H;
  if (resize_is_generic(h, w)) {
    for (int v = 0; v < H; ++v) {
      GenericTap ty = generic_tap(v, sy, h, 0);
      for (int u = 0; u < W; ++u) dst[(size_t)v * W + u] = generic_sample(src, w, generic_tap(u, sx, w, 1), ty);
    }
    return 0;
  }
  for (int v = 0; v < H; ++v) {
    AxisTap ty = axis_tap(v, sy, h);
    for (int u = 0; u < W; ++u) {
      AxisTap tx = axis_tap(u, sx, w);
      dst[(size_t)v * W + u] = bilinear_sample(src, w, tx, ty);
    }
  }
  return 0;
}

// the whole stage on the CPU through the shared header; returns number of points
long hm_depth_to_point_cloud(const uint8_t *img, int H, int W, int C, const float *depth, int h, int w,
                             int step, int invert, double scale, double cx, double cy, double f,
                             float *xyz, float *rgb, D2pcFrameParams *out) {
  const size_t P = (size_t)H * W;
  std::vector<float> d(P);
  if (h == H && w == W) std::copy(depth, depth + P, d.begin());
  else hm_resize(depth, h, w, d.data(), H, W);
  // a2: non-finite repair
  uint32_t n_nf = 0, n_nan = 0;
  for (float v : d) { if (!is_finite_f32(v)) { n_nf++; if (is_nan_f32(v)) n_nan++; } }
  NormParams np_;
  np_.median = 0.0f; np_.has_nonfinite = 0;
  std::vector<float> rep(d);
  if (n_nf) {
    std::vector<float> nn;
    nn.reserve(P);
    for (float v : d) if (!is_nan_f32(v)) nn.push_back(v);
    std::sort(nn.begin(), nn.end());
    const uint32_t m = (uint32_t)nn.size();
    float med = nan_f32();
    if (m) med = (m & 1u) ? median_from_ranks(nn[m / 2], nn[m / 2], m) : median_from_ranks(nn[m / 2 - 1], nn[m / 2], m);
    np_.median = med; np_.has_nonfinite = 1;
    for (auto &v : rep) if (!is_finite_f32(v)) v = med;
  }
  bool any_nan = false;
  for (float v : rep) if (is_nan_f32(v)) { any_nan = true; break; }
  NormParams fin;
  if (any_nan) {
    finalise_norm(0, 0, 0, 0, true, &fin);
  } else {
    std::vector<uint32_t> keys(P);
    for (size_t i = 0; i < P; ++i) keys[i] = float_to_key(rep[i]);
    std::sort(keys.begin(), keys.end());
    RankPair r2 = percentile_ranks((uint32_t)P, D2PC_Q02), r98 = percentile_ranks((uint32_t)P, D2PC_Q98);
    double p2 = lerp_percentile(key_to_float(keys[r2.lo]), key_to_float(keys[r2.hi]), r2.gamma);
    double p98 = lerp_percentile(key_to_float(keys[r98.lo]), key_to_float(keys[r98.hi]), r98.gamma);
    finalise_norm(p2, p98, key_to_float(keys.front()), key_to_float(keys.back()), false, &fin);
  }
  fin.median = np_.median; fin.has_nonfinite = np_.has_nonfinite;
  if (out) {
    out->p2 = fin.p2; out->p98 = fin.p98; out->den = fin.den; out->inv_den = fin.inv_den;
    out->lo32 = fin.lo32; out->hi32 = fin.hi32; out->den32 = fin.den32; out->median = fin.median;
    out->branch = fin.branch; out->status = D2PC_FRAME_READY; out->n_nonfinite = n_nf; out->n_nan = n_nan;
  }
  PixelConsts pc;
  pc.scale = scale; pc.cx = cx; pc.cy = cy; pc.f = f; pc.inv_f = 1.0 / f; pc.invert = invert;
  long i = 0;
  finish_norm(&fin);
  const NormParams &sn = fin;
  const bool simple = g_use_simple && fin.simple && consts_simple(pc);
  g_last_simple = simple ? 1 : 0;
  for (int v = 0; v < H; v += step)
    for (int u = 0; u < W; u += step, ++i) {
      if (simple) {  // the guard-free straight-line path the emit_fast kernel takes
        simple_point(d[(size_t)v * W + u], (double)u - cx, (double)v - cy, sn, pc, &xyz[3 * i], &xyz[3 * i + 1],
                     &xyz[3 * i + 2]);
      } else {
        double n = normalised_depth(d[(size_t)v * W + u], fin, invert);
        back_project(n, u, v, pc, &xyz[3 * i], &xyz[3 * i + 1], &xyz[3 * i + 2]);
      }
      if (C >= 3) {
        const uint8_t *cp = img + ((size_t)v * W + u) * C;
        rgb[3 * i] = cp[2]; rgb[3 * i + 1] = cp[1]; rgb[3 * i + 2] = cp[0];
      } else {
        rgb[3 * i] = rgb[3 * i + 1] = rgb[3 * i + 2] = 128.0f;
      }
    }
  return i;
}

// ax-1 in depth space: mask_interval() against the brute-force z-range test on the given values.
// Returns the number of values whose interval membership differs from the z32 comparison, or -1
// when the parameters do not describe a "simple" frame.
long hm_mask_check(double p2, double p98, int invert, double scale, float z_min, float z_max,
                   const float *d, long n, float *lo_out, float *hi_out) {
  NormParams np_;
  finalise_norm(p2, p98, 0.0f, 0.0f, false, &np_);
  np_.median = 0.0f; np_.has_nonfinite = 0;
  finish_norm(&np_);
  PixelConsts pc;
  pc.scale = scale; pc.cx = 10.0; pc.cy = 10.0; pc.f = 100.0; pc.inv_f = 1.0 / 100.0; pc.invert = invert;
  if (!np_.simple || !consts_simple(pc)) return -1;
  float lo, hi;
  mask_interval(np_, pc, z_min, z_max, &lo, &hi);
  *lo_out = lo; *hi_out = hi;
  long bad = 0;
  for (long i = 0; i < n; ++i) {
    const float z = simple_z32(d[i], np_, pc);
    const bool by_z = (z >= z_min) && (z <= z_max);
    const bool by_d = (d[i] >= lo) && (d[i] <= hi);
    if (by_z != by_d) bad++;
  }
  return bad;
}

// div_by_const(a, b, RN(1/b)) against the IEEE quotient; returns the number of mismatches
long hm_div_check(long n, unsigned long seed, int mode) {
  std::mt19937_64 rng(seed);
  long bad = 0;
  for (long i = 0; i < n; ++i) {
    double a, b;
    uint64_t r1 = rng(), r2 = rng();
    if (mode == 0) {        // arbitrary normal doubles, moderate exponents
      uint64_t ea = 1023 - 60 + (r1 >> 52) % 120, eb = 1023 - 20 + (r2 >> 52) % 60;
      uint64_t ba = (r1 & 0x000FFFFFFFFFFFFFull) | (ea << 52), bb = (r2 & 0x000FFFFFFFFFFFFFull) | (eb << 52);
      memcpy(&a, &ba, 8); memcpy(&b, &bb, 8);
      if (r1 & (1ull << 63)) a = -a;
    } else if (mode == 1) { // a = difference of float32 values, b = den-like
      float x = (float)((r1 >> 11) * (1.0 / 9007199254740992.0) * 20.0);
      float y = (float)((r2 >> 11) * (1.0 / 9007199254740992.0) * 20.0);
      a = (double)x - (double)y * 0.31; b = (double)y * 0.93 + 1e-6;
    } else if (mode == 3) { // quotients that sit on or next to representable values / midpoints
      uint64_t bq = (r1 & 0x000FFFFFFFFFFFFFull) | (1023ull << 52), bb = (r2 & 0x000FFFFFFFFFFFFFull) | (1023ull << 52);
      double q; memcpy(&q, &bq, 8); memcpy(&b, &bb, 8);
      if ((r1 >> 60) & 1) q = q + q * 1.1102230246251565e-16;  // ~half ulp up: near a midpoint
      a = q * b;
      int nudge = (int)((r2 >> 56) % 5) - 2;
      for (int k = 0; k < (nudge < 0 ? -nudge : nudge); ++k) a = nextafter(a, nudge < 0 ? 0.0 : 1e300);
    } else {                // (u - cx) * zz / f
      double zz = (r1 >> 11) * (1.0 / 9007199254740992.0) * 10.0;
      a = ((double)(int)(r2 % 3840) - 1920.0) * zz; b = 4608.0 * (1.0 + (double)((r2 >> 20) % 7) * 0.1);
    }
    if (b == 0.0) continue;
    double y = 1.0 / b;
    double q = div_by_const(a, b, y);
    double t = a / b;
    if (memcmp(&q, &t, 8) != 0) bad++;
  }
  return bad;
}

}  // extern "C"

// ---- f3: writer byte layouts (csrc/d2pc_format.h) ----------------------------------------------
#include "../../image_to_pointcloud_b200/csrc/d2pc_format.h"
extern "C" {
// all lines of save_xyz into out (capacity cap); returns total bytes, -1 on a formatting error, -2 if cap is too small
long hm_xyz_text(const float *xyz, const float *rgb, long n, char *out, long cap) {
  long pos = 0;
  char line[d2pc::kXyzMaxLine];
  for (long i = 0; i < n; ++i) {
    int m = d2pc::format_xyz_line(xyz + 3 * i, rgb + 3 * i, line);
    if (m < 0) return -1;
    if (pos + m > cap) return -2;
    memcpy(out + pos, line, m);
    pos += m;
  }
  return pos;
}
int hm_las_records(const float *xyz, const float *rgb, long n, const double *off, double scale, unsigned char *out) {
  int ok = 1;
  int32_t q[3];
  for (long i = 0; i < n; ++i)
    if (!d2pc::las_record(xyz + 3 * i, rgb + 3 * i, off, scale, out + 26 * i, q)) ok = 0;
  return ok;
}
void hm_ply_records(const float *xyz, const float *rgb, long n, unsigned char *out) {
  for (long i = 0; i < n; ++i) d2pc::ply_record(xyz + 3 * i, rgb + 3 * i, out + 27 * i);
}
unsigned hm_preview_stride(unsigned n, unsigned max_preview) { return d2pc::preview_stride(n, max_preview); }
}
