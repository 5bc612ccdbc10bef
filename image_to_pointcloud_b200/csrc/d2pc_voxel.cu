// d2pc_voxel.cu -- ax-2, voxel-grid down-sampling of each frame's emitted rows.
// No reference code exists for this step ("parity unpinned"); the frozen
// spec is Open3D's PointCloud::VoxelDownSample: on float64 copies of the float32 points,
//   vmin = min_xyz - 0.5*vs ;  idx = floor((p - vmin) / vs) per axis ;
//   one output row per occupied voxel = arithmetic mean of member points and colours.
//
// Memory plan.  One open-addressing hash table per engine, 64-byte entries (two 32 B sectors) so
// that a point touches one line of the table:
//   [ key u64 | count<<32 + sum r | sum g<<32 + sum b | sum x | sum y | sum z | pad 16 ]
//   * colours are the integral 0..255 floats emit writes: their sums are exact integers, two packed
//     64-bit adds carry count, r, g and b (N < 2^24 rows keeps every 32-bit field from overflowing);
//   * coordinates are accumulated as 64-bit fixed point of (p - vmin) with a power-of-two scale
//     chosen per frame so that N * extent * scale < 2^62: integer adds are associative, so the
//     means are deterministic (run-to-run identical) and within 2^-38 * extent of the float64 sums.
//   * insert: 4 consecutive rows per thread, equal neighbouring keys are merged in registers first
//     (raster neighbours share voxels in smooth scenes); one probe (load, CAS only on an empty
//     slot) + 5 RED.64 per run.  Slots claimed by a CTA are queued in shared memory and appended
//     to the frame's occupied-slot list with one global atomic per CTA.
//   * extract: one thread per occupied slot: mean, AoS rows, clear the entry.  The table is never
//     swept: cost scales with the voxels, not with the capacity, and the table stays clean.
#include "d2pc_device.cuh"

namespace d2pc {

constexpr unsigned long long kVoxEmpty = 0xFFFFFFFFFFFFFFFFull;
constexpr unsigned long long kVoxMagic = 0x64327063766f7832ull;  // "d2pcvox2"
constexpr int kVoxBits = 21;
constexpr int kVoxThreads = 256;
constexpr int kVoxPerThread = 4;
constexpr int kVoxTile = kVoxThreads * kVoxPerThread;
constexpr uint32_t kVoxMaxRows = 1u << 24;  // 32-bit packed count / colour fields

struct __align__(64) VoxEntry {
  unsigned long long key;
  unsigned long long cnt_r;   // count << 32 | sum of r
  unsigned long long g_b;     // sum of g << 32 | sum of b
  unsigned long long sx, sy, sz;  // fixed-point sums of (p - vmin) * scale
  unsigned long long pad[2];
};
static_assert(sizeof(VoxEntry) == 64, "entry must be one 64-byte line");

struct __align__(256) VoxHeader {
  unsigned long long magic;  // kVoxMagic once d2pc_voxel_table_init has run
  uint32_t cap;              // entries (power of two)
  uint32_t n_list;           // occupied slots of the frame being processed
  uint32_t done;             // extract CTAs finished (ticket for the reset)
  uint32_t frame_cap;        // capacity used for the current frame (pow2 >= 2 * rows, <= cap)
  double vmin[3];
  double scale;              // fixed-point scale (power of two)
  int32_t bad;               // table was not initialised
};

struct VoxTable {
  VoxHeader *hdr;
  VoxEntry *ent;    // [cap]
  uint32_t *list;   // [n_rows]
  uint32_t cap;
};

inline uint32_t vox_capacity(uint32_t n_rows) {
  uint32_t c = 1024;
  while (c < 2u * n_rows && c < 0x80000000u) c <<= 1;
  return c;
}
inline size_t vox_bytes(uint32_t cap, uint32_t n_rows) {
  return sizeof(VoxHeader) + (size_t)cap * sizeof(VoxEntry) + align_up((size_t)n_rows * 4, 256);
}
inline VoxTable vox_table(void *base, uint32_t cap) {
  VoxTable t;
  char *p = (char *)base;
  t.hdr = (VoxHeader *)p;  p += sizeof(VoxHeader);
  t.ent = (VoxEntry *)p;   p += (size_t)cap * sizeof(VoxEntry);
  t.list = (uint32_t *)p;
  t.cap = cap;
  return t;
}

__device__ __forceinline__ uint32_t hash_u64(unsigned long long k) {
  k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
  return (uint32_t)k;
}

__global__ void vox_init_kernel(VoxTable t) {
  const size_t n16 = (size_t)t.cap * (sizeof(VoxEntry) / 16);
  uint4 *p = reinterpret_cast<uint4 *>(t.ent);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
    // 16-byte word 0 of an entry holds the key: all ones = empty; everything else zero
    p[i] = ((i & 3) == 0) ? make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0u, 0u) : make_uint4(0u, 0u, 0u, 0u);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    t.hdr->magic = kVoxMagic;
    t.hdr->cap = t.cap;
    t.hdr->n_list = 0;
    t.hdr->done = 0;
    t.hdr->bad = 0;
  }
}

// per frame: origin, fixed-point scale, capacity used for this frame
__global__ void vox_begin_kernel(VoxTable t, const uint32_t *count, const float *bounds, double vs, uint32_t *ocount,
                                 int32_t *err) {
  if (threadIdx.x != 0) return;
  VoxHeader *h = t.hdr;
  *ocount = 0;
  if (h->magic != kVoxMagic || h->cap != t.cap) { h->bad = 1; *err = 2; return; }
  *err = 0;
  h->bad = 0;
  const double half = vs * 0.5;
  double ext = 0.0;
  for (int c = 0; c < 3; ++c) {
    const double vmin = (double)bounds[c] - half;
    h->vmin[c] = vmin;
    const double e = (double)bounds[3 + c] - vmin;
    if (e > ext) ext = e;
  }
  // N * ext * scale < 2^62  with  N <= 2^24: scale = 2^(38 - ceil(log2(ext)))
  int ex = 0;
  if (ext > 0.0 && ext < 1.0e300) frexp(ext, &ex);  // ext = m * 2^ex, 0.5 <= m < 1
  h->scale = ldexp(1.0, 38 - ex);
  uint32_t m = *count, c = 1024;
  while (c < 2u * m && c < t.cap) c <<= 1;
  h->frame_cap = c;
}

struct VoxRun {
  unsigned long long key;
  unsigned long long cnt_r, g_b, sx, sy, sz;
};

__device__ __forceinline__ void vox_flush(const VoxTable &t, uint32_t mask, const VoxRun &r, uint32_t *s_new,
                                          uint32_t *s_nnew) {
  uint32_t slot = hash_u64(r.key) & mask;
  while (true) {
    VoxEntry *e = t.ent + slot;
    unsigned long long k = *reinterpret_cast<volatile unsigned long long *>(&e->key);
    if (k == kVoxEmpty) {
      k = atomicCAS(&e->key, kVoxEmpty, r.key);
      if (k == kVoxEmpty) {  // this thread claimed the slot: queue it for the occupied list
        s_new[atomicAdd(s_nnew, 1u)] = slot;
        break;
      }
    }
    if (k == r.key) break;
    slot = (slot + 1) & mask;
  }
  VoxEntry *e = t.ent + slot;
  atomicAdd(&e->cnt_r, r.cnt_r);
  atomicAdd(&e->g_b, r.g_b);
  atomicAdd(&e->sx, r.sx);
  atomicAdd(&e->sy, r.sy);
  atomicAdd(&e->sz, r.sz);
}

__device__ __forceinline__ unsigned long long shfl_up_u64(unsigned long long v, int d) {
  return __shfl_up_sync(0xffffffffu, v, d);
}

// One row per lane, kVoxPerThread consecutive 32-row groups per warp.  Rows of a group that share a
// voxel with their left neighbour form a run; a segmented warp scan sums each run into its last
// lane, which alone touches the table (raster neighbours share voxels whenever the voxel is larger
// than the pixel footprint).  Every lane has one independent probe in flight.
__global__ void __launch_bounds__(kVoxThreads, 4) vox_insert_kernel(VoxTable t, const float *xyz, const float *rgb,
                                                                    const uint32_t *count, double vs, int32_t *err) {
  __shared__ uint32_t s_new[kVoxTile];
  __shared__ uint32_t s_nnew, s_base;
  const uint32_t M = *count;
  const uint32_t tile_base = blockIdx.x * (uint32_t)kVoxTile;
  if (tile_base >= M || t.hdr->bad) return;
  if (threadIdx.x == 0) s_nnew = 0;
  __syncthreads();
  const double vmin0 = t.hdr->vmin[0], vmin1 = t.hdr->vmin[1], vmin2 = t.hdr->vmin[2];
  const double scale = t.hdr->scale;
  const uint32_t mask = t.hdr->frame_cap - 1u;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const double lim = (double)(1 << kVoxBits);
  const double inv_vs = 1.0 / vs;
  static_assert(kVoxPerThread % 2 == 0, "rows are loaded two groups ahead");
#pragma unroll 1
  for (int g2 = 0; g2 < kVoxPerThread; g2 += 2) {
   float raw[2][6];
   bool live[2];
#pragma unroll
   for (int h = 0; h < 2; ++h) {  // both groups' loads are in flight before either is used
     const uint32_t i = tile_base + (uint32_t)((warp * kVoxPerThread + g2 + h) * 32 + lane);
     live[h] = i < M;
     if (live[h]) {
       const float *p = xyz + 3 * (size_t)i, *c = rgb + 3 * (size_t)i;
       raw[h][0] = __ldg(p); raw[h][1] = __ldg(p + 1); raw[h][2] = __ldg(p + 2);
       raw[h][3] = __ldg(c); raw[h][4] = __ldg(c + 1); raw[h][5] = __ldg(c + 2);
     }
   }
#pragma unroll
   for (int h = 0; h < 2; ++h) {
    VoxRun run;
    run.key = kVoxEmpty;  // rows past the end / rejected rows: no run
    run.cnt_r = run.g_b = run.sx = run.sy = run.sz = 0ull;
    if (live[h]) {
      const double o0 = (double)raw[h][0] - vmin0, o1 = (double)raw[h][1] - vmin1, o2 = (double)raw[h][2] - vmin2;
      // floor((p - vmin) / vs): correctly rounded quotient from the reciprocal (d2pc_math.h div_by_const)
      const double i0d = floor(div_by_const(o0, vs, inv_vs)), i1d = floor(div_by_const(o1, vs, inv_vs)),
                   i2d = floor(div_by_const(o2, vs, inv_vs));
      if (i0d >= 0.0 && i0d < lim && i1d >= 0.0 && i1d < lim && i2d >= 0.0 && i2d < lim) {
        run.key = ((unsigned long long)i0d << (2 * kVoxBits)) | ((unsigned long long)i1d << kVoxBits) |
                  (unsigned long long)i2d;
        run.cnt_r = (1ull << 32) | (unsigned long long)(uint32_t)raw[h][3];
        run.g_b = ((unsigned long long)(uint32_t)raw[h][4] << 32) | (unsigned long long)(uint32_t)raw[h][5];
        run.sx = (unsigned long long)__double2ll_rn(o0 * scale);
        run.sy = (unsigned long long)__double2ll_rn(o1 * scale);
        run.sz = (unsigned long long)__double2ll_rn(o2 * scale);
      } else {
        *err = 1;  // "voxel_size is too small." (or non-finite coordinates)
      }
    }
    // segmented inclusive scan over runs of equal keys
    const unsigned long long left = shfl_up_u64(run.key, 1);
    const bool head = (lane == 0) || (left != run.key);
    const unsigned heads = __ballot_sync(0xffffffffu, head);
    if (heads != 0xffffffffu) {  // some neighbours share a voxel (uniform branch)
      const int start = 31 - __clz(heads & (0xffffffffu >> (31 - lane)));  // first lane of my run
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long a0 = shfl_up_u64(run.cnt_r, d), a1 = shfl_up_u64(run.g_b, d);
        const unsigned long long a2 = shfl_up_u64(run.sx, d), a3 = shfl_up_u64(run.sy, d), a4 = shfl_up_u64(run.sz, d);
        if (lane - d >= start) { run.cnt_r += a0; run.g_b += a1; run.sx += a2; run.sy += a3; run.sz += a4; }
      }
    }
    const bool tail = (lane == 31) || ((heads >> (lane + 1)) & 1u);
    if (tail && run.key != kVoxEmpty) vox_flush(t, mask, run, s_new, &s_nnew);
   }
  }
  __syncthreads();
  const uint32_t nn = s_nnew;
  if (nn == 0) return;
  if (threadIdx.x == 0) s_base = atomicAdd(&t.hdr->n_list, nn);
  __syncthreads();
  const uint32_t base = s_base;
  for (uint32_t j = threadIdx.x; j < nn; j += kVoxThreads) t.list[base + j] = s_new[j];
}

__global__ void __launch_bounds__(kVoxThreads, 2) vox_extract_kernel(VoxTable t, float *oxyz, float *orgb, int32_t *oidx,
                                                                  uint32_t *ocount, const int32_t *err) {
  const uint32_t V = t.hdr->n_list;
  // CTAs past the occupied list have nothing to do (an unusable table has an empty list)
  if (blockIdx.x * (uint32_t)kVoxTile >= V) return;
  const bool failed = *err != 0;
  const double inv_scale = 1.0 / t.hdr->scale;  // power of two: exact
  const double vmin0 = t.hdr->vmin[0], vmin1 = t.hdr->vmin[1], vmin2 = t.hdr->vmin[2];
  constexpr int E = kVoxPerThread;  // slots per thread, all loads issued before the first use
  const uint32_t j0 = blockIdx.x * (uint32_t)kVoxTile + threadIdx.x;
  {
    uint32_t slot[E];
    uint4 w[E][3];
#pragma unroll
    for (int k = 0; k < E; ++k) {
      const uint32_t j = j0 + (uint32_t)k * kVoxThreads;
      slot[k] = j < V ? t.list[j] : 0xFFFFFFFFu;
    }
#pragma unroll
    for (int k = 0; k < E; ++k) {
      if (slot[k] == 0xFFFFFFFFu) continue;
      const uint4 *e16 = reinterpret_cast<const uint4 *>(t.ent + slot[k]);
      w[k][0] = e16[0]; w[k][1] = e16[1]; w[k][2] = e16[2];
    }
#pragma unroll
    for (int k = 0; k < E; ++k) {
      if (slot[k] == 0xFFFFFFFFu) continue;
      const uint32_t j = j0 + (uint32_t)k * kVoxThreads;
      uint4 *e16 = reinterpret_cast<uint4 *>(t.ent + slot[k]);
      // leave the table clean for the next frame
      e16[0] = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0u, 0u);
      e16[1] = make_uint4(0u, 0u, 0u, 0u);
      e16[2] = make_uint4(0u, 0u, 0u, 0u);
      if (failed) continue;
      const uint4 w0 = w[k][0], w1 = w[k][1], w2 = w[k][2];
      const unsigned long long key = ((unsigned long long)w0.y << 32) | w0.x;
      const uint32_t cnt = w0.w, sr = w0.z, sg = w1.y, sb = w1.x;
      const unsigned long long sx = ((unsigned long long)w1.w << 32) | w1.z;
      const unsigned long long sy = ((unsigned long long)w2.y << 32) | w2.x;
      const unsigned long long sz = ((unsigned long long)w2.w << 32) | w2.z;
      const double n = (double)cnt;
      float *ox = oxyz + 3 * (size_t)j, *oc = orgb + 3 * (size_t)j;
      stg_stream_f1(ox + 0, (float)(vmin0 + ((double)(long long)sx * inv_scale) / n));
      stg_stream_f1(ox + 1, (float)(vmin1 + ((double)(long long)sy * inv_scale) / n));
      stg_stream_f1(ox + 2, (float)(vmin2 + ((double)(long long)sz * inv_scale) / n));
      stg_stream_f1(oc + 0, (float)((double)sr / n));
      stg_stream_f1(oc + 1, (float)((double)sg / n));
      stg_stream_f1(oc + 2, (float)((double)sb / n));
      if (oidx) {
        oidx[3 * (size_t)j + 0] = (int32_t)(key >> (2 * kVoxBits));
        oidx[3 * (size_t)j + 1] = (int32_t)((key >> kVoxBits) & ((1u << kVoxBits) - 1u));
        oidx[3 * (size_t)j + 2] = (int32_t)(key & ((1u << kVoxBits) - 1u));
      }
    }
  }
  // the last CTA publishes the count and resets the list for the next frame
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const uint32_t ticket = atomicAdd(&t.hdr->done, 1u);
    if (ticket == (V + kVoxTile - 1) / kVoxTile - 1) {  // CTAs with work
      *ocount = failed ? 0u : V;
      t.hdr->n_list = 0;
      t.hdr->done = 0;
      __threadfence();
    }
  }
}

}  // namespace d2pc

using namespace d2pc;

extern "C" int d2pc_voxel_table_bytes(const D2pcConfig *cfg, size_t *bytes) {
  int rc = validate_config(cfg);
  if (rc) return rc;
  if (!bytes) return D2PC_ERR_INVALID_ARGUMENT;
  const Geom g = make_geom(*cfg);
  if (g.N >= kVoxMaxRows) return D2PC_ERR_UNSUPPORTED;
  *bytes = vox_bytes(vox_capacity(g.N), g.N);
  return D2PC_OK;
}

extern "C" int d2pc_voxel_table_init(const D2pcConfig *cfg, void *d_table, size_t table_bytes, void *stream) {
  int rc = validate_config(cfg);
  if (rc) return rc;
  if (!d_table || ((uintptr_t)d_table & 255u) != 0) return D2PC_ERR_INVALID_ARGUMENT;
  const Geom g = make_geom(*cfg);
  if (g.N >= kVoxMaxRows) return D2PC_ERR_UNSUPPORTED;
  const uint32_t cap = vox_capacity(g.N);
  if (table_bytes < vox_bytes(cap, g.N)) return D2PC_ERR_WORKSPACE_TOO_SMALL;
  vox_init_kernel<<<148 * 8, 256, 0, (cudaStream_t)stream>>>(vox_table(d_table, cap));
  D2PC_CHECK_LAUNCH();
  return D2PC_OK;
}

extern "C" int d2pc_voxel_enqueue(const D2pcConfig *cfg, double voxel_size, const float *d_xyz,
                                  const float *d_rgb, const uint32_t *d_count, const float *d_bounds,
                                  void *d_table, size_t table_bytes, float *d_vox_xyz, float *d_vox_rgb,
                                  int32_t *d_vox_idx, uint32_t *d_vox_count, int32_t *d_vox_error,
                                  void *stream) {
  int rc = validate_config(cfg);
  if (rc) return rc;
  if (!(voxel_size > 0.0)) return D2PC_ERR_INVALID_ARGUMENT;
  if (!d_xyz || !d_rgb || !d_count || !d_bounds || !d_table || !d_vox_xyz || !d_vox_rgb || !d_vox_count ||
      !d_vox_error)
    return D2PC_ERR_INVALID_ARGUMENT;
  if (((uintptr_t)d_table & 255u) != 0) return D2PC_ERR_INVALID_ARGUMENT;
  const Geom g = make_geom(*cfg);
  if (g.N >= kVoxMaxRows) return D2PC_ERR_UNSUPPORTED;
  const uint32_t cap = vox_capacity(g.N);
  if (table_bytes < vox_bytes(cap, g.N)) return D2PC_ERR_WORKSPACE_TOO_SMALL;
  cudaStream_t st = (cudaStream_t)stream;
  VoxTable t = vox_table(d_table, cap);
  const uint32_t tiles = (g.N + kVoxTile - 1) / kVoxTile;
  for (int b = 0; b < cfg->batch; ++b) {
    const size_t ro = (size_t)b * g.N * 3;
    vox_begin_kernel<<<1, 32, 0, st>>>(t, d_count + b, d_bounds + 6 * b, voxel_size, d_vox_count + b,
                                       d_vox_error + b);
    D2PC_CHECK_LAUNCH();
    vox_insert_kernel<<<tiles, kVoxThreads, 0, st>>>(t, d_xyz + ro, d_rgb + ro, d_count + b, voxel_size,
                                                     d_vox_error + b);
    D2PC_CHECK_LAUNCH();
    vox_extract_kernel<<<tiles, kVoxThreads, 0, st>>>(t, d_vox_xyz + ro, d_vox_rgb + ro,
                                                        d_vox_idx ? d_vox_idx + ro : nullptr, d_vox_count + b,
                                                        d_vox_error + b);
    D2PC_CHECK_LAUNCH();
  }
  return D2PC_OK;
}
