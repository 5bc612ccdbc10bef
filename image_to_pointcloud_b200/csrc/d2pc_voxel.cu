// d2pc_voxel.cu -- ax-2, voxel-grid down-sampling of each frame's emitted rows.
// No reference code exists for this step ("parity unpinned"); the frozen
// spec is Open3D's PointCloud::VoxelDownSample: on float64 copies of the float32 points,
//   vmin = min_xyz - 0.5*vs ;  idx = floor((p - vmin) / vs) per axis ;
//   one output row per occupied voxel = arithmetic mean of member points and colours.
// Round-1 implementation: one open-addressing hash table in global memory (SoA: 64-bit packed
// voxel key, count, six float64 sums), frames processed one after another on the stream;
// the extract pass compacts occupied slots and leaves the table clean for the next frame.
#include "d2pc_device.cuh"

namespace d2pc {

constexpr unsigned long long kVoxEmpty = 0xFFFFFFFFFFFFFFFFull;
constexpr int kVoxBits = 21;

struct VoxTable {
  unsigned long long *keys;  // [cap]
  uint32_t *cnt;             // [cap]
  double *sum;               // [cap][6]
  uint32_t cap;              // power of two
};

inline uint32_t vox_capacity(uint32_t n_rows) {
  uint32_t c = 1024;
  while (c < 2u * n_rows && c < 0x80000000u) c <<= 1;
  return c;
}
inline size_t vox_bytes(uint32_t cap) {
  return align_up((size_t)cap * 8, 256) + align_up((size_t)cap * 4, 256) + align_up((size_t)cap * 48, 256);
}
inline VoxTable vox_table(void *base, uint32_t cap) {
  VoxTable t;
  char *p = (char *)base;
  t.keys = (unsigned long long *)p; p += align_up((size_t)cap * 8, 256);
  t.cnt = (uint32_t *)p;            p += align_up((size_t)cap * 4, 256);
  t.sum = (double *)p;
  t.cap = cap;
  return t;
}

__device__ __forceinline__ uint32_t hash_u64(unsigned long long k) {
  k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
  return (uint32_t)k;
}

__global__ void vox_clear_kernel(VoxTable t) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < t.cap; i += (size_t)gridDim.x * blockDim.x) {
    t.keys[i] = kVoxEmpty;
    t.cnt[i] = 0;
#pragma unroll
    for (int c = 0; c < 6; ++c) t.sum[i * 6 + c] = 0.0;
  }
}

__global__ void __launch_bounds__(256) vox_insert_kernel(VoxTable t, const float *xyz, const float *rgb,
                                                         const uint32_t *count, const float *bounds,
                                                         double vs, int32_t *err) {
  const uint32_t M = *count;
  if (M == 0) return;
  const double half = vs * 0.5;
  const double vmin0 = (double)bounds[0] - half, vmin1 = (double)bounds[1] - half, vmin2 = (double)bounds[2] - half;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < M; i += gridDim.x * blockDim.x) {
    const float x = xyz[3 * (size_t)i], y = xyz[3 * (size_t)i + 1], z = xyz[3 * (size_t)i + 2];
    const double i0 = floor(((double)x - vmin0) / vs);
    const double i1 = floor(((double)y - vmin1) / vs);
    const double i2 = floor(((double)z - vmin2) / vs);
    const double lim = (double)(1 << kVoxBits);
    if (!(i0 >= 0.0 && i0 < lim && i1 >= 0.0 && i1 < lim && i2 >= 0.0 && i2 < lim)) {
      *err = 1;  // "voxel_size is too small." (or non-finite coordinates)
      continue;
    }
    const unsigned long long key = ((unsigned long long)i0 << (2 * kVoxBits)) |
                                   ((unsigned long long)i1 << kVoxBits) | (unsigned long long)i2;
    uint32_t slot = hash_u64(key) & (t.cap - 1);
    while (true) {
      unsigned long long prev = atomicCAS(&t.keys[slot], kVoxEmpty, key);
      if (prev == kVoxEmpty || prev == key) break;
      slot = (slot + 1) & (t.cap - 1);
    }
    double *s = t.sum + (size_t)slot * 6;
    atomicAdd(s + 0, (double)x); atomicAdd(s + 1, (double)y); atomicAdd(s + 2, (double)z);
    atomicAdd(s + 3, (double)rgb[3 * (size_t)i]);
    atomicAdd(s + 4, (double)rgb[3 * (size_t)i + 1]);
    atomicAdd(s + 5, (double)rgb[3 * (size_t)i + 2]);
    atomicAdd(&t.cnt[slot], 1u);
  }
}

__global__ void __launch_bounds__(256) vox_extract_kernel(VoxTable t, float *oxyz, float *orgb, int32_t *oidx,
                                                          uint32_t *ocount, const int32_t *err) {
  const bool failed = *err != 0;
  const int lane = threadIdx.x & 31;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  // cap is a multiple of the warp size, so whole warps stay converged in this loop
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < t.cap; i += stride) {
    const unsigned long long key = t.keys[i];
    const bool occ = key != kVoxEmpty;
    const unsigned m = __ballot_sync(0xffffffffu, occ && !failed);
    uint32_t base = 0;
    if (m) {
      if (lane == (__ffs(m) - 1)) base = atomicAdd(ocount, (uint32_t)__popc(m));
      base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
    }
    if (occ) {
      if (!failed) {
        const uint32_t row = base + __popc(m & ((1u << lane) - 1u));
        const double c = (double)t.cnt[i];
        const double *s = t.sum + i * 6;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          oxyz[3 * (size_t)row + k] = (float)(s[k] / c);
          orgb[3 * (size_t)row + k] = (float)(s[3 + k] / c);
        }
        if (oidx) {
          oidx[3 * (size_t)row + 0] = (int32_t)(key >> (2 * kVoxBits));
          oidx[3 * (size_t)row + 1] = (int32_t)((key >> kVoxBits) & ((1u << kVoxBits) - 1u));
          oidx[3 * (size_t)row + 2] = (int32_t)(key & ((1u << kVoxBits) - 1u));
        }
      }
      t.keys[i] = kVoxEmpty;  // leave the table clean for the next frame
      t.cnt[i] = 0;
#pragma unroll
      for (int k = 0; k < 6; ++k) t.sum[i * 6 + k] = 0.0;
    }
  }
}

}  // namespace d2pc

using namespace d2pc;

extern "C" int d2pc_voxel_table_bytes(const D2pcConfig *cfg, size_t *bytes) {
  int rc = validate_config(cfg);
  if (rc) return rc;
  if (!bytes) return D2PC_ERR_INVALID_ARGUMENT;
  *bytes = vox_bytes(vox_capacity(make_geom(*cfg).N));
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
  const Geom g = make_geom(*cfg);
  const uint32_t cap = vox_capacity(g.N);
  if (table_bytes < vox_bytes(cap)) return D2PC_ERR_WORKSPACE_TOO_SMALL;
  cudaStream_t st = (cudaStream_t)stream;
  VoxTable t = vox_table(d_table, cap);
  cudaError_t e = cudaMemsetAsync(d_vox_count, 0, sizeof(uint32_t) * cfg->batch, st);
  if (e != cudaSuccess) return record_cuda_error(e);
  e = cudaMemsetAsync(d_vox_error, 0, sizeof(int32_t) * cfg->batch, st);
  if (e != cudaSuccess) return record_cuda_error(e);
  const int blocks = 148 * 8;
  vox_clear_kernel<<<blocks, 256, 0, st>>>(t);
  D2PC_CHECK_LAUNCH();
  for (int b = 0; b < cfg->batch; ++b) {
    const size_t ro = (size_t)b * g.N * 3;
    vox_insert_kernel<<<blocks, 256, 0, st>>>(t, d_xyz + ro, d_rgb + ro, d_count + b, d_bounds + 6 * b,
                                              voxel_size, d_vox_error + b);
    D2PC_CHECK_LAUNCH();
    vox_extract_kernel<<<blocks, 256, 0, st>>>(t, d_vox_xyz + ro, d_vox_rgb + ro,
                                               d_vox_idx ? d_vox_idx + ro : nullptr, d_vox_count + b,
                                               d_vox_error + b);
    D2PC_CHECK_LAUNCH();
  }
  return D2PC_OK;
}
