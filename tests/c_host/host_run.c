/* A plain C99 host that runs the hot path on the GPU through include/d2pc.h and the CUDA runtime only
 * (no Python, no torch): the binding a non-Python maintainer would write.
 *   host_run IN OUT
 * IN : int32 H, W, C, h, w, step, invert; float64 depth_scale, cx, cy, f; float32 depth[h*w]; uint8 image[H*W*C]
 * OUT: uint32 n; float32 xyz[n*3]; float32 rgb[n*3]                                                         */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <cuda_runtime_api.h>
#include "d2pc.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 2; } } while (0)
#define DK(x) do { int r_ = (x); if (r_ != D2PC_OK) { fprintf(stderr, "%s: %s (%s)\n", #x, d2pc_error_string(r_), d2pc_last_cuda_error()); return 3; } } while (0)

int main(int argc, char **argv) {
  int32_t hd[7];
  double dd[4];
  FILE *fi, *fo;
  float *depth, *xyz, *rgb, *d_depth, *d_xyz, *d_rgb;
  uint8_t *img, *d_img;
  void *d_ws;
  int32_t *d_status, *d_any, any = 0;
  uint32_t *d_count, n = 0;
  size_t ws_bytes = 0, np_, nd, ni, rows;
  D2pcConfig cfg;
  cudaStream_t st;
  if (argc != 3) return 1;
  fi = fopen(argv[1], "rb");
  if (!fi || fread(hd, 4, 7, fi) != 7 || fread(dd, 8, 4, fi) != 4) return 1;
  nd = (size_t)hd[3] * hd[4];
  ni = (size_t)hd[0] * hd[1] * hd[2];
  depth = (float *)malloc(nd * 4);
  img = (uint8_t *)malloc(ni);
  if (fread(depth, 4, nd, fi) != nd || fread(img, 1, ni, fi) != ni) return 1;
  fclose(fi);
  memset(&cfg, 0, sizeof cfg);
  cfg.batch = 1; cfg.img_h = hd[0]; cfg.img_w = hd[1]; cfg.img_c = hd[2]; cfg.dep_h = hd[3]; cfg.dep_w = hd[4];
  cfg.step = hd[5]; cfg.invert = hd[6]; cfg.depth_scale = dd[0]; cfg.cx = dd[1]; cfg.cy = dd[2]; cfg.f = dd[3];
  rows = (size_t)((hd[0] + hd[5] - 1) / hd[5]) * (size_t)((hd[1] + hd[5] - 1) / hd[5]);
  np_ = rows * 3 * sizeof(float);
  DK(d2pc_workspace_bytes(&cfg, &ws_bytes));
  CK(cudaStreamCreate(&st));
  CK(cudaMalloc((void **)&d_depth, nd * 4));
  CK(cudaMalloc((void **)&d_img, ni));
  CK(cudaMalloc(&d_ws, ws_bytes));
  CK(cudaMalloc((void **)&d_xyz, np_));
  CK(cudaMalloc((void **)&d_rgb, np_));
  CK(cudaMalloc((void **)&d_count, 4));
  CK(cudaMalloc((void **)&d_status, 4));
  CK(cudaMalloc((void **)&d_any, 4));
  CK(cudaMemcpyAsync(d_depth, depth, nd * 4, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d_img, img, ni, cudaMemcpyHostToDevice, st));
  DK(d2pc_stats_enqueue(&cfg, d_depth, d_ws, ws_bytes, st));
  DK(d2pc_frame_status(&cfg, d_ws, d_status, d_any, st));
  DK(d2pc_emit_enqueue(&cfg, d_depth, d_img, d_ws, ws_bytes, d_xyz, d_rgb, d_count, NULL, st));
  CK(cudaMemcpyAsync(&any, d_any, 4, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  if (any) {  /* frames the fast selection declined (NaN / inf, degenerate maps): exact path, emit again */
    DK(d2pc_stats_fallback_enqueue(&cfg, d_depth, d_ws, ws_bytes, st));
    DK(d2pc_emit_enqueue(&cfg, d_depth, d_img, d_ws, ws_bytes, d_xyz, d_rgb, d_count, NULL, st));
  }
  xyz = (float *)malloc(np_);
  rgb = (float *)malloc(np_);
  CK(cudaMemcpyAsync(&n, d_count, 4, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(xyz, d_xyz, np_, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(rgb, d_rgb, np_, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  fo = fopen(argv[2], "wb");
  if (!fo) return 1;
  fwrite(&n, 4, 1, fo);
  fwrite(xyz, 12, n, fo);
  fwrite(rgb, 12, n, fo);
  fclose(fo);
  printf("rows %u fallback %d\n", n, any);
  return 0;
}
