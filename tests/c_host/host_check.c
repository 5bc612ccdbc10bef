/* A plain C99 host of the C ABI (include/d2pc.h): what a non-Python maintainer would write.
 * Calls only the host-side entry points (sizes, validation, error strings) -- no GPU needed.
 * Prints "key value" lines that tests/test_abi.py compares with the ctypes binding. */
#include <stdio.h>
#include <stddef.h>
#include <string.h>
#include "d2pc.h"

int main(void) {
  D2pcConfig cfg;
  size_t ws = 0, table = 0, smooth = 0, sor = 0, text = 0;
  int rc;
  memset(&cfg, 0, sizeof cfg);
  cfg.batch = 2; cfg.img_h = 480; cfg.img_w = 640; cfg.img_c = 3; cfg.dep_h = 518; cfg.dep_w = 686;
  cfg.step = 1; cfg.invert = 1; cfg.depth_scale = 10.0; cfg.cx = 320.0; cfg.cy = 240.0; cfg.f = 768.0;
  printf("abi_version %d\n", d2pc_abi_version());
  printf("sizeof_config %zu\n", sizeof(D2pcConfig));
  printf("sizeof_frame_params %zu\n", sizeof(D2pcFrameParams));
  rc = d2pc_workspace_bytes(&cfg, &ws);
  printf("workspace_rc %d\nworkspace_bytes %zu\n", rc, ws);
  rc = d2pc_voxel_table_bytes(&cfg, &table);
  printf("voxel_table_rc %d\nvoxel_table_bytes %zu\n", rc, table);
  rc = d2pc_smooth_scratch_bytes(&cfg, &smooth);
  printf("smooth_rc %d\nsmooth_bytes %zu\n", rc, smooth);
  rc = d2pc_sor_scratch_bytes(100000, &sor);
  printf("sor_rc %d\nsor_bytes %zu\n", rc, sor);
  rc = d2pc_xyz_text_scratch_bytes(100000, &text);
  printf("text_rc %d\ntext_bytes %zu\n", rc, text);
  cfg.step = 3;
  printf("bad_step_rc %d\n", d2pc_workspace_bytes(&cfg, &ws));
  cfg.step = 1;
  printf("null_rc %d\n", d2pc_stats_enqueue(&cfg, NULL, NULL, 0, NULL));
  printf("err1 %s\n", d2pc_error_string(D2PC_ERR_INVALID_ARGUMENT));
  return 0;
}
