// d2pc_api.cu -- argument validation, error reporting and workspace sizing of the C ABI
// (include/d2pc.h).
#include <stdio.h>
#include <string.h>

#include "d2pc_device.cuh"

namespace d2pc {

static thread_local char g_cuda_err[256] = "";

int record_cuda_error(cudaError_t e) {
  snprintf(g_cuda_err, sizeof(g_cuda_err), "%s: %s", cudaGetErrorName(e), cudaGetErrorString(e));
  return D2PC_ERR_CUDA;
}

int validate_config(const D2pcConfig *c) {
  if (!c) return D2PC_ERR_INVALID_ARGUMENT;
  if (c->batch < 1 || c->img_h < 1 || c->img_w < 1 || c->dep_h < 1 || c->dep_w < 1) return D2PC_ERR_INVALID_ARGUMENT;
  if (!(c->img_c == 1 || c->img_c == 3 || c->img_c == 4)) return D2PC_ERR_INVALID_ARGUMENT;
  if (!(c->step == 1 || c->step == 2 || c->step == 4)) return D2PC_ERR_INVALID_ARGUMENT;
  if ((unsigned long long)c->img_h * (unsigned long long)c->img_w >= (1ull << 31)) return D2PC_ERR_INVALID_ARGUMENT;
  if ((unsigned long long)c->dep_h * (unsigned long long)c->dep_w >= (1ull << 31)) return D2PC_ERR_INVALID_ARGUMENT;
  const bool resized = !(c->dep_h == c->img_h && c->dep_w == c->img_w);
  if (!(c->f == c->f) || c->f == 0.0) return D2PC_ERR_INVALID_ARGUMENT;
  return D2PC_OK;
}

}  // namespace d2pc

using namespace d2pc;

extern "C" int d2pc_abi_version(void) { return D2PC_ABI_VERSION; }

extern "C" const char *d2pc_error_string(int code) {
  switch (code) {
    case D2PC_OK: return "ok";
    case D2PC_ERR_INVALID_ARGUMENT: return "invalid argument";
    case D2PC_ERR_WORKSPACE_TOO_SMALL: return "workspace too small";
    case D2PC_ERR_CUDA: return "CUDA error";
    case D2PC_ERR_UNSUPPORTED: return "unsupported input (depth map with a 1-pixel side that needs resizing)";
    default: return "unknown error";
  }
}

extern "C" const char *d2pc_last_cuda_error(void) { return g_cuda_err; }

extern "C" int d2pc_workspace_bytes(const D2pcConfig *cfg, size_t *bytes) {
  int rc = validate_config(cfg);
  if (rc) return rc;
  if (!bytes) return D2PC_ERR_INVALID_ARGUMENT;
  *bytes = make_layout(*cfg).total;
  return D2PC_OK;
}
