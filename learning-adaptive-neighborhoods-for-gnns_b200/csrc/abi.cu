// ABI bookkeeping: version, error strings, last CUDA error.
#include "common.cuh"
#include <cstdlib>

namespace dggb {
int g_last_cuda_error = 0;
long long g_kernel_launches = 0;
bool pdl_enabled() {
  static const bool on = std::getenv("DGGB_NO_PDL") == nullptr;
  return on;
}
}

extern "C" int dggb_version(void) { return 2; }
extern "C" int dggb_last_cuda_error(void) { return dggb::g_last_cuda_error; }
extern "C" int dggb_build_arch(void) { return 1000; }
extern "C" long long dggb_kernel_launches(void) { return dggb::g_kernel_launches; }
extern "C" const char* dggb_error_string(int status) {
  switch (status) {
    case DGGB_OK: return "ok";
    case DGGB_ERR_BAD_ARG: return "bad argument (null pointer or negative size)";
    case DGGB_ERR_BAD_SHAPE: return "unsupported shape";
    case DGGB_ERR_UNSUPPORTED: return "unsupported mode";
    case DGGB_ERR_K_OVERFLOW: return "a row needed more than Kcap selected entries";
    case DGGB_ERR_CUDA: return "CUDA error (see dggb_last_cuda_error)";
    case DGGB_ERR_WORKSPACE: return "workspace too small";
    default: return "unknown status";
  }
}
