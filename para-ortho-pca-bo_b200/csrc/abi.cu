// ABI bookkeeping: version, error strings, last CUDA error.
#include "common.cuh"

static thread_local char g_last_cuda_error[256] = "";

const char *pcabo_set_cuda_error(cudaError_t e) {
  const char *s = cudaGetErrorString(e);
  int i = 0;
  for (; s[i] && i < 255; ++i) g_last_cuda_error[i] = s[i];
  g_last_cuda_error[i] = 0;
  return g_last_cuda_error;
}

extern "C" int pcabo_abi_version(void) { return PCABO_ABI_VERSION; }

extern "C" const char *pcabo_error_string(int code) {
  switch (code) {
    case PCABO_OK: return "ok";
    case PCABO_E_INVALID: return "invalid argument";
    case PCABO_E_UNSUPPORTED: return "size not supported by the sm_100a kernels";
    case PCABO_E_CUDA: return "CUDA runtime error";
    default: return "unknown error";
  }
}

extern "C" const char *pcabo_last_cuda_error(void) { return g_last_cuda_error; }
