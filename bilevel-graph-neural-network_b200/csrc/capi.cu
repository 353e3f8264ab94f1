// Library-level entry points: ABI version, error strings, launch counter.
#include "common.cuh"

namespace bignn {
long long g_launch_count = 0;

int sm_count() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached = n;
    else
      cached = 148;  // B200
  }
  return cached;
}
}  // namespace bignn

extern "C" int bignn_abi_version(void) { return BIGNN_ABI_VERSION; }

extern "C" const char* bignn_error_string(int code) {
  switch (code) {
    case 0: return "success";
    case BIGNN_EINVAL: return "bignn: invalid argument (size, null pointer or unsupported flag)";
    case BIGNN_EALIGN: return "bignn: pointer or leading dimension not aligned as required";
    case BIGNN_EWORKSPACE: return "bignn: workspace missing or too small";
    default: break;
  }
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  return "bignn: unknown error code";
}

extern "C" int64_t bignn_launch_count(void) { return (int64_t)bignn::g_launch_count; }
