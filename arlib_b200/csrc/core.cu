// core.cu -- version / error plumbing of libagcf.
#include "common.cuh"
#include <stdio.h>

namespace agcf {
thread_local int g_last_cuda_error = 0;
thread_local const char* g_last_cuda_file = "";
thread_local int g_last_cuda_line = 0;
}

extern "C" int agcf_abi_version(void) { return 4; }

extern "C" const char* agcf_strerror(int code) {
  switch (code) {
    case AGCF_OK: return "ok";
    case AGCF_EINVAL: return "invalid argument";
    case AGCF_EUNSUPPORTED: return "unsupported shape";
    case AGCF_ECUDA: return "CUDA runtime error";
    case AGCF_EWORKSPACE: return "workspace too small";
  }
  return "unknown error";
}

extern "C" int agcf_last_cuda_error(void) { return agcf::g_last_cuda_error; }

extern "C" const char* agcf_last_cuda_error_where(void) {
  static thread_local char buf[256];
  snprintf(buf, sizeof(buf), "%s (%s:%d)", cudaGetErrorString((cudaError_t)agcf::g_last_cuda_error),
           agcf::g_last_cuda_file, agcf::g_last_cuda_line);
  return buf;
}

extern "C" int agcf_device_sm_count(void) {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
  return sms;
}
