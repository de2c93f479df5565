// common.cu -- error text and device probes of the C ABI.
#include "common.cuh"

#include <cstring>

namespace b2 {
static thread_local std::string g_last_error;
void set_last_error(const std::string &msg) { g_last_error = msg; }
}  // namespace b2

extern "C" {

const char *b200dd_last_error(void) { return b2::g_last_error.c_str(); }

int b200dd_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int b200dd_device_name(int device, char *buf, int buflen) {
  if (!buf || buflen <= 0) return b2::arg_fail("b200dd_device_name: null buffer");
  cudaDeviceProp prop;
  B2_CUDA(cudaGetDeviceProperties(&prop, device));
  strncpy(buf, prop.name, (size_t)buflen - 1);
  buf[buflen - 1] = 0;
  return B200DD_OK;
}

}  // extern "C"
