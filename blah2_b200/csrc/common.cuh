// common.cuh -- error plumbing and small helpers shared by the C-ABI translation units.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <string>

#include "../../include/b200dd.h"

namespace b2 {

void set_last_error(const std::string &msg);

inline int cuda_fail(cudaError_t e, const char *what, const char *file, int line) {
  char buf[512];
  snprintf(buf, sizeof(buf), "%s failed at %s:%d: %s (%s)", what, file, line, cudaGetErrorName(e),
           cudaGetErrorString(e));
  set_last_error(buf);
  return B200DD_ERR_CUDA;
}

#define B2_CUDA(expr)                                                        \
  do {                                                                       \
    cudaError_t _e = (expr);                                                 \
    if (_e != cudaSuccess) return b2::cuda_fail(_e, #expr, __FILE__, __LINE__); \
  } while (0)

#define B2_LAUNCH_CHECK()                                                              \
  do {                                                                                 \
    cudaError_t _e = cudaGetLastError();                                               \
    if (_e != cudaSuccess) return b2::cuda_fail(_e, "kernel launch", __FILE__, __LINE__); \
  } while (0)

inline int arg_fail(const char *msg) {
  set_last_error(msg);
  return B200DD_ERR_ARG;
}
inline int geom_fail(const std::string &msg) {
  set_last_error(msg);
  return B200DD_ERR_GEOMETRY;
}

// RAII device switch (handles remember their device; callers may sit on another one)
struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
    if (dev >= 0 && dev != prev) ok = cudaSetDevice(dev) == cudaSuccess;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

template <class T> inline void free_dev(T *&p) {
  if (p) cudaFree(p);
  p = nullptr;
}

}  // namespace b2
