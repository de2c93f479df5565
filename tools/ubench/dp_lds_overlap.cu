// Micro-benchmark for round 2: do FP64 arithmetic and 16-byte shared-memory accesses overlap on one SM?
// (profiles/r01_summary.md s8: the shared-memory FP64 FFT takes the SUM of its FP64-pipe time and its
// shared-memory-port time.)  One CTA of 256 threads per SM; mode 0: every warp runs independent DFMA chains,
// mode 1: every warp streams LDS.128 / STS.128 through its own shared-memory slab, mode 2: warps 0-3 do the
// arithmetic of mode 0 and warps 4-7 the traffic of mode 1 (half the work of each).  If the two overlap,
// t(mode 2) ~ max(t0, t1) / 2; if they share a dispatch port, t(mode 2) ~ (t0 + t1) / 2.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/ubench/dp_lds_overlap \
//        tools/ubench/dp_lds_overlap.cu && tools/ubench/dp_lds_overlap
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dp_work(int iters, double &sink) {
  double a[8];
#pragma unroll
  for (int k = 0; k < 8; k++) a[k] = 1.0 + 1e-9 * (threadIdx.x + k);
  const double b = 1.0000001, c = -1e-7;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 8; k++) a[k] = fma(a[k], b, c);  // 8 independent chains: the pipe, not the latency, limits
  }
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < 8; k++) s += a[k];
  sink = s;
}

__device__ __forceinline__ void lds_work(double2 *slab, int iters, double &sink) {
  // each warp owns 32 x 17 double2 (padded rows): 8 loads + 8 stores of 16 bytes per iteration, conflict-free
  const int lane = threadIdx.x & 31;
  double2 acc = make_double2(0.0, 0.0);
  for (int i = 0; i < iters; i++) {
    double2 v[8];
#pragma unroll
    for (int k = 0; k < 8; k++) v[k] = slab[k * 33 + lane];
#pragma unroll
    for (int k = 0; k < 8; k++) { acc.x += v[k].x; slab[k * 33 + ((lane + 1) & 31)] = make_double2(v[k].y, acc.x); }
    __syncwarp();
  }
  sink = acc.x + acc.y;
}

__global__ void __launch_bounds__(256) k_mix(int mode, int dp_iters, int lds_iters, double *out) {
  __shared__ double2 smem[8][8 * 33];
  const int w = threadIdx.x >> 5;
  for (int i = threadIdx.x & 31; i < 8 * 33; i += 32) smem[w][i] = make_double2(1e-3 * i, 2e-3 * i);
  __syncthreads();
  double sink = 0.0;
  if (mode == 0) dp_work(dp_iters, sink);
  else if (mode == 1) lds_work(smem[w], lds_iters, sink);
  else if (w < 4) dp_work(dp_iters, sink);
  else lds_work(smem[w], lds_iters, sink);
  out[blockIdx.x * 256 + threadIdx.x] = sink;
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double *out;
  cudaMalloc(&out, sizeof(double) * sms * 256);
  const int dp_iters = 4000, lds_iters = 4000;
  float t[3];
  for (int mode = 0; mode < 3; mode++) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_mix<<<sms, 256>>>(mode, 10, 10, out);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k_mix<<<sms, 256>>>(mode, dp_iters, lds_iters, out);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&t[mode], e0, e1);
  }
  printf("DFMA only (8 warps): %.3f ms   LDS/STS.128 only (8 warps): %.3f ms   4 + 4 warps mixed: %.3f ms\n", t[0], t[1], t[2]);
  printf("overlap would give ~%.3f ms, a shared dispatch port ~%.3f ms   (%s)\n", 0.5f * (t[0] > t[1] ? t[0] : t[1]),
         0.5f * (t[0] + t[1]), cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
  return 0;
}
