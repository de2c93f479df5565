// Micro-benchmark: cost of a CTA barrier step (LDS -> FP64 chain -> STS -> __syncthreads) as used by the
// Toeplitz solve, for different warp counts; DFMA with three distinct source registers.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_bar(int iters, long long *clk) {
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) *clk = t1 - t0;
}

// each step: read neighbour's value from smem, `depth` dependent DFMAs, write, barrier
__global__ void k_step(int iters, int depth, double *out, long long *clk) {
  __shared__ double2 buf[2][1024];
  const int tid = threadIdx.x, n = blockDim.x;
  buf[0][tid] = make_double2(1.0 + tid * 1e-6, 0.5);
  buf[1][tid] = make_double2(0.0, 0.0);
  __syncthreads();
  double2 acc = make_double2(0.0, 0.0);
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
    const double2 v = buf[i & 1][tid ? tid - 1 : n - 1];
    double2 w = v;
    for (int d = 0; d < depth; d++) {
      w.x = fma(w.x, 0.999999, acc.y * 1e-9);
      w.y = fma(w.y, 0.999998, acc.x * 1e-9);
    }
    acc.x += w.x; acc.y += w.y;
    buf[(i + 1) & 1][tid] = w;
    __syncthreads();
  }
  long long t1 = clock64();
  out[tid] = acc.x + acc.y;
  if (tid == 0) *clk = t1 - t0;
}

__global__ void k_dfma3(double *out, int iters, long long *clk) {
  double a[8], b = 1.0000001 + threadIdx.x * 1e-9, c = 0.9999999 - threadIdx.x * 1e-9;
  for (int i = 0; i < 8; i++) a[i] = threadIdx.x * 1e-3 + i;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(a[i]) : "d"(b), "d"(c));
  }
  long long t1 = clock64();
  double s = 0;
  for (int i = 0; i < 8; i++) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) *clk = t1 - t0;
}

int main() {
  double *out; long long *clk;
  cudaMalloc(&out, sizeof(double) * 4096);
  cudaMallocManaged(&clk, sizeof(long long));
  const int iters = 4096;
  for (int threads : {32, 128, 416, 512, 1024}) {
    k_bar<<<1, threads>>>(iters, clk); cudaDeviceSynchronize();
    k_bar<<<1, threads>>>(iters, clk); cudaDeviceSynchronize();
    printf("barrier only, %4d threads: %.1f clk/iter\n", threads, double(*clk) / iters);
  }
  for (int threads : {128, 416, 1024})
    for (int depth : {0, 1, 3, 6, 12}) {
      k_step<<<1, threads>>>(iters, depth, out, clk); cudaDeviceSynchronize();
      k_step<<<1, threads>>>(iters, depth, out, clk); cudaDeviceSynchronize();
      printf("step (LDS.128 -> %2d dependent DFMA pairs -> STS.128 -> barrier), %4d threads: %.1f clk/step\n", depth, threads,
             double(*clk) / iters);
    }
  for (int warps : {4, 16, 32}) {
    k_dfma3<<<1, warps * 32>>>(out, iters, clk); cudaDeviceSynchronize();
    k_dfma3<<<1, warps * 32>>>(out, iters, clk); cudaDeviceSynchronize();
    printf("DFMA 3 distinct sources, %2d warps/SM: %.3f inst/clk/SMSP\n", warps, (iters * 8.0 * warps / 4.0) / double(*clk));
  }
  return 0;
}
