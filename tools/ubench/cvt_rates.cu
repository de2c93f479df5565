// Micro-benchmark (round 2): throughput of float <-> double conversions (F2F.F64.F32 / F2F.F32.F64) against an
// integer-ALU widening of a float to a double (exact for normal numbers and zero; denormals flush to zero) and
// against DFMA, per SM.  The WienerHopf kernels convert 32-96 values per thread per transform.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/ubench/cvt_rates tools/ubench/cvt_rates.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ double widen_int(float f) {
  const unsigned b = __float_as_uint(f);
  const unsigned e = b & 0x7f800000u;
  const unsigned hi = (b & 0x80000000u) | (((b & 0x7fffffffu) >> 3) + 0x38000000u);
  const unsigned lo = b << 29;
  return e ? __hiloint2double((int)hi, (int)lo) : 0.0;
}

template <int MODE> __global__ void __launch_bounds__(256) k(int iters, const float *in, double *out) {
  float f[8];
#pragma unroll
  for (int j = 0; j < 8; j++) f[j] = in[threadIdx.x + 256 * j];
  double acc = 0.0;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int j = 0; j < 8; j++) {
      double d;
      if (MODE == 0) d = (double)f[j];
      else if (MODE == 1) d = widen_int(f[j]);
      else d = 1.0;
      if (MODE == 2) { acc = fma(acc, 1.0000001, (double)j); }
      else if (MODE == 3) { f[j] = (float)(acc + j); acc += 1e-9; }   // double -> float
      else acc += d;
      if (MODE != 3) f[j] = f[j] * 1.0000001f + 1e-7f;  // a fresh value every iteration keeps the conversion in the loop
    }
  }
  out[blockIdx.x * 256 + threadIdx.x] = acc + f[0];
}

template <int MODE> float run(int sms, const float *in, double *out, int iters) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<sms * 4, 256>>>(16, in, out);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<MODE><<<sms * 4, 256>>>(iters, in, out);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float *in; double *out;
  cudaMalloc(&in, sizeof(float) * 2048);
  cudaMemset(in, 0x3f, sizeof(float) * 2048);
  cudaMalloc(&out, sizeof(double) * sms * 4 * 256);
  const int iters = 4000;
  const double ops = 8.0 * iters * sms * 4 * 256;
  const float t0 = run<0>(sms, in, out, iters), t1 = run<1>(sms, in, out, iters), t2 = run<2>(sms, in, out, iters), t3 = run<3>(sms, in, out, iters);
  const double clk = 1.965e9;
  printf("per SM per clock (at 1965 MHz), loop of [op + DADD + LOP3]:\n");
  printf("  F2F.F64.F32 + DADD : %.3f ms  -> %.1f values/clk/SM\n", t0, ops / (t0 * 1e-3) / clk / sms);
  printf("  integer widen + DADD: %.3f ms  -> %.1f values/clk/SM\n", t1, ops / (t1 * 1e-3) / clk / sms);
  printf("  DFMA only          : %.3f ms  -> %.1f values/clk/SM\n", t2, ops / (t2 * 1e-3) / clk / sms);
  printf("  DADD + F2F.F32.F64 : %.3f ms  -> %.1f values/clk/SM   (%s)\n", t3, ops / (t3 * 1e-3) / clk / sms, cudaGetErrorString(cudaGetLastError()));
  return 0;
}
