// Micro-benchmark: issue rate and dependent latency of scalar vs packed FP32 and of FP64 on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp_rates fp_rates.cu && ./fp_rates
#include <cstdio>
#include <cuda_runtime.h>

#define OP_FADD(a, b)  asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a.x) : "f"(b.x));
#define OP_FFMA(a, b)  asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(a.x) : "f"(b.x));
#define OP_FADD2(a, b) asm volatile("{ .reg .b64 ra, rb; mov.b64 ra, {%0,%1}; mov.b64 rb, {%2,%3}; add.rn.f32x2 ra, ra, rb; mov.b64 {%0,%1}, ra; }" : "+f"(a.x), "+f"(a.y) : "f"(b.x), "f"(b.y));
#define OP_FFMA2(a, b) asm volatile("{ .reg .b64 ra, rb; mov.b64 ra, {%0,%1}; mov.b64 rb, {%2,%3}; fma.rn.f32x2 ra, ra, rb, rb; mov.b64 {%0,%1}, ra; }" : "+f"(a.x), "+f"(a.y) : "f"(b.x), "f"(b.y));
#define OP_DADD(a, b)  asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(a) : "d"(b));
#define OP_DFMA(a, b)  asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(a) : "d"(b));

#define KERNEL_F(NAME, OP)                                                          \
  __global__ void NAME(float2 *out, int iters, int chains, long long *clk) {        \
    float2 a[8], b = make_float2(1.0000001f, 0.9999999f);                           \
    for (int i = 0; i < 8; i++) a[i] = make_float2(threadIdx.x * 1e-3f + i, 1.0f);  \
    long long t0 = clock64();                                                       \
    if (chains == 8) {                                                              \
      for (int it = 0; it < iters; it++) {                                          \
        _Pragma("unroll") for (int i = 0; i < 8; i++) { OP(a[i], b) }               \
      }                                                                             \
    } else {                                                                        \
      for (int it = 0; it < iters; it++) {                                          \
        _Pragma("unroll") for (int i = 0; i < 8; i++) { OP(a[0], b) }               \
      }                                                                             \
    }                                                                               \
    long long t1 = clock64();                                                       \
    float2 s = a[0];                                                                \
    for (int i = 1; i < 8; i++) { s.x += a[i].x; s.y += a[i].y; }                   \
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;                                 \
    if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;                        \
  }
#define KERNEL_D(NAME, OP)                                                          \
  __global__ void NAME(float2 *out, int iters, int chains, long long *clk) {        \
    double a[8], b = 1.0000001;                                                     \
    for (int i = 0; i < 8; i++) a[i] = threadIdx.x * 1e-3 + i;                      \
    long long t0 = clock64();                                                       \
    if (chains == 8) {                                                              \
      for (int it = 0; it < iters; it++) {                                          \
        _Pragma("unroll") for (int i = 0; i < 8; i++) { OP(a[i], b) }               \
      }                                                                             \
    } else {                                                                        \
      for (int it = 0; it < iters; it++) {                                          \
        _Pragma("unroll") for (int i = 0; i < 8; i++) { OP(a[0], b) }               \
      }                                                                             \
    }                                                                               \
    long long t1 = clock64();                                                       \
    double s = 0;                                                                   \
    for (int i = 0; i < 8; i++) s += a[i];                                          \
    out[blockIdx.x * blockDim.x + threadIdx.x] = make_float2((float)s, 0.f);        \
    if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;                        \
  }
KERNEL_F(k_fadd, OP_FADD)
KERNEL_F(k_ffma, OP_FFMA)
KERNEL_F(k_fadd2, OP_FADD2)
KERNEL_F(k_ffma2, OP_FFMA2)
KERNEL_D(k_dadd, OP_DADD)
KERNEL_D(k_dfma, OP_DFMA)

typedef void (*kern_t)(float2 *, int, int, long long *);

int main() {
  float2 *out; long long *clk;
  cudaMalloc(&out, sizeof(float2) * 1024 * 2048);
  cudaMallocManaged(&clk, sizeof(long long));
  struct { const char *name; kern_t k; } ks[] = {{"FADD", k_fadd}, {"FADD2", k_fadd2}, {"FFMA", k_ffma},
                                                 {"FFMA2", k_ffma2}, {"DADD", k_dadd}, {"DFMA", k_dfma}};
  const int iters = 4096;
  for (auto &e : ks) {
    // latency: one warp, one dependent chain
    e.k<<<1, 32>>>(out, iters, 1, clk); cudaDeviceSynchronize();
    e.k<<<1, 32>>>(out, iters, 1, clk); cudaDeviceSynchronize();
    double lat = double(*clk) / (iters * 8.0);
    printf("%-6s latency %.2f clk;", e.name, lat);
    // throughput per SM sub-partition: warps per SMSP = 1, 2, 4, 8 (one CTA of 128..1024 threads on one SM)
    for (int warps = 4; warps <= 32; warps *= 2) {
      e.k<<<1, warps * 32>>>(out, iters, 8, clk); cudaDeviceSynchronize();
      e.k<<<1, warps * 32>>>(out, iters, 8, clk); cudaDeviceSynchronize();
      double ipc = (iters * 8.0 * warps / 4.0) / double(*clk);  // warp instructions per clk per SMSP
      printf("  %2d warps/SM: %.3f inst/clk/SMSP", warps, ipc);
    }
    printf("\n");
  }
  return 0;
}
