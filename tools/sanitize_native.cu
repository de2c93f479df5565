// Native target for compute-sanitizer racecheck / synccheck (no Python, no PyTorch: the sanitizer instruments only
// this library's kernels).  Drives the float2 DEVICE path of the C ABI -- the path with the TMA-staged kernels -- on a
// small synthetic CPI: WienerHopf (correlation with staged + wrapped windows, solve, persistent filter kernel),
// the range / Doppler kernels, set_metrics and the detection chain, eager and as a replayed CUDA graph; then the chunk
// mode of the filter.  Prints a checksum of the map so that two runs can be compared.
//   nvcc -O2 -std=c++17 -I include -o tools/sanitize_native tools/sanitize_native.cu -L blah2_b200/lib -lb200dd -Xlinker -rpath -Xlinker '$ORIGIN/../blah2_b200/lib'
//   compute-sanitizer --tool racecheck tools/sanitize_native
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_runtime.h>

#include "b200dd.h"

#define CK(x) do { int _rc = (x); if (_rc != B200DD_OK) { printf("FAIL %s -> %d: %s\n", #x, _rc, b200dd_last_error()); return 1; } } while (0)

int main() {
  const uint32_t fs = 200000, n = 40000;
  std::vector<float2> x(n), y(n);
  unsigned s = 12345u;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (float)((int)(s >> 16) % 2001 - 1000); };
  for (uint32_t i = 0; i < n; i++) x[i] = make_float2(rnd(), rnd());
  for (uint32_t i = 0; i < n; i++) {
    float2 v = make_float2(0.5f * x[i].x + 0.1f * rnd(), 0.5f * x[i].y + 0.1f * rnd());
    if (i >= 3) { v.x += 0.2f * x[i - 3].x; v.y += 0.2f * x[i - 3].y; }
    if (i >= 17) { v.x += 0.01f * (x[i - 17].x * cosf(0.03f * i) - x[i - 17].y * sinf(0.03f * i)); v.y += 0.01f * (x[i - 17].x * sinf(0.03f * i) + x[i - 17].y * cosf(0.03f * i)); }
    y[i] = v;
  }
  float2 *dx, *dy, *dmap;
  cudaMalloc(&dx, sizeof(float2) * (n + 1));
  cudaMalloc(&dy, sizeof(float2) * (n + 1));
  b200dd_pipeline_params pp{};
  pp.caf = {-5, 60, -500, 500, fs, n, 1, -1};
  pp.clutter_enable = 1; pp.clutter_delay_min = -5; pp.clutter_delay_max = 30;
  pp.detection_enable = 1; pp.det = {1e-4, 2, 6, 3, 15.0, 4, 4, (double)fs / n, 1, 1, -1};
  b200dd_pipeline *pl;
  CK(b200dd_pipeline_create(&pp, &pl));
  b200dd_caf_geometry g;
  CK(b200dd_pipeline_get_geometry(pl, &g));
  const size_t cells = (size_t)g.n_doppler_bins * g.n_delay_bins;
  cudaMalloc(&dmap, sizeof(float2) * cells);
  std::vector<float2> map(cells);
  std::vector<double> od(4096), of(4096), os(4096);
  for (int off = 0; off < 2; off++) {  // buffers on an even and on an odd float2 boundary (TMA alignment edges)
    cudaMemcpy(dx + off, x.data(), sizeof(float2) * n, cudaMemcpyHostToDevice);
    cudaMemcpy(dy + off, y.data(), sizeof(float2) * n, cudaMemcpyHostToDevice);
    for (int rep = 0; rep < 2; rep++) {
      if (rep == 1) CK(b200dd_pipeline_prepare_device(pl, dx + off, dy + off, n, dmap, nullptr));  // graph replay
      CK(b200dd_pipeline_submit_device(pl, dx + off, dy + off, n, dmap, nullptr));
      b200dd_cpi_result r;
      CK(b200dd_pipeline_fetch(pl, &r, od.data(), of.data(), os.data(), 4096, nullptr));
      cudaMemcpy(map.data(), dmap, sizeof(float2) * cells, cudaMemcpyDeviceToHost);
      double cs = 0;
      for (size_t i = 0; i < cells; i++) cs += (double)map[i].x * ((i % 7) + 1) - (double)map[i].y * ((i % 5) + 1);
      printf("offset %d %s: status %d detections %u noise %.4f max %.4f checksum %.9e\n", off, rep ? "graph" : "eager", r.filter_status,
             r.n_detections, r.noise_power, r.max_power, cs);
    }
  }
  b200dd_pipeline_destroy(pl);
  // chunk mode: two chunks of one signal, summed correlations, replicated solve
  const int32_t dm = -5, dM = 30;
  const uint32_t nb = dM - dm;
  b200dd_wh *c[2];
  float2 *xl[2], *yl[2], *yo[2];
  double *ab[2], *absum;
  cudaMalloc(&absum, sizeof(double) * 4 * nb);
  cudaMemset(absum, 0, sizeof(double) * 4 * nb);
  std::vector<double> hsum(4 * nb, 0.0), tmp(4 * nb);
  for (int r = 0; r < 2; r++) {
    const uint32_t c0 = r * (n / 2), nc = n / 2;
    CK(b200dd_wh_create_chunk(dm, dM, n, c0, nc, -1, &c[r]));
    uint32_t l, rr, yr;
    CK(b200dd_wh_chunk_halos(c[r], &l, &rr, &yr));
    std::vector<float2> hx(l + nc + rr), hy(nc + yr);
    for (size_t j = 0; j < hx.size(); j++) hx[j] = x[(size_t)((long long)c0 - l + (long long)j + n) % n];
    for (size_t j = 0; j < hy.size(); j++) hy[j] = y[(c0 + j) % n];
    cudaMalloc(&xl[r], sizeof(float2) * hx.size());
    cudaMalloc(&yl[r], sizeof(float2) * hy.size());
    cudaMalloc(&yo[r], sizeof(float2) * nc);
    cudaMalloc(&ab[r], sizeof(double) * 4 * nb);
    cudaMemcpy(xl[r], hx.data(), sizeof(float2) * hx.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(yl[r], hy.data(), sizeof(float2) * hy.size(), cudaMemcpyHostToDevice);
    CK(b200dd_wh_chunk_corr_device(c[r], xl[r], yl[r], ab[r], nullptr));
    cudaDeviceSynchronize();
    cudaMemcpy(tmp.data(), ab[r], sizeof(double) * 4 * nb, cudaMemcpyDeviceToHost);
    for (size_t j = 0; j < tmp.size(); j++) hsum[j] += tmp[j];
  }
  cudaMemcpy(absum, hsum.data(), sizeof(double) * 4 * nb, cudaMemcpyHostToDevice);
  double cs = 0;
  for (int r = 0; r < 2; r++) {
    CK(b200dd_wh_chunk_filter_device(c[r], absum, xl[r], yl[r], yo[r], nullptr));
    cudaDeviceSynchronize();
    std::vector<float2> out(n / 2);
    cudaMemcpy(out.data(), yo[r], sizeof(float2) * out.size(), cudaMemcpyDeviceToHost);
    for (size_t i = 0; i < out.size(); i++) cs += (double)out[i].x * ((i % 3) + 1) + (double)out[i].y;
    printf("chunk %d status %d\n", r, b200dd_wh_last_status(c[r]));
    b200dd_wh_destroy(c[r]);
  }
  printf("chunk-mode checksum %.9e  (%s)\nSANITIZE_NATIVE OK\n", cs, cudaGetErrorString(cudaGetLastError()));
  return 0;
}
