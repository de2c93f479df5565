// Stand-alone driver (no Python): the drop-in harness and then the reference library, both dlopen'ed
// RTLD_LOCAL in one process like the tests do.  Usage: dropin_smoke <harness.so> <libblah2ref.so>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <dlfcn.h>
#include <vector>
typedef void *(*create_t)(int32_t, int32_t, int32_t, int32_t, uint32_t, uint32_t, int, int, int32_t, int32_t, double, int,
                          int, int, double, uint32_t);
typedef void (*destroy_t)(void *);
typedef int (*run_t)(void *, const double *, const double *, double *, double *, double *, double *, double *, uint32_t,
                     double *);
static int drive(const char *path, const std::vector<double> &x, const std::vector<double> &y, uint32_t fs, uint32_t n) {
  void *lib = dlopen(path, RTLD_NOW | RTLD_LOCAL);
  if (!lib) { printf("dlopen %s: %s\n", path, dlerror()); return 1; }
  auto create = (create_t)dlsym(lib, "refpath_chain_create");
  auto destroy = (destroy_t)dlsym(lib, "refpath_chain_destroy");
  auto run = (run_t)dlsym(lib, "refpath_chain_run");
  void *c = create(-10, 120, -5000, 5000, fs, n, 1, 1, -10, 60, 1e-5, 2, 6, 5, 15.0, 6);
  if (!c) { printf("%s: create failed\n", path); return 1; }
  const uint32_t cap = 1001 * 131;
  std::vector<double> map(2 * cap), d(cap), f(cap), s(cap);
  double metrics[2], stage[3];
  int nd = run(c, x.data(), y.data(), map.data(), metrics, d.data(), f.data(), s.data(), cap, stage);
  printf("%s: run -> %d detections, noise %.3f, stages %.2f %.2f %.2f ms\n", path, nd, metrics[0], stage[0], stage[1], stage[2]);
  fflush(stdout);
  destroy(c);
  printf("%s: destroyed\n", path);
  fflush(stdout);
  return 0;
}
int main(int argc, char **argv) {
  const uint32_t fs = 2000000, n = 200000;
  std::vector<double> x(2 * n), y(2 * n);
  srand(1);
  for (uint32_t i = 0; i < 2 * n; i++) { x[i] = rand() % 2001 - 1000; y[i] = 0.5 * x[i] + (rand() % 41 - 20); }
  for (int a = 1; a < argc; a++)
    if (drive(argv[a], x, y, fs, n)) return 1;
  return 0;
}
