// Host-side simulation of the second-generation CTA FFT (blah2_b200/csrc/fft_dit.cuh): the same
// __host__ __device__ pass functions executed by a sequential loop over "threads", checked against a
// long-double DFT (forward) and by a forward -> inverse round trip through the register hand-off.
// No GPU needed (compiled by nvcc as host code).
#include "../../blah2_b200/csrc/fft_dit.cuh"

#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <vector>

using namespace b2;

static std::vector<std::complex<long double>> ref_dft(const std::vector<std::complex<long double>> &x) {
  const int M = (int)x.size();
  const long double two_pi = 6.283185307179586476925286766559L;
  std::vector<std::complex<long double>> a = x;
  for (int i = 1, j = 0; i < M; i++) {
    int bit = M >> 1;
    for (; j & bit; bit >>= 1) j ^= bit;
    j ^= bit;
    if (i < j) std::swap(a[i], a[j]);
  }
  for (int len = 2; len <= M; len <<= 1)
    for (int i = 0; i < M; i += len)
      for (int k = 0; k < len / 2; k++) {
        long double ang = -two_pi * k / len;
        std::complex<long double> w(cosl(ang), sinl(ang));
        auto u = a[i + k], v = a[i + k + len / 2] * w;
        a[i + k] = u + v;
        a[i + k + len / 2] = u - v;
      }
  return a;
}

template <class T, int LOG2M> double run() {
  using P = dit::Plan3<LOG2M>;
  using C = cpx<T>;
  const int M = P::M;
  const long double two_pi = 6.283185307179586476925286766559L;
  std::vector<C> tw(M), s(P::MP), regs((size_t)P::NT * 16);
  std::vector<std::complex<long double>> x(M);
  for (int j = 0; j < M; j++) { tw[j].x = (T)cosl(two_pi * j / M); tw[j].y = (T)(-sinl(two_pi * j / M)); }
  srand(LOG2M * 7 + sizeof(T));
  std::vector<C> xin(M);
  for (int i = 0; i < M; i++) {
    T a = (T)(rand() / (double)RAND_MAX - 0.5), b = (T)(rand() / (double)RAND_MAX - 0.5);
    xin[i].x = a; xin[i].y = b;
    x[i] = std::complex<long double>(a, b);
  }
  auto X = ref_dft(x);
  // forward
  for (int tid = 0; tid < P::NT; tid++) {
    C v[16];
    for (int k = 0; k < 16; k++) v[k] = xin[tid + P::S2 * k];
    dit::pass0_store<T, LOG2M, -1>(s.data(), tid, v);
  }
  for (int tid = 0; tid < P::NT; tid++) {
    C v[16];
    dit::pass1_load<T, LOG2M>(s.data(), tid, v);
    dit::pass1_compute<T, LOG2M, -1>(tw.data(), tid, v);
    dit::pass1_store<T, LOG2M>(s.data(), tid, v);
  }
  long double err = 0, nrm = 0;
  for (int tid = 0; tid < P::NT; tid++) {
    C v[16];
    dit::pass2_load<T, LOG2M>(s.data(), tid, v);
    dit::pass2_compute<T, LOG2M, -1>(tw.data(), tid, v);
    for (int q = 0; q < 16; q++) {
      regs[(size_t)tid * 16 + q] = v[brev<16>(q)];
      auto d = std::complex<long double>(v[brev<16>(q)].x, v[brev<16>(q)].y) - X[tid + P::S2 * q];
      err += std::norm(d);
      nrm += std::norm(X[tid + P::S2 * q]);
    }
  }
  const double e_fwd = (double)sqrtl(err / nrm);
  // inverse through the register hand-off
  for (int tid = 0; tid < P::NT; tid++) {
    C v[16];
    for (int k = 0; k < 16; k++) v[k] = regs[(size_t)tid * 16 + k];
    dit::pass0_store<T, LOG2M, +1>(s.data(), tid, v);
  }
  for (int tid = 0; tid < P::NT; tid++) {
    C v[16];
    dit::pass1_load<T, LOG2M>(s.data(), tid, v);
    dit::pass1_compute<T, LOG2M, +1>(tw.data(), tid, v);
    dit::pass1_store<T, LOG2M>(s.data(), tid, v);
  }
  err = 0; nrm = 0;
  for (int tid = 0; tid < P::NT; tid++) {
    C v[16];
    dit::pass2_load<T, LOG2M>(s.data(), tid, v);
    dit::pass2_compute<T, LOG2M, +1>(tw.data(), tid, v);
    for (int q = 0; q < 16; q++) {
      auto d = std::complex<long double>(v[brev<16>(q)].x, v[brev<16>(q)].y) / (long double)M - x[tid + P::S2 * q];
      err += std::norm(d);
      nrm += std::norm(x[tid + P::S2 * q]);
    }
  }
  const double e_inv = (double)sqrtl(err / nrm);
  printf("DIT LOG2M=%2d %s  fwd rel-L2 %.3e  roundtrip rel-L2 %.3e\n", LOG2M, sizeof(T) == 4 ? "f32" : "f64", e_fwd, e_inv);
  return e_fwd > e_inv ? e_fwd : e_inv;
}

template <int L> int run_both() {
  double ef = run<float, L>();
  double ed = run<double, L>();
  return (ef < 2e-6 && ed < 1e-14) ? 0 : 1;
}

int main() {
  int bad = 0;
  bad += run_both<9>();
  bad += run_both<10>();
  bad += run_both<11>();
  bad += run_both<12>();
  printf(bad ? "FFT_DIT_SIM FAIL\n" : "FFT_DIT_SIM OK\n");
  return bad;
}
