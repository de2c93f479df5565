// oracle/shim/fftw_shim.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Our own double-precision, any-length CPU FFT behind the FFTW3 entry points
// declared in oracle/shim/fftw3.h.  It exists so that the reference's
// unmodified sources (Ambiguity.cpp, WienerHopf.cpp, ...) can be compiled and
// run in a container without libfftw3 and used as the parity oracle and as the
// CPU baseline ("kind": "reference" in bench.py -- with the caveat, stated in
// DESIGN.md, that stock FFTW is replaced by this FFT).
//
// Algorithm: recursive decimation-in-time mixed radix (4, 2, 3, 5, then any
// remaining prime <= kMaxDirectPrime by an O(p^2) butterfly); lengths with a
// larger prime factor (e.g. WienerHopf's N+nBins+1 = 2 000 411 = 7 x 285 773,
// WienerHopf.cpp:39-44) go through Bluestein's chirp-z with a power-of-two
// inner FFT.  Twiddles are generated in long double.
//
// Semantics match FFTW: unnormalised, sign = -1 forward / +1 backward,
// in-place allowed (in == out).

#include "fftw3.h"

#include <cmath>
#include <complex>
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

namespace {

using cd = std::complex<double>;
constexpr int kMaxDirectPrime = 128;

struct Fft {
  int n = 0;
  int sign = -1;
  std::vector<int> radices;
  std::vector<cd> tw;  // tw[k] = exp(sign * 2 pi i k / n)
  // Bluestein state
  bool bluestein = false;
  int m = 0;
  std::unique_ptr<Fft> inner_f, inner_b;
  std::vector<cd> chirp;  // exp(sign * i pi k^2 / n)
  std::vector<cd> bhat;   // FFT_m of conj(chirp) wrapped
  std::vector<cd> s0, s1, tmp;

  Fft(int n_, int sign_) : n(n_), sign(sign_) {
    int rem = n, largest = 1;
    std::vector<int> f;
    while (rem % 4 == 0) { f.push_back(4); rem /= 4; }
    while (rem % 2 == 0) { f.push_back(2); rem /= 2; }
    for (int p = 3; (int64_t)p * p <= rem; p += 2)
      while (rem % p == 0) { f.push_back(p); rem /= p; if (p > largest) largest = p; }
    if (rem > 1) { f.push_back(rem); if (rem > largest) largest = rem; }
    if (largest > kMaxDirectPrime) {
      setup_bluestein();
      return;
    }
    radices = f;
    tw.resize(n);
    const long double two_pi = 6.283185307179586476925286766559005768L;
    for (int k = 0; k < n; k++) {
      long double a = two_pi * (long double)k / (long double)n;
      tw[k] = cd((double)cosl(a), (double)(sign * sinl(a)));
    }
  }

  void setup_bluestein() {
    bluestein = true;
    m = 1;
    while (m < 2 * n - 1) m <<= 1;
    inner_f = std::make_unique<Fft>(m, -1);
    inner_b = std::make_unique<Fft>(m, +1);
    chirp.resize(n);
    const long double pi = 3.141592653589793238462643383279502884L;
    for (int64_t k = 0; k < n; k++) {
      int64_t k2 = (k * k) % (2 * (int64_t)n);
      long double a = pi * (long double)k2 / (long double)n;
      chirp[k] = cd((double)cosl(a), (double)(sign * sinl(a)));
    }
    std::vector<cd> b(m, cd(0, 0));
    b[0] = std::conj(chirp[0]);
    for (int k = 1; k < n; k++) b[k] = b[m - k] = std::conj(chirp[k]);
    bhat.resize(m);
    inner_f->exec(b.data(), bhat.data());
    s0.resize(m);
    s1.resize(m);
  }

  void exec(const cd *in, cd *out) {
    if (bluestein) {
      for (int k = 0; k < n; k++) s0[k] = in[k] * chirp[k];
      for (int k = n; k < m; k++) s0[k] = cd(0, 0);
      inner_f->exec(s0.data(), s1.data());
      for (int k = 0; k < m; k++) s1[k] *= bhat[k];
      inner_b->exec(s1.data(), s0.data());
      const double inv = 1.0 / (double)m;
      for (int k = 0; k < n; k++) out[k] = s0[k] * chirp[k] * inv;
      return;
    }
    if (n == 1) { out[0] = in[0]; return; }
    if (in == out) {
      tmp.assign(in, in + n);
      work(out, tmp.data(), 1, 0, n);
    } else {
      work(out, in, 1, 0, n);
    }
  }

  void work(cd *out, const cd *in, size_t fstride, size_t idx, int ncur) {
    const int p = radices[idx];
    const int mm = ncur / p;
    if (mm == 1) {
      for (int k = 0; k < p; k++) out[k] = in[(size_t)k * fstride];
    } else {
      for (int k = 0; k < p; k++) work(out + (size_t)k * mm, in + (size_t)k * fstride, fstride * p, idx + 1, mm);
    }
    switch (p) {
      case 2: bfly2(out, fstride, mm); break;
      case 3: bfly3(out, fstride, mm); break;
      case 4: bfly4(out, fstride, mm); break;
      case 5: bfly5(out, fstride, mm); break;
      default: bfly_generic(out, fstride, mm, p); break;
    }
  }

  void bfly2(cd *o, size_t fs, int mm) const {
    for (int u = 0; u < mm; u++) {
      cd t = o[u + mm] * tw[fs * u];
      o[u + mm] = o[u] - t;
      o[u] += t;
    }
  }

  // multiply by sign*i
  inline cd rot(cd v) const { return sign < 0 ? cd(v.imag(), -v.real()) : cd(-v.imag(), v.real()); }

  void bfly4(cd *o, size_t fs, int mm) const {
    for (int u = 0; u < mm; u++) {
      cd a0 = o[u];
      cd a1 = o[u + mm] * tw[fs * u];
      cd a2 = o[u + 2 * mm] * tw[2 * fs * u];
      cd a3 = o[u + 3 * mm] * tw[3 * fs * u];
      cd s02 = a0 + a2, d02 = a0 - a2, s13 = a1 + a3, d13 = rot(a1 - a3);
      o[u] = s02 + s13;
      o[u + mm] = d02 + d13;
      o[u + 2 * mm] = s02 - s13;
      o[u + 3 * mm] = d02 - d13;
    }
  }

  void bfly3(cd *o, size_t fs, int mm) const {
    const cd w1 = tw[fs * mm], w2 = tw[2 * fs * mm];  // exp(sign 2pi i /3), ^2
    for (int u = 0; u < mm; u++) {
      cd a0 = o[u];
      cd a1 = o[u + mm] * tw[fs * u];
      cd a2 = o[u + 2 * mm] * tw[2 * fs * u];
      o[u] = a0 + a1 + a2;
      o[u + mm] = a0 + a1 * w1 + a2 * w2;
      o[u + 2 * mm] = a0 + a1 * w2 + a2 * w1;
    }
  }

  void bfly5(cd *o, size_t fs, int mm) const {
    cd w[5];
    for (int k = 0; k < 5; k++) w[k] = tw[(fs * mm * k) % n];
    for (int u = 0; u < mm; u++) {
      cd a[5];
      a[0] = o[u];
      for (int q = 1; q < 5; q++) a[q] = o[u + q * mm] * tw[q * fs * u];
      for (int k = 0; k < 5; k++) {
        cd s = a[0];
        for (int q = 1; q < 5; q++) s += a[q] * w[(q * k) % 5];
        o[u + k * mm] = s;
      }
    }
  }

  void bfly_generic(cd *o, size_t fs, int mm, int p) const {
    std::vector<cd> a(p), w(p);
    for (int k = 0; k < p; k++) w[k] = tw[((size_t)fs * mm * k) % n];
    for (int u = 0; u < mm; u++) {
      a[0] = o[u];
      for (int q = 1; q < p; q++) a[q] = o[u + q * mm] * tw[((size_t)q * fs * u) % n];
      for (int k = 0; k < p; k++) {
        cd s = a[0];
        for (int q = 1; q < p; q++) s += a[q] * w[((size_t)q * k) % p];
        o[u + k * mm] = s;
      }
    }
  }
};

}  // namespace

struct b200dd_shim_plan_s {
  std::unique_ptr<Fft> fft;
  cd *in;
  cd *out;
};

extern "C" {

fftw_plan fftw_plan_dft_1d(int n, fftw_complex *in, fftw_complex *out, int sign, unsigned /*flags*/) {
  if (n <= 0) return nullptr;
  auto *p = new b200dd_shim_plan_s;
  p->fft = std::make_unique<Fft>(n, sign < 0 ? -1 : +1);
  p->in = reinterpret_cast<cd *>(in);
  p->out = reinterpret_cast<cd *>(out);
  return p;
}

void fftw_execute(const fftw_plan p) {
  if (p) p->fft->exec(p->in, p->out);
}

void fftw_destroy_plan(fftw_plan p) { delete p; }

int fftw_init_threads(void) { return 1; }
void fftw_plan_with_nthreads(int) {}
void fftw_cleanup_threads(void) {}

// Direct entry for tests of the shim itself (not part of FFTW's API).
__attribute__((visibility("default"))) void b200dd_shim_fft(int n, const double *in, double *out, int sign) {
  Fft f(n, sign < 0 ? -1 : +1);
  f.exec(reinterpret_cast<const cd *>(in), reinterpret_cast<cd *>(out));
}

}  // extern "C"
