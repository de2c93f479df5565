// oracle/ref_capi.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// extern "C" wrapper around the reference's UNMODIFIED hot-path classes so that
// pytest / bench.py can drive them through ctypes.  This file contains no
// reference code: it includes the reference headers where they lie under
// /root/reference/src and is compiled together with the reference .cpp files by
// oracle/Makefile into oracle/_ref/libblah2ref.so (git-ignored, travels to the
// GPU box as a built artefact).
//
// Wrapped reference entry points:
//   next_hamming                     src/process/meta/HammingNumber.cpp:38-48
//   Ambiguity::Ambiguity / process   src/process/ambiguity/Ambiguity.cpp:11-82, 92-172
//   Map::set_metrics                 src/data/Map.cpp:188-206
//   WienerHopf::process              src/process/clutter/WienerHopf.cpp:58-163
//   CfarDetector1D::process          src/process/detection/CfarDetector1D.cpp:23-100
//   Centroid::process                src/process/detection/Centroid.cpp:19-73
//   Interpolate::process             src/process/detection/Interpolate.cpp:20-91
// The call order in refpath_chain_run mirrors src/blah2.cpp:268-287.

#include "process/ambiguity/Ambiguity.h"
#include "process/clutter/WienerHopf.h"
#include "process/detection/CfarDetector1D.h"
#include "process/detection/Centroid.h"
#include "process/detection/Interpolate.h"

#include <chrono>
#include <complex>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <vector>

#define REF_API extern "C" __attribute__((visibility("default")))

// Exceptions must not cross the C boundary into ctypes: report and turn into an error value.
static thread_local char g_err[512] = "";
#define REF_TRY try {
#define REF_CATCH(retval)                                             \
  }                                                                   \
  catch (const std::exception &e) {                                   \
    snprintf(g_err, sizeof(g_err), "%s", e.what());                   \
    fprintf(stderr, "refpath: exception: %s\n", e.what());            \
    return retval;                                                    \
  }
REF_API const char *refpath_last_error() { return g_err; }

namespace {

using Complex = std::complex<double>;

double now_ms() {
  using namespace std::chrono;
  return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

void fill_iq(IqData &d, const double *v, uint32_t n) {
  for (uint32_t i = 0; i < n; i++) d.push_back({v[2 * i], v[2 * i + 1]});
}

void map_out(Map<Complex> *map, double *out) {
  const uint32_t nr = map->get_nRows(), nc = map->get_nCols();
  for (uint32_t i = 0; i < nr; i++)
    for (uint32_t j = 0; j < nc; j++) {
      out[2 * ((size_t)i * nc + j)] = map->data[i][j].real();
      out[2 * ((size_t)i * nc + j) + 1] = map->data[i][j].imag();
    }
}

std::unique_ptr<Map<Complex>> map_in(const double *m, uint32_t nDop, uint32_t nDel, const int32_t *delay,
                                     const double *doppler, double noisePower) {
  auto map = std::make_unique<Map<Complex>>(nDop, nDel);
  for (uint32_t i = 0; i < nDop; i++)
    for (uint32_t j = 0; j < nDel; j++)
      map->data[i][j] = Complex(m[2 * ((size_t)i * nDel + j)], m[2 * ((size_t)i * nDel + j) + 1]);
  for (uint32_t j = 0; j < nDel; j++) map->delay.push_back(delay[j]);
  for (uint32_t i = 0; i < nDop; i++) map->doppler.push_back(doppler[i]);
  map->noisePower = noisePower;
  map->maxPower = 0;
  return map;
}

uint32_t det_out(Detection *d, double *delay, double *doppler, double *snr, uint32_t cap) {
  auto dl = d->get_delay();
  auto dp = d->get_doppler();
  auto sn = d->get_snr();
  uint32_t n = (uint32_t)dl.size();
  for (uint32_t i = 0; i < n && i < cap; i++) {
    delay[i] = dl[i];
    doppler[i] = dp[i];
    snr[i] = sn[i];
  }
  return n;
}

struct Chain {
  uint32_t n;
  bool clutter;
  std::unique_ptr<Ambiguity> amb;
  std::unique_ptr<WienerHopf> wh;
  std::unique_ptr<CfarDetector1D> cfar;
  std::unique_ptr<Centroid> cen;
  std::unique_ptr<Interpolate> interp;
};

}  // namespace

REF_API uint32_t refpath_next_hamming(uint32_t v) { return next_hamming(v); }

REF_API int refpath_ambiguity_geometry(int32_t delayMin, int32_t delayMax, int32_t dopplerMin, int32_t dopplerMax,
                                       uint32_t fs, uint32_t n, int roundHamming, uint32_t *nDelayBins,
                                       uint32_t *nDopplerBins, uint32_t *nCorr, uint32_t *nfft, double *cpi,
                                       double *dopplerMiddle) {
  REF_TRY
  Ambiguity a(delayMin, delayMax, dopplerMin, dopplerMax, fs, n, roundHamming != 0);
  *nDelayBins = a.get_n_delay_bins();
  *nDopplerBins = a.get_n_doppler_bins();
  *nCorr = a.get_n_corr();
  *nfft = a.get_nfft();
  *cpi = a.get_cpi();
  *dopplerMiddle = a.get_doppler_middle();
  return 0;
  REF_CATCH(-1000)
}

// x, y: nIn interleaved complex128 each.  map_o: [nDop][nDel] interleaved complex128.
// metrics[0] = noisePower, metrics[1] = maxPower (Map::set_metrics as blah2.cpp:279 does).
// leftover[0..1] = samples still queued in x / y after process (the FIFOs are consumed).
REF_API int refpath_ambiguity_process(int32_t delayMin, int32_t delayMax, int32_t dopplerMin, int32_t dopplerMax,
                                      uint32_t fs, uint32_t n, int roundHamming, const double *x, const double *y,
                                      uint32_t nIn, double *map_o, int32_t *delay_o, double *doppler_o,
                                      double *metrics, uint32_t *leftover) {
  REF_TRY
  Ambiguity a(delayMin, delayMax, dopplerMin, dopplerMax, fs, n, roundHamming != 0);
  IqData xd(nIn), yd(nIn);
  fill_iq(xd, x, nIn);
  fill_iq(yd, y, nIn);
  Map<Complex> *map = a.process(&xd, &yd);
  map->set_metrics();
  map_out(map, map_o);
  for (uint32_t j = 0; j < map->get_nCols(); j++) delay_o[j] = map->delay[j];
  for (uint32_t i = 0; i < map->get_nRows(); i++) doppler_o[i] = map->doppler[i];
  metrics[0] = map->noisePower;
  metrics[1] = map->maxPower;
  leftover[0] = xd.get_length();
  leftover[1] = yd.get_length();
  return 0;
  REF_CATCH(-1000)
}

// Returns 1 when the filter succeeded (y_io overwritten with the filtered surveillance
// channel), 0 when WienerHopf::process returned false (y_io untouched).
REF_API int refpath_wienerhopf_process(int32_t delayMin, int32_t delayMax, uint32_t nSamples, const double *x,
                                       double *y_io) {
  REF_TRY
  WienerHopf wh(delayMin, delayMax, nSamples);
  IqData xd(nSamples), yd(nSamples);
  fill_iq(xd, x, nSamples);
  fill_iq(yd, y_io, nSamples);
  bool ok = wh.process(&xd, &yd);
  if (!ok) return 0;
  auto data = yd.get_data();
  for (uint32_t i = 0; i < nSamples && i < data.size(); i++) {
    y_io[2 * i] = data[i].real();
    y_io[2 * i + 1] = data[i].imag();
  }
  return 1;
  REF_CATCH(-1000)
}

REF_API void refpath_set_metrics(const double *m, uint32_t nDop, uint32_t nDel, double *metrics) {
  std::vector<int32_t> delay(nDel, 0);
  std::vector<double> doppler(nDop, 0.0);
  auto map = map_in(m, nDop, nDel, delay.data(), doppler.data(), 0.0);
  map->set_metrics();
  metrics[0] = map->noisePower;
  metrics[1] = map->maxPower;
}

REF_API uint32_t refpath_cfar(double pfa, int nGuard, int nTrain, int minDelay, double minDoppler, const double *m,
                              uint32_t nDop, uint32_t nDel, const int32_t *delay, const double *doppler,
                              double noisePower, double *o_delay, double *o_doppler, double *o_snr, uint32_t cap) {
  REF_TRY
  auto map = map_in(m, nDop, nDel, delay, doppler, noisePower);
  CfarDetector1D cfar(pfa, (int8_t)nGuard, (int8_t)nTrain, (int8_t)minDelay, minDoppler);
  auto det = cfar.process(map.get());
  return det_out(det.get(), o_delay, o_doppler, o_snr, cap);
  REF_CATCH(0xFFFFFFFFu)
}

REF_API uint32_t refpath_centroid(uint32_t nDelay, uint32_t nDoppler, double resolutionDoppler, const double *delay,
                                  const double *doppler, const double *snr, uint32_t n, double *o_delay,
                                  double *o_doppler, double *o_snr, uint32_t cap) {
  REF_TRY
  Detection in(std::vector<double>(delay, delay + n), std::vector<double>(doppler, doppler + n),
               std::vector<double>(snr, snr + n));
  Centroid cen((uint16_t)nDelay, (uint16_t)nDoppler, resolutionDoppler);
  auto det = cen.process(&in);
  return det_out(det.get(), o_delay, o_doppler, o_snr, cap);
  REF_CATCH(0xFFFFFFFFu)
}

REF_API uint32_t refpath_interpolate(int doDelay, int doDoppler, const double *delay, const double *doppler,
                                     const double *snr, uint32_t n, const double *m, uint32_t nDop, uint32_t nDel,
                                     const int32_t *mdelay, const double *mdoppler, double noisePower,
                                     double *o_delay, double *o_doppler, double *o_snr, uint32_t cap) {
  REF_TRY
  Detection in(std::vector<double>(delay, delay + n), std::vector<double>(doppler, doppler + n),
               std::vector<double>(snr, snr + n));
  auto map = map_in(m, nDop, nDel, mdelay, mdoppler, noisePower);
  Interpolate interp(doDelay != 0, doDoppler != 0);
  auto det = interp.process(&in, map.get());
  return det_out(det.get(), o_delay, o_doppler, o_snr, cap);
  REF_CATCH(0xFFFFFFFFu)
}

// ---- whole chain with persistent objects (construction outside the timed region, as in
// src/blah2.cpp:154-183), per-stage wall times named like blah2.cpp:261-288 -------------

REF_API void *refpath_chain_create(int32_t delayMin, int32_t delayMax, int32_t dopplerMin, int32_t dopplerMax,
                                   uint32_t fs, uint32_t n, int roundHamming, int clutterEnable,
                                   int32_t delayMinClutter, int32_t delayMaxClutter, double pfa, int nGuard,
                                   int nTrain, int minDelay, double minDoppler, uint32_t nCentroid) {
  REF_TRY
  auto *c = new Chain;
  c->n = n;
  c->clutter = clutterEnable != 0;
  c->amb = std::make_unique<Ambiguity>(delayMin, delayMax, dopplerMin, dopplerMax, fs, n, roundHamming != 0);
  if (c->clutter) c->wh = std::make_unique<WienerHopf>(delayMinClutter, delayMaxClutter, n);
  c->cfar = std::make_unique<CfarDetector1D>(pfa, (int8_t)nGuard, (int8_t)nTrain, (int8_t)minDelay, minDoppler);
  // blah2.cpp:183: Centroid(nCentroid, nCentroid, 1/tCpi) with tCpi = nSamples/fs
  c->cen = std::make_unique<Centroid>((uint16_t)nCentroid, (uint16_t)nCentroid, 1.0 / ((double)n / (double)fs));
  c->interp = std::make_unique<Interpolate>(true, true);
  return c;
  REF_CATCH(nullptr)
}

REF_API void refpath_chain_destroy(void *h) {
  auto *c = static_cast<Chain *>(h);
  if (!c) return;
  const bool dbg = getenv("REFPATH_DEBUG") != nullptr;
  if (dbg) fprintf(stderr, "chain_destroy: interp\n");
  c->interp.reset();
  if (dbg) fprintf(stderr, "chain_destroy: centroid\n");
  c->cen.reset();
  if (dbg) fprintf(stderr, "chain_destroy: cfar\n");
  c->cfar.reset();
  if (dbg) fprintf(stderr, "chain_destroy: wh\n");
  c->wh.reset();
  if (dbg) fprintf(stderr, "chain_destroy: amb\n");
  c->amb.reset();
  if (dbg) fprintf(stderr, "chain_destroy: done\n");
  delete c;
}

// stage_ms[0] = clutter_filter, [1] = ambiguity_processing (incl. set_metrics), [2] = detector.
// Returns -1 if the clutter filter failed (CPI skipped, blah2.cpp:270-273), else #detections.
REF_API int refpath_chain_run(void *h, const double *x, const double *y, double *map_o, double *metrics,
                              double *o_delay, double *o_doppler, double *o_snr, uint32_t cap, double *stage_ms) {
  REF_TRY
  auto *c = static_cast<Chain *>(h);
  IqData xd(c->n), yd(c->n);
  fill_iq(xd, x, c->n);
  fill_iq(yd, y, c->n);
  double t0 = now_ms();
  if (c->clutter) {
    if (!c->wh->process(&xd, &yd)) return -1;
  }
  double t1 = now_ms();
  Map<Complex> *map = c->amb->process(&xd, &yd);
  map->set_metrics();
  double t2 = now_ms();
  auto d1 = c->cfar->process(map);
  auto d2 = c->cen->process(d1.get());
  auto d3 = c->interp->process(d2.get(), map);
  double t3 = now_ms();
  if (stage_ms) {
    stage_ms[0] = t1 - t0;
    stage_ms[1] = t2 - t1;
    stage_ms[2] = t3 - t2;
  }
  if (map_o) map_out(map, map_o);
  if (metrics) {
    metrics[0] = map->noisePower;
    metrics[1] = map->maxPower;
  }
  return (int)det_out(d3.get(), o_delay, o_doppler, o_snr, cap);
  REF_CATCH(-1000)
}
