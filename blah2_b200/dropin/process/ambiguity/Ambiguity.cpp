#include "Ambiguity.h"

#include "b200dd.h"

#include <cmath>
#include <stdexcept>
#include <string>

namespace {
[[noreturn]] void fail(const char *what)
{
  throw std::runtime_error(std::string(what) + ": " + b200dd_last_error());
}
}

Ambiguity::Ambiguity(int32_t delayMin, int32_t delayMax, int32_t dopplerMin, int32_t dopplerMax,
                     uint32_t _fs, uint32_t n, bool roundHamming)
  : handle(nullptr), fs(_fs), nSamples(n)
{
  b200dd_caf_params p;
  p.delay_min = delayMin;
  p.delay_max = delayMax;
  p.doppler_min = dopplerMin;
  p.doppler_max = dopplerMax;
  p.fs = _fs;
  p.n_samples = n;
  p.round_hamming = roundHamming ? 1 : 0;
  p.device = -1;
  if (b200dd_caf_create(&p, &handle) != B200DD_OK) fail("Ambiguity");
  b200dd_caf_geometry g;
  b200dd_caf_get_geometry(handle, &g);
  nDelayBins = static_cast<uint16_t>(g.n_delay_bins);
  nDopplerBins = static_cast<uint16_t>(g.n_doppler_bins);
  nCorr = static_cast<uint16_t>(g.n_corr);
  nfft = g.nfft;
  nUsed = g.n_used;
  dopplerMiddle = g.doppler_middle;
  cpi = g.cpi;

  // the Map carries the axes exactly as the reference's constructor builds them
  map = std::make_unique<Map<Complex>>(nDopplerBins, nDelayBins);
  std::vector<int32_t> delay(nDelayBins);
  std::vector<double> doppler(nDopplerBins);
  b200dd_caf_get_axes(handle, delay.data(), doppler.data());
  for (int32_t d : delay) map->delay.push_back(d);
  for (double f : doppler) map->doppler.push_back(f);

  hostX.resize(nUsed);
  hostY.resize(nUsed);
  hostMap.resize(static_cast<size_t>(nDopplerBins) * nDelayBins);
}

Ambiguity::~Ambiguity()
{
  b200dd_caf_destroy(handle);
}

Map<std::complex<double>> *Ambiguity::process(IqData *x, IqData *y)
{
  // The pre-rotation for an off-centre Doppler window is applied on the device to the samples
  // that are consumed; samples that stay queued must be rotated too (the reference rotates the
  // whole FIFO in place), with the phase of their position in the queue.
  const uint32_t queued = x->get_length();
  for (uint32_t i = 0; i < nUsed; i++)
  {
    hostX[i] = x->pop_front();   // throws std::runtime_error on an empty FIFO, like the reference
    hostY[i] = y->pop_front();
  }
  if (dopplerMiddle != 0 && queued > nUsed)
  {
    const std::complex<double> j = {0, 1};
    for (uint32_t i = nUsed; i < queued; i++)
    {
      x->push_back(x->pop_front() * std::exp(1.0 * j * 2.0 * M_PI * dopplerMiddle * ((double)i / fs)));
    }
  }
  nSamples = nUsed;

  if (b200dd_caf_process_host(handle, reinterpret_cast<const double *>(hostX.data()),
                              reinterpret_cast<const double *>(hostY.data()), nUsed,
                              reinterpret_cast<double *>(hostMap.data())) != B200DD_OK)
    fail("Ambiguity::process");

  for (uint32_t i = 0; i < nDopplerBins; i++)
  {
    std::vector<Complex> &row = map->data[i];
    const Complex *src = hostMap.data() + static_cast<size_t>(i) * nDelayBins;
    for (uint32_t k = 0; k < nDelayBins; k++) row[k] = src[k];
  }
  return map.get();
}

double Ambiguity::get_doppler_middle() const { return dopplerMiddle; }
uint16_t Ambiguity::get_n_delay_bins() const { return nDelayBins; }
uint16_t Ambiguity::get_n_doppler_bins() const { return nDopplerBins; }
uint16_t Ambiguity::get_n_corr() const { return nCorr; }
double Ambiguity::get_cpi() const { return cpi; }
uint32_t Ambiguity::get_nfft() const { return nfft; }
uint32_t Ambiguity::get_n_samples() const { return nSamples; }
