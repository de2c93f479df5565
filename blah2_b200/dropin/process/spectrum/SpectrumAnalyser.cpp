#include "SpectrumAnalyser.h"

#include "b200dd.h"

#include <stdexcept>
#include <string>

SpectrumAnalyser::SpectrumAnalyser(uint32_t n, double bandwidth)
  : handle(nullptr), nfft(0), nSpectrum(0)
{
  if (b200dd_spectrum_create(n, bandwidth, -1, &handle) != B200DD_OK)
    throw std::runtime_error(std::string("SpectrumAnalyser: ") + b200dd_last_error());
  b200dd_spectrum_geometry g;
  b200dd_spectrum_get_geometry(handle, &g);
  nfft = g.nfft;
  nSpectrum = g.n_spectrum;
  hostX.resize(nfft);
  hostSpectrum.resize(nSpectrum);
  frequency.resize(g.n_frequency);
  b200dd_spectrum_get_frequency(handle, frequency.data(), g.n_frequency);
}

SpectrumAnalyser::~SpectrumAnalyser()
{
  b200dd_spectrum_destroy(handle);
}

void SpectrumAnalyser::process(IqData *x)
{
  {
    // snapshot of the FIFO (the reference copies it as well); only the first nfft samples are read
    const std::deque<std::complex<double>> data = x->get_data();
    if (data.size() < nfft)
      throw std::runtime_error("SpectrumAnalyser::process: fewer than nfft samples queued");
    for (uint32_t i = 0; i < nfft; i++) hostX[i] = data[i];
  }
  if (b200dd_spectrum_process_host(handle, reinterpret_cast<const double *>(hostX.data()), nfft,
                                   reinterpret_cast<double *>(hostSpectrum.data())) != B200DD_OK)
    throw std::runtime_error(std::string("SpectrumAnalyser::process: ") + b200dd_last_error());
  x->update_spectrum(hostSpectrum);
  x->update_frequency(frequency);
}
