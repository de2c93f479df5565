// Drop-in replacement for the reference's SpectrumAnalyser class
// (src/process/spectrum/SpectrumAnalyser.h:16-59): same constructor and void process(IqData *x)
// contract -- x is read (not consumed) and receives the decimated spectrum and the frequency
// vector through IqData::update_spectrum / update_frequency (SpectrumAnalyser.cpp:55,68).
#ifndef B200DD_DROPIN_SPECTRUMANALYSER_H
#define B200DD_DROPIN_SPECTRUMANALYSER_H

#include "data/IqData.h"

#include <stdint.h>
#include <complex>
#include <vector>

struct b200dd_spectrum;

class SpectrumAnalyser
{
public:
  SpectrumAnalyser(uint32_t n, double bandwidth);
  ~SpectrumAnalyser();
  SpectrumAnalyser(const SpectrumAnalyser &) = delete;
  SpectrumAnalyser &operator=(const SpectrumAnalyser &) = delete;

  void process(IqData *x);

private:
  b200dd_spectrum *handle;
  uint32_t nfft;
  uint32_t nSpectrum;
  std::vector<std::complex<double>> hostX, hostSpectrum;
  std::vector<double> frequency;
};

#endif
