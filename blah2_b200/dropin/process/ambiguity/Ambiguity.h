// Drop-in replacement for the reference's Ambiguity class: same name, constructor, process()
// and getters (src/process/ambiguity/Ambiguity.h:34-58), so src/blah2.cpp:154-160,278 and
// test/unit/process/ambiguity/TestAmbiguity.cpp compile unchanged.  The cross-ambiguity function
// itself runs on a B200 through the C ABI in include/b200dd.h; this class only marshals between
// the reference's containers (IqData deque, Map vector-of-vectors) and contiguous buffers.
#ifndef B200DD_DROPIN_AMBIGUITY_H
#define B200DD_DROPIN_AMBIGUITY_H

#include "data/IqData.h"
#include "process/PinnedBuffer.h"
#include "data/Map.h"
#include "process/meta/HammingNumber.h"

#include <stdint.h>
#include <complex>
#include <memory>
#include <vector>

struct b200dd_caf;

class Ambiguity
{
public:
  using Complex = std::complex<double>;

  Ambiguity(int32_t delayMin, int32_t delayMax, int32_t dopplerMin, int32_t dopplerMax,
            uint32_t fs, uint32_t n, bool roundHamming = false);
  ~Ambiguity();
  Ambiguity(const Ambiguity &) = delete;
  Ambiguity &operator=(const Ambiguity &) = delete;

  /// Consumes nDopplerBins*nCorr samples from both FIFOs (like the reference) and returns the
  /// map owned by this object (valid until the next call).
  Map<Complex> *process(IqData *x, IqData *y);

  double get_doppler_middle() const;
  uint16_t get_n_delay_bins() const;
  uint16_t get_n_doppler_bins() const;
  uint16_t get_n_corr() const;
  double get_cpi() const;
  uint32_t get_nfft() const;
  uint32_t get_n_samples() const;

private:
  b200dd_caf *handle;
  uint32_t fs;
  uint32_t nSamples;
  uint16_t nDelayBins, nDopplerBins, nCorr;
  uint32_t nfft, nUsed;
  double dopplerMiddle, cpi;
  PinnedBuffer hostX, hostY, hostMap;
  std::unique_ptr<Map<Complex>> map;
};

#endif
