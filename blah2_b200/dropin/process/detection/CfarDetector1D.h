// Drop-in replacement for the reference's CfarDetector1D (src/process/detection/CfarDetector1D.h:46-55).
#ifndef B200DD_DROPIN_CFARDETECTOR1D_H
#define B200DD_DROPIN_CFARDETECTOR1D_H

#include "data/Detection.h"
#include "data/Map.h"

#include <stdint.h>
#include <complex>
#include <memory>

struct b200dd_det;

class CfarDetector1D
{
public:
  CfarDetector1D(double pfa, int8_t nGuard, int8_t nTrain, int8_t minDelay, double minDoppler);
  ~CfarDetector1D();
  CfarDetector1D(const CfarDetector1D &) = delete;
  CfarDetector1D &operator=(const CfarDetector1D &) = delete;

  std::unique_ptr<Detection> process(Map<std::complex<double>> *x);

private:
  double pfa;
  int8_t nGuard, nTrain, minDelay;
  double minDoppler;
  b200dd_det *handle;
  uint32_t capDop, capDel;
};

#endif
