// Drop-in replacement for the reference's Interpolate (src/process/detection/Interpolate.h:36-45).
#ifndef B200DD_DROPIN_INTERPOLATE_H
#define B200DD_DROPIN_INTERPOLATE_H

#include "data/Detection.h"
#include "data/Map.h"

#include <complex>
#include <memory>

struct b200dd_det;

class Interpolate
{
public:
  Interpolate(bool doDelay, bool doDoppler);
  ~Interpolate();
  Interpolate(const Interpolate &) = delete;
  Interpolate &operator=(const Interpolate &) = delete;

  std::unique_ptr<Detection> process(Detection *x, Map<std::complex<double>> *y);

private:
  bool doDelay, doDoppler;
  b200dd_det *handle;
  uint32_t capDop, capDel;
};

#endif
