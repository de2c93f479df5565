// Drop-in replacement for the reference's Centroid (src/process/detection/Centroid.h:35-44).
#ifndef B200DD_DROPIN_CENTROID_H
#define B200DD_DROPIN_CENTROID_H

#include "data/Detection.h"

#include <stdint.h>
#include <memory>

struct b200dd_det;

class Centroid
{
public:
  Centroid(uint16_t nDelay, uint16_t nDoppler, double resolutionDoppler);
  ~Centroid();
  Centroid(const Centroid &) = delete;
  Centroid &operator=(const Centroid &) = delete;

  std::unique_ptr<Detection> process(Detection *x);

private:
  b200dd_det *handle;
};

#endif
