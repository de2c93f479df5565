// Drop-in replacement for the reference's WienerHopf class (src/process/clutter/WienerHopf.h:68-78):
// same constructor and bool process(IqData*, IqData*) contract -- false = filter failed, y untouched.
#ifndef B200DD_DROPIN_WIENERHOPF_H
#define B200DD_DROPIN_WIENERHOPF_H

#include "data/IqData.h"
#include "process/PinnedBuffer.h"

#include <stdint.h>
#include <complex>
#include <vector>

struct b200dd_wh;

class WienerHopf
{
public:
  WienerHopf(int32_t delayMin, int32_t delayMax, uint32_t nSamples);
  ~WienerHopf();
  WienerHopf(const WienerHopf &) = delete;
  WienerHopf &operator=(const WienerHopf &) = delete;

  /// x is read only; on success y is replaced by the filtered surveillance channel.
  bool process(IqData *x, IqData *y);

private:
  b200dd_wh *handle;
  uint32_t nSamples;
  PinnedBuffer hostX, hostY;
};

#endif
