// oracle/ref_spectrum_capi.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// extern "C" wrapper around the reference's UNMODIFIED SpectrumAnalyser
// (src/process/spectrum/SpectrumAnalyser.cpp:9-74) for ctypes.  Contains no reference code.
// SpectrumAnalyser::process stores its result in PRIVATE members of IqData (IqData.h:38,41,
// written by update_spectrum / update_frequency, IqData.cpp:83-91) whose only reader is to_json
// (IqData.cpp:92-125: 10 log10 |.|, two decimals) -- far too coarse to pin arithmetic.  This one
// translation unit therefore includes the reference's IqData.h with `private` spelt `public`
// (standard headers first, so only that class is affected; the class layout is unchanged and the
// reference's own IqData.cpp is compiled normally).
//
// Like oracle/ref_capi.cpp the very same file is also compiled against blah2_b200/dropin's
// headers by tests/native/Makefile, where it drives the drop-in SpectrumAnalyser class.
#include <complex>
#include <cstdint>
#include <cstdio>
#include <deque>
#include <mutex>
#include <string>
#include <vector>

#define private public
#include "data/IqData.h"
#undef private
#include "process/spectrum/SpectrumAnalyser.h"

#define REF_API extern "C" __attribute__((visibility("default")))

// x: nIn interleaved complex128 (the FIFO is filled with all of them, capacity nIn);
// spectrum_o: cap complex128 slots; frequency_o: cap doubles.  counts[0] = spectrum entries,
// counts[1] = frequency entries, counts[2] = samples left in x afterwards (process must not consume).
// Returns 0, or -1000 when an exception was caught.
REF_API int refpath_spectrum_process(uint32_t n, double bandwidth, const double *x, uint32_t nIn, double *spectrum_o,
                                     double *frequency_o, uint32_t cap, uint32_t *counts) {
  try {
    SpectrumAnalyser sa(n, bandwidth);
    IqData xd(nIn);
    for (uint32_t i = 0; i < nIn; i++) xd.push_back({x[2 * i], x[2 * i + 1]});
    sa.process(&xd);
    counts[0] = (uint32_t)xd.spectrum.size();
    counts[1] = (uint32_t)xd.frequency.size();
    counts[2] = xd.get_length();
    for (uint32_t i = 0; i < counts[0] && i < cap; i++) {
      spectrum_o[2 * i] = xd.spectrum[i].real();
      spectrum_o[2 * i + 1] = xd.spectrum[i].imag();
    }
    for (uint32_t i = 0; i < counts[1] && i < cap; i++) frequency_o[i] = xd.frequency[i];
    return 0;
  } catch (const std::exception &e) {
    fprintf(stderr, "refpath: exception: %s\n", e.what());
    return -1000;
  }
}
