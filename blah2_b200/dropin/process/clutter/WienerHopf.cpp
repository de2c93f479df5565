#include "WienerHopf.h"

#include "b200dd.h"

#include <algorithm>
#include <iostream>
#include <stdexcept>
#include <string>

WienerHopf::WienerHopf(int32_t delayMin, int32_t delayMax, uint32_t _nSamples)
  : handle(nullptr), nSamples(_nSamples)
{
  if (b200dd_wh_create(delayMin, delayMax, _nSamples, -1, &handle) != B200DD_OK)
    throw std::runtime_error(std::string("WienerHopf: ") + b200dd_last_error());
  hostX.resize(nSamples);
  hostY.resize(nSamples);
}

WienerHopf::~WienerHopf()
{
  b200dd_wh_destroy(handle);
}

bool WienerHopf::process(IqData *x, IqData *y)
{
  // snapshot both FIFOs (the reference copies them as well) into contiguous staging
  {
    const std::deque<std::complex<double>> xd = x->get_data();
    const std::deque<std::complex<double>> yd = y->get_data();
    if (xd.size() < nSamples || yd.size() < nSamples)
      throw std::runtime_error("WienerHopf::process: fewer than nSamples queued");
    // sequential deque iteration (operator[] on a deque recomputes the block address for every element)
    std::copy(xd.begin(), xd.begin() + nSamples, hostX.data());
    std::copy(yd.begin(), yd.begin() + nSamples, hostY.data());
  }
  const int rc = b200dd_wh_process_host(handle, reinterpret_cast<const double *>(hostX.data()),
                                        reinterpret_cast<double *>(hostY.data()));
  if (rc == B200DD_FILTER_FAILED)
  {
    std::cerr << "Chol decomposition failed, skip clutter filter" << std::endl;
    return false;
  }
  if (rc != B200DD_OK)
    throw std::runtime_error(std::string("WienerHopf::process: ") + b200dd_last_error());

  y->clear();
  for (uint32_t i = 0; i < nSamples; i++) y->push_back(hostY[i]);
  return true;
}
