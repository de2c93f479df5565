#include "Centroid.h"

#include "DetCommon.h"

Centroid::Centroid(uint16_t nDelay, uint16_t nDoppler, double resolutionDoppler) : handle(nullptr)
{
  b200dd_det_params p = b200dd_dropin::blank_params();
  p.n_centroid_delay = nDelay;
  p.n_centroid_doppler = nDoppler;
  p.resolution_doppler = resolutionDoppler;
  handle = b200dd_dropin::make_handle(p, 512, 512);  // sizes only bound the detection-list capacity (2^18)
}

Centroid::~Centroid()
{
  b200dd_det_destroy(handle);
}

std::unique_ptr<Detection> Centroid::process(Detection *x)
{
  std::vector<double> d = x->get_delay(), f = x->get_doppler(), s = x->get_snr();
  const uint32_t n = static_cast<uint32_t>(s.size());
  std::vector<double> od(n ? n : 1), of(n ? n : 1), os(n ? n : 1);
  uint32_t m = 0;
  if (b200dd_det_centroid_host(handle, d.data(), f.data(), s.data(), n, od.data(), of.data(), os.data(),
                               n ? n : 1, &m) != B200DD_OK)
    throw std::runtime_error(std::string("Centroid::process: ") + b200dd_last_error());
  return b200dd_dropin::to_detection(od, of, os, m);
}
