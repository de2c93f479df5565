#include "Interpolate.h"

#include "DetCommon.h"

Interpolate::Interpolate(bool _doDelay, bool _doDoppler)
  : doDelay(_doDelay), doDoppler(_doDoppler), handle(nullptr), capDop(0), capDel(0)
{
}

Interpolate::~Interpolate()
{
  b200dd_det_destroy(handle);
}

std::unique_ptr<Detection> Interpolate::process(Detection *x, Map<std::complex<double>> *y)
{
  b200dd_dropin::FlatMap m(y);
  if (!handle || m.nDop > capDop || m.nDel > capDel)
  {
    b200dd_det_destroy(handle);
    b200dd_det_params p = b200dd_dropin::blank_params();
    p.interp_delay = doDelay ? 1 : 0;
    p.interp_doppler = doDoppler ? 1 : 0;
    handle = b200dd_dropin::make_handle(p, m.nDop, m.nDel);
    capDop = m.nDop;
    capDel = m.nDel;
  }
  std::vector<double> d = x->get_delay(), f = x->get_doppler(), s = x->get_snr();
  const uint32_t n = static_cast<uint32_t>(s.size());
  std::vector<double> od(n ? n : 1), of(n ? n : 1), os(n ? n : 1);
  uint32_t k = 0;
  if (b200dd_det_interpolate_host(handle, d.data(), f.data(), s.data(), n,
                                  reinterpret_cast<const double *>(m.cells.data()), m.nDop, m.nDel, m.delay.data(),
                                  m.doppler.data(), y->noisePower, od.data(), of.data(), os.data(), n ? n : 1,
                                  &k) != B200DD_OK)
    throw std::runtime_error(std::string("Interpolate::process: ") + b200dd_last_error());
  return b200dd_dropin::to_detection(od, of, os, k);
}
