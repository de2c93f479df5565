#include "CfarDetector1D.h"

#include "DetCommon.h"

CfarDetector1D::CfarDetector1D(double _pfa, int8_t _nGuard, int8_t _nTrain, int8_t _minDelay, double _minDoppler)
  : pfa(_pfa), nGuard(_nGuard), nTrain(_nTrain), minDelay(_minDelay), minDoppler(_minDoppler),
    handle(nullptr), capDop(0), capDel(0)
{
}

CfarDetector1D::~CfarDetector1D()
{
  b200dd_det_destroy(handle);
}

std::unique_ptr<Detection> CfarDetector1D::process(Map<std::complex<double>> *x)
{
  b200dd_dropin::FlatMap m(x);
  if (!handle || m.nDop > capDop || m.nDel > capDel)
  {
    b200dd_det_destroy(handle);
    b200dd_det_params p = b200dd_dropin::blank_params();
    p.pfa = pfa;
    p.n_guard = nGuard;
    p.n_train = nTrain;
    p.min_delay = minDelay;
    p.min_doppler = minDoppler;
    handle = b200dd_dropin::make_handle(p, m.nDop, m.nDel);
    capDop = m.nDop;
    capDel = m.nDel;
  }
  const uint32_t cap = m.nDop * m.nDel;
  std::vector<double> d(cap), f(cap), s(cap);
  uint32_t n = 0;
  const int rc = b200dd_det_process_host(handle, B200DD_DET_CFAR, reinterpret_cast<const double *>(m.cells.data()),
                                         m.nDop, m.nDel, m.delay.data(), m.doppler.data(), x->noisePower, d.data(),
                                         f.data(), s.data(), cap, &n);
  // B200DD_ERR_CAPACITY: more detections than the device lists hold (min(cells, 2^18)) -- the list would be
  // truncated, which the reference never does, so that is an error here too
  if (rc != B200DD_OK)
    throw std::runtime_error(std::string("CfarDetector1D::process: ") + b200dd_last_error());
  if (n > cap) n = cap;
  return b200dd_dropin::to_detection(d, f, s, n);
}
