// Shared helpers of the detection drop-ins (not part of the reference's interface).
#ifndef B200DD_DROPIN_DETCOMMON_H
#define B200DD_DROPIN_DETCOMMON_H

#include "data/Detection.h"
#include "data/Map.h"

#include "b200dd.h"

#include <complex>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace b200dd_dropin
{

struct FlatMap
{
  std::vector<std::complex<double>> cells;
  std::vector<int32_t> delay;
  std::vector<double> doppler;
  uint32_t nDop = 0, nDel = 0;

  explicit FlatMap(Map<std::complex<double>> *m)
  {
    nDop = m->get_nRows();
    nDel = m->get_nCols();
    cells.resize(static_cast<size_t>(nDop) * nDel);
    for (uint32_t i = 0; i < nDop; i++)
      for (uint32_t j = 0; j < nDel; j++) cells[static_cast<size_t>(i) * nDel + j] = m->data[i][j];
    delay.assign(m->delay.begin(), m->delay.end());
    doppler.assign(m->doppler.begin(), m->doppler.end());
  }
};

inline b200dd_det *make_handle(const b200dd_det_params &p, uint32_t nDop, uint32_t nDel)
{
  b200dd_det *h = nullptr;
  if (b200dd_det_create(&p, nDop, nDel, &h) != B200DD_OK)
    throw std::runtime_error(std::string("detection: ") + b200dd_last_error());
  return h;
}

inline b200dd_det_params blank_params()
{
  b200dd_det_params p;
  p.pfa = 1e-5;
  p.n_guard = 0;
  p.n_train = 0;
  p.min_delay = 0;
  p.min_doppler = 0.0;
  p.n_centroid_delay = 0;
  p.n_centroid_doppler = 0;
  p.resolution_doppler = 1.0;
  p.interp_delay = 1;
  p.interp_doppler = 1;
  p.device = -1;
  return p;
}

inline std::unique_ptr<Detection> to_detection(std::vector<double> &d, std::vector<double> &f, std::vector<double> &s,
                                               uint32_t n)
{
  d.resize(n);
  f.resize(n);
  s.resize(n);
  return std::make_unique<Detection>(d, f, s);
}

}  // namespace b200dd_dropin

#endif
