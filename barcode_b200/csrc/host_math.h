// barcode_b200/csrc/host_math.h -- host-side cosmology scalars both precision modes derive their constants from.
#pragma once
#include <cmath>

namespace bgpu {

// E_Hubble_a, fgrow (cosmo.cc:26-31,182-217)
inline double host_E_Hubble_a(double a, double OM, double OL) {
  const double OK = 1. - OM - OL;
  return std::sqrt(OM / (a * a * a) + OK / (a * a) + OL);
}
inline double host_fgrow(double a, double OM, double OL) {
  const double E = host_E_Hubble_a(a, OM, OL);
  const double Omega = OM / ((E * E) * (a * a * a));
  return std::pow(Omega, 5. / 9.);
}

}  // namespace bgpu
