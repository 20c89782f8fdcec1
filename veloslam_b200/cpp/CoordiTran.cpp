#include "CoordiTran.h"

#include <cmath>

namespace {
const double kA = 6378137.0000;  // WGS-84 semi-major axis (m)
const double kB = 6356752.3142;  // semi-minor axis (m)
}

// Closed-form geodetic -> ECEF with the reference's operation order (CoordiTran.cpp:51-80), so
// that doubles agree with it bit for bit on the host.
void llh2xyz(double llh[3], double xyz[3]) {
  const double phi = llh[0], lambda = llh[1], h = llh[2];
  const double e = std::sqrt(1 - (kB / kA) * (kB / kA));
  const double sinphi = std::sin(phi), cosphi = std::cos(phi);
  const double coslam = std::cos(lambda), sinlam = std::sin(lambda);
  const double tan2phi = (std::tan(phi)) * (std::tan(phi));
  const double tmp = 1 - e * e;
  const double tmpden = std::sqrt(1 + tmp * tan2phi);
  xyz[0] = (kA * coslam) / tmpden + h * coslam * cosphi;
  xyz[1] = (kA * sinlam) / tmpden + h * sinlam * cosphi;
  const double tmp2 = std::sqrt(1 - e * e * sinphi * sinphi);
  xyz[2] = (kA * tmp * sinphi) / tmp2 + h * sinphi;
}

// ECEF -> geodetic, closed form (CoordiTran.cpp:82-150).
void xyz2llh(double xyz[3], double llh[3]) {
  const double pi = 3.141592653589793;
  const double x = xyz[0], y = xyz[1], z = xyz[2];
  const double x2 = x * x, y2 = y * y, z2 = z * z;
  const double a = kA, b = kB;
  const double e = std::sqrt(1 - (b / a) * (b / a));
  const double b2 = b * b, e2 = e * e, ep = e * (a / b);
  const double r = std::sqrt(x2 + y2), r2 = r * r;
  const double E2 = a * a - b * b;
  const double F = 54 * b2 * z2;
  const double G = r2 + (1 - e2) * z2 - e2 * E2;
  const double c = (e2 * e2 * F * r2) / (G * G * G);
  const double s = std::pow(double(1 + c + std::sqrt(c * c + 2 * c)), double(1.0 / 3.0));
  const double P = F / (3 * (s + 1 / s + 1) * (s + 1 / s + 1) * G * G);
  const double Q = std::sqrt(1 + 2 * e2 * e2 * P);
  const double ro = -(P * e2 * r) / (1 + Q) +
                    std::sqrt((a * a / 2) * (1 + 1 / Q) - (P * (1 - e2) * z2) / (Q * (1 + Q)) - P * r2 / 2);
  const double tmp = (r - e2 * ro) * (r - e2 * ro);
  const double U = std::sqrt(tmp + z2);
  const double V = std::sqrt(tmp + (1 - e2) * z2);
  const double zo = (b2 * z) / (a * V);
  llh[2] = U * (a * V - b2) / (a * V);
  llh[0] = std::atan((z + ep * ep * zo) / r);
  const double t = std::atan(y / x);
  if (x >= 0)
    llh[1] = t;
  else if ((x < 0) & (y >= 0))
    llh[1] = pi + t;
  else
    llh[1] = t - pi;
}

// ECEF -> ENU about orgxyz (CoordiTran.cpp:152-187).
void xyz2enu(double xyz[3], double orgxyz[3], double enu[3]) {
  double dif[3], orgllh[3];
  for (int i = 0; i < 3; ++i) dif[i] = xyz[i] - orgxyz[i];
  xyz2llh(orgxyz, orgllh);
  const double sinphi = std::sin(orgllh[0]), cosphi = std::cos(orgllh[0]);
  const double sinlam = std::sin(orgllh[1]), coslam = std::cos(orgllh[1]);
  const double R[3][3] = {{-sinlam, coslam, 0},
                          {-sinphi * coslam, -sinphi * sinlam, cosphi},
                          {cosphi * coslam, cosphi * sinlam, sinphi}};
  enu[0] = enu[1] = enu[2] = 0;
  for (int i = 0; i < 3; ++i) {
    enu[0] = enu[0] + R[0][i] * dif[i];
    enu[1] = enu[1] + R[1][i] * dif[i];
    enu[2] = enu[2] + R[2][i] * dif[i];
  }
}

void llh2enu(double llh[3], double orgxyz[3], double enu[3]) {
  double xyz[3] = {0, 0, 0};
  llh2xyz(llh, xyz);
  xyz2enu(xyz, orgxyz, enu);
}
