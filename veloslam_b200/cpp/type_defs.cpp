#include "type_defs.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>

PoseTransform::PoseTransform() {
  for (int i = 0; i < 3; ++i) {
    T[i] = 0;
    R[i] = 0;
    V[i] = 0;
  }
  week_number = 0;
  milliseconds = week_number_pos = 0;
  seconds_pos = -1;
}

PoseTransform PoseTransform::operator+(PoseTransform delta) const {
  PoseTransform result;
  for (int i = 0; i < 3; ++i) {
    result.T[i] = T[i] + delta.T[i];
    result.R[i] = R[i] + delta.R[i];
    result.V[i] = V[i] + delta.V[i];
  }
  return result;
}
PoseTransform PoseTransform::operator*(double ratio) const {
  PoseTransform result;
  for (int i = 0; i < 3; ++i) {
    result.T[i] = T[i] * ratio;
    result.R[i] = R[i] * ratio;
    result.V[i] = V[i] * ratio;
  }
  return result;
}
PoseTransform PoseTransform::operator-(PoseTransform delta) const {
  PoseTransform result;
  for (int i = 0; i < 3; ++i) {
    result.T[i] = T[i] - delta.T[i];
    result.R[i] = R[i] - delta.R[i];
    result.V[i] = V[i] - delta.V[i];
  }
  return result;
}

namespace {
// L <- L * AngleAxis(angle, unit axis).toRotationMatrix()  (Rodrigues form, as Eigen 3.x)
void rotate(double L[3][3], double angle, int axis) {
  const double s = std::sin(angle), c = std::cos(angle);
  const double ax[3] = {axis == 0 ? 1.0 : 0.0, axis == 1 ? 1.0 : 0.0, axis == 2 ? 1.0 : 0.0};
  double sa[3], ca[3], Rm[3][3];
  for (int k = 0; k < 3; ++k) {
    sa[k] = s * ax[k];
    ca[k] = (1.0 - c) * ax[k];
  }
  double tmp = ca[0] * ax[1];
  Rm[0][1] = tmp - sa[2];
  Rm[1][0] = tmp + sa[2];
  tmp = ca[0] * ax[2];
  Rm[0][2] = tmp + sa[1];
  Rm[2][0] = tmp - sa[1];
  tmp = ca[1] * ax[2];
  Rm[1][2] = tmp - sa[0];
  Rm[2][1] = tmp + sa[0];
  Rm[0][0] = ca[0] * ax[0] + c;
  Rm[1][1] = ca[1] * ax[1] + c;
  Rm[2][2] = ca[2] * ax[2] + c;
  double out[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) out[i][j] = L[i][0] * Rm[0][j] + L[i][1] * Rm[1][j] + L[i][2] * Rm[2][j];
  std::memcpy(L, out, sizeof(out));
}
}  // namespace

Affine3d PoseTransform::getMatrix() const {
  Affine3d a;
  double L[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  rotate(L, TO_RADIUS(R[0]), 1);  // UnitY
  rotate(L, TO_RADIUS(R[1]), 0);  // UnitX
  rotate(L, TO_RADIUS(R[2]), 2);  // UnitZ
  std::memcpy(a.L, L, sizeof(L));
  a.t[0] = T[0];
  a.t[1] = T[1];
  a.t[2] = T[2];
  return a;
}

void transformPoint(double pt0[3], const Affine3d& m) {
  const double x = pt0[0], y = pt0[1], z = pt0[2];
  pt0[0] = m.L[0][0] * x + m.L[0][1] * y + m.L[0][2] * z + m.t[0];
  pt0[1] = m.L[1][0] * x + m.L[1][1] * y + m.L[1][2] * z + m.t[1];
  pt0[2] = m.L[2][0] * x + m.L[2][1] * y + m.L[2][2] * z + m.t[2];
}

std::string to_iso_string(const ptime& t) {
  if (t.is_special()) return "not-a-date-time";
  int64_t us = t.us;
  int64_t sec = us / 1000000;
  int64_t frac = us % 1000000;
  if (frac < 0) {
    frac += 1000000;
    --sec;
  }
  std::time_t tt = (std::time_t)sec;
  std::tm m;
  gmtime_r(&tt, &m);
  char buf[64];
  if (frac)
    std::snprintf(buf, sizeof(buf), "%04d%02d%02dT%02d%02d%02d.%06lld", m.tm_year + 1900,
                  m.tm_mon + 1, m.tm_mday, m.tm_hour, m.tm_min, m.tm_sec, (long long)frac);
  else
    std::snprintf(buf, sizeof(buf), "%04d%02d%02dT%02d%02d%02d", m.tm_year + 1900, m.tm_mon + 1,
                  m.tm_mday, m.tm_hour, m.tm_min, m.tm_sec);
  return buf;
}

bool from_iso_string(const std::string& s, ptime* out) {
  int y, mo, d, h, mi, se;
  if (s.size() < 15 || s[8] != 'T' ||
      std::sscanf(s.c_str(), "%4d%2d%2dT%2d%2d%2d", &y, &mo, &d, &h, &mi, &se) != 6)
    return false;
  long frac = 0;
  if (s.size() > 16 && s[15] == '.') {
    std::string f = s.substr(16);
    f.resize(6, '0');
    frac = std::atol(f.c_str());
  }
  std::tm m = std::tm();
  m.tm_year = y - 1900;
  m.tm_mon = mo - 1;
  m.tm_mday = d;
  m.tm_hour = h;
  m.tm_min = mi;
  m.tm_sec = se;
  const std::time_t tt = timegm(&m);
  *out = ptime((int64_t)tt * 1000000ll + frac);
  return true;
}
