#include "CalibrationFile.h"

#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>

namespace {
// text of the first <name ...>...</name> element inside [from, to); npos-safe
bool element(const std::string& x, const std::string& name, size_t from, size_t to, size_t* body,
             size_t* body_end, size_t* after) {
  const std::string open = "<" + name;
  size_t p = from;
  while (true) {
    p = x.find(open, p);
    if (p == std::string::npos || p >= to) return false;
    const char c = x[p + open.size()];
    if (c == '>' || c == ' ' || c == '\t' || c == '\n' || c == '\r') break;
    p += open.size();
  }
  const size_t gt = x.find('>', p);
  if (gt == std::string::npos || gt >= to) return false;
  const std::string close = "</" + name + ">";
  const size_t e = x.find(close, gt);
  if (e == std::string::npos || e > to) return false;
  *body = gt + 1;
  *body_end = e;
  *after = e + close.size();
  return true;
}
double number(const std::string& x, const std::string& name, size_t from, size_t to, bool* found) {
  size_t b, e, a;
  if (!element(x, name, from, to, &b, &e, &a)) {
    *found = false;
    return 0.0;
  }
  *found = true;
  return std::atof(x.substr(b, e - b).c_str());
}
}  // namespace

bool CalibrationFile::load(const std::string& filename, std::string* error) {
  std::ifstream f(filename.c_str());
  if (!f) {
    if (error) *error = "cannot open " + filename;
    return false;
  }
  std::stringstream ss;
  ss << f.rdbuf();
  const std::string x = ss.str();
  std::memset(rows, 0, sizeof(rows));
  n_rows = 0;
  n_enabled = 0;
  size_t b, e, a;
  if (element(x, "enabled_", 0, x.size(), &b, &e, &a)) {
    size_t p = b, ib, ie, ia;
    while (element(x, "item", p, e, &ib, &ie, &ia)) {
      if (std::atoi(x.substr(ib, ie - ib).c_str()) == 1) ++n_enabled;
      p = ia;
    }
  }
  size_t p = 0;
  bool any = false;
  while (element(x, "px", p, x.size(), &b, &e, &a)) {
    bool ok;
    const int id = (int)number(x, "id_", b, e, &ok);
    if (ok && id >= 0 && id < VS_MAX_LASERS) {
      bool dummy;
      rows[id].rot_correction_deg = number(x, "rotCorrection_", b, e, &dummy);
      rows[id].vert_correction_deg = number(x, "vertCorrection_", b, e, &dummy);
      rows[id].dist_correction_cm = number(x, "distCorrection_", b, e, &dummy);
      rows[id].vert_offset_correction_cm = number(x, "vertOffsetCorrection_", b, e, &dummy);
      rows[id].horiz_offset_correction_cm = number(x, "horizOffsetCorrection_", b, e, &dummy);
      if (id + 1 > n_rows) n_rows = id + 1;
      any = true;
    }
    p = a;
  }
  if (!any && error) *error = "no <px> calibration items in " + filename;
  return any;
}
