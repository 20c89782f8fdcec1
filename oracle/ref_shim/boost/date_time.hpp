// ref_shim: the slice of Boost.DateTime the reference's hot-path sources use, with ptime as
// int64 microseconds since 1970-01-01 plus a not_a_date_time state (test infrastructure).
#pragma once
#include <climits>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <ostream>
#include <string>
#include <sys/time.h>
#include <boost/shared_ptr.hpp>

namespace boost {
namespace date_detail {
inline int64_t days_from_civil(int64_t y, unsigned m, unsigned d) {
  y -= m <= 2;
  const int64_t era = (y >= 0 ? y : y - 399) / 400;
  const unsigned yoe = static_cast<unsigned>(y - era * 400);
  const unsigned doy = (153 * (m + (m > 2 ? -3 : 9)) + 2) / 5 + d - 1;
  const unsigned doe = yoe * 365 + yoe / 4 - yoe / 100 + doy;
  return era * 146097 + static_cast<int64_t>(doe) - 719468;
}
inline void civil_from_days(int64_t z, int64_t& y, unsigned& m, unsigned& d) {
  z += 719468;
  const int64_t era = (z >= 0 ? z : z - 146096) / 146097;
  const unsigned doe = static_cast<unsigned>(z - era * 146097);
  const unsigned yoe = (doe - doe / 1460 + doe / 36524 - doe / 146096) / 365;
  y = static_cast<int64_t>(yoe) + era * 400;
  const unsigned doy = doe - (365 * yoe + yoe / 4 - yoe / 100);
  const unsigned mp = (5 * doy + 2) / 153;
  d = doy - (153 * mp + 2) / 5 + 1;
  m = mp + (mp < 10 ? 3 : -9);
  y += (m <= 2);
}
}  // namespace date_detail

// test hook: when set (!= INT64_MIN) every clock below reads this instead of the wall clock, so
// that the reference's wall-clock dependent code (TimeSolver.cxx) is reproducible
namespace shim_clock {
inline int64_t& fake_now_us() {
  static int64_t v = INT64_MIN;
  return v;
}
inline int64_t now_us() {
  if (fake_now_us() != INT64_MIN) return fake_now_us();
  timeval tv;
  gettimeofday(&tv, nullptr);
  return static_cast<int64_t>(tv.tv_sec) * 1000000ll + tv.tv_usec;
}
}  // namespace shim_clock

namespace gregorian {
struct days {
  int64_t n;
  explicit days(int64_t v) : n(v) {}
};
class date {
 public:
  date() : d_(0) {}
  explicit date(int64_t days_since_epoch) : d_(days_since_epoch) {}
  date(int y, int m, int d) : d_(date_detail::days_from_civil(y, m, d)) {}
  int64_t day_count() const { return d_; }
  int year() const {
    int64_t y; unsigned m, d;
    date_detail::civil_from_days(d_, y, m, d);
    return static_cast<int>(y);
  }
  date operator-(const days& o) const { return date(d_ - o.n); }
  date operator+(const days& o) const { return date(d_ + o.n); }
  int week_number() const {  // ISO 8601 week number
    const int wd = static_cast<int>(((d_ % 7) + 10) % 7);  // Monday = 0 (1970-01-01 was a Thursday)
    const int64_t thursday = d_ - wd + 3;
    int64_t ty; unsigned tm_, td;
    date_detail::civil_from_days(thursday, ty, tm_, td);
    const int64_t jan1 = date_detail::days_from_civil(ty, 1, 1);
    return static_cast<int>((thursday - jan1) / 7) + 1;
  }
 private:
  int64_t d_;
};
struct day_clock {
  static date local_day() {
    const int64_t us = shim_clock::now_us();
    int64_t d = us / 86400000000ll;
    if (us % 86400000000ll < 0) --d;
    return date(d);
  }
};
inline std::tm to_tm(const date& d) {
  std::tm t = std::tm();
  int64_t y; unsigned m, dd;
  date_detail::civil_from_days(d.day_count(), y, m, dd);
  t.tm_year = static_cast<int>(y - 1900);
  t.tm_mon = static_cast<int>(m) - 1;
  t.tm_mday = static_cast<int>(dd);
  t.tm_wday = static_cast<int>(((d.day_count() % 7) + 11) % 7);  // Sunday = 0
  t.tm_yday = static_cast<int>(d.day_count() - date_detail::days_from_civil(y, 1, 1));
  t.tm_isdst = -1;
  return t;
}
}  // namespace gregorian

namespace posix_time {
class time_duration {
 public:
  time_duration() : us_(0) {}
  time_duration(int64_t h, int64_t m, int64_t s, int64_t frac = 0)
      : us_(((h * 60 + m) * 60 + s) * 1000000ll + frac) {}
  static time_duration from_us(int64_t us) { time_duration d; d.us_ = us; return d; }
  int64_t total_microseconds() const { return us_; }
  int64_t total_milliseconds() const { return us_ / 1000; }
  int64_t total_seconds() const { return us_ / 1000000; }
  int64_t fractional_seconds() const { return us_ % 1000000; }
  int64_t hours() const { return us_ / 3600000000ll; }
  int64_t minutes() const { return (us_ / 60000000ll) % 60; }
  int64_t seconds() const { return (us_ / 1000000ll) % 60; }
  bool operator<(const time_duration& o) const { return us_ < o.us_; }
  bool operator>(const time_duration& o) const { return us_ > o.us_; }
  bool operator<=(const time_duration& o) const { return us_ <= o.us_; }
  bool operator>=(const time_duration& o) const { return us_ >= o.us_; }
  bool operator==(const time_duration& o) const { return us_ == o.us_; }
  bool operator!=(const time_duration& o) const { return us_ != o.us_; }
  time_duration operator+(const time_duration& o) const { return from_us(us_ + o.us_); }
  time_duration operator-(const time_duration& o) const { return from_us(us_ - o.us_); }
  time_duration operator-() const { return from_us(-us_); }
 private:
  int64_t us_;
};
inline time_duration hours(int64_t h) { return time_duration(h, 0, 0, 0); }
inline time_duration minutes(int64_t m) { return time_duration(0, m, 0, 0); }
inline time_duration seconds(int64_t s) { return time_duration(0, 0, s, 0); }
inline time_duration milliseconds(int64_t v) { return time_duration::from_us(v * 1000); }
inline time_duration microseconds(int64_t v) { return time_duration::from_us(v); }

class ptime {
 public:
  ptime() : us_(0), special_(true) {}
  explicit ptime(const gregorian::date& d) : us_(d.day_count() * 86400000000ll), special_(false) {}
  ptime(const gregorian::date& d, const time_duration& t)
      : us_(d.day_count() * 86400000000ll + t.total_microseconds()), special_(false) {}
  static ptime from_us(int64_t us) { ptime p; p.us_ = us; p.special_ = false; return p; }
  int64_t us() const { return us_; }
  bool is_special() const { return special_; }
  bool is_not_a_date_time() const { return special_; }
  gregorian::date date() const {
    int64_t d = us_ / 86400000000ll;
    if (us_ % 86400000000ll < 0) --d;
    return gregorian::date(d);
  }
  time_duration time_of_day() const {
    return time_duration::from_us(us_ - date().day_count() * 86400000000ll);
  }
  // not_a_date_time sorts before every real time and equals only itself
  bool operator==(const ptime& o) const { return special_ == o.special_ && (special_ || us_ == o.us_); }
  bool operator!=(const ptime& o) const { return !(*this == o); }
  bool operator<(const ptime& o) const { return key() < o.key(); }
  bool operator>(const ptime& o) const { return key() > o.key(); }
  bool operator<=(const ptime& o) const { return !(*this > o); }
  bool operator>=(const ptime& o) const { return !(*this < o); }
  time_duration operator-(const ptime& o) const { return time_duration::from_us(us_ - o.us_); }
  ptime operator+(const time_duration& d) const { ptime p(*this); p.us_ += d.total_microseconds(); return p; }
  ptime operator-(const time_duration& d) const { ptime p(*this); p.us_ -= d.total_microseconds(); return p; }
  ptime& operator+=(const time_duration& d) { us_ += d.total_microseconds(); return *this; }
  ptime& operator-=(const time_duration& d) { us_ -= d.total_microseconds(); return *this; }
 private:
  __int128 key() const { return special_ ? -(((__int128)1) << 100) : (__int128)us_; }
  int64_t us_;
  bool special_;
};

inline ptime from_time_t(std::time_t t) { return ptime::from_us(static_cast<int64_t>(t) * 1000000ll); }
inline std::tm to_tm(const ptime& t) {
  std::tm r = gregorian::to_tm(t.date());
  const time_duration d = t.time_of_day();
  r.tm_hour = static_cast<int>(d.hours());
  r.tm_min = static_cast<int>(d.minutes());
  r.tm_sec = static_cast<int>(d.seconds());
  return r;
}
struct bad_time_string : boost::exception {};
inline std::string to_iso_string(const ptime& t) {
  if (t.is_special()) return "not-a-date-time";
  std::tm m = to_tm(t);
  char buf[64];
  const long frac = static_cast<long>(t.time_of_day().fractional_seconds());
  if (frac)
    std::snprintf(buf, sizeof(buf), "%04d%02d%02dT%02d%02d%02d.%06ld", m.tm_year + 1900, m.tm_mon + 1,
                  m.tm_mday, m.tm_hour, m.tm_min, m.tm_sec, frac);
  else
    std::snprintf(buf, sizeof(buf), "%04d%02d%02dT%02d%02d%02d", m.tm_year + 1900, m.tm_mon + 1,
                  m.tm_mday, m.tm_hour, m.tm_min, m.tm_sec);
  return buf;
}
inline ptime from_iso_string(const std::string& s) {
  int y, mo, d, h, mi, se;
  long frac = 0;
  if (s.size() < 15 || s[8] != 'T' ||
      std::sscanf(s.c_str(), "%4d%2d%2dT%2d%2d%2d", &y, &mo, &d, &h, &mi, &se) != 6)
    throw bad_time_string();
  if (s.size() > 16 && s[15] == '.') {
    std::string f = s.substr(16);
    f.resize(6, '0');
    frac = std::atol(f.c_str());
  }
  return ptime(gregorian::date(y, mo, d), time_duration(h, mi, se, frac));
}
inline std::ostream& operator<<(std::ostream& os, const ptime& t) { return os << to_iso_string(t); }
inline std::ostream& operator<<(std::ostream& os, const time_duration& d) {
  return os << d.total_microseconds() << "us";
}
struct microsec_clock {
  static ptime local_time() { return ptime::from_us(shim_clock::now_us()); }
  static ptime universal_time() { return local_time(); }
};
}  // namespace posix_time
}  // namespace boost
