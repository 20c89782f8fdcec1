#include "TimeSolver.h"

#include <sys/time.h>

namespace {
int64_t wallClockUs() {
  timeval tv;
  gettimeofday(&tv, nullptr);
  return (int64_t)tv.tv_sec * 1000000ll + tv.tv_usec;
}
const int64_t kHourUs = 3600ll * 1000000ll;
}  // namespace

TimeSolver::TimeSolver() : now_(wallClockUs), hdlBaseUs_(0), hdlInited_(false), lastHdlReport_(0) {}

// TimeSolver.cxx:20-33.  insInited is never set in the reference, so the offset is re-taken
// from the clock on every call: the result is now + (time of pose - time of packet send).
ptime TimeSolver::calcTimestamp(InsPVA const* d) {
  const int64_t sent = (int64_t)(d->week_number * 168) * kHourUs + (int64_t)(double(d->milliseconds) * 1e3);
  const int64_t pose = (int64_t)(d->week_number_pos * 168) * kHourUs + (int64_t)(double(d->seconds_pos) * 1e6);
  return ptime(now_() + (pose - sent));
}

// TimeSolver.cxx:34-49
ptime TimeSolver::calcTimestamp(uint32_t microsecToHour) {
  if (!hdlInited_) {
    hdlBaseUs_ = now_() - (int64_t)microsecToHour;
    hdlInited_ = true;
  }
  if (lastHdlReport_ > microsecToHour) hdlBaseUs_ += kHourUs;  // one hour wrapped
  lastHdlReport_ = microsecToHour;
  return ptime(hdlBaseUs_ + (int64_t)microsecToHour);
}
