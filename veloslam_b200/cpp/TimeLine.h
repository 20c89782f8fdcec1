// TimeLine.h -- time-ordered container with the reference's TimeLine<T> interface
// (/root/reference/TimeLine.h:23-137), stored as one sorted vector instead of time buckets plus
// a circular buffer: the pose timeline is uploaded to HBM as flat sorted arrays, so the flat
// layout is the native one.  Net lookup semantics are the reference's:
//   getBoundaryData(t): i = clamp(lower_bound(t), 1, N-1) -> (item[i-1], item[i])
//   (TimeLine.h:384-468; on an exact hit the reference may return (hit, next) instead, which
//   interpolates to the same pose to <= 1 ulp)
#ifndef VELOSLAM_B200_TIMELINE_H
#define VELOSLAM_B200_TIMELINE_H

#include <algorithm>
#include <memory>
#include <utility>
#include <vector>

#include "type_defs.h"

template <typename T_>
class TimeLine {
 public:
  typedef std::shared_ptr<T_> Ptr;
  size_t size() const { return items_.size(); }
  Ptr back() const { return items_.empty() ? Ptr() : items_.back(); }
  void clear() { items_.clear(); }
  void unload() { std::vector<Ptr>().swap(items_); }

  // keeps items time-sorted; an item with an existing timestamp overwrites the old one
  // (reference TimeLine.h:140-226)
  void addData(Ptr data) {
    if (items_.empty() || data->timestamp > items_.back()->timestamp) {
      items_.push_back(data);
      return;
    }
    auto it = lower(data->timestamp);
    if (it != items_.end() && (*it)->timestamp == data->timestamp)
      *it = data;
    else
      items_.insert(it, data);
  }

  Ptr getExactDataAt(const ptime& t) const {
    auto it = lower(t);
    return (it != items_.end() && (*it)->timestamp == t) ? *it : Ptr();
  }

  Ptr getNearestData(const ptime& t) const {
    if (items_.empty()) return Ptr();
    auto it = lower(t);
    if (it == items_.begin()) return *it;
    if (it == items_.end()) return items_.back();
    auto prev = it - 1;
    return ((t - (*prev)->timestamp).us <= ((*it)->timestamp - t).us) ? *prev : *it;
  }
  ptime getNearestTime(const ptime& t) const {
    Ptr p = getNearestData(t);
    return p ? p->timestamp : ptime();
  }

  std::pair<Ptr, Ptr> getBoundaryData(const ptime& t) const {
    const size_t n = items_.size();
    if (n == 0) return std::make_pair(Ptr(), Ptr());
    if (n == 1) return std::make_pair(items_[0], Ptr());
    size_t i = lower(t) - items_.begin();
    if (i < 1) i = 1;
    if (i > n - 1) i = n - 1;
    return std::make_pair(items_[i - 1], items_[i]);
  }
  std::pair<ptime, ptime> getBoundaryTime(const ptime& t) const {
    auto b = getBoundaryData(t);
    return std::make_pair(b.first ? b.first->timestamp : ptime(),
                          b.second ? b.second->timestamp : ptime());
  }

  // items with a <= timestamp < b
  std::vector<Ptr> getRangeBetween(const ptime& a, const ptime& b) const {
    auto lo = lower(a), hi = lower(b);
    return lo < hi ? std::vector<Ptr>(lo, hi) : std::vector<Ptr>();
  }
  std::vector<Ptr> getAll() const { return items_; }
  const std::vector<Ptr>& items() const { return items_; }

 private:
  typename std::vector<Ptr>::const_iterator lower(const ptime& t) const {
    return std::lower_bound(items_.begin(), items_.end(), t,
                            [](const Ptr& p, const ptime& v) { return p->timestamp < v; });
  }
  typename std::vector<Ptr>::iterator lower(const ptime& t) {
    return std::lower_bound(items_.begin(), items_.end(), t,
                            [](const Ptr& p, const ptime& v) { return p->timestamp < v; });
  }
  std::vector<Ptr> items_;
};

#endif
