#pragma once
#include <deque>
#include <mutex>
namespace boost { namespace lockfree {
template <class T> class queue {
 public:
  queue() {}
  explicit queue(size_t) {}
  bool push(const T& v) { std::lock_guard<std::mutex> l(m_); q_.push_back(v); return true; }
  bool pop(T& v) { std::lock_guard<std::mutex> l(m_); if (q_.empty()) return false; v = q_.front(); q_.pop_front(); return true; }
  bool empty() { std::lock_guard<std::mutex> l(m_); return q_.empty(); }
 private:
  std::deque<T> q_; std::mutex m_;
};
} }
