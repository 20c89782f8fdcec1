// ref_shim: boost::circular_buffer restricted to what TimeLine.h / HDLManager.h use, with the
// overwrite rules documented by Boost.CircularBuffer (test infrastructure).
#pragma once
#include <deque>
#include <cstddef>
namespace boost {
template <class T> class circular_buffer {
 public:
  typedef typename std::deque<T>::iterator iterator;
  circular_buffer() : cap_(0) {}
  explicit circular_buffer(std::size_t cap) : cap_(cap) {}
  std::size_t size() const { return d_.size(); }
  std::size_t capacity() const { return cap_; }
  bool empty() const { return d_.empty(); }
  bool full() const { return d_.size() == cap_; }
  void clear() { d_.clear(); }
  T& operator[](std::size_t i) { return d_[i]; }
  T& back() { return d_.back(); }
  T& front() { return d_.front(); }
  iterator begin() { return d_.begin(); }
  iterator end() { return d_.end(); }
  void push_back(const T& v) {           // full: the first element is overwritten
    if (cap_ == 0) return;
    if (d_.size() == cap_) d_.pop_front();
    d_.push_back(v);
  }
  void push_front(const T& v) {          // full: the last element is overwritten
    if (cap_ == 0) return;
    if (d_.size() == cap_) d_.pop_back();
    d_.push_front(v);
  }
  iterator insert(iterator pos, const T& v) {
    // full: the first element is overwritten; full and pos == begin(): nothing is inserted
    if (cap_ == 0) return d_.end();
    std::size_t idx = pos - d_.begin();
    if (d_.size() == cap_) {
      if (idx == 0) return d_.begin();
      d_.pop_front();
      --idx;
    }
    return d_.insert(d_.begin() + idx, v);
  }
  void swap(circular_buffer& o) { d_.swap(o.d_); std::size_t c = cap_; cap_ = o.cap_; o.cap_ = c; }
 private:
  std::deque<T> d_;
  std::size_t cap_;
};
}  // namespace boost
