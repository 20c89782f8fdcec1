// ref_shim: boost::shared_ptr as an alias of std::shared_ptr (test infrastructure)
#pragma once
#include <memory>
#include <exception>
namespace boost {
template <class T> using shared_ptr = std::shared_ptr<T>;
template <class T> using weak_ptr = std::weak_ptr<T>;
using std::make_shared;
// minimal intrusive_ptr: calls intrusive_ptr_add_ref / intrusive_ptr_release found by ADL
template <class T> class intrusive_ptr {
 public:
  intrusive_ptr() : p_(nullptr) {}
  intrusive_ptr(T* p, bool add_ref = true) : p_(p) { if (p_ && add_ref) intrusive_ptr_add_ref(p_); }
  intrusive_ptr(const intrusive_ptr& o) : p_(o.p_) { if (p_) intrusive_ptr_add_ref(p_); }
  ~intrusive_ptr() { if (p_) intrusive_ptr_release(p_); }
  intrusive_ptr& operator=(const intrusive_ptr& o) { intrusive_ptr(o).swap(*this); return *this; }
  void swap(intrusive_ptr& o) { T* t = p_; p_ = o.p_; o.p_ = t; }
  T* get() const { return p_; }
  T& operator*() const { return *p_; }
  T* operator->() const { return p_; }
  explicit operator bool() const { return p_ != nullptr; }
 private:
  T* p_;
};
struct exception { virtual ~exception() {} };
}  // namespace boost
