// FrameArena.h -- host memory of one decoded rotation, and the allocator that lets the
// reference's containers sit on top of it without a copy.
//
// The GPU lays a frame out exactly as the reference's HDLFrame holds it: one contiguous run of
// ready-made pcl::PointXYZI / PointMeta records per points[row] / pointsMeta[row] list
// (include/veloslam_b200.h, vs_layout_frames).  The facade copies that run device -> host ONCE,
// straight into a page-locked FrameArena, and every `points[row]->points` vector ADOPTS its slice
// of the arena: no per-point push_back, no second host copy.  The arena lives as long as any
// vector (or frame) that points into it and then goes back to a process-wide pool of pinned
// blocks (page-locking memory costs far more than a frame takes to decode).
//
// The vectors stay ordinary std::vector's in every other respect (vs::AdoptableVector derives
// publicly from std::vector): push_back / resize / assign beyond the adopted capacity allocate
// from the heap as usual.
#ifndef VELOSLAM_B200_FRAMEARENA_H
#define VELOSLAM_B200_FRAMEARENA_H

#include <cstddef>
#include <cstdint>
#include <memory>
#include <new>
#include <type_traits>
#include <utility>
#include <vector>

namespace vs {

// One page-locked block from the pool (FrameArena.cpp); returned to the pool on destruction.
class Arena {
 public:
  // nullptr when page-locked memory cannot be had (no CUDA device / out of memory)
  static std::shared_ptr<Arena> acquire(size_t bytes);
  // a plain heap block outside the pool: host-only tests and tools (nothing on the GPU path
  // uses it -- device -> host copies into pageable memory are staged and synchronous)
  static std::shared_ptr<Arena> heap(size_t bytes);
  ~Arena();
  uint8_t* data() const { return base_; }
  size_t capacity() const { return bytes_; }
  bool contains(const void* p) const {
    const uint8_t* q = static_cast<const uint8_t*>(p);
    return q >= base_ && q < base_ + bytes_;
  }
  // bytes page-locked by the pool right now (in use + cached), for tests and diagnostics
  static size_t pooledBytes();
  static void trimPool();

 private:
  Arena(uint8_t* b, size_t n, bool pooled) : base_(b), bytes_(n), pooled_(pooled) {}
  Arena(const Arena&);
  void operator=(const Arena&);
  uint8_t* base_;
  size_t bytes_;
  bool pooled_;
};

// std::allocator in every respect but one: storage that lies inside the arena belongs to the
// arena (deallocate leaves it alone), and the allocator keeps the arena alive.
template <class T>
class ArenaAllocator {
 public:
  typedef T value_type;
  typedef std::true_type propagate_on_container_move_assignment;
  typedef std::true_type propagate_on_container_copy_assignment;
  typedef std::true_type propagate_on_container_swap;
  typedef std::false_type is_always_equal;

  ArenaAllocator() {}
  explicit ArenaAllocator(std::shared_ptr<Arena> a) : arena_(std::move(a)) {}
  template <class U>
  ArenaAllocator(const ArenaAllocator<U>& o) : arena_(o.arena()) {}

  T* allocate(size_t n) { return static_cast<T*>(::operator new(n * sizeof(T))); }
  void deallocate(T* p, size_t) {
    if (arena_ && arena_->contains(p)) return;  // the arena owns it
    ::operator delete(p);
  }
  // a copy of an adopted vector is an ordinary heap vector
  ArenaAllocator select_on_container_copy_construction() const { return ArenaAllocator(); }
  const std::shared_ptr<Arena>& arena() const { return arena_; }

 private:
  std::shared_ptr<Arena> arena_;
};
template <class T, class U>
bool operator==(const ArenaAllocator<T>& a, const ArenaAllocator<U>& b) { return a.arena() == b.arena(); }
template <class T, class U>
bool operator!=(const ArenaAllocator<T>& a, const ArenaAllocator<U>& b) { return !(a == b); }

// A std::vector (publicly: every member, every conversion to the base) that can also be pointed
// at n records which already sit in an arena, in O(1): what `points[row]->points` and
// `*pointsMeta[row]` of a frame decoded on the GPU are.
template <class T>
class AdoptableVector : public std::vector<T, ArenaAllocator<T> > {
  typedef std::vector<T, ArenaAllocator<T> > Base;

 public:
  using Base::Base;
  AdoptableVector() {}

  void adoptStorage(const std::shared_ptr<Arena>& arena, T* data, size_t n) {
    {
      Base fresh((ArenaAllocator<T>(arena)));  // empty, allocator bound to the arena
      Base::swap(fresh);                       // old contents (and allocator) die with `fresh`
    }
    if (n == 0) return;
#if defined(__GLIBCXX__)
    // libstdc++: begin / end / end-of-storage live in the protected _M_impl of the base
    static_assert(std::is_trivially_destructible<T>::value && std::is_trivially_copyable<T>::value,
                  "adopted records are plain data");
    this->_M_impl._M_start = data;
    this->_M_impl._M_finish = data + n;
    this->_M_impl._M_end_of_storage = data + n;
#else
    Base::assign(data, data + n);  // other standard libraries: one copy per row
#endif
  }
};

// Make `v` the n records at `data` (inside `arena`) without touching them.
template <class T>
void adopt(AdoptableVector<T>& v, const std::shared_ptr<Arena>& arena, T* data, size_t n) {
  v.adoptStorage(arena, data, n);
}

}  // namespace vs

#endif
