// FrameArena.cpp -- process-wide pool of page-locked blocks behind vs::Arena.
#include "FrameArena.h"

#include <map>
#include <mutex>

#include "../../include/veloslam_b200.h"

namespace vs {
namespace {

struct Pool {
  std::mutex m;
  std::multimap<size_t, uint8_t*> free;  // cached blocks by size
  size_t cachedBytes = 0;
  size_t liveBytes = 0;
  // blocks kept for re-use; beyond this they are unlocked and freed
  static constexpr size_t kMaxCached = 2ull << 30;
};
// leaked on purpose: frames (and their arenas) may be destroyed after static destructors ran
Pool& pool() {
  static Pool* p = new Pool;
  return *p;
}

size_t roundUp(size_t bytes) {
  size_t g = 64u << 10;
  while (g < bytes && g < (1u << 20)) g <<= 1;
  if (bytes <= g) return g;
  return (bytes + (1u << 20) - 1) & ~(size_t)((1u << 20) - 1);
}

}  // namespace

std::shared_ptr<Arena> Arena::acquire(size_t bytes) {
  const size_t want = roundUp(bytes ? bytes : 1);
  Pool& P = pool();
  uint8_t* block = nullptr;
  size_t got = 0;
  {
    std::lock_guard<std::mutex> lock(P.m);
    auto it = P.free.lower_bound(want);
    if (it != P.free.end() && it->first <= want + want / 2 + (1u << 20)) {
      block = it->second;
      got = it->first;
      P.cachedBytes -= got;
      P.free.erase(it);
    }
  }
  if (!block) {
    void* p = nullptr;
    if (vs_host_alloc(want, &p) != VS_OK || !p) return std::shared_ptr<Arena>();
    block = static_cast<uint8_t*>(p);
    got = want;
  }
  {
    std::lock_guard<std::mutex> lock(P.m);
    P.liveBytes += got;
  }
  return std::shared_ptr<Arena>(new Arena(block, got, true));
}

std::shared_ptr<Arena> Arena::heap(size_t bytes) {
  uint8_t* block = static_cast<uint8_t*>(::operator new(bytes ? bytes : 1));
  return std::shared_ptr<Arena>(new Arena(block, bytes ? bytes : 1, false));
}

Arena::~Arena() {
  if (!pooled_) {
    ::operator delete(base_);
    return;
  }
  Pool& P = pool();
  bool keep = false;
  {
    std::lock_guard<std::mutex> lock(P.m);
    P.liveBytes -= bytes_;
    if (P.cachedBytes + bytes_ <= Pool::kMaxCached) {
      P.free.insert(std::make_pair(bytes_, base_));
      P.cachedBytes += bytes_;
      keep = true;
    }
  }
  if (!keep) vs_host_free(base_);
}

size_t Arena::pooledBytes() {
  Pool& P = pool();
  std::lock_guard<std::mutex> lock(P.m);
  return P.liveBytes + P.cachedBytes;
}

void Arena::trimPool() {
  Pool& P = pool();
  std::multimap<size_t, uint8_t*> drop;
  {
    std::lock_guard<std::mutex> lock(P.m);
    drop.swap(P.free);
    P.cachedBytes = 0;
  }
  for (auto& kv : drop) vs_host_free(kv.second);
}

}  // namespace vs
