#pragma once
#include <ostream>
namespace ref_shim { struct NullStream { template <class T> NullStream& operator<<(const T&) { return *this; }
  NullStream& operator<<(std::ostream& (*)(std::ostream&)) { return *this; } }; }
#define DLOG(x) ::ref_shim::NullStream()
#define LOG(x) ::ref_shim::NullStream()
#define DLOG_IF(x, c) ::ref_shim::NullStream()
#define CHECK(x) ::ref_shim::NullStream()
