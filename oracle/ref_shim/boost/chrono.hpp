#pragma once
#include <chrono>
namespace boost { namespace chrono {
using std::chrono::microseconds; using std::chrono::milliseconds; using std::chrono::seconds;
using std::chrono::duration_cast;
struct thread_clock : std::chrono::steady_clock {};
} }
