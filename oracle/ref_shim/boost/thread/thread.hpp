#pragma once
#include <condition_variable>
#include <mutex>
#include <thread>
#include <boost/chrono.hpp>
namespace boost {
using std::mutex; using std::thread; using std::condition_variable; using std::cv_status;
template <class M> using unique_lock = std::unique_lock<M>;
template <class M> using lock_guard = std::lock_guard<M>;
namespace this_thread { using std::this_thread::sleep_for; }
}
