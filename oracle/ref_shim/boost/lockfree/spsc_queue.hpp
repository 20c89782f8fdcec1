#pragma once
#include <boost/lockfree/queue.hpp>
namespace boost { namespace lockfree { template <class T> using spsc_queue = queue<T>; } }
