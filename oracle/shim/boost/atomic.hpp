#pragma once
#include <atomic>
namespace boost { template <class T> using atomic = std::atomic<T>; }
