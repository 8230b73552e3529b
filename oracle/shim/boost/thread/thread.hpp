// stand-in for boost::thread / thread_group / mutex / unique_lock / ref / this_thread over std::
#pragma once
// (the real Boost headers pull in most of the standard library; the reference relies on that)
#include <array>
#include <atomic>
#include <chrono>
#include <functional>
#include <iomanip>
#include <mutex>
#include <thread>
#include <vector>
namespace boost {
class thread : public std::thread {
 public:
  using std::thread::thread;
};
class thread_group {
  std::vector<thread*> threads_;
 public:
  thread_group() = default;
  thread_group(const thread_group&) = delete;
  ~thread_group() { for (auto* t : threads_) delete t; }
  void add_thread(thread* t) { threads_.push_back(t); }
  void join_all() { for (auto* t : threads_) if (t->joinable()) t->join(); }
  size_t size() const { return threads_.size(); }
};
using mutex = std::mutex;
template <class M> using unique_lock = std::unique_lock<M>;
template <class M> using lock_guard = std::lock_guard<M>;
template <class T> std::reference_wrapper<T> ref(T& t) { return std::ref(t); }
template <class T> std::reference_wrapper<const T> cref(const T& t) { return std::cref(t); }
namespace chrono {
using milliseconds = std::chrono::milliseconds;
using seconds = std::chrono::seconds;
using microseconds = std::chrono::microseconds;
}  // namespace chrono
namespace this_thread {
template <class D> void sleep_for(const D& d) { std::this_thread::sleep_for(d); }
inline void yield() { std::this_thread::yield(); }
}  // namespace this_thread
}  // namespace boost
