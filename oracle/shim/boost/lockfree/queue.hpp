// stand-in for boost::lockfree::queue<T, fixed_sized<true>>: bounded MPMC queue, push/pop return false when full/empty
#pragma once
#include <deque>
#include <mutex>
namespace boost { namespace lockfree {
template <bool B> struct fixed_sized {};
template <size_t N> struct capacity {};
template <class T, class... Options>
class queue {
  std::deque<T> q_; mutable std::mutex m_; size_t cap_;
 public:
  explicit queue(size_t n = 0) : cap_(n) {}
  bool push(const T& v) { std::lock_guard<std::mutex> g(m_); if (cap_ && q_.size() >= cap_) return false; q_.push_back(v); return true; }
  bool bounded_push(const T& v) { return push(v); }
  bool unsynchronized_push(const T& v) { return push(v); }
  bool pop(T& out) { std::lock_guard<std::mutex> g(m_); if (q_.empty()) return false; out = q_.front(); q_.pop_front(); return true; }
  bool unsynchronized_pop(T& out) { return pop(out); }
  bool empty() const { std::lock_guard<std::mutex> g(m_); return q_.empty(); }
};
} }  // namespace boost::lockfree
