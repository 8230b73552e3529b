// stand-in for tbb::concurrent_queue (unbounded MPMC FIFO; try_pop)
#pragma once
#include <deque>
#include <mutex>
namespace tbb {
template <class T>
class concurrent_queue {
  std::deque<T> q_; mutable std::mutex m_;
 public:
  void push(const T& v) { std::lock_guard<std::mutex> g(m_); q_.push_back(v); }
  bool try_pop(T& out) { std::lock_guard<std::mutex> g(m_); if (q_.empty()) return false; out = q_.front(); q_.pop_front(); return true; }
  bool empty() const { std::lock_guard<std::mutex> g(m_); return q_.empty(); }
  size_t unsafe_size() const { std::lock_guard<std::mutex> g(m_); return q_.size(); }
  void clear() { std::lock_guard<std::mutex> g(m_); q_.clear(); }
};
}  // namespace tbb
