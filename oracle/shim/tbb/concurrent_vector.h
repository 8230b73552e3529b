// stand-in for tbb::concurrent_vector: push_back returns an iterator, elements never move (std::deque), size() is safe
#pragma once
#include <deque>
#include <mutex>
namespace tbb {
template <class T>
class concurrent_vector {
  std::deque<T> d_; mutable std::mutex m_;
 public:
  typedef typename std::deque<T>::iterator iterator;
  typedef typename std::deque<T>::const_iterator const_iterator;
  typedef T value_type;
  typedef size_t size_type;
  concurrent_vector() {}
  explicit concurrent_vector(size_t n) : d_(n) {}
  iterator push_back(const T& v) { std::lock_guard<std::mutex> g(m_); d_.push_back(v); return d_.end() - 1; }
  iterator push_back(T&& v) { std::lock_guard<std::mutex> g(m_); d_.push_back(std::move(v)); return d_.end() - 1; }
  template <class... A> iterator emplace_back(A&&... a) { std::lock_guard<std::mutex> g(m_); d_.emplace_back(std::forward<A>(a)...); return d_.end() - 1; }
  iterator grow_by(size_t n) { std::lock_guard<std::mutex> g(m_); size_t o = d_.size(); d_.resize(o + n); return d_.begin() + o; }
  iterator grow_to_at_least(size_t n) { std::lock_guard<std::mutex> g(m_); size_t o = d_.size(); if (n > o) d_.resize(n); return d_.begin() + o; }
  size_t size() const { std::lock_guard<std::mutex> g(m_); return d_.size(); }
  bool empty() const { return size() == 0; }
  void resize(size_t n) { std::lock_guard<std::mutex> g(m_); d_.resize(n); }
  void reserve(size_t) {}
  void clear() { std::lock_guard<std::mutex> g(m_); d_.clear(); }
  void shrink_to_fit() {}
  T& operator[](size_t i) { return d_[i]; }
  const T& operator[](size_t i) const { return d_[i]; }
  T& at(size_t i) { return d_.at(i); }
  const T& at(size_t i) const { return d_.at(i); }
  T& back() { return d_.back(); }
  T& front() { return d_.front(); }
  iterator begin() { return d_.begin(); }
  iterator end() { return d_.end(); }
  const_iterator begin() const { return d_.begin(); }
  const_iterator end() const { return d_.end(); }
  const_iterator cbegin() const { return d_.cbegin(); }
  const_iterator cend() const { return d_.cend(); }
};
}  // namespace tbb
