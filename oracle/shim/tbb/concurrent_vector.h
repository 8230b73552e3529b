// stand-in for tbb::concurrent_vector: growing never moves elements and never invalidates iterators (the reference
// keeps an iterator to "its" contig across push_back calls of other walkers and subtracts begin() from it to get the
// contig id), push_back returns an iterator, size() is safe to call concurrently.
#pragma once
#include <cstddef>
#include <deque>
#include <iterator>
#include <mutex>
namespace tbb {
template <class T>
class concurrent_vector {
  std::deque<T> d_; mutable std::mutex m_;   // deque: references stay valid on push_back; iterators below are index based
  template <class V, class Ref>
  class iter {
    V* v_; size_t i_;
   public:
    typedef std::random_access_iterator_tag iterator_category;
    typedef T value_type; typedef ptrdiff_t difference_type; typedef Ref* pointer; typedef Ref& reference;
    iter() : v_(nullptr), i_(0) {}
    iter(V* v, size_t i) : v_(v), i_(i) {}
    template <class V2, class R2> iter(const iter<V2, R2>& o) : v_(o.container()), i_(o.index()) {}
    V* container() const { return v_; }
    size_t index() const { return i_; }
    Ref& operator*() const { return (*v_)[i_]; }
    Ref* operator->() const { return &(*v_)[i_]; }
    Ref& operator[](ptrdiff_t n) const { return (*v_)[i_ + n]; }
    iter& operator++() { ++i_; return *this; }
    iter operator++(int) { iter t = *this; ++i_; return t; }
    iter& operator--() { --i_; return *this; }
    iter operator--(int) { iter t = *this; --i_; return t; }
    iter& operator+=(ptrdiff_t n) { i_ += n; return *this; }
    iter& operator-=(ptrdiff_t n) { i_ -= n; return *this; }
    iter operator+(ptrdiff_t n) const { return iter(v_, i_ + n); }
    iter operator-(ptrdiff_t n) const { return iter(v_, i_ - n); }
    template <class V2, class R2> ptrdiff_t operator-(const iter<V2, R2>& o) const { return (ptrdiff_t)i_ - (ptrdiff_t)o.index(); }
    template <class V2, class R2> bool operator==(const iter<V2, R2>& o) const { return i_ == o.index(); }
    template <class V2, class R2> bool operator!=(const iter<V2, R2>& o) const { return i_ != o.index(); }
    template <class V2, class R2> bool operator<(const iter<V2, R2>& o) const { return i_ < o.index(); }
    template <class V2, class R2> bool operator>(const iter<V2, R2>& o) const { return i_ > o.index(); }
    template <class V2, class R2> bool operator<=(const iter<V2, R2>& o) const { return i_ <= o.index(); }
    template <class V2, class R2> bool operator>=(const iter<V2, R2>& o) const { return i_ >= o.index(); }
  };
 public:
  typedef iter<concurrent_vector, T> iterator;
  typedef iter<const concurrent_vector, const T> const_iterator;
  typedef T value_type;
  typedef size_t size_type;
  concurrent_vector() {}
  explicit concurrent_vector(size_t n) : d_(n) {}
  iterator push_back(const T& v) { std::lock_guard<std::mutex> g(m_); d_.push_back(v); return iterator(this, d_.size() - 1); }
  template <class... A> iterator emplace_back(A&&... a) { std::lock_guard<std::mutex> g(m_); d_.emplace_back(std::forward<A>(a)...); return iterator(this, d_.size() - 1); }
  iterator grow_by(size_t n) { std::lock_guard<std::mutex> g(m_); size_t o = d_.size(); d_.resize(o + n); return iterator(this, o); }
  iterator grow_to_at_least(size_t n) { std::lock_guard<std::mutex> g(m_); size_t o = d_.size(); if (n > o) d_.resize(n); return iterator(this, o); }
  size_t size() const { std::lock_guard<std::mutex> g(m_); return d_.size(); }
  bool empty() const { return size() == 0; }
  void resize(size_t n) { std::lock_guard<std::mutex> g(m_); d_.resize(n); }
  void reserve(size_t) {}
  void clear() { std::lock_guard<std::mutex> g(m_); d_.clear(); }
  void shrink_to_fit() {}
  // element access takes the lock too: std::deque's map of blocks may be reallocated by a concurrent push_back
  T& operator[](size_t i) { std::lock_guard<std::mutex> g(m_); return d_[i]; }
  const T& operator[](size_t i) const { std::lock_guard<std::mutex> g(m_); return d_[i]; }
  T& at(size_t i) { std::lock_guard<std::mutex> g(m_); return d_.at(i); }
  const T& at(size_t i) const { std::lock_guard<std::mutex> g(m_); return d_.at(i); }
  T& back() { std::lock_guard<std::mutex> g(m_); return d_.back(); }
  T& front() { std::lock_guard<std::mutex> g(m_); return d_.front(); }
  iterator begin() { return iterator(this, 0); }
  iterator end() { return iterator(this, size()); }
  const_iterator begin() const { return const_iterator(this, 0); }
  const_iterator end() const { return const_iterator(this, size()); }
  const_iterator cbegin() const { return const_iterator(this, 0); }
  const_iterator cend() const { return const_iterator(this, size()); }
};
}  // namespace tbb
