// stand-in for tbb::concurrent_unordered_map
#pragma once
#include <mutex>
#include <unordered_map>
namespace tbb {
#ifndef SHIM_TBB_HASH
#define SHIM_TBB_HASH
template <class Key> struct tbb_hash { size_t operator()(const Key& k) const { return std::hash<Key>()(k); } };
#endif
template <class Key, class T, class Hash = tbb_hash<Key>, class Eq = std::equal_to<Key>>
class concurrent_unordered_map {
  std::unordered_map<Key, T, Hash, Eq> s_; mutable std::mutex m_;
 public:
  typedef typename std::unordered_map<Key, T, Hash, Eq>::iterator iterator;
  typedef typename std::unordered_map<Key, T, Hash, Eq>::const_iterator const_iterator;
  typedef std::pair<const Key, T> value_type;
  std::pair<iterator, bool> insert(const value_type& v) { std::lock_guard<std::mutex> g(m_); return s_.insert(v); }
  iterator find(const Key& k) { std::lock_guard<std::mutex> g(m_); return s_.find(k); }
  const_iterator find(const Key& k) const { std::lock_guard<std::mutex> g(m_); return s_.find(k); }
  T& operator[](const Key& k) { std::lock_guard<std::mutex> g(m_); return s_[k]; }
  size_t count(const Key& k) const { std::lock_guard<std::mutex> g(m_); return s_.count(k); }
  size_t size() const { std::lock_guard<std::mutex> g(m_); return s_.size(); }
  iterator begin() { return s_.begin(); }
  iterator end() { return s_.end(); }
  const_iterator begin() const { return s_.begin(); }
  const_iterator end() const { return s_.end(); }
  void clear() { std::lock_guard<std::mutex> g(m_); s_.clear(); }
};
}  // namespace tbb
