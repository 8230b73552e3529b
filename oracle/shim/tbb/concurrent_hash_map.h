// stand-in for tbb::concurrent_hash_map<Key, T, HashCompare> with accessor / const_accessor (one global lock; an
// accessor holds it for its lifetime, which is what gives the reference its per-key critical sections)
#pragma once
#include <functional>
#include <mutex>
#include <unordered_map>
#include <utility>
namespace tbb {
template <class Key> struct tbb_hash_compare {
  static size_t hash(const Key& k) { return std::hash<Key>()(k); }
  static bool equal(const Key& a, const Key& b) { return a == b; }
};
template <class Key, class T, class HashCompare = tbb_hash_compare<Key>>
class concurrent_hash_map {
  struct H { HashCompare hc; size_t operator()(const Key& k) const { return hc.hash(k); } };
  struct E { HashCompare hc; bool operator()(const Key& a, const Key& b) const { return hc.equal(a, b); } };
  typedef std::unordered_map<Key, T, H, E> map_t;
  map_t m_; mutable std::recursive_mutex mu_;
 public:
  typedef std::pair<const Key, T> value_type;
  typedef typename map_t::iterator iterator;
  typedef typename map_t::const_iterator const_iterator;
  class const_accessor {
    friend class concurrent_hash_map;
   protected:
    std::unique_lock<std::recursive_mutex> lk_; value_type* p_ = nullptr;
   public:
    const_accessor() {}
    ~const_accessor() { release(); }
    void release() { if (lk_.owns_lock()) lk_.unlock(); p_ = nullptr; }
    bool empty() const { return p_ == nullptr; }
    const value_type& operator*() const { return *p_; }
    const value_type* operator->() const { return p_; }
  };
  class accessor : public const_accessor {
   public:
    value_type& operator*() const { return *this->p_; }
    value_type* operator->() const { return this->p_; }
  };
  concurrent_hash_map() {}
  explicit concurrent_hash_map(size_t n) { m_.reserve(n); }
  bool find(const_accessor& a, const Key& k) const {
    a.release();
    a.lk_ = std::unique_lock<std::recursive_mutex>(mu_);
    auto it = const_cast<map_t&>(m_).find(k);
    if (it == m_.end()) { a.lk_.unlock(); return false; }
    a.p_ = &*it; return true;
  }
  bool find(accessor& a, const Key& k) { return static_cast<const concurrent_hash_map*>(this)->find(static_cast<const_accessor&>(a), k); }
  bool insert(const_accessor& a, const Key& k) {
    a.release();
    a.lk_ = std::unique_lock<std::recursive_mutex>(mu_);
    auto r = m_.emplace(k, T());
    a.p_ = &*r.first; return r.second;
  }
  bool insert(accessor& a, const Key& k) { return insert(static_cast<const_accessor&>(a), k); }
  bool insert(const_accessor& a, const value_type& v) {
    a.release();
    a.lk_ = std::unique_lock<std::recursive_mutex>(mu_);
    auto r = m_.insert(v);
    a.p_ = &*r.first; return r.second;
  }
  bool insert(accessor& a, const value_type& v) { return insert(static_cast<const_accessor&>(a), v); }
  bool insert(const value_type& v) { std::lock_guard<std::recursive_mutex> g(mu_); return m_.insert(v).second; }
  bool erase(const Key& k) { std::lock_guard<std::recursive_mutex> g(mu_); return m_.erase(k) != 0; }
  bool erase(const_accessor& a) { if (a.empty()) return false; Key k = a.p_->first; a.p_ = nullptr; m_.erase(k); a.lk_.unlock(); return true; }
  size_t count(const Key& k) const { std::lock_guard<std::recursive_mutex> g(mu_); return m_.count(k); }
  size_t size() const { std::lock_guard<std::recursive_mutex> g(mu_); return m_.size(); }
  bool empty() const { return size() == 0; }
  void clear() { std::lock_guard<std::recursive_mutex> g(mu_); m_.clear(); }
  void rehash(size_t n = 0) { std::lock_guard<std::recursive_mutex> g(mu_); m_.rehash(n); }
  iterator begin() { return m_.begin(); }
  iterator end() { return m_.end(); }
  const_iterator begin() const { return m_.begin(); }
  const_iterator end() const { return m_.end(); }
};
}  // namespace tbb
