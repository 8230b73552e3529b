// stand-in for boost::tokenizer<boost::char_separator<char>> (drops empty tokens, like the default char_separator)
#pragma once
#include <string>
#include <vector>
namespace boost {
template <class Char = char>
class char_separator {
 public:
  std::basic_string<Char> dropped;
  explicit char_separator(const Char* d = " ") : dropped(d) {}
};
template <class Sep = char_separator<char>>
class tokenizer {
  std::vector<std::string> toks_;
 public:
  typedef std::vector<std::string>::const_iterator iterator;
  typedef iterator const_iterator;
  tokenizer(const std::string& s, const Sep& sep = Sep()) {
    std::string cur;
    for (char c : s) {
      if (sep.dropped.find(c) != std::string::npos) { if (!cur.empty()) { toks_.push_back(cur); cur.clear(); } }
      else cur.push_back(c);
    }
    if (!cur.empty()) toks_.push_back(cur);
  }
  iterator begin() const { return toks_.begin(); }
  iterator end() const { return toks_.end(); }
};
}  // namespace boost
