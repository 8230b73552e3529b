// stand-in for the slice of boost::program_options the reference CLIs use:
//   options_description(caption).add_options()("long,s", value<T>()->required()/default_value(v), "help") ...
//   parse_command_line / store / notify / variables_map[name].as<T>() / count(name)
// Naming rule reproduced: an option declared as ",k" (no long name) is stored under "-k".
#pragma once
#include <iostream>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>
namespace boost { namespace program_options {
class error : public std::logic_error { public: using std::logic_error::logic_error; };
class required_option : public error { public: explicit required_option(const std::string& n) : error("the option '" + n + "' is required but missing") {} };
class unknown_option : public error { public: explicit unknown_option(const std::string& n) : error("unrecognised option '" + n + "'") {} };

struct value_base {
  bool required_ = false, has_default_ = false, is_switch_ = false;
  std::string default_text_;
  virtual ~value_base() {}
  virtual std::shared_ptr<void> parse(const std::string& s) const = 0;
  virtual std::shared_ptr<void> default_value_any() const = 0;
};
template <class T> struct from_text { static T get(const std::string& s) { std::istringstream is(s); T v; is >> v; if (is.fail()) throw error("bad option value '" + s + "'"); return v; } };
template <> struct from_text<std::string> { static std::string get(const std::string& s) { return s; } };
template <> struct from_text<char> { static char get(const std::string& s) { if (s.size() != 1) throw error("bad option value '" + s + "'"); return s[0]; } };
template <class T>
struct typed_value : value_base {
  T default_{};
  typed_value* required() { required_ = true; return this; }
  typed_value* default_value(const T& v) { default_ = v; has_default_ = true; std::ostringstream os; os << v; default_text_ = os.str(); return this; }
  typed_value* default_value(const T& v, const std::string& text) { default_ = v; has_default_ = true; default_text_ = text; return this; }
  std::shared_ptr<void> parse(const std::string& s) const override { return std::make_shared<T>(from_text<T>::get(s)); }
  std::shared_ptr<void> default_value_any() const override { return std::make_shared<T>(default_); }
};
template <class T> typed_value<T>* value() { return new typed_value<T>(); }
inline typed_value<bool>* bool_switch() { auto* v = new typed_value<bool>(); v->is_switch_ = true; v->default_value(false); return v; }

struct option_description {
  std::string long_name, key; char short_name = 0; std::shared_ptr<value_base> sem; std::string help;
};
class options_description;
class options_easy_init {
  options_description* owner_;
 public:
  explicit options_easy_init(options_description* o) : owner_(o) {}
  options_easy_init& operator()(const char* name, const char* help);
  options_easy_init& operator()(const char* name, value_base* sem, const char* help = "");
};
class options_description {
 public:
  std::string caption;
  std::vector<option_description> opts;
  explicit options_description(const std::string& c = "") : caption(c) {}
  options_easy_init add_options() { return options_easy_init(this); }
  void add(const char* name, value_base* sem, const char* help) {
    option_description d;
    std::string n(name);
    size_t c = n.find(',');
    if (c == std::string::npos) d.long_name = n; else { d.long_name = n.substr(0, c); if (c + 1 < n.size()) d.short_name = n[c + 1]; }
    d.key = d.long_name.empty() ? std::string("-") + d.short_name : d.long_name;
    d.sem.reset(sem); d.help = help ? help : "";
    opts.push_back(d);
  }
  const option_description* find_long(const std::string& n) const { for (auto& o : opts) if (!o.long_name.empty() && o.long_name == n) return &o; return nullptr; }
  const option_description* find_short(char c) const { for (auto& o : opts) if (o.short_name == c) return &o; return nullptr; }
};
inline options_easy_init& options_easy_init::operator()(const char* name, const char* help) {
  auto* v = new typed_value<bool>(); v->is_switch_ = true; owner_->add(name, v, help); return *this;
}
inline options_easy_init& options_easy_init::operator()(const char* name, value_base* sem, const char* help) { owner_->add(name, sem, help); return *this; }
inline std::ostream& operator<<(std::ostream& os, const options_description& d) {
  os << d.caption << "\n";
  for (auto& o : d.opts) {
    std::string left = "  ";
    if (o.short_name) { left += "-"; left += o.short_name; if (!o.long_name.empty()) left += " [ --" + o.long_name + " ]"; }
    else left += "--" + o.long_name;
    if (!o.sem->is_switch_) { left += " arg"; if (o.sem->has_default_) left += " (=" + o.sem->default_text_ + ")"; }
    os << left; if (left.size() < 38) os << std::string(38 - left.size(), ' '); else os << " ";
    os << o.help << "\n";
  }
  return os;
}
class variable_value {
  std::shared_ptr<void> v_; bool defaulted_ = false;
 public:
  variable_value() {}
  variable_value(std::shared_ptr<void> v, bool d) : v_(v), defaulted_(d) {}
  template <class T> const T& as() const { if (!v_) throw error("option has no value"); return *static_cast<const T*>(v_.get()); }
  bool empty() const { return !v_; }
  bool defaulted() const { return defaulted_; }
};
class variables_map : public std::map<std::string, variable_value> {
 public:
  std::vector<std::string> required_missing;
  const options_description* desc = nullptr;
  size_t count(const std::string& k) const { return std::map<std::string, variable_value>::count(k); }
  const variable_value& operator[](const std::string& k) const {
    static const variable_value none; auto it = find(k); return it == end() ? none : it->second;
  }
  void put(const std::string& k, const variable_value& v) { std::map<std::string, variable_value>::operator[](k) = v; }
};
struct parsed_options { const options_description* desc; std::vector<std::pair<const option_description*, std::string>> found; };
inline parsed_options parse_command_line(int argc, const char* const* argv, const options_description& d) {
  parsed_options p; p.desc = &d;
  for (int i = 1; i < argc; i++) {
    std::string a(argv[i]);
    const option_description* o = nullptr; std::string val; bool have_val = false;
    if (a.size() > 2 && a[0] == '-' && a[1] == '-') {
      std::string n = a.substr(2); size_t eq = n.find('=');
      if (eq != std::string::npos) { val = n.substr(eq + 1); n = n.substr(0, eq); have_val = true; }
      o = d.find_long(n);
      if (!o) throw unknown_option(a);
    } else if (a.size() >= 2 && a[0] == '-') {
      o = d.find_short(a[1]);
      if (!o) throw unknown_option(a);
      if (a.size() > 2) { val = a.substr(2); have_val = true; }
    } else throw error("too many positional options have been specified on the command line");
    if (o->sem->is_switch_) { p.found.push_back({o, "1"}); continue; }
    if (!have_val) { if (i + 1 >= argc) throw error("the required argument for option '" + a + "' is missing"); val = argv[++i]; }
    p.found.push_back({o, val});
  }
  return p;
}
inline parsed_options parse_command_line(int argc, char* argv[], const options_description& d) { return parse_command_line(argc, (const char* const*)argv, d); }
inline void store(const parsed_options& p, variables_map& vm) {
  vm.desc = p.desc;
  for (auto& f : p.found) {
    if (vm.count(f.first->key) && !vm[f.first->key].defaulted()) continue;
    vm.put(f.first->key, variable_value(f.first->sem->is_switch_ ? std::make_shared<bool>(true) : f.first->sem->parse(f.second), false));
  }
  for (auto& o : p.desc->opts) {
    if (vm.count(o.key)) continue;
    if (o.sem->has_default_) vm.put(o.key, variable_value(o.sem->default_value_any(), true));     // bool_switch() defaults to false
  }
}
inline void notify(variables_map& vm) {
  if (!vm.desc) return;
  for (auto& o : vm.desc->opts)
    if (o.sem->required_ && !vm.count(o.key)) throw required_option(o.long_name.empty() ? std::string("-") + o.short_name : "--" + o.long_name);
}
} }  // namespace boost::program_options
