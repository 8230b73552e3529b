// cli_options.h -- a small stand-in for boost::program_options (absent in this image) that accepts
// what the reference CLIs accept: "--name value", "--name=value", "-x value", "-xvalue", bool switches.
#pragma once
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

struct OptSpec {
  std::string long_name;   // "" if none
  char short_name;         // 0 if none
  bool has_value;
  bool required;
  std::string def;         // default ("" = none)
  bool has_default;
  std::string help;
};

class Options {
 public:
  std::vector<OptSpec> specs;
  std::map<std::string, std::string> values;   // key = canonical name (long name, or "-x")
  void add(const std::string& l, char s, bool has_value, bool required, const char* def, const std::string& help) {
    specs.push_back({l, s, has_value, required, def ? def : "", def != nullptr, help});
  }
  static std::string key(const OptSpec& o) { return o.long_name.empty() ? std::string("-") + o.short_name : o.long_name; }
  const OptSpec* find_long(const std::string& n) const { for (auto& o : specs) if (!o.long_name.empty() && o.long_name == n) return &o; return nullptr; }
  const OptSpec* find_short(char c) const { for (auto& o : specs) if (o.short_name == c) return &o; return nullptr; }
  void parse(int argc, char** argv) {
    for (int i = 1; i < argc; i++) {
      std::string a = argv[i];
      const OptSpec* o = nullptr; std::string val; bool have = false;
      if (a.size() > 2 && a[0] == '-' && a[1] == '-') {
        size_t eq = a.find('=');
        o = find_long(a.substr(2, eq == std::string::npos ? std::string::npos : eq - 2));
        if (eq != std::string::npos) { val = a.substr(eq + 1); have = true; }
      } else if (a.size() >= 2 && a[0] == '-') {
        o = find_short(a[1]);
        if (a.size() > 2) { val = a.substr(a[2] == '=' ? 3 : 2); have = true; }
      }
      if (!o) throw std::runtime_error("unrecognised option '" + a + "'");
      if (o->has_value) {
        if (!have) { if (i + 1 >= argc) throw std::runtime_error("the required argument for option '" + a + "' is missing"); val = argv[++i]; }
        values[key(*o)] = val;
      } else {
        values[key(*o)] = "1";
      }
    }
  }
  void notify() const {
    for (auto& o : specs)
      if (o.required && !values.count(key(o)))
        throw std::runtime_error("the option '" + (o.long_name.empty() ? std::string("-") + o.short_name : "--" + o.long_name) + "' is required but missing");
  }
  bool count(const std::string& k) const { return values.count(k) != 0; }
  std::string get(const std::string& k) const {
    auto it = values.find(k);
    if (it != values.end()) return it->second;
    for (auto& o : specs) if (key(o) == k && o.has_default) return o.def;
    return "";
  }
  std::string describe(const std::string& caption) const {
    std::string out = caption + ":\n";
    for (auto& o : specs) {
      std::string n = "  ";
      if (o.short_name) { n += "-"; n += o.short_name; }
      if (!o.long_name.empty()) { if (o.short_name) n += " [ --" + o.long_name + " ]"; else n += "--" + o.long_name; }
      if (o.has_value) { n += " arg"; if (o.has_default && !o.def.empty()) n += " (=" + o.def + ")"; }
      if (n.size() < 30) n.resize(30, ' '); else n += " ";
      out += n + o.help + "\n";
    }
    return out;
  }
};
