// cli_options.hpp -- the slice of cxxopts behaviour raytrace's main() relies on
// (source.cpp:197-218), re-implemented because third-party sources are not copied here:
//   short/long options, "--opt=value" and "--opt value", positional capture of the first free
//   argument as --output, repeated -l collected with the LAST value winning, and an unknown
//   option or a second positional raising an exception nobody catches (the reference aborts
//   with exit 134 on `raytrace --bogus`; SURVEY.md probe C.5).
#pragma once
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace uecrt_host {

struct option_not_exists_exception : std::runtime_error {
  explicit option_not_exists_exception(const std::string& o) : std::runtime_error("Option '" + o + "' does not exist") {}
};
struct missing_argument_exception : std::runtime_error {
  explicit missing_argument_exception(const std::string& o) : std::runtime_error("Option '" + o + "' is missing an argument") {}
};

class options {
 public:
  options(std::string program, std::string description) : program_(std::move(program)), description_(std::move(description)) {}
  options& add(char short_name, const std::string& long_name, const std::string& help, bool takes_value) {
    specs_.push_back({short_name, long_name, help, takes_value});
    return *this;
  }
  void positional(const std::string& long_name) { positional_ = long_name; }

  void parse(int argc, char** argv) {
    bool positional_used = false;
    for (int i = 1; i < argc; ++i) {
      const std::string a = argv[i];
      const spec* sp = nullptr;
      std::string value;
      bool has_value = false;
      if (a.rfind("--", 0) == 0 && a.size() > 2) {
        const auto eq = a.find('=');
        const std::string name = a.substr(2, eq == std::string::npos ? std::string::npos : eq - 2);
        sp = find_long(name);
        if (!sp) throw option_not_exists_exception(name);
        if (eq != std::string::npos) { value = a.substr(eq + 1); has_value = true; }
      } else if (a.size() >= 2 && a[0] == '-' && a != "--") {
        sp = find_short(a[1]);
        if (!sp) throw option_not_exists_exception(std::string(1, a[1]));
        if (a.size() > 2) { value = a.substr(2); has_value = true; }
      } else {
        if (positional_.empty() || positional_used) throw option_not_exists_exception("positional");
        positional_used = true;
        values_[positional_].push_back(a);
        continue;
      }
      if (sp->takes_value) {
        if (!has_value) {
          if (i + 1 >= argc) throw missing_argument_exception(sp->long_name);
          value = argv[++i];
        }
        values_[sp->long_name].push_back(value);
      } else {
        values_[sp->long_name].push_back("true");
      }
    }
  }
  std::size_t count(const std::string& long_name) const {
    auto it = values_.find(long_name);
    return it == values_.end() ? 0 : it->second.size();
  }
  const std::string& last(const std::string& long_name) const { return values_.at(long_name).back(); }
  template <class T>
  T as(const std::string& long_name) const {
    std::istringstream is(last(long_name));
    T v{};
    is >> v;
    if (is.fail()) throw std::runtime_error("Argument '" + last(long_name) + "' failed to parse");
    return v;
  }
  std::string help() const {
    std::ostringstream os;
    os << description_ << "\nUsage:\n  " << program_ << " [OPTION...] positional parameters\n\n";
    std::size_t w = 0;
    std::vector<std::string> left;
    for (const auto& s : specs_) {
      std::string l = "  ";
      l += s.short_name ? std::string("-") + s.short_name + ", " : std::string("    ");
      l += "--" + s.long_name + (s.takes_value ? " arg" : "");
      left.push_back(l);
      w = std::max(w, l.size());
    }
    for (std::size_t i = 0; i < specs_.size(); ++i) os << left[i] << std::string(w - left[i].size() + 2, ' ') << specs_[i].help << "\n";
    return os.str();
  }

 private:
  struct spec { char short_name; std::string long_name, help; bool takes_value; };
  const spec* find_long(const std::string& n) const { for (auto& s : specs_) if (s.long_name == n) return &s; return nullptr; }
  const spec* find_short(char c) const { for (auto& s : specs_) if (s.short_name == c) return &s; return nullptr; }
  std::string program_, description_, positional_;
  std::vector<spec> specs_;
  std::map<std::string, std::vector<std::string>> values_;
};

}  // namespace uecrt_host
