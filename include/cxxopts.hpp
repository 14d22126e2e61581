/**
 * @file cxxopts.hpp
 * @brief The subset of the cxxopts command-line interface that Gunrock's example programs use
 * (`examples/algorithms/{tc,mst,bc,geo,hits,...}/*.cu`, `include/gunrock/io/parameters.hxx:35-120`),
 * written from scratch so that those translation units build with no third-party fetch (the reference
 * downloads jarro2783/cxxopts v3.0.0 at configure time, cmake/FetchCXXOpts.cmake).
 *
 * Supported: `Options(program, description)`, `add_options()("s,long", "help"[, value<T>()])...`,
 * `value<T>()` with `->default_value("text")`, `parse(argc, argv)` with `--long value`, `--long=value`,
 * `-s value`, boolean switches (`--flag`, `--flag=true|false`), `result.count("long")`,
 * `result["long"].as<T>()` for bool / int / float / double / std::string, and `help()`.
 * Unknown options and missing arguments raise `cxxopts::OptionException` like the original.
 */
#pragma once

#include <cstdlib>
#include <initializer_list>
#include <iostream>  // callers print `options.help()` with std::cout and include nothing else (examples/tools/cmd.cu)
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <type_traits>
#include <string>
#include <vector>

namespace cxxopts {

class OptionException : public std::runtime_error {
 public:
  explicit OptionException(const std::string& m) : std::runtime_error(m) {}
};

class Value {
 public:
  bool is_bool = false;
  bool has_default = false;
  std::string default_text;
  std::shared_ptr<Value> default_value(const std::string& text) {
    has_default = true;
    default_text = text;
    return self.lock();
  }
  std::weak_ptr<Value> self;
};

template <typename T>
std::shared_ptr<Value> value() {
  auto v = std::make_shared<Value>();
  v->self = v;
  v->is_bool = std::is_same<T, bool>::value;
  return v;
}

namespace detail {
template <typename T>
inline T convert(const std::string& text) {
  std::istringstream in(text);
  T out{};
  in >> out;
  if (in.fail())
    throw OptionException("Argument '" + text + "' failed to parse");
  return out;
}
template <>
inline std::string convert<std::string>(const std::string& text) {
  return text;
}
template <>
inline bool convert<bool>(const std::string& text) {
  if (text == "true" || text == "True" || text == "1" || text == "t" || text == "T")
    return true;
  if (text == "false" || text == "False" || text == "0" || text == "f" || text == "F" || text.empty())
    return false;
  throw OptionException("Argument '" + text + "' failed to parse");
}
}  // namespace detail

class OptionValue {
 public:
  std::size_t occurrences = 0;
  bool has_text = false;  // given on the command line or by a default
  std::string text;
  std::string name;
  std::size_t count() const { return occurrences; }
  template <typename T>
  T as() const {
    if (!has_text)
      throw OptionException("Option '" + name + "' has no value");
    return detail::convert<T>(text);
  }
};

class ParseResult {
 public:
  std::map<std::string, OptionValue> values;  // by long name
  std::size_t count(const std::string& name) const {
    auto it = values.find(name);
    return it == values.end() ? 0 : it->second.occurrences;
  }
  const OptionValue& operator[](const std::string& name) const {
    auto it = values.find(name);
    if (it == values.end())
      throw OptionException("Option '" + name + "' does not exist");
    return it->second;
  }
};

class Options {
  struct option_t {
    std::string short_name, long_name, description;
    std::shared_ptr<Value> value;  // null: a plain switch
  };
  std::string program_, description_;
  std::vector<option_t> options_;

  const option_t* find_long(const std::string& n) const {
    for (auto& o : options_)
      if (o.long_name == n)
        return &o;
    return nullptr;
  }
  const option_t* find_short(const std::string& n) const {
    for (auto& o : options_)
      if (!o.short_name.empty() && o.short_name == n)
        return &o;
    return nullptr;
  }

 public:
  class OptionAdder {
    Options& owner_;

   public:
    explicit OptionAdder(Options& o) : owner_(o) {}
    OptionAdder& operator()(const std::string& names, const std::string& description,
                            std::shared_ptr<Value> v = nullptr, std::string /*arg_help*/ = "") {
      option_t o;
      auto comma = names.find(',');
      if (comma == std::string::npos) {
        o.long_name = names;
      } else {
        o.short_name = names.substr(0, comma);
        o.long_name = names.substr(comma + 1);
        if (o.short_name.size() > 1)  // "long,s" order is accepted by cxxopts as well
          std::swap(o.short_name, o.long_name);
      }
      o.description = description;
      o.value = v;
      owner_.options_.push_back(o);
      return *this;
    }
  };

  Options(std::string program, std::string description = "")
      : program_(std::move(program)), description_(std::move(description)) {}

  OptionAdder add_options(std::string /*group*/ = "") { return OptionAdder(*this); }

  ParseResult parse(int argc, const char* const* argv) const {
    ParseResult r;
    for (auto& o : options_) {
      OptionValue v;
      v.name = o.long_name;
      if (o.value && o.value->has_default) {
        v.has_text = true;
        v.text = o.value->default_text;
      } else if (!o.value || o.value->is_bool) {
        v.has_text = true;  // an absent switch reads as false
        v.text = "false";
      }
      r.values[o.long_name] = v;
    }
    for (int i = 1; i < argc; ++i) {
      std::string arg = argv[i];
      if (arg == "--")
        break;
      const option_t* opt = nullptr;
      std::string inline_value;
      bool has_inline = false;
      if (arg.rfind("--", 0) == 0) {
        std::string body = arg.substr(2);
        auto eq = body.find('=');
        if (eq != std::string::npos) {
          inline_value = body.substr(eq + 1);
          body = body.substr(0, eq);
          has_inline = true;
        }
        opt = find_long(body);
        if (!opt)
          throw OptionException("Option '" + body + "' does not exist");
      } else if (arg.size() >= 2 && arg[0] == '-') {
        std::string body = arg.substr(1, 1);
        opt = find_short(body);
        if (!opt)
          throw OptionException("Option '" + body + "' does not exist");
        if (arg.size() > 2) {
          inline_value = arg.substr(2);
          has_inline = true;
        }
      } else {
        continue;  // positional arguments are not used by the examples
      }
      OptionValue& v = r.values[opt->long_name];
      v.occurrences += 1;
      const bool is_switch = !opt->value || opt->value->is_bool;
      if (has_inline) {
        v.text = inline_value;
        v.has_text = true;
      } else if (is_switch) {
        v.text = "true";
        v.has_text = true;
      } else {
        if (i + 1 >= argc)
          throw OptionException("Option '" + opt->long_name + "' is missing an argument");
        v.text = argv[++i];
        v.has_text = true;
      }
    }
    return r;
  }
  ParseResult parse(int argc, char** argv) const {
    return parse(argc, const_cast<const char* const*>(argv));
  }

  std::string help(std::initializer_list<std::string> /*groups*/ = {}) const {
    std::ostringstream out;
    out << description_ << "\nUsage:\n  " << program_ << " [OPTION...]\n\n";
    for (auto& o : options_) {
      std::string left = "  ";
      left += o.short_name.empty() ? "    " : ("-" + o.short_name + ", ");
      left += "--" + o.long_name;
      if (o.value && !o.value->is_bool)
        left += " arg";
      if (left.size() < 30)
        left.resize(30, ' ');
      else
        left += "  ";
      out << left << o.description;
      if (o.value && o.value->has_default)
        out << " (default: " << o.value->default_text << ")";
      out << "\n";
    }
    return out.str();
  }
};

}  // namespace cxxopts
