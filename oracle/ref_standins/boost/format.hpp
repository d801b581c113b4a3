// TEST INFRASTRUCTURE - stand-in for boost::format as lsq_registration_impl.hpp:151-155 uses it (the LM debug table):
// printf-style directives filled left to right by operator%. See oracle/ref_standins/Eigen/Core.
#pragma once
#include <cstdio>
#include <ostream>
#include <string>
#include <type_traits>

namespace boost {

// 0 = honour the directive's own precision (boost's behaviour); > 0 = print floating-point arguments with that many
// significant digits instead, so that a test can read the reference's own LM debug table back at full precision
inline int& format_stand_in_precision() { static int p = 0; return p; }

class format {
  std::string fmt_, out_;
  size_t pos_ = 0;
  void copy_literal() {   // copy up to the next directive
    while (pos_ < fmt_.size()) {
      if (fmt_[pos_] == '%') {
        if (pos_ + 1 < fmt_.size() && fmt_[pos_ + 1] == '%') { out_ += '%'; pos_ += 2; continue; }
        return;
      }
      out_ += fmt_[pos_++];
    }
  }
  std::string next_directive() {
    copy_literal();
    size_t e = pos_;
    if (e >= fmt_.size()) return std::string();
    e++;
    while (e < fmt_.size() && std::string("diouxXeEfgGcs").find(fmt_[e]) == std::string::npos) e++;
    std::string d = fmt_.substr(pos_, e + 1 - pos_);
    pos_ = e + 1;
    return d;
  }
  template <typename... A> void emit(const std::string& d, A... a) {
    char buf[256];
    std::snprintf(buf, sizeof buf, d.c_str(), a...);
    out_ += buf;
  }

public:
  explicit format(const char* f) : fmt_(f) {}
  template <typename V> format& operator%(const V& v) {
    std::string d = next_directive();
    if (d.empty()) return *this;
    const char conv = d.back();
    if (std::is_same<V, char>::value && (conv == 'c' || conv == 's')) { d.back() = 'c'; emit(d, (int)v); }
    else if (std::is_floating_point<V>::value) {
      if (format_stand_in_precision() > 0) { char full[32]; std::snprintf(full, sizeof full, "%%%d.%dg", format_stand_in_precision() + 8, format_stand_in_precision()); emit(full, (double)v); }
      else { if (std::string("eEfgG").find(conv) == std::string::npos) d.back() = 'g'; emit(d, (double)v); }
    }
    else { d.back() = 'd'; emit(d, (int)v); }
    return *this;
  }
  format& operator%(const char* v) { std::string d = next_directive(); if (!d.empty()) { d.back() = 's'; emit(d, v); } return *this; }
  std::string str() { copy_literal(); return out_; }
  friend std::ostream& operator<<(std::ostream& os, format f) { return os << f.str(); }
};

}  // namespace boost
