// Stand-in for <boost/format.hpp>: printf-style formatting of ONE integer argument, which is all the reference's extractors
// use it for (MulranPointCloudSelect.cpp:98-99, :245-246: "%010ld" / "%06d" file names).  See ../README.md.
#pragma once
#include <cstdio>
#include <string>

namespace boost {
class format {
  std::string spec_, out_;
 public:
  explicit format(const char* spec) : spec_(spec) {}
  explicit format(const std::string& spec) : spec_(spec) {}
  template <class T> format& operator%(const T& v) {
    char buf[128];
    std::snprintf(buf, sizeof buf, spec_.c_str(), v);
    out_ = buf;
    return *this;
  }
  std::string str() const { return out_.empty() ? spec_ : out_; }
};
}  // namespace boost
