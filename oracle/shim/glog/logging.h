// glog/logging.h — stand-in for google-glog so that the reference's translation units compile here
// unmodified (test infrastructure, oracle/).  INFO/WARNING text is dropped without being formatted
// (the per-step mu dump must be off when timing); ERROR/FATAL go to stderr.
#ifndef REKF_ORACLE_GLOG_SHIM_H
#define REKF_ORACLE_GLOG_SHIM_H
#include <cstdlib>
#include <iostream>

namespace glogshim
{
struct Null
{
  template <class T> Null &operator<<(const T &) { return *this; }
  Null &operator<<(std::ostream &(*)(std::ostream &)) { return *this; }
};
struct Err
{
  bool fatal;
  explicit Err(bool f) : fatal(f) {}
  ~Err()
  {
    std::cerr << std::endl;
    if (fatal) std::abort();
  }
  template <class T> Err &operator<<(const T &v)
  {
    std::cerr << v;
    return *this;
  }
};
} // namespace glogshim

#define GLOGSHIM_INFO for (; false;) ::glogshim::Null()
#define GLOGSHIM_WARNING for (; false;) ::glogshim::Null()
#define GLOGSHIM_ERROR ::glogshim::Err(false)
#define GLOGSHIM_FATAL ::glogshim::Err(true)
#define LOG(severity) GLOGSHIM_##severity
#define CHECK(cond) for (; !(cond);) ::glogshim::Err(true) << "Check failed: " #cond " "
#endif
