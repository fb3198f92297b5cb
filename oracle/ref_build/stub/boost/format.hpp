// Minimal stand-in for <boost/format.hpp>, used ONLY to compile the reference's
// logger (reference src/logger.h:26,78) when building oracle/_ref. Boost is not
// installed in this image. Test infrastructure, not product code.
#pragma once
#include <string>
#include <sstream>
#include <ostream>
namespace boost {
class wformat {
 public:
  wformat() {}
  explicit wformat(const wchar_t* s) : s_(s ? s : L"") {}
  template <typename T> wformat& operator%(const T& v) {
    std::wostringstream o; o << L" [" << v << L"]"; args_ += o.str(); return *this;
  }
  std::wstring str() const { return s_ + args_; }
 private:
  std::wstring s_, args_;
};
inline std::wostream& operator<<(std::wostream& o, const wformat& f) { return o << f.str(); }
}  // namespace boost
