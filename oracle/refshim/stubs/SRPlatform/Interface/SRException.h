// TEST INFRASTRUCTURE stub for SRException/SRString/SRMessageBuilder: the numeric headers only need them to
// format one alignment error (SRSimd.h:88). Minimal std::string-based stand-ins.
#pragma once
#include <string>
#include <sstream>
#include <exception>
namespace SRPlat {
class SRString {
  std::string _s;
public:
  SRString() {}
  explicit SRString(const std::string& s) : _s(s) {}
  static SRString MakeUnowned(const char *p) { return SRString(std::string(p)); }
  static SRString MakeClone(const char *p) { return SRString(std::string(p)); }
  const std::string& ToStd() const { return _s; }
  size_t GetData(const char *&p) const { p = _s.c_str(); return _s.size(); }
};
class SRException : public std::exception {
  SRString _msg;
public:
  explicit SRException(SRString&& m) : _msg(std::move(m)) {}
  explicit SRException(const SRString& m) : _msg(m) {}
  const char *what() const noexcept override { return _msg.ToStd().c_str(); }
  SRString ToString() const { return _msg; }
};
class SRMessageBuilder {
  std::ostringstream _os;
public:
  SRMessageBuilder() {}
  template<typename T> explicit SRMessageBuilder(const T& v) { _os << v; }
  template<typename T> SRMessageBuilder& operator()(const T& v) { _os << v; return *this; }
  SRMessageBuilder& AppendChar(char c) { _os << c; return *this; }
  SRString GetOwnedSRString() { return SRString(_os.str()); }
};
} // namespace SRPlat
