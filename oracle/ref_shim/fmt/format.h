// Minimal fmt stand-in — TEST INFRASTRUCTURE (oracle/_ref build only).  Supports "{}".
#pragma once
#include <cstdio>
#include <sstream>
#include <string>
#include <string_view>
namespace fmt {
namespace detail {
inline void emit(std::ostringstream& os, std::string_view& f)
{
  os << f;
  f = {};
}
template <class A, class... R> void emit(std::ostringstream& os, std::string_view& f, const A& a, const R&... r)
{
  const auto p = f.find("{}");
  if (p == std::string_view::npos) {
    os << f;
    f = {};
    return;
  }
  os << f.substr(0, p) << a;
  f.remove_prefix(p + 2);
  emit(os, f, r...);
}
} // namespace detail
template <class... A> std::string format(std::string_view f, const A&... a)
{
  std::ostringstream os;
  detail::emit(os, f, a...);
  return os.str();
}
template <class... A> void print(std::FILE* out, std::string_view f, const A&... a)
{
  const std::string s = format(f, a...);
  std::fwrite(s.data(), 1, s.size(), out);
}
template <class... A> void print(std::string_view f, const A&... a) { print(stdout, f, a...); }
} // namespace fmt
