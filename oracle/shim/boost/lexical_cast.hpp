// Minimal stand-in for boost::lexical_cast so the UNMODIFIED reference sources
// compile in an image without Boost. Test infrastructure only (used by
// oracle/Makefile when building oracle/_ref from /root/reference/src).
// The reference uses exactly one instantiation: lexical_cast<bool>(std::string)
// (src/global_faldoi.cpp:1884), which accepts "0" / "1" and throws otherwise.
#pragma once
#include <sstream>
#include <stdexcept>
#include <string>
namespace boost {
struct bad_lexical_cast : std::runtime_error {
    bad_lexical_cast() : std::runtime_error("bad lexical cast") {}
};
template <class T>
T lexical_cast(const std::string &s) {
    std::istringstream in(s);
    T v;
    if (!(in >> v) || in.peek() != std::istringstream::traits_type::eof())
        throw bad_lexical_cast();
    return v;
}
}  // namespace boost
