// TEST INFRASTRUCTURE: the two boost string helpers VX3.cuh touches.
#pragma once
#include <algorithm>
#include <cctype>
#include <string>
namespace boost {
inline void to_upper(std::string &s) { std::transform(s.begin(), s.end(), s.begin(), [](unsigned char c) { return (char)std::toupper(c); }); }
inline void trim_right(std::string &s) {
    while (!s.empty() && std::isspace((unsigned char)s.back()))
        s.pop_back();
}
} // namespace boost
