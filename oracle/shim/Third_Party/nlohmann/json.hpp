// TEST INFRASTRUCTURE ONLY -- stand-in for nlohmann/json.hpp (un-vendored by the reference) so that shapes/shapes.cpp,
// which holds the bmap block-file reader next to the JSON one, compiles where it lies under /root/reference/src
// (oracle/Makefile.ref).  JSON block files are outside the path: parse() aborts, the other members only have to compile.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <istream>
#include <string>
#include <utility>
#include <vector>

namespace nlohmann
{
class json
{
  public:
    static json parse(std::istream&)
    {
        std::fprintf(stderr, "oracle/shim: JSON block files are outside the path and not supported by the stand-in\n");
        std::abort();
    }
    bool contains(std::string const&) const { return false; }
    json const& operator[](std::string const&) const { return *this; }
    template <typename T>
    T get() const
    {
        return T();
    }
    std::vector<std::pair<std::string, json>> items() { return {}; }
};
} // namespace nlohmann
