// TEST INFRASTRUCTURE ONLY -- stand-in for nlohmann/json.hpp (un-vendored by the reference) so that shapes/shapes.cpp,
// which holds the JSON block-file reader next to the bmap one, compiles where it lies under /root/reference/src
// (oracle/Makefile.ref) and reads JSON block files.  Written from the library's published behaviour, only as far as
// shapes.cpp:229-395 uses it: parse(istream), contains, operator[], get<T>() for strings, booleans, numbers and (nested)
// arrays of numbers, items() over an object in key order (the library's default object type is std::map).  get<T>() is
// as strict as the library's: a string only from a string, a bool only from a boolean, a number from a number or a
// boolean; anything else throws (type_error there, std::runtime_error here -- shapes.cpp catches std::exception).
#pragma once
#include <cstdlib>
#include <istream>
#include <iterator>
#include <map>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

namespace nlohmann
{
class json
{
  public:
    enum kind_t
    {
        null_k,
        bool_k,
        int_k,
        float_k,
        string_k,
        array_k,
        object_k
    };

    static json parse(std::istream& in)
    {
        const std::string text((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
        size_t pos = 0;
        json v = value(text, pos);
        skip(text, pos);
        if (pos != text.size())
            fail("unexpected text after the document", pos);
        return v;
    }
    bool contains(std::string const& key) const { return kind == object_k && obj.count(key) != 0; }
    json const& operator[](std::string const& key) const
    {
        if (kind != object_k)
            throw std::runtime_error("cannot use operator[] with a string argument on a non-object");
        return obj.at(key);
    }
    std::map<std::string, json>& items()
    {
        if (kind != object_k)
            throw std::runtime_error("cannot use items() on a non-object");
        return obj;
    }
    template <typename T>
    T get() const
    {
        T out;
        from(*this, out);
        return out;
    }

  private:
    kind_t kind = null_k;
    bool b = false;
    long long i = 0;
    double f = 0.0;
    std::string s;
    std::vector<json> arr;
    std::map<std::string, json> obj;

    static void from(json const& j, std::string& out)
    {
        if (j.kind != string_k)
            throw std::runtime_error("type must be string");
        out = j.s;
    }
    static void from(json const& j, bool& out)
    {
        if (j.kind != bool_k)
            throw std::runtime_error("type must be boolean");
        out = j.b;
    }
    template <typename T, typename std::enable_if<std::is_arithmetic<T>::value && !std::is_same<T, bool>::value, int>::type = 0>
    static void from(json const& j, T& out)
    {
        if (j.kind == int_k)
            out = static_cast<T>(j.i);
        else if (j.kind == float_k)
            out = static_cast<T>(j.f);
        else if (j.kind == bool_k)
            out = static_cast<T>(j.b);
        else
            throw std::runtime_error("type must be number");
    }
    template <typename T>
    static void from(json const& j, std::vector<T>& out)
    {
        if (j.kind != array_k)
            throw std::runtime_error("type must be array");
        out.clear();
        for (json const& e : j.arr)
        {
            T v;
            from(e, v);
            out.push_back(v);
        }
    }

    [[noreturn]] static void fail(const char* what, size_t pos)
    {
        throw std::runtime_error(std::string("parse error at byte ") + std::to_string(pos) + ": " + what);
    }
    static void skip(std::string const& t, size_t& p)
    {
        while (p < t.size() && (t[p] == ' ' || t[p] == '\t' || t[p] == '\n' || t[p] == '\r')) ++p;
    }
    static std::string string_token(std::string const& t, size_t& p)
    {
        std::string out;
        ++p; /* opening quote */
        while (p < t.size() && t[p] != '"')
        {
            char c = t[p++];
            if (c == '\\')
            {
                if (p >= t.size())
                    fail("unterminated escape", p);
                const char e = t[p++];
                switch (e)
                {
                case 'n': c = '\n'; break;
                case 't': c = '\t'; break;
                case 'r': c = '\r'; break;
                case 'b': c = '\b'; break;
                case 'f': c = '\f'; break;
                case 'u': fail("\\u escapes are not supported by the stand-in", p);
                default: c = e; /* quote, backslash, slash */
                }
            }
            out.push_back(c);
        }
        if (p >= t.size())
            fail("unterminated string", p);
        ++p;
        return out;
    }
    static json value(std::string const& t, size_t& p)
    {
        skip(t, p);
        if (p >= t.size())
            fail("unexpected end of input", p);
        json v;
        const char c = t[p];
        if (c == '{')
        {
            v.kind = object_k;
            ++p;
            skip(t, p);
            if (p < t.size() && t[p] == '}')
            {
                ++p;
                return v;
            }
            for (;;)
            {
                skip(t, p);
                if (p >= t.size() || t[p] != '"')
                    fail("expected a key", p);
                const std::string key = string_token(t, p);
                skip(t, p);
                if (p >= t.size() || t[p] != ':')
                    fail("expected ':'", p);
                ++p;
                v.obj[key] = value(t, p); /* a repeated key keeps its last value */
                skip(t, p);
                if (p < t.size() && t[p] == ',')
                {
                    ++p;
                    continue;
                }
                if (p < t.size() && t[p] == '}')
                {
                    ++p;
                    return v;
                }
                fail("expected ',' or '}'", p);
            }
        }
        if (c == '[')
        {
            v.kind = array_k;
            ++p;
            skip(t, p);
            if (p < t.size() && t[p] == ']')
            {
                ++p;
                return v;
            }
            for (;;)
            {
                v.arr.push_back(value(t, p));
                skip(t, p);
                if (p < t.size() && t[p] == ',')
                {
                    ++p;
                    continue;
                }
                if (p < t.size() && t[p] == ']')
                {
                    ++p;
                    return v;
                }
                fail("expected ',' or ']'", p);
            }
        }
        if (c == '"')
        {
            v.kind = string_k;
            v.s = string_token(t, p);
            return v;
        }
        if (t.compare(p, 4, "true") == 0)
        {
            v.kind = bool_k;
            v.b = true;
            p += 4;
            return v;
        }
        if (t.compare(p, 5, "false") == 0)
        {
            v.kind = bool_k;
            p += 5;
            return v;
        }
        if (t.compare(p, 4, "null") == 0)
        {
            p += 4;
            return v;
        }
        /* number: an integer unless it has a fraction or an exponent (the library keeps the two apart) */
        size_t q = p;
        if (q < t.size() && t[q] == '-')
            ++q;
        const size_t digits = q;
        while (q < t.size() && t[q] >= '0' && t[q] <= '9') ++q;
        if (q == digits)
            fail("invalid literal", p);
        bool integral = true;
        if (q < t.size() && t[q] == '.')
        {
            integral = false;
            ++q;
            while (q < t.size() && t[q] >= '0' && t[q] <= '9') ++q;
        }
        if (q < t.size() && (t[q] == 'e' || t[q] == 'E'))
        {
            integral = false;
            ++q;
            if (q < t.size() && (t[q] == '+' || t[q] == '-'))
                ++q;
            while (q < t.size() && t[q] >= '0' && t[q] <= '9') ++q;
        }
        const std::string tok = t.substr(p, q - p);
        if (integral)
        {
            v.kind = int_k;
            v.i = std::strtoll(tok.c_str(), nullptr, 10);
        }
        else
        {
            v.kind = float_k;
            v.f = std::strtod(tok.c_str(), nullptr);
        }
        p = q;
        return v;
    }
};
} // namespace nlohmann
