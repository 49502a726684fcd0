// toml.hpp -- a small TOML reader, enough for gr_raytracer scene files.
//
// The reference deserialises `RenderConfig` with serde + the `toml` crate
// (reference src/main.rs:74-77, src/configuration.rs).  Supported here: comments, bare / quoted /
// dotted keys, [tables], [[arrays of tables]], basic and literal strings, integers, floats
// (incl. exponents, inf, nan, underscores), booleans, arrays (multi-line) and inline tables.
// Not supported (unused by scene files): dates/times, multi-line strings.
#pragma once

#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace grrt_host {

struct TomlError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

struct TomlValue {
    enum Kind { NONE, STRING, INT, FLOAT, BOOL, ARRAY, TABLE } kind = NONE;
    std::string s;
    long long i = 0;
    double f = 0.0;
    bool b = false;
    std::vector<TomlValue> arr;
    std::map<std::string, TomlValue> tab;
    bool defined_inline = false;

    bool is_table() const { return kind == TABLE; }
    bool is_number() const { return kind == INT || kind == FLOAT; }
    double as_double() const {
        if (kind == FLOAT) return f;
        if (kind == INT) return (double)i;
        throw TomlError("expected a number");
    }
    const TomlValue* find(const std::string& key) const {
        if (kind != TABLE) return nullptr;
        auto it = tab.find(key);
        return it == tab.end() ? nullptr : &it->second;
    }
};

TomlValue parse_toml(const std::string& text);

}  // namespace grrt_host
