// toml.cpp -- see toml.hpp
#include "toml.hpp"

#include <cctype>
#include <cmath>
#include <cstdlib>

namespace grrt_host {
namespace {

struct Parser {
    const std::string& t;
    size_t p = 0;
    int line = 1;
    explicit Parser(const std::string& text) : t(text) {}

    [[noreturn]] void fail(const std::string& msg) const {
        throw TomlError("TOML parse error at line " + std::to_string(line) + ": " + msg);
    }
    bool eof() const { return p >= t.size(); }
    char peek() const { return eof() ? '\0' : t[p]; }
    char get() {
        char c = peek();
        ++p;
        if (c == '\n') ++line;
        return c;
    }
    void skip_ws() {
        while (!eof() && (peek() == ' ' || peek() == '\t')) ++p;
    }
    void skip_comment() {
        if (peek() == '#')
            while (!eof() && peek() != '\n') ++p;
    }
    void skip_ws_nl() {  // whitespace, newlines and comments (inside arrays)
        for (;;) {
            skip_ws();
            if (peek() == '#') skip_comment();
            else if (peek() == '\n' || peek() == '\r') get();
            else break;
        }
    }
    void end_of_line() {
        skip_ws();
        skip_comment();
        if (eof()) return;
        if (peek() == '\r') get();
        if (peek() == '\n') {
            get();
            return;
        }
        fail("unexpected characters after value");
    }

    std::string parse_basic_string() {
        std::string out;
        get();  // opening quote
        for (;;) {
            if (eof() || peek() == '\n') fail("unterminated string");
            char c = get();
            if (c == '"') break;
            if (c == '\\') {
                char e = get();
                switch (e) {
                    case 'n': out += '\n'; break;
                    case 't': out += '\t'; break;
                    case 'r': out += '\r'; break;
                    case '"': out += '"'; break;
                    case '\\': out += '\\'; break;
                    case 'b': out += '\b'; break;
                    case 'f': out += '\f'; break;
                    case 'u':
                    case 'U': {
                        int n = e == 'u' ? 4 : 8;
                        unsigned long cp = 0;
                        for (int k = 0; k < n; ++k) {
                            char h = get();
                            if (!std::isxdigit((unsigned char)h)) fail("bad unicode escape");
                            cp = cp * 16 + (unsigned long)(std::isdigit((unsigned char)h) ? h - '0' : (std::tolower(h) - 'a' + 10));
                        }
                        if (cp < 0x80) out += (char)cp;
                        else if (cp < 0x800) { out += (char)(0xC0 | (cp >> 6)); out += (char)(0x80 | (cp & 0x3F)); }
                        else if (cp < 0x10000) { out += (char)(0xE0 | (cp >> 12)); out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F)); }
                        else { out += (char)(0xF0 | (cp >> 18)); out += (char)(0x80 | ((cp >> 12) & 0x3F)); out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F)); }
                        break;
                    }
                    default: fail("unknown escape sequence");
                }
            } else {
                out += c;
            }
        }
        return out;
    }
    std::string parse_literal_string() {
        std::string out;
        get();
        for (;;) {
            if (eof() || peek() == '\n') fail("unterminated literal string");
            char c = get();
            if (c == '\'') break;
            out += c;
        }
        return out;
    }
    std::string parse_key_part() {
        skip_ws();
        if (peek() == '"') return parse_basic_string();
        if (peek() == '\'') return parse_literal_string();
        std::string out;
        while (!eof() && (std::isalnum((unsigned char)peek()) || peek() == '_' || peek() == '-')) out += get();
        if (out.empty()) fail("expected a key");
        return out;
    }
    std::vector<std::string> parse_dotted_key() {
        std::vector<std::string> parts;
        for (;;) {
            parts.push_back(parse_key_part());
            skip_ws();
            if (peek() == '.') {
                get();
                continue;
            }
            break;
        }
        return parts;
    }

    TomlValue parse_number_or_bool() {
        size_t start = p;
        while (!eof() && (std::isalnum((unsigned char)peek()) || peek() == '+' || peek() == '-' || peek() == '.' || peek() == '_')) ++p;
        std::string tok = t.substr(start, p - start);
        if (tok.empty()) fail("expected a value");
        TomlValue v;
        if (tok == "true" || tok == "false") {
            v.kind = TomlValue::BOOL;
            v.b = tok == "true";
            return v;
        }
        std::string clean;
        for (char c : tok)
            if (c != '_') clean += c;
        std::string body = clean;
        double sign = 1.0;
        if (!body.empty() && (body[0] == '+' || body[0] == '-')) {
            if (body[0] == '-') sign = -1.0;
            body = body.substr(1);
        }
        if (body == "inf") { v.kind = TomlValue::FLOAT; v.f = sign * INFINITY; return v; }
        if (body == "nan") { v.kind = TomlValue::FLOAT; v.f = NAN; return v; }
        bool is_float = false, hex = body.size() > 2 && body[0] == '0' && (body[1] == 'x' || body[1] == 'o' || body[1] == 'b');
        if (!hex)
            for (char c : body)
                if (c == '.' || c == 'e' || c == 'E') is_float = true;
        char* end = nullptr;
        if (is_float) {
            v.kind = TomlValue::FLOAT;
            v.f = std::strtod(clean.c_str(), &end);  // correctly rounded, as Rust's f64::from_str
        } else {
            v.kind = TomlValue::INT;
            int base = 10;
            std::string digits = clean;
            if (hex) {
                base = body[1] == 'x' ? 16 : (body[1] == 'o' ? 8 : 2);
                digits = body.substr(2);
            }
            v.i = std::strtoll(digits.c_str(), &end, base);
        }
        if (!end || *end != '\0') fail("malformed number '" + tok + "'");
        return v;
    }

    TomlValue parse_value() {
        skip_ws();
        TomlValue v;
        char c = peek();
        if (c == '"') {
            if (t.compare(p, 3, "\"\"\"") == 0) fail("multi-line strings are not supported");
            v.kind = TomlValue::STRING;
            v.s = parse_basic_string();
            return v;
        }
        if (c == '\'') {
            v.kind = TomlValue::STRING;
            v.s = parse_literal_string();
            return v;
        }
        if (c == '[') {
            get();
            v.kind = TomlValue::ARRAY;
            for (;;) {
                skip_ws_nl();
                if (peek() == ']') {
                    get();
                    break;
                }
                v.arr.push_back(parse_value());
                skip_ws_nl();
                if (peek() == ',') {
                    get();
                    continue;
                }
                if (peek() == ']') {
                    get();
                    break;
                }
                fail("expected ',' or ']' in array");
            }
            return v;
        }
        if (c == '{') {
            get();
            v.kind = TomlValue::TABLE;
            v.defined_inline = true;
            skip_ws();
            if (peek() == '}') {
                get();
                return v;
            }
            for (;;) {
                std::vector<std::string> key = parse_dotted_key();
                skip_ws();
                if (get() != '=') fail("expected '=' in inline table");
                TomlValue val = parse_value();
                insert(v, key, std::move(val));
                skip_ws();
                if (peek() == ',') {
                    get();
                    skip_ws();
                    continue;
                }
                if (peek() == '}') {
                    get();
                    break;
                }
                fail("expected ',' or '}' in inline table");
            }
            return v;
        }
        return parse_number_or_bool();
    }

    // descend (creating tables) along all but the last key part; arrays of tables resolve to their last element
    TomlValue* descend(TomlValue* cur, const std::vector<std::string>& path, size_t n) {
        for (size_t k = 0; k < n; ++k) {
            TomlValue& next = cur->tab[path[k]];
            if (next.kind == TomlValue::NONE) next.kind = TomlValue::TABLE;
            if (next.kind == TomlValue::ARRAY) {
                if (next.arr.empty() || !next.arr.back().is_table()) fail("key '" + path[k] + "' is not a table");
                cur = &next.arr.back();
            } else if (next.kind == TomlValue::TABLE) {
                cur = &next;
            } else {
                fail("key '" + path[k] + "' is not a table");
            }
        }
        return cur;
    }
    void insert(TomlValue& table, const std::vector<std::string>& key, TomlValue&& val) {
        TomlValue* cur = descend(&table, key, key.size() - 1);
        if (cur->tab.count(key.back())) fail("duplicate key '" + key.back() + "'");
        cur->tab[key.back()] = std::move(val);
    }

    TomlValue parse_document() {
        TomlValue root;
        root.kind = TomlValue::TABLE;
        TomlValue* current = &root;
        while (!eof()) {
            skip_ws();
            char c = peek();
            if (c == '\n' || c == '\r') {
                get();
                continue;
            }
            if (c == '#') {
                skip_comment();
                continue;
            }
            if (eof()) break;
            if (c == '[') {
                get();
                bool array_table = peek() == '[';
                if (array_table) get();
                std::vector<std::string> path = parse_dotted_key();
                skip_ws();
                if (get() != ']') fail("expected ']'");
                if (array_table && get() != ']') fail("expected ']]'");
                end_of_line();
                TomlValue* parent = descend(&root, path, path.size() - 1);
                TomlValue& slot = parent->tab[path.back()];
                if (array_table) {
                    if (slot.kind == TomlValue::NONE) slot.kind = TomlValue::ARRAY;
                    if (slot.kind != TomlValue::ARRAY) fail("'" + path.back() + "' is not an array of tables");
                    TomlValue t2;
                    t2.kind = TomlValue::TABLE;
                    slot.arr.push_back(std::move(t2));
                    current = &slot.arr.back();
                } else {
                    if (slot.kind == TomlValue::NONE) slot.kind = TomlValue::TABLE;
                    if (slot.kind == TomlValue::ARRAY && !slot.arr.empty()) current = &slot.arr.back();
                    else if (slot.kind == TomlValue::TABLE) current = &slot;
                    else fail("'" + path.back() + "' is not a table");
                }
                continue;
            }
            std::vector<std::string> key = parse_dotted_key();
            skip_ws();
            if (get() != '=') fail("expected '=' after key");
            TomlValue val = parse_value();
            insert(*current, key, std::move(val));
            end_of_line();
        }
        return root;
    }
};

}  // namespace

TomlValue parse_toml(const std::string& text) {
    Parser p(text);
    return p.parse_document();
}

}  // namespace grrt_host
