// json.hpp — a small recursive-descent JSON reader, enough for glTF 2.0 files.
#pragma once

#include <cstdlib>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace ptb {

struct Json {
    enum Kind { Null, Bool, Number, String, Array, Object } kind = Null;
    bool b = false;
    double num = 0;
    std::string str;
    std::vector<Json> arr;
    std::vector<std::pair<std::string, Json>> obj;

    const Json* find(const std::string& key) const {
        if (kind != Object) return nullptr;
        for (const auto& kv : obj)
            if (kv.first == key) return &kv.second;
        return nullptr;
    }
    bool has(const std::string& key) const { return find(key) != nullptr; }
    const Json& at(const std::string& key) const {
        const Json* j = find(key);
        if (!j) throw std::runtime_error("glTF: missing key '" + key + "'");
        return *j;
    }
    const Json& at(size_t i) const {
        if (kind != Array || i >= arr.size()) throw std::runtime_error("glTF: array index out of range");
        return arr[i];
    }
    size_t size() const { return kind == Array ? arr.size() : (kind == Object ? obj.size() : 0); }
    double number(double fallback) const { return kind == Number ? num : fallback; }
    double get(const std::string& key, double fallback) const {
        const Json* j = find(key);
        return j && j->kind == Number ? j->num : fallback;
    }
    std::string get(const std::string& key, const std::string& fallback) const {
        const Json* j = find(key);
        return j && j->kind == String ? j->str : fallback;
    }
    long long index(const std::string& key) const { // -1 when absent
        const Json* j = find(key);
        return j && j->kind == Number ? static_cast<long long>(j->num) : -1;
    }
};

class JsonParser {
public:
    explicit JsonParser(const std::string& text) : s(text) {}
    Json parse() {
        // UTF-8 byte-order mark
        if (s.size() >= 3 && (unsigned char)s[0] == 0xEF && (unsigned char)s[1] == 0xBB && (unsigned char)s[2] == 0xBF) i = 3;
        Json v = value(0);
        ws();
        if (i != s.size()) fail("trailing characters");
        return v;
    }

private:
    const std::string& s;
    size_t i = 0;

    [[noreturn]] void fail(const char* why) const {
        throw std::runtime_error(std::string("JSON parse error at byte ") + std::to_string(i) + ": " + why);
    }
    void ws() {
        while (i < s.size() && (s[i] == ' ' || s[i] == '\t' || s[i] == '\n' || s[i] == '\r')) i++;
    }
    bool lit(const char* w) {
        size_t n = 0;
        while (w[n]) n++;
        if (s.compare(i, n, w) == 0) {
            i += n;
            return true;
        }
        return false;
    }
    Json value(int depth) {
        if (depth > 200) fail("nesting too deep");
        ws();
        if (i >= s.size()) fail("unexpected end");
        Json v;
        const char c = s[i];
        if (c == '{') {
            v.kind = Json::Object;
            i++;
            ws();
            if (i < s.size() && s[i] == '}') {
                i++;
                return v;
            }
            for (;;) {
                ws();
                if (i >= s.size() || s[i] != '"') fail("expected string key");
                std::string key = string();
                ws();
                if (i >= s.size() || s[i] != ':') fail("expected ':'");
                i++;
                v.obj.emplace_back(std::move(key), value(depth + 1));
                ws();
                if (i < s.size() && s[i] == ',') {
                    i++;
                    continue;
                }
                if (i < s.size() && s[i] == '}') {
                    i++;
                    return v;
                }
                fail("expected ',' or '}'");
            }
        }
        if (c == '[') {
            v.kind = Json::Array;
            i++;
            ws();
            if (i < s.size() && s[i] == ']') {
                i++;
                return v;
            }
            for (;;) {
                v.arr.push_back(value(depth + 1));
                ws();
                if (i < s.size() && s[i] == ',') {
                    i++;
                    continue;
                }
                if (i < s.size() && s[i] == ']') {
                    i++;
                    return v;
                }
                fail("expected ',' or ']'");
            }
        }
        if (c == '"') {
            v.kind = Json::String;
            v.str = string();
            return v;
        }
        if (lit("true")) {
            v.kind = Json::Bool;
            v.b = true;
            return v;
        }
        if (lit("false")) {
            v.kind = Json::Bool;
            return v;
        }
        if (lit("null")) return v;
        // number
        const char* begin = s.c_str() + i;
        char* end = nullptr;
        const double d = std::strtod(begin, &end);
        if (end == begin) fail("unexpected character");
        i += static_cast<size_t>(end - begin);
        v.kind = Json::Number;
        v.num = d;
        return v;
    }
    std::string string() {
        std::string out;
        i++; // opening quote
        while (i < s.size() && s[i] != '"') {
            char c = s[i++];
            if (c != '\\') {
                out.push_back(c);
                continue;
            }
            if (i >= s.size()) fail("bad escape");
            c = s[i++];
            switch (c) {
            case '"': out.push_back('"'); break;
            case '\\': out.push_back('\\'); break;
            case '/': out.push_back('/'); break;
            case 'b': out.push_back('\b'); break;
            case 'f': out.push_back('\f'); break;
            case 'n': out.push_back('\n'); break;
            case 'r': out.push_back('\r'); break;
            case 't': out.push_back('\t'); break;
            case 'u': {
                if (i + 4 > s.size()) fail("bad \\u escape");
                unsigned cp = static_cast<unsigned>(std::strtoul(s.substr(i, 4).c_str(), nullptr, 16));
                i += 4;
                if (cp < 0x80) {
                    out.push_back(static_cast<char>(cp));
                } else if (cp < 0x800) {
                    out.push_back(static_cast<char>(0xC0 | (cp >> 6)));
                    out.push_back(static_cast<char>(0x80 | (cp & 0x3F)));
                } else {
                    out.push_back(static_cast<char>(0xE0 | (cp >> 12)));
                    out.push_back(static_cast<char>(0x80 | ((cp >> 6) & 0x3F)));
                    out.push_back(static_cast<char>(0x80 | (cp & 0x3F)));
                }
                break;
            }
            default: fail("bad escape");
            }
        }
        if (i >= s.size()) fail("unterminated string");
        i++; // closing quote
        return out;
    }
};

} // namespace ptb
