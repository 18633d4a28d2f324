// Minimal JSON reader for DeepMimic's JSON-in-.txt assets (characters, controllers,
// motions, datasets, terrain).  Replaces the role of the vendored jsoncpp used by the
// reference (R/DeepMimicCore/util/json/, call sites e.g. anim/KinTree.cpp:Load,
// anim/Motion.cpp:LoadJson).  Only the subset those files use: objects, arrays, numbers,
// strings, true/false/null; C/C++ comments are skipped like jsoncpp's default reader.
#pragma once
#include <cmath>
#include <cstdlib>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace dmh {

class Json {
public:
    enum Type { Null, Bool, Number, String, Array, Object };
    Type type = Null;
    bool b = false;
    double num = 0;
    std::string str;
    std::vector<Json> arr;
    std::vector<std::pair<std::string, Json>> obj;  // keeps file order

    bool isNull() const { return type == Null; }
    bool isNumeric() const { return type == Number || type == Bool; }
    bool isArray() const { return type == Array; }
    bool isObject() const { return type == Object; }
    bool isString() const { return type == String; }
    size_t size() const { return type == Array ? arr.size() : (type == Object ? obj.size() : 0); }

    const Json& operator[](const std::string& key) const {
        static const Json null_json;
        if (type != Object) return null_json;
        for (const auto& kv : obj)
            if (kv.first == key) return kv.second;
        return null_json;
    }
    const Json& operator[](size_t i) const { return arr.at(i); }
    bool has(const std::string& key) const { return !(*this)[key].isNull(); }

    double asDouble(double def = 0) const {
        if (type == Number) return num;
        if (type == Bool) return b ? 1.0 : 0.0;
        return def;
    }
    int asInt(int def = 0) const { return isNumeric() ? static_cast<int>(asDouble()) : def; }
    bool asBool(bool def = false) const {
        if (type == Bool) return b;
        if (type == Number) return num != 0;
        return def;
    }
    std::string asString(const std::string& def = "") const { return type == String ? str : def; }

    double get(const std::string& key, double def) const {
        const Json& v = (*this)[key];
        return v.isNull() ? def : v.asDouble(def);
    }
    bool getBool(const std::string& key, bool def) const {
        const Json& v = (*this)[key];
        return v.isNull() ? def : v.asBool(def);
    }

    static Json parse(const std::string& text) {
        Parser p{text, 0};
        Json v = p.value();
        p.ws();
        return v;
    }
    static Json parseFile(const std::string& path) {
        std::ifstream f(path);
        if (!f.good()) throw std::runtime_error("cannot open JSON file: " + path);
        std::stringstream ss;
        ss << f.rdbuf();
        return parse(ss.str());
    }

private:
    struct Parser {
        const std::string& s;
        size_t i;
        [[noreturn]] void fail(const char* what) const {
            throw std::runtime_error(std::string("JSON parse error: ") + what + " at offset " + std::to_string(i));
        }
        void ws() {
            for (;;) {
                while (i < s.size() && (s[i] == ' ' || s[i] == '\t' || s[i] == '\n' || s[i] == '\r')) ++i;
                if (i + 1 < s.size() && s[i] == '/' && s[i + 1] == '/') {
                    while (i < s.size() && s[i] != '\n') ++i;
                } else if (i + 1 < s.size() && s[i] == '/' && s[i + 1] == '*') {
                    i += 2;
                    while (i + 1 < s.size() && !(s[i] == '*' && s[i + 1] == '/')) ++i;
                    i += 2;
                } else {
                    break;
                }
            }
        }
        Json value() {
            ws();
            if (i >= s.size()) fail("unexpected end");
            char c = s[i];
            if (c == '{') return object();
            if (c == '[') return array();
            if (c == '"') {
                Json j;
                j.type = String;
                j.str = string();
                return j;
            }
            if (s.compare(i, 4, "true") == 0) { i += 4; Json j; j.type = Bool; j.b = true; return j; }
            if (s.compare(i, 5, "false") == 0) { i += 5; Json j; j.type = Bool; j.b = false; return j; }
            if (s.compare(i, 4, "null") == 0) { i += 4; return Json(); }
            return number();
        }
        Json number() {
            const char* beg = s.c_str() + i;
            char* end = nullptr;
            double v = std::strtod(beg, &end);
            if (end == beg) fail("bad number");
            i += static_cast<size_t>(end - beg);
            Json j;
            j.type = Number;
            j.num = v;
            return j;
        }
        std::string string() {
            std::string out;
            ++i;  // opening quote
            while (i < s.size() && s[i] != '"') {
                if (s[i] == '\\' && i + 1 < s.size()) {
                    char e = s[i + 1];
                    switch (e) {
                        case 'n': out += '\n'; break;
                        case 't': out += '\t'; break;
                        case 'r': out += '\r'; break;
                        case 'b': out += '\b'; break;
                        case 'f': out += '\f'; break;
                        default: out += e; break;
                    }
                    i += 2;
                } else {
                    out += s[i++];
                }
            }
            if (i >= s.size()) fail("unterminated string");
            ++i;
            return out;
        }
        Json array() {
            Json j;
            j.type = Array;
            ++i;
            ws();
            if (i < s.size() && s[i] == ']') { ++i; return j; }
            for (;;) {
                j.arr.push_back(value());
                ws();
                if (i >= s.size()) fail("unterminated array");
                if (s[i] == ',') { ++i; ws(); if (i < s.size() && s[i] == ']') { ++i; break; } continue; }
                if (s[i] == ']') { ++i; break; }
                fail("expected , or ]");
            }
            return j;
        }
        Json object() {
            Json j;
            j.type = Object;
            ++i;
            ws();
            if (i < s.size() && s[i] == '}') { ++i; return j; }
            for (;;) {
                ws();
                if (i >= s.size() || s[i] != '"') fail("expected key string");
                std::string k = string();
                ws();
                if (i >= s.size() || s[i] != ':') fail("expected :");
                ++i;
                Json v = value();
                j.obj.emplace_back(std::move(k), std::move(v));
                ws();
                if (i >= s.size()) fail("unterminated object");
                if (s[i] == ',') { ++i; ws(); if (i < s.size() && s[i] == '}') { ++i; break; } continue; }
                if (s[i] == '}') { ++i; break; }
                fail("expected , or }");
            }
            return j;
        }
    };
};

}  // namespace dmh
