// Minimal JSON reader for the config surface of the per-keyframe path.
// The reference reads resources/config.json through jsoncpp (src/config.cpp:9-30); only the value kinds
// that file uses are needed: objects, arrays, strings, numbers, booleans, null (plus // and /* */ comments,
// which jsoncpp tolerates).
#pragma once
#include <cstdlib>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace rss {

struct JsonValue {
    enum Kind { Null, Bool, Number, String, Array, Object } kind = Null;
    bool b = false;
    double num = 0.0;
    std::string str;
    std::vector<JsonValue> arr;
    std::vector<std::pair<std::string, JsonValue>> obj;

    const JsonValue* find(const std::string& key) const {
        if (kind != Object) return nullptr;
        for (auto& kv : obj)
            if (kv.first == key) return &kv.second;
        return nullptr;
    }
    // jsoncpp-like conversions (config.cpp:104-202): bools and numbers convert into each other
    double asDouble() const { return kind == Number ? num : (kind == Bool ? (b ? 1.0 : 0.0) : 0.0); }
    int asInt() const { return (int)asDouble(); }
    bool asBool() const { return kind == Bool ? b : (kind == Number ? num != 0.0 : false); }
};

class JsonParser {
    const char* p_;
    const char* end_;
    void skip() {
        for (;;) {
            while (p_ < end_ && (*p_ == ' ' || *p_ == '\t' || *p_ == '\n' || *p_ == '\r')) p_++;
            if (p_ + 1 < end_ && p_[0] == '/' && p_[1] == '/') {
                while (p_ < end_ && *p_ != '\n') p_++;
            } else if (p_ + 1 < end_ && p_[0] == '/' && p_[1] == '*') {
                p_ += 2;
                while (p_ + 1 < end_ && !(p_[0] == '*' && p_[1] == '/')) p_++;
                p_ += 2;
            } else
                return;
        }
    }
    [[noreturn]] void fail(const char* what) { throw std::runtime_error(std::string("JSON: ") + what); }
    std::string parseString() {
        std::string s;
        p_++;  // opening quote
        while (p_ < end_ && *p_ != '"') {
            if (*p_ == '\\' && p_ + 1 < end_) {
                p_++;
                switch (*p_) {
                    case 'n': s.push_back('\n'); break;
                    case 't': s.push_back('\t'); break;
                    case 'r': s.push_back('\r'); break;
                    case 'b': s.push_back('\b'); break;
                    case 'f': s.push_back('\f'); break;
                    case 'u': {  // keep the low byte; config strings are ASCII
                        if (p_ + 4 >= end_) fail("bad \\u escape");
                        s.push_back((char)strtol(std::string(p_ + 1, p_ + 5).c_str(), nullptr, 16));
                        p_ += 4;
                        break;
                    }
                    default: s.push_back(*p_);
                }
                p_++;
            } else
                s.push_back(*p_++);
        }
        if (p_ >= end_) fail("unterminated string");
        p_++;
        return s;
    }
    JsonValue parseValue() {
        skip();
        if (p_ >= end_) fail("unexpected end");
        JsonValue v;
        if (*p_ == '{') {
            v.kind = JsonValue::Object;
            p_++;
            skip();
            if (p_ < end_ && *p_ == '}') { p_++; return v; }
            for (;;) {
                skip();
                if (p_ >= end_ || *p_ != '"') fail("expected key");
                std::string k = parseString();
                skip();
                if (p_ >= end_ || *p_ != ':') fail("expected ':'");
                p_++;
                v.obj.emplace_back(k, parseValue());
                skip();
                if (p_ < end_ && *p_ == ',') { p_++; continue; }
                if (p_ < end_ && *p_ == '}') { p_++; break; }
                fail("expected ',' or '}'");
            }
        } else if (*p_ == '[') {
            v.kind = JsonValue::Array;
            p_++;
            skip();
            if (p_ < end_ && *p_ == ']') { p_++; return v; }
            for (;;) {
                v.arr.push_back(parseValue());
                skip();
                if (p_ < end_ && *p_ == ',') { p_++; continue; }
                if (p_ < end_ && *p_ == ']') { p_++; break; }
                fail("expected ',' or ']'");
            }
        } else if (*p_ == '"') {
            v.kind = JsonValue::String;
            v.str = parseString();
        } else if (end_ - p_ >= 4 && std::string(p_, p_ + 4) == "true") {
            v.kind = JsonValue::Bool; v.b = true; p_ += 4;
        } else if (end_ - p_ >= 5 && std::string(p_, p_ + 5) == "false") {
            v.kind = JsonValue::Bool; v.b = false; p_ += 5;
        } else if (end_ - p_ >= 4 && std::string(p_, p_ + 4) == "null") {
            p_ += 4;
        } else {
            char* e = nullptr;
            v.num = strtod(p_, &e);
            if (e == p_) fail("unexpected character");
            v.kind = JsonValue::Number;
            p_ = e;
        }
        return v;
    }

public:
    static JsonValue parse(const std::string& text) {
        JsonParser ps;
        ps.p_ = text.data();
        ps.end_ = text.data() + text.size();
        JsonValue v = ps.parseValue();
        ps.skip();
        return v;
    }
};

}  // namespace rss
