// json.hpp -- small DOM JSON parser for the host mirror (stand-in for std.json.parseFromSlice(std.json.Value, ...),
// src/config.zig:60).  Objects keep insertion order; a duplicate key is an error (std.json's default
// duplicate_field_behavior for a dynamic Value); numbers that look like integers and fit i64 are Integer, all others
// Float (the loader only ever accepts Integer, src/config.zig:162-166).
#pragma once
#include <cerrno>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

namespace tkzjson {

struct Value;
using ValuePtr = std::unique_ptr<Value>;

struct Value {
    enum Kind { Null, Bool, Integer, Float, String, Array, Object } kind = Null;
    bool b = false;
    int64_t i = 0;
    double f = 0;
    std::string s;
    std::vector<ValuePtr> arr;
    std::vector<std::pair<std::string, ValuePtr>> obj;            // insertion order
    std::unordered_map<std::string, size_t> index;                // key -> position in obj

    const Value* get(const std::string& key) const {
        if (kind != Object) return nullptr;
        auto it = index.find(key);
        return it == index.end() ? nullptr : obj[it->second].second.get();
    }
};

class Parser {
   public:
    Parser(const char* p, size_t n) : p_(p), end_(p + n) {}
    // returns nullptr on any syntax error (-> ConfigError.InvalidJson)
    ValuePtr parse() {
        ValuePtr v = value(0);
        if (!v) return nullptr;
        ws();
        if (p_ != end_) return nullptr;
        return v;
    }

   private:
    const char* p_;
    const char* end_;

    void ws() { while (p_ < end_ && (*p_ == ' ' || *p_ == '\t' || *p_ == '\n' || *p_ == '\r')) p_++; }
    bool lit(const char* w) { size_t n = strlen(w); if ((size_t)(end_ - p_) < n || memcmp(p_, w, n) != 0) return false; p_ += n; return true; }

    static void put_utf8(std::string& o, uint32_t cp) {
        if (cp < 0x80) o.push_back((char)cp);
        else if (cp < 0x800) { o.push_back((char)(0xC0 | (cp >> 6))); o.push_back((char)(0x80 | (cp & 0x3F))); }
        else if (cp < 0x10000) { o.push_back((char)(0xE0 | (cp >> 12))); o.push_back((char)(0x80 | ((cp >> 6) & 0x3F))); o.push_back((char)(0x80 | (cp & 0x3F))); }
        else { o.push_back((char)(0xF0 | (cp >> 18))); o.push_back((char)(0x80 | ((cp >> 12) & 0x3F))); o.push_back((char)(0x80 | ((cp >> 6) & 0x3F))); o.push_back((char)(0x80 | (cp & 0x3F))); }
    }
    bool hex4(uint32_t& out) {
        if (end_ - p_ < 4) return false;
        out = 0;
        for (int k = 0; k < 4; k++) {
            char c = *p_++; out <<= 4;
            if (c >= '0' && c <= '9') out |= (uint32_t)(c - '0');
            else if (c >= 'a' && c <= 'f') out |= (uint32_t)(c - 'a' + 10);
            else if (c >= 'A' && c <= 'F') out |= (uint32_t)(c - 'A' + 10);
            else return false;
        }
        return true;
    }
    bool string(std::string& out) {
        if (p_ >= end_ || *p_ != '"') return false;
        p_++;
        for (;;) {
            if (p_ >= end_) return false;
            unsigned char c = (unsigned char)*p_++;
            if (c == '"') return true;
            if (c < 0x20) return false;
            if (c != '\\') { out.push_back((char)c); continue; }
            if (p_ >= end_) return false;
            char e = *p_++;
            switch (e) {
                case '"': out.push_back('"'); break;
                case '\\': out.push_back('\\'); break;
                case '/': out.push_back('/'); break;
                case 'b': out.push_back('\b'); break;
                case 'f': out.push_back('\f'); break;
                case 'n': out.push_back('\n'); break;
                case 'r': out.push_back('\r'); break;
                case 't': out.push_back('\t'); break;
                case 'u': {
                    uint32_t cp;
                    if (!hex4(cp)) return false;
                    if (cp >= 0xD800 && cp <= 0xDBFF) {
                        uint32_t lo;
                        if (end_ - p_ >= 6 && p_[0] == '\\' && p_[1] == 'u') {
                            const char* save = p_; p_ += 2;
                            if (!hex4(lo)) return false;
                            if (lo >= 0xDC00 && lo <= 0xDFFF) cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
                            else p_ = save;                       // lone high surrogate: kept as is (WTF-8)
                        }
                    }
                    put_utf8(out, cp);
                    break;
                }
                default: return false;
            }
        }
    }
    ValuePtr number() {
        const char* s = p_;
        bool intlike = true;
        if (p_ < end_ && *p_ == '-') p_++;
        if (p_ >= end_ || !(*p_ >= '0' && *p_ <= '9')) return nullptr;
        if (*p_ == '0') p_++; else while (p_ < end_ && *p_ >= '0' && *p_ <= '9') p_++;
        if (p_ < end_ && *p_ == '.') { intlike = false; p_++; if (p_ >= end_ || !(*p_ >= '0' && *p_ <= '9')) return nullptr; while (p_ < end_ && *p_ >= '0' && *p_ <= '9') p_++; }
        if (p_ < end_ && (*p_ == 'e' || *p_ == 'E')) {
            intlike = false; p_++;
            if (p_ < end_ && (*p_ == '+' || *p_ == '-')) p_++;
            if (p_ >= end_ || !(*p_ >= '0' && *p_ <= '9')) return nullptr;
            while (p_ < end_ && *p_ >= '0' && *p_ <= '9') p_++;
        }
        std::string t(s, p_);
        ValuePtr v(new Value());
        if (intlike && t != "-0") {
            errno = 0; char* e = nullptr;
            long long x = strtoll(t.c_str(), &e, 10);
            if (errno == 0 && e && *e == 0) { v->kind = Value::Integer; v->i = x; return v; }
        }
        v->kind = Value::Float; v->f = strtod(t.c_str(), nullptr);
        return v;
    }
    ValuePtr value(int depth) {
        if (depth > 512) return nullptr;
        ws();
        if (p_ >= end_) return nullptr;
        char c = *p_;
        if (c == '{') {
            p_++;
            ValuePtr v(new Value()); v->kind = Value::Object;
            ws();
            if (p_ < end_ && *p_ == '}') { p_++; return v; }
            for (;;) {
                ws();
                std::string k;
                if (!string(k)) return nullptr;
                ws();
                if (p_ >= end_ || *p_ != ':') return nullptr;
                p_++;
                ValuePtr x = value(depth + 1);
                if (!x) return nullptr;
                if (v->index.count(k)) return nullptr;            // duplicate field -> error
                v->index.emplace(k, v->obj.size());
                v->obj.emplace_back(std::move(k), std::move(x));
                ws();
                if (p_ >= end_) return nullptr;
                if (*p_ == ',') { p_++; continue; }
                if (*p_ == '}') { p_++; return v; }
                return nullptr;
            }
        }
        if (c == '[') {
            p_++;
            ValuePtr v(new Value()); v->kind = Value::Array;
            ws();
            if (p_ < end_ && *p_ == ']') { p_++; return v; }
            for (;;) {
                ValuePtr x = value(depth + 1);
                if (!x) return nullptr;
                v->arr.push_back(std::move(x));
                ws();
                if (p_ >= end_) return nullptr;
                if (*p_ == ',') { p_++; continue; }
                if (*p_ == ']') { p_++; return v; }
                return nullptr;
            }
        }
        if (c == '"') { ValuePtr v(new Value()); v->kind = Value::String; if (!string(v->s)) return nullptr; return v; }
        if (c == 't') { if (!lit("true")) return nullptr; ValuePtr v(new Value()); v->kind = Value::Bool; v->b = true; return v; }
        if (c == 'f') { if (!lit("false")) return nullptr; ValuePtr v(new Value()); v->kind = Value::Bool; v->b = false; return v; }
        if (c == 'n') { if (!lit("null")) return nullptr; return ValuePtr(new Value()); }
        return number();
    }
};

}  // namespace tkzjson
