// minijson.h -- small JSON value/parser/printer for the test-driver protocol of the host shim
// (scheduler_host.cpp) and of the object-level oracle (oracle/sched_oracle.cpp).  Not part of the hot path.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace mj {

struct Value {
    enum Type { Null, Bool, Int, Real, Str, Arr, Obj } type = Null;
    bool b = false;
    int64_t i = 0;
    double d = 0;
    std::string s;
    std::vector<Value> a;
    std::vector<std::pair<std::string, Value>> o;

    Value() {}
    static Value boolean(bool v) { Value x; x.type = Bool; x.b = v; return x; }
    static Value integer(int64_t v) { Value x; x.type = Int; x.i = v; return x; }
    static Value string(const std::string &v) { Value x; x.type = Str; x.s = v; return x; }
    static Value array() { Value x; x.type = Arr; return x; }
    static Value object() { Value x; x.type = Obj; return x; }

    bool is_null() const { return type == Null; }
    const Value *find(const std::string &k) const {
        if (type != Obj) return nullptr;
        for (auto &kv : o) if (kv.first == k) return &kv.second;
        return nullptr;
    }
    const Value &at(const std::string &k) const {
        static const Value none;
        const Value *v = find(k);
        return v ? *v : none;
    }
    void set(const std::string &k, Value v) { o.emplace_back(k, std::move(v)); }
    void push(Value v) { a.push_back(std::move(v)); }
    int64_t as_int(int64_t def = 0) const { return type == Int ? i : type == Real ? (int64_t)d : type == Bool ? (b ? 1 : 0) : def; }
    std::string as_str(const std::string &def = "") const { return type == Str ? s : def; }
    bool as_bool(bool def = false) const { return type == Bool ? b : type == Int ? i != 0 : def; }
};

struct Parser {
    const std::string &t;
    size_t p = 0;
    explicit Parser(const std::string &text) : t(text) {}
    [[noreturn]] void fail(const char *m) { throw std::runtime_error(std::string("json: ") + m + " at " + std::to_string(p)); }
    void ws() { while (p < t.size() && (t[p] == ' ' || t[p] == '\n' || t[p] == '\t' || t[p] == '\r')) p++; }
    static void utf8(std::string &out, uint32_t c) {
        if (c < 0x80) out += (char)c;
        else if (c < 0x800) { out += (char)(0xC0 | (c >> 6)); out += (char)(0x80 | (c & 0x3F)); }
        else if (c < 0x10000) { out += (char)(0xE0 | (c >> 12)); out += (char)(0x80 | ((c >> 6) & 0x3F)); out += (char)(0x80 | (c & 0x3F)); }
        else { out += (char)(0xF0 | (c >> 18)); out += (char)(0x80 | ((c >> 12) & 0x3F)); out += (char)(0x80 | ((c >> 6) & 0x3F)); out += (char)(0x80 | (c & 0x3F)); }
    }
    uint32_t hex4() {
        if (p + 4 > t.size()) fail("bad \\u");
        uint32_t v = (uint32_t)std::strtoul(t.substr(p, 4).c_str(), nullptr, 16);
        p += 4;
        return v;
    }
    std::string str() {
        if (t[p] != '"') fail("expected string");
        p++;
        std::string out;
        while (p < t.size() && t[p] != '"') {
            char c = t[p++];
            if (c != '\\') { out += c; continue; }
            if (p >= t.size()) fail("bad escape");
            char e = t[p++];
            switch (e) {
                case 'n': out += '\n'; break; case 't': out += '\t'; break; case 'r': out += '\r'; break;
                case 'b': out += '\b'; break; case 'f': out += '\f'; break;
                case 'u': {
                    uint32_t c1 = hex4();
                    if (c1 >= 0xD800 && c1 < 0xDC00 && p + 1 < t.size() && t[p] == '\\' && t[p + 1] == 'u') {
                        p += 2;
                        uint32_t c2 = hex4();
                        c1 = 0x10000 + ((c1 - 0xD800) << 10) + (c2 - 0xDC00);
                    }
                    utf8(out, c1);
                    break;
                }
                default: out += e;
            }
        }
        if (p >= t.size()) fail("unterminated string");
        p++;
        return out;
    }
    Value value() {
        ws();
        if (p >= t.size()) fail("eof");
        char c = t[p];
        Value v;
        if (c == '{') {
            v.type = Value::Obj; p++; ws();
            if (t[p] == '}') { p++; return v; }
            for (;;) {
                ws(); std::string k = str(); ws();
                if (t[p] != ':') fail("expected :");
                p++;
                v.o.emplace_back(k, value());
                ws();
                if (t[p] == ',') { p++; continue; }
                if (t[p] == '}') { p++; break; }
                fail("expected , or }");
            }
        } else if (c == '[') {
            v.type = Value::Arr; p++; ws();
            if (t[p] == ']') { p++; return v; }
            for (;;) {
                v.a.push_back(value());
                ws();
                if (t[p] == ',') { p++; continue; }
                if (t[p] == ']') { p++; break; }
                fail("expected , or ]");
            }
        } else if (c == '"') {
            v.type = Value::Str; v.s = str();
        } else if (t.compare(p, 4, "true") == 0) { v = Value::boolean(true); p += 4; }
        else if (t.compare(p, 5, "false") == 0) { v = Value::boolean(false); p += 5; }
        else if (t.compare(p, 4, "null") == 0) { p += 4; }
        else {
            size_t q = p;
            bool real = false;
            if (t[q] == '-') q++;
            while (q < t.size() && ((t[q] >= '0' && t[q] <= '9') || t[q] == '.' || t[q] == 'e' || t[q] == 'E' || t[q] == '+' || t[q] == '-')) {
                if (t[q] == '.' || t[q] == 'e' || t[q] == 'E') real = true;
                q++;
            }
            if (q == p) fail("unexpected character");
            std::string num = t.substr(p, q - p);
            if (real) { v.type = Value::Real; v.d = std::strtod(num.c_str(), nullptr); }
            else { v.type = Value::Int; v.i = std::strtoll(num.c_str(), nullptr, 10); }
            p = q;
        }
        return v;
    }
};

inline Value parse(const std::string &text) {
    Parser ps(text);
    Value v = ps.value();
    return v;
}

inline void dump_str(const std::string &s, std::string &out) {
    out += '"';
    for (unsigned char c : s) {
        if (c == '"') out += "\\\"";
        else if (c == '\\') out += "\\\\";
        else if (c == '\n') out += "\\n";
        else if (c == '\t') out += "\\t";
        else if (c == '\r') out += "\\r";
        else if (c < 0x20) { char b[8]; snprintf(b, sizeof b, "\\u%04x", c); out += b; }
        else out += (char)c;
    }
    out += '"';
}

inline void dump(const Value &v, std::string &out) {
    switch (v.type) {
        case Value::Null: out += "null"; break;
        case Value::Bool: out += v.b ? "true" : "false"; break;
        case Value::Int: out += std::to_string(v.i); break;
        case Value::Real: { char b[40]; snprintf(b, sizeof b, "%.17g", v.d); out += b; break; }
        case Value::Str: dump_str(v.s, out); break;
        case Value::Arr: {
            out += '[';
            for (size_t i = 0; i < v.a.size(); i++) { if (i) out += ','; dump(v.a[i], out); }
            out += ']';
            break;
        }
        case Value::Obj: {
            out += '{';
            for (size_t i = 0; i < v.o.size(); i++) { if (i) out += ','; dump_str(v.o[i].first, out); out += ':'; dump(v.o[i].second, out); }
            out += '}';
            break;
        }
    }
}

inline std::string dump(const Value &v) { std::string s; dump(v, s); return s; }

}  // namespace mj
