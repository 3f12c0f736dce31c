// rg_json.h — a small strict JSON reader, enough for the GameConfig schema
// (reference: serde_json behind GameConfig::from_json, core/src/lib.rs:144-146).
// Host only. Integers keep full 128-bit precision because `seed` is a u128.
#pragma once
#include <cstdint>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace rgjson {

struct Value;
using ValuePtr = std::shared_ptr<Value>;

struct Value {
  enum Kind { Null, Bool, Int, Real, String, Array, Object } kind = Null;
  bool b = false;
  bool negative = false;
  unsigned __int128 mag = 0;  // |integer|
  double real = 0.0;
  std::string str;
  std::vector<ValuePtr> arr;
  std::vector<std::pair<std::string, ValuePtr>> obj;  // insertion order, duplicates: last wins on lookup

  const Value* get(const std::string& key) const {
    const Value* hit = nullptr;
    for (const auto& kv : obj)
      if (kv.first == key) hit = kv.second.get();
    return hit;
  }
  bool is_int() const { return kind == Int; }
  int64_t as_i64() const {
    if (kind != Int) throw std::runtime_error("expected an integer");
    if (mag > (unsigned __int128)INT64_MAX + (negative ? 1 : 0)) throw std::runtime_error("integer out of range");
    return negative ? -(int64_t)mag : (int64_t)mag;
  }
  uint64_t as_u64() const {
    if (kind != Int || negative || mag > (unsigned __int128)UINT64_MAX) throw std::runtime_error("expected an unsigned integer");
    return (uint64_t)mag;
  }
};

class Parser {
 public:
  explicit Parser(const std::string& s) : s_(s) {}
  ValuePtr parse() {
    ValuePtr v = value();
    ws();
    if (p_ != s_.size()) fail("trailing characters");
    return v;
  }

 private:
  const std::string& s_;
  size_t p_ = 0;
  int depth_ = 0;

  [[noreturn]] void fail(const char* m) const {
    size_t line = 1, col = 1;
    for (size_t i = 0; i < p_ && i < s_.size(); ++i) {
      if (s_[i] == '\n') { ++line; col = 1; } else ++col;
    }
    throw std::runtime_error(std::string(m) + " at line " + std::to_string(line) + " column " + std::to_string(col));
  }
  void ws() {
    while (p_ < s_.size() && (s_[p_] == ' ' || s_[p_] == '\t' || s_[p_] == '\n' || s_[p_] == '\r')) ++p_;
  }
  bool lit(const char* w) {
    size_t n = 0;
    while (w[n]) ++n;
    if (s_.compare(p_, n, w) == 0) { p_ += n; return true; }
    return false;
  }
  ValuePtr value() {
    if (++depth_ > 128) fail("recursion limit exceeded");
    ws();
    if (p_ >= s_.size()) fail("EOF while parsing a value");
    auto v = std::make_shared<Value>();
    char c = s_[p_];
    if (c == '{') {
      ++p_;
      v->kind = Value::Object;
      ws();
      if (p_ < s_.size() && s_[p_] == '}') { ++p_; --depth_; return v; }
      for (;;) {
        ws();
        if (p_ >= s_.size() || s_[p_] != '"') fail("key must be a string");
        std::string k = string();
        ws();
        if (p_ >= s_.size() || s_[p_] != ':') fail("expected `:`");
        ++p_;
        v->obj.emplace_back(k, value());
        ws();
        if (p_ < s_.size() && s_[p_] == ',') { ++p_; continue; }
        if (p_ < s_.size() && s_[p_] == '}') { ++p_; break; }
        fail("expected `,` or `}`");
      }
    } else if (c == '[') {
      ++p_;
      v->kind = Value::Array;
      ws();
      if (p_ < s_.size() && s_[p_] == ']') { ++p_; --depth_; return v; }
      for (;;) {
        v->arr.push_back(value());
        ws();
        if (p_ < s_.size() && s_[p_] == ',') { ++p_; continue; }
        if (p_ < s_.size() && s_[p_] == ']') { ++p_; break; }
        fail("expected `,` or `]`");
      }
    } else if (c == '"') {
      v->kind = Value::String;
      v->str = string();
    } else if (lit("true")) {
      v->kind = Value::Bool;
      v->b = true;
    } else if (lit("false")) {
      v->kind = Value::Bool;
    } else if (lit("null")) {
      v->kind = Value::Null;
    } else if (c == '-' || (c >= '0' && c <= '9')) {
      number(*v);
    } else {
      fail("expected value");
    }
    --depth_;
    return v;
  }
  void number(Value& v) {
    size_t start = p_;
    if (s_[p_] == '-') { v.negative = true; ++p_; }
    if (p_ >= s_.size() || s_[p_] < '0' || s_[p_] > '9') fail("invalid number");
    if (s_[p_] == '0' && p_ + 1 < s_.size() && s_[p_ + 1] >= '0' && s_[p_ + 1] <= '9') fail("invalid number");
    unsigned __int128 m = 0;
    bool overflow = false;
    while (p_ < s_.size() && s_[p_] >= '0' && s_[p_] <= '9') {
      unsigned __int128 nm = m * 10 + (unsigned)(s_[p_] - '0');
      if (nm / 10 != m) overflow = true;
      m = nm;
      ++p_;
    }
    bool real = false;
    if (p_ < s_.size() && s_[p_] == '.') {
      real = true;
      ++p_;
      if (p_ >= s_.size() || s_[p_] < '0' || s_[p_] > '9') fail("invalid number");
      while (p_ < s_.size() && s_[p_] >= '0' && s_[p_] <= '9') ++p_;
    }
    if (p_ < s_.size() && (s_[p_] == 'e' || s_[p_] == 'E')) {
      real = true;
      ++p_;
      if (p_ < s_.size() && (s_[p_] == '+' || s_[p_] == '-')) ++p_;
      if (p_ >= s_.size() || s_[p_] < '0' || s_[p_] > '9') fail("invalid number");
      while (p_ < s_.size() && s_[p_] >= '0' && s_[p_] <= '9') ++p_;
    }
    if (real || overflow) {
      v.kind = Value::Real;
      v.real = std::stod(s_.substr(start, p_ - start));
    } else {
      v.kind = Value::Int;
      v.mag = m;
      if (m == 0) v.negative = false;
    }
  }
  std::string string() {
    ++p_;  // opening quote
    std::string out;
    for (;;) {
      if (p_ >= s_.size()) fail("EOF while parsing a string");
      unsigned char c = (unsigned char)s_[p_++];
      if (c == '"') break;
      if (c < 0x20) fail("control character in string");
      if (c != '\\') { out.push_back((char)c); continue; }
      if (p_ >= s_.size()) fail("EOF while parsing a string");
      char e = s_[p_++];
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
          unsigned cp = hex4();
          if (cp >= 0xD800 && cp < 0xDC00 && p_ + 1 < s_.size() && s_[p_] == '\\' && s_[p_ + 1] == 'u') {
            p_ += 2;
            unsigned lo = hex4();
            cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
          }
          if (cp < 0x80) out.push_back((char)cp);
          else if (cp < 0x800) { out.push_back((char)(0xC0 | (cp >> 6))); out.push_back((char)(0x80 | (cp & 0x3F))); }
          else if (cp < 0x10000) {
            out.push_back((char)(0xE0 | (cp >> 12)));
            out.push_back((char)(0x80 | ((cp >> 6) & 0x3F)));
            out.push_back((char)(0x80 | (cp & 0x3F)));
          } else {
            out.push_back((char)(0xF0 | (cp >> 18)));
            out.push_back((char)(0x80 | ((cp >> 12) & 0x3F)));
            out.push_back((char)(0x80 | ((cp >> 6) & 0x3F)));
            out.push_back((char)(0x80 | (cp & 0x3F)));
          }
          break;
        }
        default: fail("invalid escape");
      }
    }
    return out;
  }
  unsigned hex4() {
    if (p_ + 4 > s_.size()) fail("invalid unicode escape");
    unsigned v = 0;
    for (int i = 0; i < 4; ++i) {
      char c = s_[p_++];
      v <<= 4;
      if (c >= '0' && c <= '9') v |= (unsigned)(c - '0');
      else if (c >= 'a' && c <= 'f') v |= (unsigned)(c - 'a' + 10);
      else if (c >= 'A' && c <= 'F') v |= (unsigned)(c - 'A' + 10);
      else fail("invalid unicode escape");
    }
    return v;
  }
};

inline ValuePtr parse(const std::string& s) { return Parser(s).parse(); }

}  // namespace rgjson
