// Minimal JSON reader for glTF 2.0 documents (RFC 8259 values; numbers as double; \uXXXX escapes to UTF-8).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace ptrs_host {

struct Json {
  enum Kind { Null, Bool, Number, String, Array, Object } kind = Null;
  bool b = false;
  double num = 0.0;
  std::string str;
  std::vector<Json> arr;
  std::vector<std::pair<std::string, Json>> obj;  // document order

  bool is(Kind k) const { return kind == k; }
  const Json* get(const std::string& key) const {
    if (kind != Object) return nullptr;
    for (const auto& kv : obj)
      if (kv.first == key) return &kv.second;
    return nullptr;
  }
  const Json& at(const std::string& key) const {
    const Json* v = get(key);
    if (!v) throw std::runtime_error("JSON: missing key '" + key + "'");
    return *v;
  }
  const Json& at(size_t i) const {
    if (kind != Array || i >= arr.size()) throw std::runtime_error("JSON: array index out of range");
    return arr[i];
  }
  size_t size() const { return kind == Array ? arr.size() : (kind == Object ? obj.size() : 0); }
  double number_or(const std::string& key, double dflt) const {
    const Json* v = get(key);
    return v && v->kind == Number ? v->num : dflt;
  }
  long index_or(const std::string& key, long dflt) const {
    const Json* v = get(key);
    return v && v->kind == Number ? (long)v->num : dflt;
  }
  std::string string_or(const std::string& key, const std::string& dflt) const {
    const Json* v = get(key);
    return v && v->kind == String ? v->str : dflt;
  }
};

class JsonParser {
 public:
  JsonParser(const char* p, size_t n) : p_(p), n_(n) {}
  Json parse_document() {
    Json v = value();
    ws();
    if (i_ != n_) fail("trailing characters");
    return v;
  }

 private:
  const char* p_;
  size_t n_, i_ = 0;
  [[noreturn]] void fail(const std::string& what) const { throw std::runtime_error("JSON: " + what + " at byte " + std::to_string(i_)); }
  void ws() {
    while (i_ < n_ && (p_[i_] == ' ' || p_[i_] == '\t' || p_[i_] == '\n' || p_[i_] == '\r')) ++i_;
  }
  bool lit(const char* s) {
    const size_t k = std::char_traits<char>::length(s);
    if (i_ + k <= n_ && std::char_traits<char>::compare(p_ + i_, s, k) == 0) {
      i_ += k;
      return true;
    }
    return false;
  }
  static void utf8(std::string& out, unsigned cp) {
    if (cp < 0x80) out.push_back((char)cp);
    else if (cp < 0x800) {
      out.push_back((char)(0xC0 | (cp >> 6)));
      out.push_back((char)(0x80 | (cp & 0x3F)));
    } else if (cp < 0x10000) {
      out.push_back((char)(0xE0 | (cp >> 12)));
      out.push_back((char)(0x80 | ((cp >> 6) & 0x3F)));
      out.push_back((char)(0x80 | (cp & 0x3F)));
    } else {
      out.push_back((char)(0xF0 | (cp >> 18)));
      out.push_back((char)(0x80 | ((cp >> 12) & 0x3F)));
      out.push_back((char)(0x80 | ((cp >> 6) & 0x3F)));
      out.push_back((char)(0x80 | (cp & 0x3F)));
    }
  }
  unsigned hex4() {
    if (i_ + 4 > n_) fail("truncated \\u escape");
    unsigned v = 0;
    for (int k = 0; k < 4; ++k) {
      const char c = p_[i_++];
      v <<= 4;
      if (c >= '0' && c <= '9') v |= (unsigned)(c - '0');
      else if (c >= 'a' && c <= 'f') v |= (unsigned)(c - 'a' + 10);
      else if (c >= 'A' && c <= 'F') v |= (unsigned)(c - 'A' + 10);
      else fail("bad hex digit");
    }
    return v;
  }
  std::string string() {
    if (i_ >= n_ || p_[i_] != '"') fail("expected a string");
    ++i_;
    std::string out;
    for (;;) {
      if (i_ >= n_) fail("unterminated string");
      const char c = p_[i_++];
      if (c == '"') return out;
      if (c != '\\') {
        out.push_back(c);
        continue;
      }
      if (i_ >= n_) fail("unterminated escape");
      const char e = p_[i_++];
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
          if (cp >= 0xD800 && cp < 0xDC00 && i_ + 1 < n_ && p_[i_] == '\\' && p_[i_ + 1] == 'u') {
            i_ += 2;
            const unsigned lo = hex4();
            cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
          }
          utf8(out, cp);
          break;
        }
        default: fail("bad escape");
      }
    }
  }
  Json value() {
    ws();
    if (i_ >= n_) fail("unexpected end");
    Json v;
    const char c = p_[i_];
    if (c == '{') {
      ++i_;
      v.kind = Json::Object;
      ws();
      if (i_ < n_ && p_[i_] == '}') {
        ++i_;
        return v;
      }
      for (;;) {
        ws();
        std::string key = string();
        ws();
        if (i_ >= n_ || p_[i_] != ':') fail("expected ':'");
        ++i_;
        v.obj.emplace_back(std::move(key), value());
        ws();
        if (i_ < n_ && p_[i_] == ',') {
          ++i_;
          continue;
        }
        if (i_ < n_ && p_[i_] == '}') {
          ++i_;
          return v;
        }
        fail("expected ',' or '}'");
      }
    }
    if (c == '[') {
      ++i_;
      v.kind = Json::Array;
      ws();
      if (i_ < n_ && p_[i_] == ']') {
        ++i_;
        return v;
      }
      for (;;) {
        v.arr.push_back(value());
        ws();
        if (i_ < n_ && p_[i_] == ',') {
          ++i_;
          continue;
        }
        if (i_ < n_ && p_[i_] == ']') {
          ++i_;
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
    {
      const std::string tok(p_ + i_, std::min<size_t>(n_ - i_, 64));
      char* end = nullptr;
      const double d = std::strtod(tok.c_str(), &end);
      if (end == tok.c_str()) fail("unexpected character");
      i_ += (size_t)(end - tok.c_str());
      v.kind = Json::Number;
      v.num = d;
      return v;
    }
  }
};

}  // namespace ptrs_host
