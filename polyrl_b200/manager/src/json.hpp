// Minimal JSON value (parse + serialise) for the rollout manager.  Self-contained, header-only.
#pragma once
#include <cctype>
#include <cmath>
#include <cstring>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace prl {

class Json {
 public:
  enum Type { Null, Bool, Int, Double, String, Array, Object };
  using array_t = std::vector<Json>;
  using object_t = std::vector<std::pair<std::string, Json>>;   // insertion-ordered

  Json() : type_(Null) {}
  Json(std::nullptr_t) : type_(Null) {}
  Json(bool b) : type_(Bool), b_(b) {}
  Json(int v) : type_(Int), i_(v) {}
  Json(int64_t v) : type_(Int), i_(v) {}
  Json(uint64_t v) : type_(Int), i_(static_cast<int64_t>(v)) {}
  Json(double v) : type_(Double), d_(v) {}
  Json(const char* s) : type_(String), s_(s) {}
  Json(std::string s) : type_(String), s_(std::move(s)) {}
  static Json array() { Json j; j.type_ = Array; return j; }
  static Json object() { Json j; j.type_ = Object; return j; }

  Type type() const { return type_; }
  bool is_null() const { return type_ == Null; }
  bool is_bool() const { return type_ == Bool; }
  bool is_number() const { return type_ == Int || type_ == Double; }
  bool is_string() const { return type_ == String; }
  bool is_array() const { return type_ == Array; }
  bool is_object() const { return type_ == Object; }

  bool as_bool(bool dflt = false) const { return type_ == Bool ? b_ : dflt; }
  int64_t as_int(int64_t dflt = 0) const {
    if (type_ == Int) return i_;
    if (type_ == Double) return static_cast<int64_t>(d_);
    return dflt;
  }
  double as_double(double dflt = 0.0) const {
    if (type_ == Double) return d_;
    if (type_ == Int) return static_cast<double>(i_);
    return dflt;
  }
  const std::string& as_string() const { static const std::string e; return type_ == String ? s_ : e; }

  array_t& arr() { if (type_ != Array) { *this = array(); } return a_; }
  const array_t& arr() const { static const array_t e; return type_ == Array ? a_ : e; }
  object_t& obj() { if (type_ != Object) { *this = object(); } return o_; }
  const object_t& obj() const { static const object_t e; return type_ == Object ? o_ : e; }
  size_t size() const { return type_ == Array ? a_.size() : (type_ == Object ? o_.size() : 0); }

  void push_back(Json v) { arr().push_back(std::move(v)); }
  bool contains(const std::string& k) const { return find(k) != nullptr; }
  const Json* find(const std::string& k) const {
    if (type_ != Object) return nullptr;
    for (auto& kv : o_) if (kv.first == k) return &kv.second;
    return nullptr;
  }
  Json* find(const std::string& k) {
    if (type_ != Object) return nullptr;
    for (auto& kv : o_) if (kv.first == k) return &kv.second;
    return nullptr;
  }
  const Json& operator[](const std::string& k) const { static const Json n; auto* p = find(k); return p ? *p : n; }
  Json& operator[](const std::string& k) {
    if (type_ != Object) *this = object();
    if (auto* p = find(k)) return *p;
    o_.emplace_back(k, Json());
    return o_.back().second;
  }
  const Json& operator[](size_t i) const { static const Json n; return (type_ == Array && i < a_.size()) ? a_[i] : n; }
  Json& operator[](size_t i) { return a_[i]; }
  // path lookup: j.at_path({"meta_info","finish_reason","type"})
  const Json& at_path(std::initializer_list<const char*> keys) const {
    const Json* cur = this;
    static const Json n;
    for (auto k : keys) { cur = cur->find(k); if (!cur) return n; }
    return *cur;
  }

  std::string dump() const { std::string out; dump_to(out); return out; }
  static Json parse(const std::string& s) {
    size_t i = 0;
    Json v = parse_value(s, i);
    skip_ws(s, i);
    if (i != s.size()) throw std::runtime_error("json: trailing characters");
    return v;
  }
  static bool try_parse(const std::string& s, Json& out) {
    try { out = parse(s); return true; } catch (...) { return false; }
  }

 private:
  Type type_;
  bool b_ = false;
  int64_t i_ = 0;
  double d_ = 0.0;
  std::string s_;
  array_t a_;
  object_t o_;

  static void skip_ws(const std::string& s, size_t& i) {
    while (i < s.size() && (s[i] == ' ' || s[i] == '\n' || s[i] == '\t' || s[i] == '\r')) ++i;
  }
  static void dump_string(const std::string& s, std::string& out) {
    out.push_back('"');
    for (unsigned char c : s) {
      switch (c) {
        case '"': out += "\\\""; break;
        case '\\': out += "\\\\"; break;
        case '\n': out += "\\n"; break;
        case '\r': out += "\\r"; break;
        case '\t': out += "\\t"; break;
        default:
          if (c < 0x20) { char buf[8]; snprintf(buf, sizeof buf, "\\u%04x", c); out += buf; }
          else out.push_back(static_cast<char>(c));
      }
    }
    out.push_back('"');
  }
  void dump_to(std::string& out) const {
    switch (type_) {
      case Null: out += "null"; break;
      case Bool: out += b_ ? "true" : "false"; break;
      case Int: out += std::to_string(i_); break;
      case Double: {
        if (!std::isfinite(d_)) { out += "null"; break; }
        char buf[40]; snprintf(buf, sizeof buf, "%.17g", d_);
        out += buf;
        if (!strpbrk(buf, ".eE")) out += ".0";
        break;
      }
      case String: dump_string(s_, out); break;
      case Array:
        out.push_back('[');
        for (size_t k = 0; k < a_.size(); ++k) { if (k) out.push_back(','); a_[k].dump_to(out); }
        out.push_back(']');
        break;
      case Object:
        out.push_back('{');
        for (size_t k = 0; k < o_.size(); ++k) {
          if (k) out.push_back(',');
          dump_string(o_[k].first, out);
          out.push_back(':');
          o_[k].second.dump_to(out);
        }
        out.push_back('}');
        break;
    }
  }
  static void append_utf8(std::string& out, uint32_t cp) {
    if (cp < 0x80) out.push_back(static_cast<char>(cp));
    else if (cp < 0x800) { out.push_back(static_cast<char>(0xC0 | (cp >> 6))); out.push_back(static_cast<char>(0x80 | (cp & 0x3F))); }
    else if (cp < 0x10000) {
      out.push_back(static_cast<char>(0xE0 | (cp >> 12))); out.push_back(static_cast<char>(0x80 | ((cp >> 6) & 0x3F)));
      out.push_back(static_cast<char>(0x80 | (cp & 0x3F)));
    } else {
      out.push_back(static_cast<char>(0xF0 | (cp >> 18))); out.push_back(static_cast<char>(0x80 | ((cp >> 12) & 0x3F)));
      out.push_back(static_cast<char>(0x80 | ((cp >> 6) & 0x3F))); out.push_back(static_cast<char>(0x80 | (cp & 0x3F)));
    }
  }
  static std::string parse_string(const std::string& s, size_t& i) {
    std::string out;
    ++i;   // opening quote
    while (i < s.size() && s[i] != '"') {
      if (s[i] == '\\') {
        if (++i >= s.size()) throw std::runtime_error("json: bad escape");
        switch (s[i]) {
          case 'n': out.push_back('\n'); break;
          case 't': out.push_back('\t'); break;
          case 'r': out.push_back('\r'); break;
          case 'b': out.push_back('\b'); break;
          case 'f': out.push_back('\f'); break;
          case 'u': {
            if (i + 4 >= s.size()) throw std::runtime_error("json: bad \\u");
            uint32_t cp = static_cast<uint32_t>(strtoul(s.substr(i + 1, 4).c_str(), nullptr, 16));
            i += 4;
            if (cp >= 0xD800 && cp <= 0xDBFF && i + 6 < s.size() && s[i + 1] == '\\' && s[i + 2] == 'u') {
              uint32_t lo = static_cast<uint32_t>(strtoul(s.substr(i + 3, 4).c_str(), nullptr, 16));
              cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
              i += 6;
            }
            append_utf8(out, cp);
            break;
          }
          default: out.push_back(s[i]);
        }
        ++i;
      } else {
        out.push_back(s[i++]);
      }
    }
    if (i >= s.size()) throw std::runtime_error("json: unterminated string");
    ++i;
    return out;
  }
  static Json parse_value(const std::string& s, size_t& i) {
    skip_ws(s, i);
    if (i >= s.size()) throw std::runtime_error("json: unexpected end");
    char c = s[i];
    if (c == '{') {
      Json o = object();
      ++i; skip_ws(s, i);
      if (i < s.size() && s[i] == '}') { ++i; return o; }
      for (;;) {
        skip_ws(s, i);
        if (i >= s.size() || s[i] != '"') throw std::runtime_error("json: expected key");
        std::string k = parse_string(s, i);
        skip_ws(s, i);
        if (i >= s.size() || s[i] != ':') throw std::runtime_error("json: expected ':'");
        ++i;
        o.o_.emplace_back(std::move(k), parse_value(s, i));
        skip_ws(s, i);
        if (i < s.size() && s[i] == ',') { ++i; continue; }
        if (i < s.size() && s[i] == '}') { ++i; return o; }
        throw std::runtime_error("json: expected ',' or '}'");
      }
    }
    if (c == '[') {
      Json a = array();
      ++i; skip_ws(s, i);
      if (i < s.size() && s[i] == ']') { ++i; return a; }
      for (;;) {
        a.a_.push_back(parse_value(s, i));
        skip_ws(s, i);
        if (i < s.size() && s[i] == ',') { ++i; continue; }
        if (i < s.size() && s[i] == ']') { ++i; return a; }
        throw std::runtime_error("json: expected ',' or ']'");
      }
    }
    if (c == '"') return Json(parse_string(s, i));
    if (s.compare(i, 4, "true") == 0) { i += 4; return Json(true); }
    if (s.compare(i, 5, "false") == 0) { i += 5; return Json(false); }
    if (s.compare(i, 4, "null") == 0) { i += 4; return Json(); }
    if (s.compare(i, 3, "NaN") == 0) { i += 3; return Json(); }
    size_t j = i;
    bool is_float = false;
    if (j < s.size() && (s[j] == '-' || s[j] == '+')) ++j;
    if (s.compare(j, 8, "Infinity") == 0) { i = j + 8; return Json(); }
    while (j < s.size() && (isdigit(static_cast<unsigned char>(s[j])) || s[j] == '.' || s[j] == 'e' || s[j] == 'E' || s[j] == '-' || s[j] == '+')) {
      if (s[j] == '.' || s[j] == 'e' || s[j] == 'E') is_float = true;
      ++j;
    }
    if (j == i) throw std::runtime_error(std::string("json: unexpected character '") + c + "'");
    std::string num = s.substr(i, j - i);
    i = j;
    if (is_float) return Json(strtod(num.c_str(), nullptr));
    return Json(static_cast<int64_t>(strtoll(num.c_str(), nullptr, 10)));
  }
};

}  // namespace prl
