// json.h -- a small ordered JSON value, parser and compact writer for the C++ front end.
//
// Stands where the reference uses triton::common::TritonJson (a RapidJSON wrapper that is
// not part of the reference tree).  Objects keep insertion order (the reference's headers
// are written in Add() order); numbers keep their integer-ness (int64 / uint64 / double);
// doubles are written as the shortest round-trip digits in RapidJSON's layout ("1.0",
// "0.001", "1e21", "1.5e-7").
#ifndef TB200_CPP_JSON_H_
#define TB200_CPP_JSON_H_

#include <charconv>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

namespace tb200 { namespace json {

class Value {
 public:
  enum class Type { Null, Bool, Int, UInt, Double, String, Array, Object };

  Value() = default;
  static Value Bool(bool v) { Value x; x.type_ = Type::Bool; x.b_ = v; return x; }
  static Value Int(int64_t v) { Value x; x.type_ = Type::Int; x.i_ = v; return x; }
  static Value UInt(uint64_t v) { Value x; x.type_ = Type::UInt; x.u_ = v; return x; }
  static Value Double(double v) { Value x; x.type_ = Type::Double; x.d_ = v; return x; }
  static Value String(std::string v) { Value x; x.type_ = Type::String; x.s_ = std::move(v); return x; }
  static Value Array() { Value x; x.type_ = Type::Array; return x; }
  static Value Object() { Value x; x.type_ = Type::Object; return x; }

  Type type() const { return type_; }
  bool is_null() const { return type_ == Type::Null; }
  bool is_object() const { return type_ == Type::Object; }
  bool is_array() const { return type_ == Type::Array; }
  bool is_string() const { return type_ == Type::String; }
  bool is_number() const { return type_ == Type::Int || type_ == Type::UInt || type_ == Type::Double; }

  // object
  Value& Add(const std::string& key, Value v) {
    members_.emplace_back(key, std::move(v));
    return members_.back().second;
  }
  const Value* Find(const std::string& key) const {
    for (const auto& kv : members_) {
      if (kv.first == key) return &kv.second;
    }
    return nullptr;
  }
  const std::vector<std::pair<std::string, Value>>& members() const { return members_; }

  // array
  Value& Append(Value v) {
    items_.push_back(std::move(v));
    return items_.back();
  }
  size_t size() const { return type_ == Type::Array ? items_.size() : members_.size(); }
  const Value& operator[](size_t i) const { return items_[i]; }
  void Reserve(size_t n) { items_.reserve(n); }

  // scalars (loose accessors: any numeric kind converts)
  bool AsBool(bool* out) const {
    if (type_ != Type::Bool) return false;
    *out = b_;
    return true;
  }
  bool AsInt(int64_t* out) const {
    if (type_ == Type::Int) *out = i_;
    else if (type_ == Type::UInt) *out = static_cast<int64_t>(u_);
    else if (type_ == Type::Double) *out = static_cast<int64_t>(d_);
    else return false;
    return true;
  }
  bool AsUInt(uint64_t* out) const {
    if (type_ == Type::UInt) *out = u_;
    else if (type_ == Type::Int) *out = static_cast<uint64_t>(i_);
    else if (type_ == Type::Double) *out = static_cast<uint64_t>(d_);
    else return false;
    return true;
  }
  bool AsDouble(double* out) const {
    if (type_ == Type::Double) *out = d_;
    else if (type_ == Type::Int) *out = static_cast<double>(i_);
    else if (type_ == Type::UInt) *out = static_cast<double>(u_);
    else return false;
    return true;
  }
  const std::string& str() const { return s_; }

  // ---- writer -------------------------------------------------------------------
  void Write(std::string* out) const {
    switch (type_) {
      case Type::Null: out->append("null"); break;
      case Type::Bool: out->append(b_ ? "true" : "false"); break;
      case Type::Int: out->append(std::to_string(i_)); break;
      case Type::UInt: out->append(std::to_string(u_)); break;
      case Type::Double: WriteDouble(d_, out); break;
      case Type::String: WriteString(s_, out); break;
      case Type::Array: {
        out->push_back('[');
        for (size_t i = 0; i < items_.size(); ++i) {
          if (i) out->push_back(',');
          items_[i].Write(out);
        }
        out->push_back(']');
        break;
      }
      case Type::Object: {
        out->push_back('{');
        for (size_t i = 0; i < members_.size(); ++i) {
          if (i) out->push_back(',');
          WriteString(members_[i].first, out);
          out->push_back(':');
          members_[i].second.Write(out);
        }
        out->push_back('}');
        break;
      }
    }
  }
  std::string Dump() const {
    std::string s;
    Write(&s);
    return s;
  }

  static void WriteString(const std::string& s, std::string* out) {
    static const char* hex = "0123456789ABCDEF";
    out->push_back('"');
    for (unsigned char c : s) {
      switch (c) {
        case '"': out->append("\\\""); break;
        case '\\': out->append("\\\\"); break;
        case '\b': out->append("\\b"); break;
        case '\f': out->append("\\f"); break;
        case '\n': out->append("\\n"); break;
        case '\r': out->append("\\r"); break;
        case '\t': out->append("\\t"); break;
        default:
          if (c < 0x20) {
            out->append("\\u00");
            out->push_back(hex[c >> 4]);
            out->push_back(hex[c & 15]);
          } else {
            out->push_back(static_cast<char>(c));
          }
      }
    }
    out->push_back('"');
  }

  // shortest round-trip digits, laid out like RapidJSON's Writer (dtoa + Prettify):
  // integral values keep a ".0", plain notation while the decimal point falls within
  // (-6, 21], exponent form "de-7" / "d.ddde21" outside; non-finite -> null
  static void WriteDouble(double v, std::string* out) {
    if (!std::isfinite(v)) {
      out->append("null");
      return;
    }
    if (v == 0.0) {
      out->append(std::signbit(v) ? "-0.0" : "0.0");
      return;
    }
    char buf[64];
    auto res = std::to_chars(buf, buf + sizeof(buf), v, std::chars_format::scientific);
    std::string sci(buf, res.ptr);  // [-]d[.ddd]e[+-]XX
    size_t pos = 0;
    if (sci[0] == '-') {
      out->push_back('-');
      pos = 1;
    }
    const size_t e = sci.find('e');
    std::string digits;
    for (size_t i = pos; i < e; ++i) {
      if (sci[i] != '.') digits.push_back(sci[i]);
    }
    const int exp10 = std::atoi(sci.c_str() + e + 1);
    const int length = static_cast<int>(digits.size());
    const int kk = exp10 + 1;  // position of the decimal point relative to the digits
    if (length <= kk && kk <= 21) {
      out->append(digits);
      out->append(static_cast<size_t>(kk - length), '0');
      out->append(".0");
    } else if (0 < kk && kk <= 21) {
      out->append(digits, 0, static_cast<size_t>(kk));
      out->push_back('.');
      out->append(digits, static_cast<size_t>(kk), std::string::npos);
    } else if (-6 < kk && kk <= 0) {
      out->append("0.");
      out->append(static_cast<size_t>(-kk), '0');
      out->append(digits);
    } else {
      out->push_back(digits[0]);
      if (length > 1) {
        out->push_back('.');
        out->append(digits, 1, std::string::npos);
      }
      out->push_back('e');
      out->append(std::to_string(kk - 1));
    }
  }

  // ---- parser -------------------------------------------------------------------
  // Parses [text, text+n).  On failure returns false and leaves a message in *err.
  static bool Parse(const char* text, size_t n, Value* out, std::string* err) {
    Parser p{text, text + n, err};
    p.SkipWs();
    if (!p.ParseValue(out, 0)) return false;
    p.SkipWs();
    if (p.cur != p.end) return p.Fail("trailing characters after the JSON document");
    return true;
  }

 private:
  struct Parser {
    const char* cur;
    const char* end;
    std::string* err;

    bool Fail(const char* msg) {
      if (err) *err = std::string("failed to parse the JSON: ") + msg;
      return false;
    }
    void SkipWs() {
      while (cur < end && (*cur == ' ' || *cur == '\t' || *cur == '\n' || *cur == '\r')) ++cur;
    }
    bool Literal(const char* lit) {
      const size_t n = strlen(lit);
      if (static_cast<size_t>(end - cur) < n || memcmp(cur, lit, n) != 0) return Fail("invalid literal");
      cur += n;
      return true;
    }
    static void AppendUtf8(uint32_t cp, std::string* s) {
      if (cp < 0x80) {
        s->push_back(static_cast<char>(cp));
      } else if (cp < 0x800) {
        s->push_back(static_cast<char>(0xC0 | (cp >> 6)));
        s->push_back(static_cast<char>(0x80 | (cp & 0x3F)));
      } else if (cp < 0x10000) {
        s->push_back(static_cast<char>(0xE0 | (cp >> 12)));
        s->push_back(static_cast<char>(0x80 | ((cp >> 6) & 0x3F)));
        s->push_back(static_cast<char>(0x80 | (cp & 0x3F)));
      } else {
        s->push_back(static_cast<char>(0xF0 | (cp >> 18)));
        s->push_back(static_cast<char>(0x80 | ((cp >> 12) & 0x3F)));
        s->push_back(static_cast<char>(0x80 | ((cp >> 6) & 0x3F)));
        s->push_back(static_cast<char>(0x80 | (cp & 0x3F)));
      }
    }
    bool Hex4(uint32_t* out) {
      if (end - cur < 4) return Fail("truncated \\u escape");
      uint32_t v = 0;
      for (int i = 0; i < 4; ++i) {
        const char c = *cur++;
        v <<= 4;
        if (c >= '0' && c <= '9') v |= static_cast<uint32_t>(c - '0');
        else if (c >= 'a' && c <= 'f') v |= static_cast<uint32_t>(c - 'a' + 10);
        else if (c >= 'A' && c <= 'F') v |= static_cast<uint32_t>(c - 'A' + 10);
        else return Fail("bad \\u escape");
      }
      *out = v;
      return true;
    }
    bool ParseString(std::string* s) {
      ++cur;  // opening quote
      for (;;) {
        if (cur >= end) return Fail("unterminated string");
        const char c = *cur++;
        if (c == '"') return true;
        if (c != '\\') {
          s->push_back(c);
          continue;
        }
        if (cur >= end) return Fail("unterminated escape");
        const char e = *cur++;
        switch (e) {
          case '"': s->push_back('"'); break;
          case '\\': s->push_back('\\'); break;
          case '/': s->push_back('/'); break;
          case 'b': s->push_back('\b'); break;
          case 'f': s->push_back('\f'); break;
          case 'n': s->push_back('\n'); break;
          case 'r': s->push_back('\r'); break;
          case 't': s->push_back('\t'); break;
          case 'u': {
            uint32_t cp = 0;
            if (!Hex4(&cp)) return false;
            if (cp >= 0xD800 && cp <= 0xDBFF && end - cur >= 6 && cur[0] == '\\' && cur[1] == 'u') {
              cur += 2;
              uint32_t lo = 0;
              if (!Hex4(&lo)) return false;
              cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
            }
            AppendUtf8(cp, s);
            break;
          }
          default: return Fail("unknown escape");
        }
      }
    }
    bool ParseNumber(Value* out) {
      const char* start = cur;
      bool neg = false, integral = true;
      if (cur < end && *cur == '-') {
        neg = true;
        ++cur;
      }
      while (cur < end && ((*cur >= '0' && *cur <= '9') || *cur == '.' || *cur == 'e' || *cur == 'E' || *cur == '+' || *cur == '-')) {
        if (*cur == '.' || *cur == 'e' || *cur == 'E') integral = false;
        ++cur;
      }
      if (cur == start || (neg && cur == start + 1)) return Fail("invalid number");
      if (integral) {
        if (neg) {
          int64_t v = 0;
          auto r = std::from_chars(start, cur, v);
          if (r.ec == std::errc() && r.ptr == cur) {
            *out = Value::Int(v);
            return true;
          }
        } else {
          uint64_t v = 0;
          auto r = std::from_chars(start, cur, v);
          if (r.ec == std::errc() && r.ptr == cur) {
            *out = v <= static_cast<uint64_t>(INT64_MAX) ? Value::Int(static_cast<int64_t>(v)) : Value::UInt(v);
            return true;
          }
        }
      }
      double d = 0.0;
      auto r = std::from_chars(start, cur, d);
      if (r.ec != std::errc() || r.ptr != cur) return Fail("invalid number");
      *out = Value::Double(d);
      return true;
    }
    bool ParseValue(Value* out, int depth) {
      if (depth > 256) return Fail("nesting too deep");
      if (cur >= end) return Fail("unexpected end of document");
      switch (*cur) {
        case '{': {
          ++cur;
          *out = Value::Object();
          SkipWs();
          if (cur < end && *cur == '}') {
            ++cur;
            return true;
          }
          for (;;) {
            SkipWs();
            if (cur >= end || *cur != '"') return Fail("expected a member name");
            std::string key;
            if (!ParseString(&key)) return false;
            SkipWs();
            if (cur >= end || *cur != ':') return Fail("expected ':'");
            ++cur;
            SkipWs();
            Value v;
            if (!ParseValue(&v, depth + 1)) return false;
            out->Add(key, std::move(v));
            SkipWs();
            if (cur < end && *cur == ',') {
              ++cur;
              continue;
            }
            if (cur < end && *cur == '}') {
              ++cur;
              return true;
            }
            return Fail("expected ',' or '}'");
          }
        }
        case '[': {
          ++cur;
          *out = Value::Array();
          SkipWs();
          if (cur < end && *cur == ']') {
            ++cur;
            return true;
          }
          for (;;) {
            SkipWs();
            Value v;
            if (!ParseValue(&v, depth + 1)) return false;
            out->Append(std::move(v));
            SkipWs();
            if (cur < end && *cur == ',') {
              ++cur;
              continue;
            }
            if (cur < end && *cur == ']') {
              ++cur;
              return true;
            }
            return Fail("expected ',' or ']'");
          }
        }
        case '"': {
          std::string s;
          if (!ParseString(&s)) return false;
          *out = Value::String(std::move(s));
          return true;
        }
        case 't':
          if (!Literal("true")) return false;
          *out = Value::Bool(true);
          return true;
        case 'f':
          if (!Literal("false")) return false;
          *out = Value::Bool(false);
          return true;
        case 'n':
          if (!Literal("null")) return false;
          *out = Value();
          return true;
        default:
          return ParseNumber(out);
      }
    }
  };

  Type type_ = Type::Null;
  bool b_ = false;
  int64_t i_ = 0;
  uint64_t u_ = 0;
  double d_ = 0.0;
  std::string s_;
  std::vector<Value> items_;
  std::vector<std::pair<std::string, Value>> members_;
};

}}  // namespace tb200::json

#endif  // TB200_CPP_JSON_H_
