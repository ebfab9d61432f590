// Minimal JSON reader for the proto3-JSON IIR wire format (no third-party dependency).
// proto3 JSON omits default-valued scalars, so every getter takes a default.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace b200 {
namespace json {

class Value {
public:
  enum Kind { Null, Bool, Number, String, Array, Object };
  Kind kind = Null;
  bool b = false;
  double num = 0;
  std::string str;
  std::vector<Value> arr;
  std::vector<std::pair<std::string, Value>> obj; // insertion order preserved

  bool isObject() const { return kind == Object; }
  bool isArray() const { return kind == Array; }
  bool isNull() const { return kind == Null; }

  const Value* find(const std::string& key) const {
    if(kind != Object)
      return nullptr;
    for(auto const& kv : obj)
      if(kv.first == key)
        return &kv.second;
    return nullptr;
  }
  bool has(const std::string& key) const { return find(key) != nullptr; }
  const Value& at(const std::string& key) const {
    const Value* v = find(key);
    if(!v)
      throw std::runtime_error("IIR JSON: missing key '" + key + "'");
    return *v;
  }
  // proto3 int64 may be serialised as a string
  long getInt(const std::string& key, long def = 0) const {
    const Value* v = find(key);
    if(!v || v->isNull())
      return def;
    if(v->kind == Number)
      return static_cast<long>(v->num);
    if(v->kind == String)
      return std::strtol(v->str.c_str(), nullptr, 10);
    if(v->kind == Bool)
      return v->b ? 1 : 0;
    return def;
  }
  double getDouble(const std::string& key, double def = 0) const {
    const Value* v = find(key);
    if(!v || v->isNull())
      return def;
    if(v->kind == Number)
      return v->num;
    if(v->kind == String)
      return std::strtod(v->str.c_str(), nullptr);
    if(v->kind == Bool)
      return v->b ? 1 : 0;
    return def;
  }
  bool getBool(const std::string& key, bool def = false) const {
    const Value* v = find(key);
    if(!v || v->isNull())
      return def;
    if(v->kind == Bool)
      return v->b;
    if(v->kind == Number)
      return v->num != 0;
    return def;
  }
  std::string getString(const std::string& key, const std::string& def = "") const {
    const Value* v = find(key);
    if(!v || v->kind != String)
      return def;
    return v->str;
  }
  long asInt() const {
    if(kind == Number)
      return static_cast<long>(num);
    if(kind == String)
      return std::strtol(str.c_str(), nullptr, 10);
    return 0;
  }
};

class Parser {
  const std::string& s_;
  size_t p_ = 0;

  [[noreturn]] void fail(const std::string& msg) const {
    throw std::runtime_error("JSON parse error at offset " + std::to_string(p_) + ": " + msg);
  }
  void ws() {
    while(p_ < s_.size() && (s_[p_] == ' ' || s_[p_] == '\n' || s_[p_] == '\t' || s_[p_] == '\r'))
      ++p_;
  }
  std::string parseString() {
    if(s_[p_] != '"')
      fail("expected string");
    ++p_;
    std::string out;
    while(p_ < s_.size() && s_[p_] != '"') {
      char c = s_[p_++];
      if(c == '\\') {
        if(p_ >= s_.size())
          fail("bad escape");
        char e = s_[p_++];
        switch(e) {
        case 'n': out += '\n'; break;
        case 't': out += '\t'; break;
        case 'r': out += '\r'; break;
        case 'b': out += '\b'; break;
        case 'f': out += '\f'; break;
        case 'u': {
          if(p_ + 4 > s_.size())
            fail("bad \\u escape");
          unsigned cp = std::strtoul(s_.substr(p_, 4).c_str(), nullptr, 16);
          p_ += 4;
          if(cp < 0x80)
            out += static_cast<char>(cp);
          else if(cp < 0x800) {
            out += static_cast<char>(0xC0 | (cp >> 6));
            out += static_cast<char>(0x80 | (cp & 0x3F));
          } else {
            out += static_cast<char>(0xE0 | (cp >> 12));
            out += static_cast<char>(0x80 | ((cp >> 6) & 0x3F));
            out += static_cast<char>(0x80 | (cp & 0x3F));
          }
          break;
        }
        default: out += e;
        }
      } else
        out += c;
    }
    if(p_ >= s_.size())
      fail("unterminated string");
    ++p_;
    return out;
  }
  Value parseValue() {
    ws();
    if(p_ >= s_.size())
      fail("unexpected end");
    Value v;
    char c = s_[p_];
    if(c == '{') {
      v.kind = Value::Object;
      ++p_;
      ws();
      if(s_[p_] == '}') {
        ++p_;
        return v;
      }
      while(true) {
        ws();
        std::string key = parseString();
        ws();
        if(s_[p_] != ':')
          fail("expected ':'");
        ++p_;
        v.obj.emplace_back(std::move(key), parseValue());
        ws();
        if(s_[p_] == ',') {
          ++p_;
          continue;
        }
        if(s_[p_] == '}') {
          ++p_;
          break;
        }
        fail("expected ',' or '}'");
      }
    } else if(c == '[') {
      v.kind = Value::Array;
      ++p_;
      ws();
      if(s_[p_] == ']') {
        ++p_;
        return v;
      }
      while(true) {
        v.arr.push_back(parseValue());
        ws();
        if(s_[p_] == ',') {
          ++p_;
          continue;
        }
        if(s_[p_] == ']') {
          ++p_;
          break;
        }
        fail("expected ',' or ']'");
      }
    } else if(c == '"') {
      v.kind = Value::String;
      v.str = parseString();
    } else if(s_.compare(p_, 4, "true") == 0) {
      v.kind = Value::Bool;
      v.b = true;
      p_ += 4;
    } else if(s_.compare(p_, 5, "false") == 0) {
      v.kind = Value::Bool;
      v.b = false;
      p_ += 5;
    } else if(s_.compare(p_, 4, "null") == 0) {
      p_ += 4;
    } else {
      char* end = nullptr;
      v.kind = Value::Number;
      v.num = std::strtod(s_.c_str() + p_, &end);
      if(end == s_.c_str() + p_)
        fail("unexpected character");
      p_ = static_cast<size_t>(end - s_.c_str());
    }
    return v;
  }

public:
  explicit Parser(const std::string& s) : s_(s) {}
  Value parse() {
    Value v = parseValue();
    ws();
    if(p_ != s_.size())
      fail("trailing characters");
    return v;
  }
};

inline Value parse(const std::string& text) { return Parser(text).parse(); }

// Compact JSON text of a DOM (numbers with 17 significant digits round-trip doubles exactly).
inline void writeTo(const Value& v, std::string& o) {
  switch(v.kind) {
  case Value::Null: o += "null"; break;
  case Value::Bool: o += v.b ? "true" : "false"; break;
  case Value::Number: {
    char buf[40];
    if(v.num == static_cast<double>(static_cast<long long>(v.num)) && v.num > -1e15 && v.num < 1e15)
      std::snprintf(buf, sizeof buf, "%lld", static_cast<long long>(v.num));
    else if(v.num != v.num)
      std::snprintf(buf, sizeof buf, "\"NaN\"");
    else if(v.num - v.num != 0)
      std::snprintf(buf, sizeof buf, v.num > 0 ? "\"Infinity\"" : "\"-Infinity\"");
    else
      std::snprintf(buf, sizeof buf, "%.17g", v.num);
    o += buf;
    break;
  }
  case Value::String:
    o += '"';
    for(char c : v.str) {
      switch(c) {
      case '"': o += "\\\""; break;
      case '\\': o += "\\\\"; break;
      case '\n': o += "\\n"; break;
      case '\t': o += "\\t"; break;
      case '\r': o += "\\r"; break;
      default:
        if(static_cast<unsigned char>(c) < 0x20) {
          char buf[8];
          std::snprintf(buf, sizeof buf, "\\u%04x", c);
          o += buf;
        } else
          o += c;
      }
    }
    o += '"';
    break;
  case Value::Array:
    o += '[';
    for(size_t i = 0; i < v.arr.size(); ++i) {
      if(i)
        o += ',';
      writeTo(v.arr[i], o);
    }
    o += ']';
    break;
  case Value::Object:
    o += '{';
    for(size_t i = 0; i < v.obj.size(); ++i) {
      if(i)
        o += ',';
      Value k;
      k.kind = Value::String;
      k.str = v.obj[i].first;
      writeTo(k, o);
      o += ':';
      writeTo(v.obj[i].second, o);
    }
    o += '}';
    break;
  }
}
inline std::string write(const Value& v) {
  std::string o;
  writeTo(v, o);
  return o;
}

} // namespace json
} // namespace b200
