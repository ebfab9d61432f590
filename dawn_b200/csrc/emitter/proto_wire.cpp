#include "proto_wire.hpp"

#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <sstream>
#include <stdexcept>
#include <vector>

namespace b200 {
namespace proto {
namespace {

using json::Value;

const char* const kSchemaText =
#include "proto_schema.inc"
    ;

enum class T { I32, Bool, F64, F32, Str, Enum, Msg, W32, Map };

struct TypeRef {
  T t = T::I32;
  std::string ref; // enum / message name
};

struct Field {
  int number = 0;
  std::string name;
  bool repeated = false;
  bool isMap = false;
  TypeRef type; // for maps: the value type
  TypeRef key;  // maps only
};

struct Message {
  std::string name;
  std::vector<Field> fields;
  const Field* byNumber(int n) const {
    for(auto const& f : fields)
      if(f.number == n)
        return &f;
    return nullptr;
  }
  const Field* byName(const std::string& n) const {
    for(auto const& f : fields)
      if(f.name == n)
        return &f;
    return nullptr;
  }
};

struct Enum {
  std::map<int, std::string> byNumber;
  std::map<std::string, int> byName;
};

struct Schema {
  std::map<std::string, Message> messages;
  std::map<std::string, Enum> enums;
  std::vector<std::string> order;
};

TypeRef parseType(const std::string& s) {
  TypeRef r;
  if(s == "i32")
    r.t = T::I32;
  else if(s == "bool")
    r.t = T::Bool;
  else if(s == "f64")
    r.t = T::F64;
  else if(s == "f32")
    r.t = T::F32;
  else if(s == "str")
    r.t = T::Str;
  else if(s == "w32")
    r.t = T::W32;
  else if(s.compare(0, 2, "e:") == 0)
    r.t = T::Enum, r.ref = s.substr(2);
  else if(s.compare(0, 2, "m:") == 0)
    r.t = T::Msg, r.ref = s.substr(2);
  else
    throw std::logic_error("proto schema: bad type '" + s + "'");
  return r;
}

const Schema& schema() {
  static const Schema S = [] {
    Schema s;
    std::istringstream in(kSchemaText);
    std::string line;
    Message* cur = nullptr;
    while(std::getline(in, line)) {
      if(line.empty())
        continue;
      std::istringstream ls(line);
      if(line[0] == '#') {
        std::string name, item;
        ls >> name;
        Enum& e = s.enums[name.substr(1)];
        while(ls >> item) {
          size_t eq = item.find('=');
          int num = std::stoi(item.substr(eq + 1));
          e.byNumber[num] = item.substr(0, eq);
          e.byName[item.substr(0, eq)] = num;
        }
      } else if(line[0] == '@') {
        std::string name = line.substr(1);
        cur = &s.messages[name];
        cur->name = name;
        s.order.push_back(name);
      } else {
        Field f;
        std::string type;
        ls >> f.number >> f.name >> type;
        if(!type.empty() && type[0] == '*') {
          f.repeated = true;
          type = type.substr(1);
        }
        if(type.compare(0, 4, "map:") == 0) {
          size_t c = type.find(':', 4);
          f.isMap = true;
          f.key = parseType(type.substr(4, c - 4));
          f.type = parseType(type.substr(c + 1));
        } else
          f.type = parseType(type);
        cur->fields.push_back(f);
      }
    }
    return s;
  }();
  return S;
}

bool isMap(const Field& f) { return f.isMap; }

const Message& messageOf(const std::string& name) {
  auto it = schema().messages.find(name);
  if(it == schema().messages.end())
    throw std::runtime_error("IIR bytes: unknown message type '" + name + "'");
  return it->second;
}

// ---------------------------------------------------------------------------------------- reader
struct Reader {
  const unsigned char* p;
  const unsigned char* end;

  [[noreturn]] static void fail(const std::string& m) {
    throw std::runtime_error("IIR bytes: " + m);
  }
  bool done() const { return p >= end; }
  uint64_t varint() {
    uint64_t v = 0;
    for(int shift = 0; shift < 70; shift += 7) {
      if(p >= end)
        fail("truncated varint");
      unsigned char c = *p++;
      v |= static_cast<uint64_t>(c & 0x7F) << shift;
      if(!(c & 0x80))
        return v;
    }
    fail("varint too long");
  }
  Reader sub() {
    uint64_t n = varint();
    if(n > static_cast<uint64_t>(end - p))
      fail("length-delimited field runs past the end of its message");
    Reader r{p, p + n};
    p += n;
    return r;
  }
  uint64_t fixed(int bytes) {
    if(end - p < bytes)
      fail("truncated fixed-width field");
    uint64_t v = 0;
    for(int i = 0; i < bytes; ++i)
      v |= static_cast<uint64_t>(p[i]) << (8 * i);
    p += bytes;
    return v;
  }
  void skip(int wireType) {
    switch(wireType) {
    case 0: varint(); break;
    case 1: fixed(8); break;
    case 2: sub(); break;
    case 5: fixed(4); break;
    default: fail("unsupported wire type " + std::to_string(wireType));
    }
  }
};

Value number(double d) {
  Value v;
  v.kind = Value::Number;
  v.num = d;
  return v;
}
Value boolean(bool b) {
  Value v;
  v.kind = Value::Bool;
  v.b = b;
  return v;
}
Value string(std::string s) {
  Value v;
  v.kind = Value::String;
  v.str = std::move(s);
  return v;
}

Value decodeMessage(Reader r, const Message& msg);

// one scalar / message occurrence of type `t`; `wireType` as found on the wire
Value decodeOne(Reader& r, const TypeRef& t, int wireType) {
  switch(t.t) {
  case T::I32:
    if(wireType != 0)
      Reader::fail("int32 field with wire type " + std::to_string(wireType));
    return number(static_cast<double>(static_cast<int32_t>(static_cast<int64_t>(r.varint()))));
  case T::Bool:
    if(wireType != 0)
      Reader::fail("bool field with wire type " + std::to_string(wireType));
    return boolean(r.varint() != 0);
  case T::Enum: {
    if(wireType != 0)
      Reader::fail("enum field with wire type " + std::to_string(wireType));
    int n = static_cast<int>(static_cast<int64_t>(r.varint()));
    auto const& e = schema().enums.at(t.ref);
    auto it = e.byNumber.find(n);
    return it != e.byNumber.end() ? string(it->second) : number(n); // unknown values stay numeric
  }
  case T::F64: {
    if(wireType != 1)
      Reader::fail("double field with wire type " + std::to_string(wireType));
    uint64_t bits = r.fixed(8);
    double d;
    std::memcpy(&d, &bits, 8);
    return number(d);
  }
  case T::F32: {
    if(wireType != 5)
      Reader::fail("float field with wire type " + std::to_string(wireType));
    uint32_t bits = static_cast<uint32_t>(r.fixed(4));
    float f;
    std::memcpy(&f, &bits, 4);
    return number(f);
  }
  case T::Str: {
    if(wireType != 2)
      Reader::fail("string field with wire type " + std::to_string(wireType));
    Reader s = r.sub();
    return string(std::string(reinterpret_cast<const char*>(s.p), s.end - s.p));
  }
  case T::Msg:
    if(wireType != 2)
      Reader::fail("message field with wire type " + std::to_string(wireType));
    return decodeMessage(r.sub(), messageOf(t.ref));
  case T::W32: { // google.protobuf.Int32Value { int32 value = 1; } -> plain number in JSON
    if(wireType != 2)
      Reader::fail("wrapper field with wire type " + std::to_string(wireType));
    Reader s = r.sub();
    int32_t v = 0;
    while(!s.done()) {
      uint64_t tag = s.varint();
      if((tag >> 3) == 1 && (tag & 7) == 0)
        v = static_cast<int32_t>(static_cast<int64_t>(s.varint()));
      else
        s.skip(static_cast<int>(tag & 7));
    }
    return number(v);
  }
  case T::Map: break;
  }
  Reader::fail("bad schema type");
}

bool packable(const TypeRef& t) {
  return t.t == T::I32 || t.t == T::Bool || t.t == T::Enum || t.t == T::F64 || t.t == T::F32;
}
int scalarWireType(const TypeRef& t) {
  switch(t.t) {
  case T::F64: return 1;
  case T::F32: return 5;
  case T::I32:
  case T::Bool:
  case T::Enum: return 0;
  default: return 2;
  }
}

std::string keyString(const Value& k) {
  if(k.kind == Value::String)
    return k.str;
  if(k.kind == Value::Bool)
    return k.b ? "true" : "false";
  return std::to_string(static_cast<long long>(k.num));
}

Value* slot(Value& obj, const std::string& name) {
  for(auto& kv : obj.obj)
    if(kv.first == name)
      return &kv.second;
  obj.obj.emplace_back(name, Value());
  return &obj.obj.back().second;
}

Value decodeMessage(Reader r, const Message& msg) {
  Value out;
  out.kind = Value::Object;
  while(!r.done()) {
    uint64_t tag = r.varint();
    int num = static_cast<int>(tag >> 3), wt = static_cast<int>(tag & 7);
    const Field* f = msg.byNumber(num);
    if(!f) {
      r.skip(wt); // unknown field: newer schema revision
      continue;
    }
    if(isMap(*f)) {
      if(wt != 2)
        Reader::fail("map field '" + f->name + "' with wire type " + std::to_string(wt));
      Reader e = r.sub();
      Value key, val;
      bool haveKey = false, haveVal = false;
      while(!e.done()) {
        uint64_t t2 = e.varint();
        int n2 = static_cast<int>(t2 >> 3), w2 = static_cast<int>(t2 & 7);
        const TypeRef& kt = f->key;
        if(n2 == 1)
          key = decodeOne(e, kt, w2), haveKey = true;
        else if(n2 == 2)
          val = decodeOne(e, f->type, w2), haveVal = true;
        else
          e.skip(w2);
      }
      if(!haveKey) // proto3 default key
        key = f->key.t == T::Str ? string("") : number(0);
      if(!haveVal) {
        if(f->type.t == T::Msg)
          val.kind = Value::Object;
        else if(f->type.t == T::Str)
          val = string("");
        else
          val = number(0);
      }
      Value* m = slot(out, f->name);
      m->kind = Value::Object;
      *slot(*m, keyString(key)) = std::move(val);
      continue;
    }
    if(f->repeated) {
      Value* a = slot(out, f->name);
      a->kind = Value::Array;
      if(wt == 2 && packable(f->type)) {
        Reader pk = r.sub();
        while(!pk.done())
          a->arr.push_back(decodeOne(pk, f->type, scalarWireType(f->type)));
      } else
        a->arr.push_back(decodeOne(r, f->type, wt));
      continue;
    }
    Value v = decodeOne(r, f->type, wt);
    Value* s = slot(out, f->name);
    if(f->type.t == T::Msg && s->kind == Value::Object && v.kind == Value::Object) {
      // a singular message field seen twice merges (protobuf semantics)
      for(auto& kv : v.obj)
        *slot(*s, kv.first) = std::move(kv.second);
    } else
      *s = std::move(v);
  }
  return out;
}

// ---------------------------------------------------------------------------------------- writer
void putVarint(std::string& o, uint64_t v) {
  while(v >= 0x80) {
    o.push_back(static_cast<char>((v & 0x7F) | 0x80));
    v >>= 7;
  }
  o.push_back(static_cast<char>(v));
}
void putTag(std::string& o, int num, int wt) { putVarint(o, (static_cast<uint64_t>(num) << 3) | wt); }
void putFixed(std::string& o, uint64_t v, int bytes) {
  for(int i = 0; i < bytes; ++i)
    o.push_back(static_cast<char>((v >> (8 * i)) & 0xFF));
}

std::string encodeMessage(const Value& v, const Message& msg);

long long intOf(const Value& v) {
  if(v.kind == Value::Number)
    return static_cast<long long>(v.num);
  if(v.kind == Value::String)
    return std::strtoll(v.str.c_str(), nullptr, 10);
  if(v.kind == Value::Bool)
    return v.b ? 1 : 0;
  return 0;
}

void encodeScalarPayload(std::string& o, const Value& v, const TypeRef& t) {
  switch(t.t) {
  case T::I32: putVarint(o, static_cast<uint64_t>(static_cast<int64_t>(intOf(v)))); break;
  case T::Bool: putVarint(o, (v.kind == Value::Bool ? v.b : intOf(v) != 0) ? 1 : 0); break;
  case T::Enum: {
    long long n = 0;
    if(v.kind == Value::String) {
      auto const& e = schema().enums.at(t.ref);
      auto it = e.byName.find(v.str);
      if(it == e.byName.end())
        throw std::runtime_error("IIR: unknown enum value '" + v.str + "' of " + t.ref);
      n = it->second;
    } else
      n = intOf(v);
    putVarint(o, static_cast<uint64_t>(static_cast<int64_t>(n)));
    break;
  }
  case T::F64: {
    double d = v.kind == Value::String ? std::strtod(v.str.c_str(), nullptr)
               : v.kind == Value::Bool ? (v.b ? 1.0 : 0.0)
                                       : v.num;
    uint64_t bits;
    std::memcpy(&bits, &d, 8);
    putFixed(o, bits, 8);
    break;
  }
  case T::F32: {
    float f = static_cast<float>(v.kind == Value::String ? std::strtod(v.str.c_str(), nullptr) : v.num);
    uint32_t bits;
    std::memcpy(&bits, &f, 4);
    putFixed(o, bits, 4);
    break;
  }
  default: throw std::logic_error("not a scalar");
  }
}

void encodeOne(std::string& o, int num, const Value& v, const TypeRef& t) {
  switch(t.t) {
  case T::Str:
    putTag(o, num, 2);
    putVarint(o, v.str.size());
    o += v.str;
    break;
  case T::Msg: {
    std::string body = encodeMessage(v, messageOf(t.ref));
    putTag(o, num, 2);
    putVarint(o, body.size());
    o += body;
    break;
  }
  case T::W32: {
    std::string body;
    if(intOf(v) != 0) {
      putTag(body, 1, 0);
      putVarint(body, static_cast<uint64_t>(static_cast<int64_t>(intOf(v))));
    }
    putTag(o, num, 2);
    putVarint(o, body.size());
    o += body;
    break;
  }
  default:
    putTag(o, num, scalarWireType(t));
    encodeScalarPayload(o, v, t);
  }
}

bool isDefault(const Value& v, const TypeRef& t) {
  switch(t.t) {
  case T::Str: return v.str.empty();
  case T::Msg:
  case T::W32: return false;
  case T::Bool: return !(v.kind == Value::Bool ? v.b : intOf(v) != 0);
  case T::F64:
  case T::F32: return v.kind == Value::Number && v.num == 0 && !std::signbit(v.num);
  case T::Enum:
    if(v.kind == Value::String) {
      auto const& e = schema().enums.at(t.ref);
      auto it = e.byName.find(v.str);
      return it != e.byName.end() && it->second == 0;
    }
    return intOf(v) == 0;
  default: return intOf(v) == 0;
  }
}

// oneof members keep explicit presence: a default value is still written
bool inOneof(const Message& m, const Field& f) {
  static const std::map<std::string, std::vector<std::string>> oneofs = {
      {"FieldDimensions", {"cartesian_horizontal_dimension", "unstructured_horizontal_dimension"}},
      {"StencilFunctionArg", {"field_value", "direction_value", "offset_value"}},
      {"Interval", {"special_lower_level", "lower_level", "special_upper_level", "upper_level"}},
      {"Type", {"name", "builtin_type"}},
      {"GlobalVariableValue",
       {"boolean_value", "integer_value", "double_value", "string_value", "float_value"}},
  };
  auto it = oneofs.find(m.name);
  if(it == oneofs.end())
    return false;
  for(auto const& n : it->second)
    if(n == f.name)
      return true;
  return false;
}

std::string encodeMessage(const Value& v, const Message& msg) {
  std::string o;
  if(v.kind != Value::Object)
    return o;
  // fields in field-number order, like the official serializers
  std::vector<const Field*> fs;
  for(auto const& f : msg.fields)
    fs.push_back(&f);
  for(size_t a = 0; a < fs.size(); ++a)
    for(size_t b = a + 1; b < fs.size(); ++b)
      if(fs[b]->number < fs[a]->number)
        std::swap(fs[a], fs[b]);
  for(const Field* f : fs) {
    const Value* x = v.find(f->name);
    if(!x || x->isNull())
      continue;
    if(isMap(*f)) {
      const TypeRef& kt = f->key;
      // deterministic serialization: entries ordered by key (numerically for integer keys)
      std::vector<const std::pair<std::string, Value>*> entries;
      for(auto const& kv : x->obj)
        entries.push_back(&kv);
      std::stable_sort(entries.begin(), entries.end(), [&](auto const* a, auto const* b) {
        if(kt.t == T::Str)
          return a->first < b->first;
        return std::strtoll(a->first.c_str(), nullptr, 10) < std::strtoll(b->first.c_str(), nullptr, 10);
      });
      for(auto const* kvp : entries) {
        auto const& kv = *kvp;
        std::string entry;
        Value key = kt.t == T::Str ? string(kv.first) : number(std::strtod(kv.first.c_str(), nullptr));
        // map entries always carry both key and value, defaults included
        encodeOne(entry, 1, key, kt);
        encodeOne(entry, 2, kv.second, f->type);
        putTag(o, f->number, 2);
        putVarint(o, entry.size());
        o += entry;
      }
    } else if(f->repeated) {
      if(x->arr.empty())
        continue;
      if(packable(f->type)) {
        std::string body;
        for(auto const& e : x->arr)
          encodeScalarPayload(body, e, f->type);
        putTag(o, f->number, 2);
        putVarint(o, body.size());
        o += body;
      } else
        for(auto const& e : x->arr)
          encodeOne(o, f->number, e, f->type);
    } else {
      if(!inOneof(msg, *f) && isDefault(*x, f->type))
        continue;
      encodeOne(o, f->number, *x, f->type);
    }
  }
  return o;
}

} // namespace

namespace {
// protobuf's default JSON name of a field: underscores dropped, the following letter upper-cased
std::string jsonName(const std::string& protoName) {
  std::string o;
  bool up = false;
  for(char c : protoName) {
    if(c == '_') {
      up = true;
      continue;
    }
    o += up ? static_cast<char>(std::toupper(static_cast<unsigned char>(c))) : c;
    up = false;
  }
  return o;
}

void normalizeMessage(Value& v, const Message& msg) {
  if(v.kind != Value::Object)
    return;
  for(auto& kv : v.obj) {
    const Field* f = msg.byName(kv.first);
    if(!f)
      for(auto const& cand : msg.fields)
        if(jsonName(cand.name) == kv.first) {
          f = &cand;
          kv.first = cand.name;
          break;
        }
    if(!f || f->type.t != T::Msg)
      continue;
    const Message& sub = messageOf(f->type.ref);
    if(isMap(*f)) {
      for(auto& e : kv.second.obj)
        normalizeMessage(e.second, sub);
    } else if(f->repeated) {
      for(auto& e : kv.second.arr)
        normalizeMessage(e, sub);
    } else
      normalizeMessage(kv.second, sub);
  }
}
} // namespace

void normalizeJson(json::Value& v, const std::string& message) { normalizeMessage(v, messageOf(message)); }

json::Value decode(const std::string& bytes, const std::string& message) {
  Reader r{reinterpret_cast<const unsigned char*>(bytes.data()),
           reinterpret_cast<const unsigned char*>(bytes.data()) + bytes.size()};
  return decodeMessage(r, messageOf(message));
}

std::string encode(const json::Value& v, const std::string& message) {
  return encodeMessage(v, messageOf(message));
}

bool looksLikeJson(const std::string& data) {
  for(char c : data) {
    if(c == ' ' || c == '\n' || c == '\t' || c == '\r')
      continue;
    return c == '{';
  }
  return false;
}

std::string schemaDump() {
  std::ostringstream o;
  auto typeName = [](const TypeRef& t) -> std::string {
    switch(t.t) {
    case T::I32: return "int32";
    case T::Bool: return "bool";
    case T::F64: return "double";
    case T::F32: return "float";
    case T::Str: return "string";
    case T::W32: return "Int32Value";
    case T::Enum: return "enum " + t.ref;
    case T::Msg: return t.ref;
    case T::Map: break;
    }
    return "?";
  };
  for(auto const& name : schema().order) {
    auto const& m = schema().messages.at(name);
    for(auto const& f : m.fields) {
      o << m.name << " " << f.number << " " << f.name << " ";
      if(isMap(f)) {
        o << "map<" << typeName(f.key) << "," << typeName(f.type) << ">";
      } else
        o << (f.repeated ? "repeated " : "") << typeName(f.type);
      o << "\n";
    }
  }
  for(auto const& e : schema().enums)
    for(auto const& kv : e.second.byNumber)
      o << "enum " << e.first << " " << kv.second << " " << kv.first << "\n";
  return o.str();
}

} // namespace proto
} // namespace b200
