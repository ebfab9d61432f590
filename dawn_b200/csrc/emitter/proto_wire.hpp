// Protobuf wire-format reader for serialized StencilInstantiations (IIRSerializer::Format::Byte,
// dawn/src/dawn/Serialization/IIRSerializer.h): decodes the bytes into the same DOM the proto3-JSON
// reader produces (original field names, enums as names, maps as objects with string keys), so one
// IIR front-end serves both formats.  No libprotobuf dependency: the schema is a table
// (proto_schema.inc) restating IIR.proto / statements.proto / enums.proto.
#pragma once
#include "json.hpp"

#include <string>

namespace b200 {
namespace proto {

// Decodes `bytes` as message `message` (e.g. "StencilInstantiation").
json::Value decode(const std::string& bytes, const std::string& message);

// Re-encodes a DOM (as produced by decode() or by the JSON reader) into wire format; used by the
// round-trip tests and by tools that hand IIR bytes to other dawn binaries.
std::string encode(const json::Value& v, const std::string& message);

// proto3 JSON parsers accept a field under its proto name (i_range, block_stmt: what IIRSerializer
// writes) and under its lowerCamelCase JSON name (iRange, blockStmt: what other protobuf runtimes write
// by default).  Renames the latter to the former throughout a DOM of message type `message`.
void normalizeJson(json::Value& v, const std::string& message);

// True if `data` looks like proto3 JSON text (first non-blank byte is '{').
bool looksLikeJson(const std::string& data);

// Schema dump ("Message.number name type" lines) for the schema-vs-.proto test.
std::string schemaDump();

} // namespace proto
} // namespace b200
