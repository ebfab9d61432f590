// Minimal in-memory model of Dawn's optimized IIR, read from the proto3-JSON wire format that
// `dawn-opt` writes and `dawn-codegen` reads (schema: dawn/src/dawn/IIR/proto/IIR/IIR.proto,
// dawn/src/dawn/AST/proto/AST/statements.proto; deserialisation in the reference:
// dawn/src/dawn/Serialization/IIRSerializer.cpp:611-804).
//
// Only what a code generator needs is modelled.  Derived information the wire format does not
// carry (stage extents, field intents, accumulated field extents) is recomputed in analysis.hpp,
// mirroring iir::StencilInstantiation::computeDerivedInfo (IIR/StencilInstantiation.cpp:304-373).
#pragma once
#include "json.hpp"

#include <array>
#include <map>
#include <memory>
#include <set>
#include <string>
#include <vector>

namespace b200 {
namespace iir {

enum class LoopOrder { Forward, Backward, Parallel };
enum class LocationType { None = -1, Cells = 0, Edges = 1, Vertices = 2 };

// iir::FieldAccessType (IIR/FieldAccessMetadata.h:130-138)
enum class AccessType {
  GlobalVariable = 0,
  Literal = 1,
  LocalVariable = 2,
  StencilTemporary = 3,
  InterStencilTemporary = 4,
  Field = 5,
  APIField = 6
};

inline LocationType parseLocation(const std::string& s) {
  if(s == "Cell" || s == "Cells")
    return LocationType::Cells;
  if(s == "Edge" || s == "Edges")
    return LocationType::Edges;
  if(s == "Vertex" || s == "Vertices")
    return LocationType::Vertices;
  return LocationType::None;
}
inline const char* locationName(LocationType l) {
  switch(l) {
  case LocationType::Cells: return "Cells";
  case LocationType::Edges: return "Edges";
  case LocationType::Vertices: return "Vertices";
  default: return "None";
  }
}

// Level marks of ast::Interval: Start = 0, End = 1<<20 (AST/Interval.h); bound = level + offset.
struct Interval {
  static constexpr int End = 1 << 20;
  int lowerLevel = 0, upperLevel = End, lowerOffset = 0, upperOffset = 0;
  int lowerBound() const { return lowerLevel + lowerOffset; }
  int upperBound() const { return upperLevel + upperOffset; }
  bool overlaps(int lo, int hi) const { return lowerBound() <= hi && lo <= upperBound(); }
};

struct Expr;
using ExprPtr = std::shared_ptr<Expr>;

struct Expr {
  enum Kind { Literal, Var, Field, Unary, Binary, Ternary, Assign, FunCall, Reduce };
  Kind kind = Literal;
  // Literal
  std::string value;   // literal text
  std::string typeId;  // Integer | Float | Double | Boolean | Auto
  // Var / Field
  std::string name;
  int accessID = 0;
  bool isExternal = false;    // Var: global
  int di = 0, dj = 0, dk = 0; // Field: cartesian offsets + vertical shift
  bool hasOffset = false;     // Field (unstructured): read at the neighbour inside a reduction
  // Field: vertical indirection (statements.proto FieldAccessExpr.vertical_indirection): the level
  // is the value of this field at the element and level k, plus dk
  std::string kFrom;
  int kFromID = 0;
  // Unary / Binary / Assign / Reduce
  std::string op;
  // children: Unary{0}, Binary{l,r}, Ternary{c,l,r}, Assign{l,r}, FunCall{args...},
  // Reduce{rhs, init, weights...}
  std::vector<ExprPtr> kids;
  // Reduce
  std::vector<LocationType> chain;
  bool includeCenter = false;
  // StencilCall (control flow of the stencil description only): accessID = stencil ID
};

struct Stmt;
using StmtPtr = std::shared_ptr<Stmt>;
struct Stmt {
  enum Kind { ExprStmt, VarDecl, If, Block, Loop, StencilCall };
  Kind kind = ExprStmt;
  ExprPtr expr;               // ExprStmt; If: condition; VarDecl: initialiser (may be null)
  std::string name;           // VarDecl
  std::string typeId;         // VarDecl builtin type
  int accessID = 0;           // VarDecl
  std::vector<StmtPtr> body;  // Block statements / If then-branch
  std::vector<StmtPtr> orelse; // If else-branch
  bool hasElse = false;
  // Loop over the neighbours of a chain (statements.proto LoopStmt / LoopDescriptorChain); body = body
  std::vector<LocationType> chain;
  bool includeCenter = false;
};

struct DoMethod {
  int id = 0;
  Interval interval;
  std::vector<StmtPtr> stmts;
};

struct Extent {
  int iMinus = 0, iPlus = 0, jMinus = 0, jPlus = 0, kMinus = 0, kPlus = 0;
  void merge(const Extent& o) {
    iMinus = std::min(iMinus, o.iMinus), iPlus = std::max(iPlus, o.iPlus);
    jMinus = std::min(jMinus, o.jMinus), jPlus = std::max(jPlus, o.jPlus);
    kMinus = std::min(kMinus, o.kMinus), kPlus = std::max(kPlus, o.kPlus);
  }
  Extent add(const Extent& o) const {
    return {iMinus + o.iMinus, iPlus + o.iPlus, jMinus + o.jMinus,
            jPlus + o.jPlus,   kMinus + o.kMinus, kPlus + o.kPlus};
  }
  bool horizontalZero() const { return !iMinus && !iPlus && !jMinus && !jPlus; }
};

struct Stage {
  int id = 0;
  LocationType location = LocationType::None;
  std::vector<DoMethod> doMethods;
  bool hasIRange = false, hasJRange = false;
  Interval iRange, jRange;
  // derived (analysis.hpp)
  Extent extent;
};

struct Cache {
  std::string type;   // CT_IJ | CT_K | CT_IJK | CT_Bypass
  std::string policy; // CP_Fill ...
  int accessID = 0;
  bool hasWindow = false;
  int windowMinus = 0, windowPlus = 0;
};

struct MultiStage {
  int id = 0;
  LoopOrder loopOrder = LoopOrder::Forward;
  std::vector<Stage> stages;
  std::map<int, Cache> caches;
};

struct Stencil {
  int id = 0;
  std::vector<MultiStage> multiStages;
};

struct FieldDims {
  bool maskI = true, maskJ = true, maskK = true; // Cartesian
  bool unstructured = false;
  std::vector<LocationType> chain; // unstructured: dense location = chain[0]; sparse if size()>1
  bool includeCenter = false;
  bool isSparse() const { return chain.size() > 1; }
};

struct GlobalValue {
  std::string type; // Boolean | Integer | Double | Float
  double value = 0;
  bool valueIsSet = false;
};

struct Instantiation {
  std::string name;
  std::string filename;
  bool unstructured = false;
  std::map<int, std::string> accessIDToName;
  std::map<int, AccessType> accessIDToType;
  std::map<int, FieldDims> fieldDims;
  std::vector<int> apiFieldIDs;       // ordered: run() argument order
  std::vector<int> temporaryFieldIDs;
  std::vector<int> allocatedFieldIDs; // inter-stencil temporaries owned by the wrapper
  std::vector<int> globalVariableIDs;
  std::vector<std::pair<std::string, GlobalValue>> globals; // declaration order
  std::vector<Stencil> stencils;
  // statements of the stencil description around the stencil calls (IIR.controlFlowStatements); empty
  // or a plain list of calls = run every stencil in order
  std::vector<StmtPtr> controlFlow;

  bool isTemporary(int id) const {
    for(int t : temporaryFieldIDs)
      if(t == id)
        return true;
    return false;
  }
  bool isAllocated(int id) const {
    for(int t : allocatedFieldIDs)
      if(t == id)
        return true;
    return false;
  }
  bool isField(int id) const {
    auto it = accessIDToType.find(id);
    if(it == accessIDToType.end())
      return fieldDims.count(id) && (isTemporary(id) || isAllocated(id));
    return it->second == AccessType::StencilTemporary ||
           it->second == AccessType::InterStencilTemporary || it->second == AccessType::Field ||
           it->second == AccessType::APIField;
  }
  const std::string& nameOf(int id) const { return accessIDToName.at(id); }
};

// Parses the serialized StencilInstantiation (JSON text).  Throws std::runtime_error on malformed
// input or on constructs a CUDA-class backend cannot take (un-inlined stencil function calls, cf.
// dawn/src/dawn/CodeGen/Cuda/ASTStencilBody.cpp:75-81).
Instantiation parseInstantiation(const std::string& jsonText);

} // namespace iir
} // namespace b200
