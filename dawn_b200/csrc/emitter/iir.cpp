#include "iir.hpp"
#include "proto_wire.hpp"

#include <algorithm>
#include <stdexcept>

namespace b200 {
namespace iir {
namespace {

using json::Value;

Interval parseInterval(const Value& v) {
  Interval iv;
  iv.lowerOffset = static_cast<int>(v.getInt("lower_offset"));
  iv.upperOffset = static_cast<int>(v.getInt("upper_offset"));
  // oneof LowerLevel { SpecialLevel special_lower_level; int32 lower_level; }  (statements.proto:120-147)
  // proto3 JSON drops a defaulted enum inside a oneof only if the oneof is unset, so a missing
  // key means level 0 (Start).
  if(v.has("special_lower_level"))
    iv.lowerLevel = v.getString("special_lower_level") == "End" ? Interval::End : 0;
  else
    iv.lowerLevel = static_cast<int>(v.getInt("lower_level"));
  if(v.has("special_upper_level"))
    iv.upperLevel = v.getString("special_upper_level") == "End" ? Interval::End : 0;
  else
    iv.upperLevel = static_cast<int>(v.getInt("upper_level"));
  return iv;
}

const std::pair<std::string, Value>& only(const Value& v, const char* what) {
  if(!v.isObject() || v.obj.empty())
    throw std::runtime_error(std::string("IIR JSON: empty ") + what + " node");
  return v.obj.front();
}

ExprPtr parseExpr(const Value& node);

ExprPtr parseExprPayload(const std::string& kind, const Value& v) {
  auto e = std::make_shared<Expr>();
  if(kind == "literal_access_expr") {
    e->kind = Expr::Literal;
    e->value = v.getString("value");
    if(v.has("type"))
      e->typeId = v.at("type").getString("type_id", "Float");
    else
      e->typeId = "Float";
    if(v.has("data"))
      e->accessID = static_cast<int>(v.at("data").getInt("accessID"));
  } else if(kind == "var_access_expr") {
    e->kind = Expr::Var;
    e->name = v.getString("name");
    e->isExternal = v.getBool("is_external");
    if(v.has("data"))
      e->accessID = static_cast<int>(v.at("data").getInt("accessID"));
    if(v.has("index") && !v.at("index").isNull() && v.at("index").isObject() &&
       !v.at("index").obj.empty())
      throw std::runtime_error("IIR: array-indexed variable access is not supported: " + e->name);
  } else if(kind == "field_access_expr") {
    e->kind = Expr::Field;
    e->name = v.getString("name");
    e->dk = static_cast<int>(v.getInt("vertical_shift"));
    e->kFrom = v.getString("vertical_indirection");
    if(!e->kFrom.empty() && v.has("vertical_indirection_data"))
      e->kFromID = static_cast<int>(v.at("vertical_indirection_data").getInt("accessID"));
    if(v.has("cartesian_offset")) {
      e->di = static_cast<int>(v.at("cartesian_offset").getInt("i_offset"));
      e->dj = static_cast<int>(v.at("cartesian_offset").getInt("j_offset"));
    } else if(v.has("unstructured_offset")) {
      e->hasOffset = v.at("unstructured_offset").getBool("has_offset");
    }
    if(v.has("data"))
      e->accessID = static_cast<int>(v.at("data").getInt("accessID"));
  } else if(kind == "unary_operator") {
    e->kind = Expr::Unary;
    e->op = v.getString("op");
    e->kids.push_back(parseExpr(v.at("operand")));
  } else if(kind == "binary_operator") {
    e->kind = Expr::Binary;
    e->op = v.getString("op");
    e->kids.push_back(parseExpr(v.at("left")));
    e->kids.push_back(parseExpr(v.at("right")));
  } else if(kind == "assignment_expr") {
    e->kind = Expr::Assign;
    e->op = v.getString("op", "=");
    e->kids.push_back(parseExpr(v.at("left")));
    e->kids.push_back(parseExpr(v.at("right")));
  } else if(kind == "ternary_operator") {
    e->kind = Expr::Ternary;
    e->kids.push_back(parseExpr(v.at("cond")));
    e->kids.push_back(parseExpr(v.at("left")));
    e->kids.push_back(parseExpr(v.at("right")));
  } else if(kind == "fun_call_expr") {
    e->kind = Expr::FunCall;
    e->name = v.getString("callee");
    if(const Value* args = v.find("arguments"))
      for(auto const& a : args->arr)
        e->kids.push_back(parseExpr(a));
  } else if(kind == "reduction_over_neighbor_expr") {
    e->kind = Expr::Reduce;
    e->op = v.getString("op", "+");
    e->kids.push_back(parseExpr(v.at("rhs")));
    e->kids.push_back(parseExpr(v.at("init")));
    if(const Value* w = v.find("weights"))
      for(auto const& a : w->arr)
        e->kids.push_back(parseExpr(a));
    if(v.has("iter_space")) {
      auto const& is = v.at("iter_space");
      if(const Value* ch = is.find("chain"))
        for(auto const& c : ch->arr)
          e->chain.push_back(parseLocation(c.str));
      e->includeCenter = is.getBool("include_center");
    } else if(const Value* ch = v.find("chain")) { // older schema revision
      for(auto const& c : ch->arr)
        e->chain.push_back(parseLocation(c.str));
    }
  } else if(kind == "stencil_fun_call_expr" || kind == "stencil_fun_arg_expr") {
    throw std::runtime_error(
        "IIR: stencil function calls must be inlined for the b200 backend (run PassGroup::Inlining, "
        "as for --backend=cuda)");
  } else {
    throw std::runtime_error("IIR: unsupported expression node '" + kind + "'");
  }
  return e;
}

ExprPtr parseExpr(const Value& node) {
  auto const& kv = only(node, "expression");
  return parseExprPayload(kv.first, kv.second);
}

StmtPtr parseStmt(const Value& node);

void parseBlockInto(const Value& node, std::vector<StmtPtr>& out) {
  auto const& kv = only(node, "statement");
  if(kv.first == "block_stmt") {
    if(const Value* st = kv.second.find("statements"))
      for(auto const& s : st->arr)
        out.push_back(parseStmt(s));
  } else {
    out.push_back(parseStmt(node));
  }
}

StmtPtr parseStmt(const Value& node) {
  auto const& kv = only(node, "statement");
  const std::string& kind = kv.first;
  const Value& v = kv.second;
  auto s = std::make_shared<Stmt>();
  if(kind == "expr_stmt") {
    s->kind = Stmt::ExprStmt;
    s->expr = parseExpr(v.at("expr"));
  } else if(kind == "var_decl_stmt") {
    s->kind = Stmt::VarDecl;
    s->name = v.getString("name");
    if(v.getInt("dimension") != 0)
      throw std::runtime_error("IIR: array variable declarations are not supported: " + s->name);
    s->typeId = "Float";
    if(v.has("type") && v.at("type").has("builtin_type"))
      s->typeId = v.at("type").at("builtin_type").getString("type_id", "Float");
    if(v.has("var_decl_stmt_data"))
      s->accessID = static_cast<int>(v.at("var_decl_stmt_data").getInt("accessID"));
    if(const Value* il = v.find("init_list"))
      if(!il->arr.empty())
        s->expr = parseExpr(il->arr.front());
  } else if(kind == "if_stmt") {
    s->kind = Stmt::If;
    auto const& cond = only(v.at("cond_part"), "if condition");
    if(cond.first != "expr_stmt")
      throw std::runtime_error("IIR: if-condition must be an expression statement");
    s->expr = parseExpr(cond.second.at("expr"));
    parseBlockInto(v.at("then_part"), s->body);
    if(v.has("else_part") && v.at("else_part").isObject() && !v.at("else_part").obj.empty()) {
      s->hasElse = true;
      parseBlockInto(v.at("else_part"), s->orelse);
    }
  } else if(kind == "stencil_call_decl_stmt") {
    s->kind = Stmt::StencilCall;
    std::string callee = v.has("stencil_call") ? v.at("stencil_call").getString("callee") : "";
    size_t us = callee.rfind('_');
    if(callee.compare(0, 11, "__code_gen_") != 0 || us == std::string::npos)
      throw std::runtime_error("IIR: unexpected stencil call '" + callee + "'");
    s->accessID = std::stoi(callee.substr(us + 1));
  } else if(kind == "boundary_condition_decl_stmt") {
    throw std::runtime_error("IIR: boundary condition declarations are not supported by the b200 "
                             "backends (the reference's CUDA backend does not generate them either)");
  } else if(kind == "loop_stmt") {
    s->kind = Stmt::Loop;
    const Value* space = nullptr;
    if(v.has("loop_descriptor") && v.at("loop_descriptor").has("loop_descriptor_chain")) {
      auto const& d = v.at("loop_descriptor").at("loop_descriptor_chain");
      if(d.has("iter_space"))
        space = &d.at("iter_space");
      else if(d.has("chain")) // older schema revision: chain stored directly
        space = &d;
    }
    if(!space)
      throw std::runtime_error("IIR: only loops over a neighbour chain are supported (the reference "
                               "asserts \"general loop concept not implemented yet\")");
    if(const Value* ch = space->find("chain"))
      for(auto const& c : ch->arr)
        s->chain.push_back(parseLocation(c.str));
    s->includeCenter = space->getBool("include_center");
    if(s->chain.size() < 2)
      throw std::runtime_error("IIR: loop statement with a neighbour chain shorter than 2");
    parseBlockInto(v.at("statements"), s->body);
  } else if(kind == "block_stmt") {
    s->kind = Stmt::Block;
    if(const Value* st = v.find("statements"))
      for(auto const& c : st->arr)
        s->body.push_back(parseStmt(c));
  } else {
    throw std::runtime_error("IIR: unsupported statement node '" + kind + "'");
  }
  return s;
}

FieldDims parseDims(const Value& v) {
  FieldDims d;
  d.maskK = v.getInt("mask_k") != 0;
  if(v.has("cartesian_horizontal_dimension")) {
    auto const& c = v.at("cartesian_horizontal_dimension");
    d.maskI = c.getInt("mask_cart_i") != 0;
    d.maskJ = c.getInt("mask_cart_j") != 0;
  } else if(v.has("unstructured_horizontal_dimension")) {
    d.unstructured = true;
    auto const& u = v.at("unstructured_horizontal_dimension");
    if(u.has("iter_space")) {
      if(const Value* ch = u.at("iter_space").find("chain"))
        for(auto const& c : ch->arr)
          d.chain.push_back(parseLocation(c.str));
      d.includeCenter = u.at("iter_space").getBool("include_center");
    } else {
      // older schema: dense_location_type + sparse_part
      if(u.has("sparse_part") && !u.at("sparse_part").arr.empty())
        for(auto const& c : u.at("sparse_part").arr)
          d.chain.push_back(parseLocation(c.str));
      else
        d.chain.push_back(parseLocation(u.getString("dense_location_type")));
    }
  } else {
    // neither: a purely vertical field
    d.maskI = d.maskJ = false;
  }
  return d;
}

} // namespace

Instantiation parseInstantiation(const std::string& data) {
  // IIRSerializer::Format::Json (proto3 JSON text) or ::Byte (protobuf wire format), told apart by
  // the first byte: a serialized StencilInstantiation starts with tag 0x0A, never with '{'
  Value root = proto::looksLikeJson(data) ? json::parse(data) : proto::decode(data, "StencilInstantiation");
  if(proto::looksLikeJson(data))
    proto::normalizeJson(root, "StencilInstantiation");
  Instantiation inst;
  const Value& meta = root.at("metadata");
  const Value& ir = root.at("internalIR");
  inst.filename = root.getString("filename");
  inst.name = meta.getString("stencilName");
  inst.unstructured = ir.getString("gridType", "Cartesian") == "Unstructured";

  if(const Value* m = meta.find("accessIDToName"))
    for(auto const& kv : m->obj)
      inst.accessIDToName[std::stoi(kv.first)] = kv.second.str;
  if(const Value* m = meta.find("accessIDToType"))
    for(auto const& kv : m->obj)
      inst.accessIDToType[std::stoi(kv.first)] = static_cast<AccessType>(kv.second.asInt());
  if(const Value* m = meta.find("fieldIDtoDimensions"))
    for(auto const& kv : m->obj)
      inst.fieldDims[std::stoi(kv.first)] = parseDims(kv.second);
  auto readIds = [&](const char* key, std::vector<int>& out) {
    if(const Value* a = meta.find(key))
      for(auto const& x : a->arr)
        out.push_back(static_cast<int>(x.asInt()));
  };
  readIds("APIFieldIDs", inst.apiFieldIDs);
  readIds("temporaryFieldIDs", inst.temporaryFieldIDs);
  readIds("allocatedFieldIDs", inst.allocatedFieldIDs);
  readIds("globalVariableIDs", inst.globalVariableIDs);

  if(const Value* g = ir.find("globalVariableToValue"))
    for(auto const& kv : g->obj) {
      GlobalValue gv;
      gv.type = kv.second.getString("type", "Double");
      gv.value = kv.second.getDouble("value");
      if(kv.second.has("value") && kv.second.at("value").kind == Value::Bool)
        gv.value = kv.second.at("value").b ? 1 : 0;
      gv.valueIsSet = kv.second.getBool("valueIsSet");
      inst.globals.emplace_back(kv.first, gv);
    }

  // the reference keeps globals in a std::map keyed by name (AST/GlobalVariableMap); protobuf map order
  // is unspecified, so order them the same way
  std::stable_sort(inst.globals.begin(), inst.globals.end(),
                   [](auto const& a, auto const& b) { return a.first < b.first; });

  if(const Value* sts = ir.find("stencils"))
    for(auto const& sj : sts->arr) {
      Stencil st;
      st.id = static_cast<int>(sj.getInt("stencilID"));
      if(const Value* mss = sj.find("multiStages"))
        for(auto const& mj : mss->arr) {
          MultiStage ms;
          ms.id = static_cast<int>(mj.getInt("multiStageID"));
          std::string lo = mj.getString("loopOrder", "Forward");
          ms.loopOrder = lo == "Parallel"   ? LoopOrder::Parallel
                         : lo == "Backward" ? LoopOrder::Backward
                                            : LoopOrder::Forward;
          if(const Value* cs = mj.find("Caches"))
            for(auto const& kv : cs->obj) {
              Cache c;
              c.accessID = std::stoi(kv.first);
              c.type = kv.second.getString("type", "CT_IJ");
              c.policy = kv.second.getString("policy", "CP_Unknown");
              if(kv.second.has("cacheWindow")) {
                c.hasWindow = true;
                c.windowMinus = static_cast<int>(kv.second.at("cacheWindow").getInt("minus"));
                c.windowPlus = static_cast<int>(kv.second.at("cacheWindow").getInt("plus"));
              }
              ms.caches[c.accessID] = c;
            }
          if(const Value* stages = mj.find("stages"))
            for(auto const& sgj : stages->arr) {
              Stage sg;
              sg.id = static_cast<int>(sgj.getInt("stageID"));
              sg.location = parseLocation(sgj.getString("locationType"));
              if(sgj.has("i_range")) {
                sg.hasIRange = true;
                sg.iRange = parseInterval(sgj.at("i_range"));
              }
              if(sgj.has("j_range")) {
                sg.hasJRange = true;
                sg.jRange = parseInterval(sgj.at("j_range"));
              }
              if(const Value* dms = sgj.find("doMethods"))
                for(auto const& dj : dms->arr) {
                  DoMethod dm;
                  dm.id = static_cast<int>(dj.getInt("doMethodID"));
                  dm.interval = parseInterval(dj.at("interval"));
                  if(dj.has("ast"))
                    parseBlockInto(dj.at("ast"), dm.stmts);
                  sg.doMethods.push_back(std::move(dm));
                }
              ms.stages.push_back(std::move(sg));
            }
          st.multiStages.push_back(std::move(ms));
        }
      inst.stencils.push_back(std::move(st));
    }
  if(const Value* cf = ir.find("controlFlowStatements"))
    for(auto const& s : cf->arr)
      inst.controlFlow.push_back(parseStmt(s));
  if(inst.stencils.empty())
    throw std::runtime_error("IIR: instantiation '" + inst.name + "' contains no stencil");
  return inst;
}

} // namespace iir
} // namespace b200
