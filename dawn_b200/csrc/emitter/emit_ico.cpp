// B200 (sm_100a) unstructured emitter ("b200-ico"): optimized unstructured IIR -> CUDA.
//
// Replaces dawn/src/dawn/CodeGen/Cuda-ico (CudaIcoCodeGen.cpp:148-160,261-453,1316-1464 and
// Cuda-ico/ASTStencilBody.cpp:135-210,309-378).  Device layout is the reference's: dense fields
// level-major `[k][element]`, sparse fields `[k][nbh][element]`, neighbour tables SoA
// `[nbh][element]` int32 with -1 for missing neighbours.  What is new:
//  * stage fusion: consecutive stages on the same location type whose neighbour reads do not touch
//    anything written earlier in the group become ONE kernel (the five edge stages of ICON's nabla2
//    run as one kernel; the reference launches one kernel per stage, CudaIcoCodeGen.cpp:1316-1464);
//  * values written at the element itself are forwarded in registers to later statements of the
//    group, and temporaries that never leave the group are not stored at all;
//  * each thread loads its neighbour indices ONCE and sweeps `levels_per_thread` levels with them
//    (the reference re-reads the table row for every level unless LEVELS_PER_THREAD is raised);
//  * reductions are fully unrolled over the chain size (ICOChainSize twin below), accumulating in
//    neighbour-table order from `init`, so results are bit-identical to naive-ico.
#include "analysis.hpp"
#include "codegen.hpp"

#include <algorithm>
#include <cctype>
#include <functional>
#include <set>
#include <sstream>
#include <stdexcept>

namespace b200 {
namespace codegen {

// ------------------------------------------------------------------------------------------------
// Chain sizes on the regular triangular lattice (twin of dawn::ICOChainSize,
// dawn/src/dawn/CodeGen/IcoChainSizes.cpp:199-264; known answers TestCuda-IcoCodeGen.cpp:31-87).
// Elements are lattice triples (x, y, kind); one hop maps an element to its adjacent elements of
// another type through offset tables; the size is the number of distinct elements reached after
// the last hop (minus the origin when the chain returns to its own type).
// ------------------------------------------------------------------------------------------------
namespace {
struct Lat {
  int x, y, c;
  bool operator<(const Lat& o) const {
    return x != o.x ? x < o.x : (y != o.y ? y < o.y : c < o.c);
  }
};
struct Off {
  int dx, dy, c;
};
using iir::LocationType;

// adjacency of the lattice; row = kind of the source element
const std::vector<std::vector<Off>>& hopTable(LocationType from, LocationType to) {
  static const std::vector<std::vector<Off>> V2E = {
      {{0, 0, 0}, {0, 0, 2}, {-1, 0, 0}, {-1, 0, 1}, {0, -1, 1}, {0, -1, 2}}};
  static const std::vector<std::vector<Off>> V2C = {
      {{0, 0, 0}, {-1, 0, 0}, {-1, 0, 1}, {0, -1, 0}, {0, -1, 1}, {-1, -1, 1}}};
  static const std::vector<std::vector<Off>> E2V = {
      {{0, 0, 0}, {1, 0, 0}}, {{1, 0, 0}, {0, 1, 0}}, {{0, 0, 0}, {0, 1, 0}}};
  static const std::vector<std::vector<Off>> E2C = {
      {{0, 0, 0}, {0, -1, 1}}, {{0, 0, 0}, {0, 0, 1}}, {{0, 0, 0}, {-1, 0, 1}}};
  static const std::vector<std::vector<Off>> C2V = {{{0, 0, 0}, {1, 0, 0}, {0, 1, 0}},
                                                    {{1, 1, 0}, {1, 0, 0}, {0, 1, 0}}};
  static const std::vector<std::vector<Off>> C2E = {{{0, 0, 0}, {0, 0, 1}, {0, 0, 2}},
                                                    {{0, 0, 1}, {1, 0, 2}, {0, 1, 0}}};
  if(from == LocationType::Vertices && to == LocationType::Edges)
    return V2E;
  if(from == LocationType::Vertices && to == LocationType::Cells)
    return V2C;
  if(from == LocationType::Edges && to == LocationType::Vertices)
    return E2V;
  if(from == LocationType::Edges && to == LocationType::Cells)
    return E2C;
  if(from == LocationType::Cells && to == LocationType::Vertices)
    return C2V;
  if(from == LocationType::Cells && to == LocationType::Edges)
    return C2E;
  throw std::invalid_argument("invalid neighbour chain: consecutive locations must differ");
}
} // namespace

int icoChainSize(const std::vector<LocationType>& chain, bool includeCenter) {
  if(chain.size() < 2)
    throw std::invalid_argument("neighbour chain needs at least two locations");
  std::set<Lat> cur{{0, 0, 0}};
  for(size_t h = 0; h + 1 < chain.size(); ++h) {
    auto const& table = hopTable(chain[h], chain[h + 1]);
    std::set<Lat> next;
    for(auto const& e : cur) {
      auto const& row = table[table.size() == 1 ? 0 : e.c];
      for(auto const& o : row)
        next.insert({e.x + o.dx, e.y + o.dy, o.c});
    }
    cur.swap(next);
  }
  int n = static_cast<int>(cur.size());
  if(chain.front() == chain.back())
    --n;
  return n + (includeCenter ? 1 : 0);
}

namespace {
using namespace iir;
using analysis::StageInfo;

std::string cIdent(const std::string& s) {
  std::string o;
  for(char c : s)
    o += (std::isalnum(static_cast<unsigned char>(c)) || c == '_') ? c : '_';
  return o;
}
char locChar(LocationType l) { return l == LocationType::Cells ? 'c' : l == LocationType::Edges ? 'e' : 'v'; }
std::string chainTag(const std::vector<LocationType>& ch, bool center) {
  std::string s;
  for(auto l : ch)
    s += locChar(l);
  return s + (center ? "_o" : "");
}
std::string locName(LocationType l) {
  return l == LocationType::Cells ? "cells" : l == LocationType::Edges ? "edges" : "vertices";
}

// Calls fn(field access, nesting) for every field access of a stage; nesting = number of enclosing
// neighbour loops / reductions (weights and init of a reduction belong to the enclosing level).
void forEachAccess(const Stage& stage,
                   const std::function<void(const Expr&, int nesting, bool inLoop)>& fn) {
  std::function<void(const ExprPtr&, int, bool)> walk = [&](const ExprPtr& e, int nest, bool loop) {
    if(!e)
      return;
    if(e->kind == Expr::Field)
      fn(*e, nest, loop);
    if(e->kind == Expr::Reduce) {
      walk(e->kids[0], nest + 1, loop);
      for(size_t n = 1; n < e->kids.size(); ++n)
        walk(e->kids[n], nest, loop);
      return;
    }
    for(auto const& k : e->kids)
      walk(k, nest, loop);
  };
  std::function<void(const StmtPtr&, int, bool)> ws = [&](const StmtPtr& s, int nest, bool loop) {
    walk(s->expr, nest, loop);
    const bool isLoop = s->kind == Stmt::Loop;
    for(auto const& c : s->body)
      ws(c, nest + (isLoop ? 1 : 0), loop || isLoop);
    for(auto const& c : s->orelse)
      ws(c, nest, loop);
  };
  for(auto const& dm : stage.doMethods)
    for(auto const& s : dm.stmts)
      ws(s, 0, false);
}

// A stage runs as one thread per element, all its statements back to back: a statement may not gather (read at
// another element) what the stage writes - the neighbour's thread may or may not have stored yet - nor may a
// later statement overwrite what an earlier one gathered.  The reference's cuda-ico kernels have the same race
// (one kernel per stage, CudaIcoCodeGen.cpp:1316-1464) and its naive loops give a result that depends on the
// element order; Dawn's PassStageSplitter / PassFieldVersioning remove such stages before code generation.
// IIR that still has one (GenerateUnstructuredStencils.cpp:900-937 `tempField`, built without the splitter)
// is rejected.
void checkStageInternalHazards(const Stage& stage, const std::function<LocationType(int)>& fieldLoc,
                               const std::function<std::string(int)>& nameOf) {
  for(auto const& dm : stage.doMethods) {
    std::set<int> written, gathered;
    for(auto const& s : dm.stmts) {
      Stage one;
      one.location = stage.location;
      one.doMethods.emplace_back();
      one.doMethods.back().stmts.push_back(s);
      std::set<int> away, writes;
      forEachAccess(one, [&](const Expr& e, int nest, bool) {
        if(nest > 0 && (e.hasOffset || nest > 1 || fieldLoc(e.accessID) != stage.location))
          away.insert(e.accessID);
      });
      analysis::visitStmts({s}, [&](const Expr& e, bool w) {
        if(w && e.kind == Expr::Field)
          writes.insert(e.accessID);
      });
      for(int f : away)
        if(written.count(f) || writes.count(f))
          throw std::runtime_error("b200-ico: stage " + std::to_string(stage.id) + " gathers field '" + nameOf(f) +
                                   "' from other elements and writes it as well; the accesses of neighbouring "
                                   "elements cannot be ordered inside one stage (split the stage or version the field: "
                                   "PassStageSplitter / PassFieldVersioning)");
      for(int f : writes)
        if(gathered.count(f))
          throw std::runtime_error("b200-ico: stage " + std::to_string(stage.id) + " overwrites field '" + nameOf(f) +
                                   "' that an earlier statement of the same stage gathers from other elements "
                                   "(split the stage: PassStageSplitter)");
      written.insert(writes.begin(), writes.end());
      gathered.insert(away.begin(), away.end());
    }
  }
}

struct Group {
  std::vector<const Stage*> stages;
  const MultiStage* ms;
  LocationType loc;
  std::string name;
  std::set<int> fieldsRead, fieldsWritten;
  std::vector<std::pair<std::vector<LocationType>, bool>> chains; // tables needed
  std::set<std::string> topChains; // tags of chains walked from the thread's own element: their
                                   // neighbour indices are loaded once per thread
  bool usesGlobals = false;
  // unstructured iteration space of the stages (Stage::i_range holds it: levels are the subdomain
  // magic numbers of driver-includes/unstructured_domain.hpp:23, offsets index into the splitters)
  bool hasSpace = false;
  Interval space;
  // co-scheduling: two independent parallel groups that gather the same field share one launch whose
  // blocks alternate between them in proportion to their sizes, so that what one fetches the other
  // finds in L2 (ICON nabla2: `vec` is gathered by the vertex and by the cell stage)
  int coWith = -1;      // index of the group launched together with this one
  bool coMember = false; // launched by an earlier group
  // chaining (ico_chain): the launch of this group (and its co-scheduled partner) also runs the NEXT group,
  // which gathers through a chain what this launch writes: consumer blocks follow their producers in a
  // host-built block order and read the handed-over fields from L2
  int chainNext = -1;       // index of the consumer group run by this launch
  bool chainMember = false; // run by the launch of an earlier group
};

class IcoEmitter {
  const Instantiation& inst_;
  const Options& opt_;
  std::ostringstream os_;
  std::set<int> registerOnly_; // temporaries that never leave their kernel group
  std::set<std::pair<const Group*, int>> laterReads_; // (launching group, field) read by a later group of the stencil
  std::set<int> scratch_;      // temporaries that need storage
  int block_, lpt_;

public:
  IcoEmitter(const Instantiation& i, const Options& o) : inst_(i), opt_(o) {
    block_ = o.BlockSizeIco > 0 ? o.BlockSizeIco : 128;
    lpt_ = o.LevelsPerThread > 0 ? o.LevelsPerThread : 8;
  }
  std::string generate();

private:
  std::string fname(int id) const { return cIdent(inst_.nameOf(id)); }
  const FieldDims& dims(int id) const { return inst_.fieldDims.at(id); }
  LocationType fieldLoc(int id) const {
    auto const& d = dims(id);
    return d.chain.empty() ? LocationType::None : d.chain.front();
  }
  bool isSparse(int id) const { return dims(id).isSparse(); }
  int sparseSize(int id) const { return icoChainSize(dims(id).chain, dims(id).includeCenter); }
  bool isVertical(int id) const { return dims(id).chain.empty(); }
  bool hasK(int id) const { return dims(id).maskK; }
  // does `stage` read, at an element other than its own, any field in `written`?  (has_offset reads,
  // fields living on another location type, and everything inside a nested reduction)
  bool readsThroughChain(const Stage& stage, const std::set<int>& written) const {
    bool hit = false;
    forEachAccess(stage, [&](const Expr& e, int nest, bool) {
      if(!written.count(e.accessID) || nest == 0)
        return;
      if(e.hasOffset || nest > 1 || fieldLoc(e.accessID) != stage.location)
        hit = true;
    });
    return hit;
  }
  // fields `stage` touches at an element other than the thread's own (same rule as above)
  void fieldsTouchedAway(const Stage& stage, std::set<int>& out) const {
    forEachAccess(stage, [&](const Expr& e, int nest, bool) {
      if(nest > 0 && (e.hasOffset || nest > 1 || fieldLoc(e.accessID) != stage.location))
        out.insert(e.accessID);
    });
  }

  std::vector<Group> plan(const Stencil& st);
  // how one role of a launch is emitted
  struct RoleOpts {
    std::string levelBlock = "blockIdx.y"; // index of the block of levels
    bool noReturn = false;                 // threads past the element range skip the body instead of returning
    std::set<int> cg;                      // fields another role of the launch wrote: ld.global.cg
    std::set<int> streamOnce;              // read-only fields every role reads at its own element only: evict-first
    std::set<int> storeStream;             // written fields nobody in the launch reads again: st.global.cs
    std::set<int> plainLoad;               // fields another role wrote, complete line-wise: ordinary (L1-cached) loads
    std::set<int> stage;                   // ico_stage: fields no role of the launch writes (candidates for staged rows)
  };
  // what the staged variant of the kernel being emitted needs (emitKernel -> emitHost)
  struct StageMeta {
    size_t smemBytes = 0;
    std::set<int> fields;
  };
  StageMeta curStage_;
  std::map<std::string, StageMeta> stageMeta_; // by kernel name
  bool emitRoleStaged(const Group& g, const std::string& blockExpr, const std::string& lo, const std::string& hi,
                      const RoleOpts& ro);
  void emitRolePlain(const Group& g, const std::string& blockExpr, const std::string& lo, const std::string& hi,
                     const RoleOpts& ro);
  void emitKernel(const Group& g, const Group* partner, const Group* consumer);
  void emitRole(const Group& g, const std::string& blockExpr, const std::string& lo,
                const std::string& hi, const RoleOpts& ro);
  void planCoScheduling(std::vector<Group>& groups) const;
  void planChains(std::vector<Group>& groups) const;
  // tables through which consumer `c` gathers fields written by producer role `p` (top-level chains only);
  // false if it reaches them in a way the chain plan does not cover
  bool chainTables(const Group& c, const Group& p, std::vector<std::pair<std::vector<LocationType>, bool>>& out) const;
  void emitHost(const std::vector<Group>& groups);
  void collect(Group& g);
};

void IcoEmitter::collect(Group& g) {
  std::set<std::string> seen;
  // nested = inside the rhs of a reduction or the body of a neighbour loop
  std::function<void(const ExprPtr&, bool)> walk = [&](const ExprPtr& e, bool nested) {
    if(!e)
      return;
    if(e->kind == Expr::Var && e->isExternal)
      g.usesGlobals = true;
    if(e->kind == Expr::Reduce) {
      std::string t = chainTag(e->chain, e->includeCenter);
      if(seen.insert(t).second)
        g.chains.push_back({e->chain, e->includeCenter});
      if(!nested)
        g.topChains.insert(t);
      walk(e->kids[0], true);
      for(size_t n = 1; n < e->kids.size(); ++n)
        walk(e->kids[n], nested);
      return;
    }
    for(auto const& k : e->kids)
      walk(k, nested);
  };
  for(auto const* st : g.stages)
    for(auto const& dm : st->doMethods) {
      analysis::visitStmts(dm.stmts, [&](const Expr& e, bool w) {
        if(e.kind != Expr::Field)
          return;
        (w ? g.fieldsWritten : g.fieldsRead).insert(e.accessID);
      });
      std::function<void(const StmtPtr&, bool)> ws = [&](const StmtPtr& s, bool nested) {
        walk(s->expr, nested);
        const bool isLoop = s->kind == Stmt::Loop;
        if(isLoop) {
          std::string t = chainTag(s->chain, s->includeCenter);
          if(seen.insert(t).second)
            g.chains.push_back({s->chain, s->includeCenter});
          g.topChains.insert(t);
        }
        for(auto const& c : s->body)
          ws(c, nested || isLoop);
        for(auto const& c : s->orelse)
          ws(c, nested);
      };
      for(auto const& s : dm.stmts)
        ws(s, false);
    }
}

std::vector<Group> IcoEmitter::plan(const Stencil& st) {
  std::vector<Group> groups;
  int n = 0;
  for(auto const& ms : st.multiStages) {
    analysis::checkParallelLevels(inst_, ms);
    Group cur;
    std::set<int> written, readAway;
    auto flush = [&]() {
      if(cur.stages.empty())
        return;
      cur.name = cIdent(inst_.name) + "_stencil" + std::to_string(st.id) + "_k" + std::to_string(n++) +
                 "_kernel";
      collect(cur);
      groups.push_back(cur);
      cur = Group{};
      written.clear();
      readAway.clear();
    };
    for(auto const& stage : ms.stages) {
      if(stage.location == LocationType::None)
        throw std::runtime_error("b200-ico: stage " + std::to_string(stage.id) +
                                 " has no location type");
      checkStageInternalHazards(stage, [this](int id) { return fieldLoc(id); },
                                [this](int id) { return inst_.nameOf(id); });
      bool sameSpace = cur.hasSpace == stage.hasIRange &&
                       (!stage.hasIRange ||
                        (cur.space.lowerLevel == stage.iRange.lowerLevel &&
                         cur.space.upperLevel == stage.iRange.upperLevel &&
                         cur.space.lowerOffset == stage.iRange.lowerOffset &&
                         cur.space.upperOffset == stage.iRange.upperOffset));
      StageInfo info = analysis::stageInfo(inst_, stage);
      // write-after-read: what the group gathered from other elements must not be overwritten inside
      // the same launch (the reference runs one kernel per stage, CudaIcoCodeGen.cpp:1316-1464, so the
      // gathers of stage A are complete before stage B stores)
      bool overwritesGathered = false;
      for(auto const& fp : info.fields)
        overwritesGathered = overwritesGathered || (fp.second.written && readAway.count(fp.first));
      bool fits = !cur.stages.empty() && cur.loc == stage.location && sameSpace &&
                  !readsThroughChain(stage, written) && !overwritesGathered && opt_.FuseColumns != 0;
      if(!fits)
        flush();
      cur.ms = &ms;
      cur.loc = stage.location;
      cur.hasSpace = stage.hasIRange;
      cur.space = stage.iRange;
      cur.stages.push_back(&stage);
      fieldsTouchedAway(stage, readAway);
      for(auto const& fp : info.fields)
        if(fp.second.written)
          written.insert(fp.first);
    }
    flush();
  }
  // temporaries: register-only if every access happens at the element itself, without vertical
  // offset, inside ONE group, and the first access there is a write
  for(int t : inst_.temporaryFieldIDs) {
    int users = 0;
    bool ok = true;
    for(auto const& g : groups) {
      bool used = g.fieldsRead.count(t) || g.fieldsWritten.count(t);
      if(!used)
        continue;
      ++users;
      bool firstIsWrite = false, decided = false;
      for(auto const* stg : g.stages)
        for(auto const& dm : stg->doMethods)
          analysis::visitStmts(dm.stmts, [&](const Expr& e, bool w) {
            if(e.kind != Expr::Field || e.accessID != t)
              return;
            if(e.hasOffset || e.dk != 0 || !e.kFrom.empty())
              ok = false;
            if(!decided) {
              // visitStmts reports the write of `t = rhs` before the reads inside rhs
              firstIsWrite = w;
              decided = true;
            }
          });
      if(!firstIsWrite)
        ok = false;
      for(auto const* stg : g.stages)
        forEachAccess(*stg, [&](const Expr& e, int nest, bool inLoop) {
          if(e.kFromID == t && !e.kFrom.empty())
            ok = false; // used as a level index
          if(e.accessID == t && (inLoop || nest > 1 || (nest == 1 && fieldLoc(t) != stg->location)))
            ok = false; // touched inside a neighbour loop or away from the thread's own element
        });
      if(g.ms->loopOrder != LoopOrder::Parallel)
        ok = false;
      // conditional writes cannot be forwarded
      for(auto const* stg : g.stages)
        for(auto const& dm : stg->doMethods)
          for(auto const& s : dm.stmts)
            if(s->kind == Stmt::If)
              ok = false;
    }
    if(users == 1 && ok)
      registerOnly_.insert(t);
    else if(users > 0)
      scratch_.insert(t);
  }
  return groups;
}

// ------------------------------------------------------------------------------------------------
class IcoBody {
public:
  const Instantiation& inst;
  const Group& g;
  const std::set<int>& registerOnly;
  std::function<std::string(int)> fname;
  std::function<const FieldDims&(int)> dims;
  std::function<int(int)> sparseSize;
  std::map<int, std::string> forwarded; // field -> variable with its current value at (pidx, k)
  int uid = 0;
  // chained launches (IcoEmitter::RoleOpts): cache policy per field
  std::set<int> cg, streamOnce, storeStream, plainLoad;
  // ico_stage: own-element reads of these fields come from the block's staged rows.  A row is one (field, sparse
  // position) pair, `rows` lists them in slot order; a slot holds rowElems elements per level, the first
  // `alignElems - 1` of which may be lead-in (the copy starts on a 16-byte boundary)
  const std::set<int>* stage = nullptr;
  std::vector<std::pair<int, int>>* rows = nullptr;
  int rowElems = 0, alignElems = 1;

  struct Ctx {
    std::string center = "pidx"; // index of the element the reduction runs on
    std::string nbh;             // neighbour index variable ("" outside reductions)
    std::string sparseIdx;       // position in the neighbour list
    LocationType centerLoc;
  };

  static std::string stride(LocationType l) {
    return l == LocationType::Cells ? "n_cells" : l == LocationType::Edges ? "n_edges" : "n_vertices";
  }
  std::string literal(const Expr& e) const {
    if(e.typeId == "Integer")
      return "(int)" + e.value;
    if(e.typeId == "Boolean")
      return (e.value == "true" || e.value == "1") ? "true" : "false";
    std::string v = e.value;
    if(v.find_first_of(".eE") == std::string::npos && v.find("inf") == std::string::npos &&
       v.find("nan") == std::string::npos)
      v += ".0";
    return "(::dawn::float_type)" + v;
  }
  // level expression of an access: k + shift, or (int)index_field(element, k) + shift for a vertical
  // indirection (reference: Cuda-ico/ASTStencilBody.cpp:193-210)
  std::string kexpr(const Expr& e, const Ctx& c) {
    if(e.kFrom.empty())
      return e.dk ? "(k + (" + std::to_string(e.dk) + "))" : "k";
    Expr idx;
    idx.kind = Expr::Field;
    idx.name = e.kFrom;
    idx.accessID = e.kFromID;
    return "((int)(" + read(idx, c) + ") + (" + std::to_string(e.dk) + "))";
  }

  // Which element does a dense access address?  Outside neighbour loops / reductions the thread's own
  // element.  Inside: the neighbour if the access carries an offset or the field lives on another
  // location type than the element the loop runs around (the reference's naive-ico backend resolves
  // un-flagged accesses by location type, CXXNaive-ico/ASTStencilBody.cpp:237-252; its cuda-ico
  // backend by the flag alone, Cuda-ico/ASTStencilBody.cpp:153-163 - both agree on valid programs).
  bool atNeighbour(const Expr& e, const Ctx& c) const {
    if(c.nbh.empty()) {
      if(e.hasOffset)
        throw std::runtime_error("b200-ico: neighbour access to '" + e.name +
                                 "' outside a reduction or neighbour loop");
      return false;
    }
    auto const& d = dims(e.accessID);
    return e.hasOffset || (!d.chain.empty() && d.chain.front() != c.centerLoc);
  }

  // global element index of the first element of a staged row at level k: [k][element] or [k][nbh][element]
  std::string stagedIndex(const std::pair<int, int>& row) {
    auto const& d = dims(row.first);
    const std::string n = stride(d.chain.front());
    if(row.second < 0)
      return "(size_t)k * " + n + " + blk_lo";
    return "((size_t)k * " + std::to_string(sparseSize(row.first)) + " + " + std::to_string(row.second) + ") * " + n +
           " + blk_lo";
  }

  std::string fieldRef(const Expr& e, const Ctx& c) {
    auto const& d = dims(e.accessID);
    std::string p = fname(e.accessID) + "_p";
    if(d.chain.empty())
      return p + "[" + kexpr(e, c) + "]";
    LocationType loc = d.chain.front();
    // horizontal-only fields (mask_k = 0) have no level term
    if(d.isSparse()) {
      if(c.sparseIdx.empty())
        throw std::runtime_error("b200-ico: sparse field '" + e.name +
                                 "' used outside a reduction or neighbour loop");
      if(!d.maskK)
        return p + "[(size_t)" + c.sparseIdx + " * " + stride(loc) + " + " + c.center + "]";
      return p + "[((size_t)" + kexpr(e, c) + " * " + std::to_string(sparseSize(e.accessID)) + " + " +
             c.sparseIdx + ") * " + stride(loc) + " + " + c.center + "]";
    }
    const std::string elem = atNeighbour(e, c) ? c.nbh : c.center;
    if(!d.maskK)
      return p + "[" + elem + "]";
    return p + "[(size_t)" + kexpr(e, c) + " * " + stride(loc) + " + " + elem + "]";
  }

  std::string read(const Expr& e, const Ctx& c) {
    const bool own = c.center == "pidx" && !dims(e.accessID).isSparse() &&
                     (dims(e.accessID).chain.empty() || !atNeighbour(e, c));
    if(own && e.dk == 0 && e.kFrom.empty()) {
      auto it = forwarded.find(e.accessID);
      if(it != forwarded.end())
        return it->second;
    }
    if(registerOnly.count(e.accessID))
      throw std::runtime_error("b200-ico: internal error, register-only temporary read from memory");
    if(stage && stage->count(e.accessID) && c.center == "pidx" && e.dk == 0 && e.kFrom.empty()) {
      auto const& d = dims(e.accessID);
      if(d.maskK && !d.chain.empty() && (d.isSparse() ? !c.sparseIdx.empty() : !atNeighbour(e, c))) {
        const std::pair<int, int> row{e.accessID, d.isSparse() ? std::stoi(c.sparseIdx) : -1};
        size_t slot = 0;
        while(slot < rows->size() && (*rows)[slot] != row)
          ++slot;
        if(slot == rows->size())
          rows->push_back(row);
        return "stg[(kk * @ROWS@ + " + std::to_string(slot) + ") * " + std::to_string(rowElems) + " + (unsigned)((" +
               stagedIndex(row) + ") % " + std::to_string(alignElems) + ") + threadIdx.x]";
      }
    }
    std::string ref = fieldRef(e, c);
    if(cg.count(e.accessID))
      return "__ldcg(&" + ref + ")"; // written by another role of this launch: L2 is the point of coherence
    if(plainLoad.count(e.accessID)) {
      // same, but every line the block touches is complete, so L1 may keep it - except the lines at either end of
      // a level, which a level of ANOTHER block of levels shares: those elements are always read through L2
      auto const& d = dims(e.accessID);
      if(d.isSparse() || d.chain.empty() || !d.maskK)
        return "__ldcg(&" + ref + ")";
      const std::string elem = atNeighbour(e, c) ? c.nbh : c.center;
      return "::dawn_b200::chain::load(&" + fname(e.accessID) + "_p[(size_t)" + kexpr(e, c) + " * " +
             stride(d.chain.front()) + "], " + elem + ", " + stride(d.chain.front()) + ")";
    }
    if(!g.fieldsWritten.count(e.accessID))
      return (streamOnce.count(e.accessID) ? "__ldcs(&" : "__ldg(&") + ref + ")";
    return ref;
  }

  // emits the reduction as statements into `out`, returns the accumulator variable
  std::string reduce(const Expr& e, const Ctx& outer, std::vector<std::string>& out, int indent) {
    std::string pad(indent * 2, ' ');
    int id = uid++;
    std::string acc = "red" + std::to_string(id);
    int size = icoChainSize(e.chain, e.includeCenter);
    std::string tag = chainTag(e.chain, e.includeCenter);
    bool top = outer.center == "pidx";
    out.push_back(pad + "::dawn::float_type " + acc + " = " + expr(e.kids[1], outer, out, indent) + ";");
    bool weighted = e.kids.size() > 2;
    // the reference's own fixtures pass longer weight lists (reference_gradient.hpp: 4 weights for
    // the 3 edges of a cell); only the first `size` are ever indexed
    if(weighted && (int)e.kids.size() - 2 < size)
      throw std::runtime_error("b200-ico: reduction over " + tag + " needs " + std::to_string(size) +
                               " weights");
    for(int n = 0; n < size; ++n) {
      Ctx c;
      c.center = outer.center;
      c.centerLoc = e.chain.front();
      c.sparseIdx = std::to_string(n);
      std::string nb;
      if(top)
        nb = "nb_" + tag + "_" + std::to_string(n);
      else {
        nb = "nbi" + std::to_string(id) + "_" + std::to_string(n);
        out.push_back(pad + "const int " + nb + " = __ldg(&tbl_" + tag + "[" + outer.center + " + " +
                      stride(e.chain.front()) + " * " + std::to_string(n) + "]);");
      }
      c.nbh = nb;
      out.push_back(pad + "if(" + nb + " >= 0) {");
      // the rhs of a nested reduction iterates around the neighbour
      Ctx inner = c;
      std::vector<std::string> body;
      std::string rhs = exprIn(e.kids[0], inner, body, indent + 1);
      for(auto const& l : body)
        out.push_back(l);
      if(weighted)
        rhs = "(" + expr(e.kids[2 + n], outer, out, indent + 1) + " * " + rhs + ")";
      if(e.op == "min" || e.op == "max")
        out.push_back(pad + "  " + acc + " = ::dawn_b200::math::" + e.op + "(" + acc + ", " + rhs + ");");
      else
        out.push_back(pad + "  " + acc + " " + e.op + "= " + rhs + ";");
      out.push_back(pad + "}");
    }
    return acc;
  }

  // expression inside a reduction body: nested reductions run around the neighbour element
  std::string exprIn(const ExprPtr& e, const Ctx& c, std::vector<std::string>& out, int indent) {
    if(e->kind == Expr::Reduce) {
      if(c.nbh.empty()) // not inside a neighbour loop / reduction: around the context's own element
        return reduce(*e, c, out, indent);
      Ctx around;
      around.center = c.nbh;
      around.centerLoc = e->chain.front();
      return reduce(*e, around, out, indent);
    }
    return expr(e, c, out, indent);
  }

  std::string expr(const ExprPtr& e, const Ctx& c, std::vector<std::string>& out, int indent) {
    switch(e->kind) {
    case Expr::Literal: return literal(*e);
    case Expr::Var: return e->isExternal ? "g." + cIdent(e->name) : cIdent(e->name);
    case Expr::Field: return read(*e, c);
    case Expr::Unary: return "(" + e->op + expr(e->kids[0], c, out, indent) + ")";
    case Expr::Binary: {
      std::string l = exprIn(e->kids[0], c, out, indent);
      std::string r = exprIn(e->kids[1], c, out, indent);
      return "(" + l + " " + e->op + " " + r + ")";
    }
    case Expr::Ternary:
      return "(" + expr(e->kids[0], c, out, indent) + " ? " + expr(e->kids[1], c, out, indent) +
             " : " + expr(e->kids[2], c, out, indent) + ")";
    case Expr::FunCall: {
      size_t p = e->name.rfind("::");
      std::string s = "::dawn_b200::math::" + (p == std::string::npos ? e->name : e->name.substr(p + 2)) + "(";
      for(size_t n = 0; n < e->kids.size(); ++n)
        s += (n ? ", " : "") + expr(e->kids[n], c, out, indent);
      return s + ")";
    }
    case Expr::Reduce: return exprIn(e, c, out, indent);
    case Expr::Assign: throw std::runtime_error("b200-ico: nested assignment expressions are not supported");
    }
    return "";
  }

  void stmt(const StmtPtr& s, std::vector<std::string>& out, int indent, bool conditional) {
    Ctx top;
    top.centerLoc = g.loc;
    stmt(s, out, indent, conditional, top);
  }

  void stmt(const StmtPtr& s, std::vector<std::string>& out, int indent, bool conditional,
            const Ctx& top) {
    std::string pad(indent * 2, ' ');
    switch(s->kind) {
    case Stmt::ExprStmt: {
      if(s->expr->kind != Expr::Assign) {
        out.push_back(pad + expr(s->expr, top, out, indent) + ";");
        break;
      }
      const Expr& lhs = *s->expr->kids[0];
      std::string rhs = expr(s->expr->kids[1], top, out, indent);
      if(lhs.kind == Expr::Var) {
        out.push_back(pad + cIdent(lhs.name) + " " + s->expr->op + " " + rhs + ";");
        break;
      }
      if(lhs.hasOffset || lhs.dk != 0 || !lhs.kFrom.empty())
        throw std::runtime_error("b200-ico: writes with offsets are not supported (" + lhs.name + ")");
      if(!top.nbh.empty() && !dims(lhs.accessID).isSparse() && atNeighbour(lhs, top))
        throw std::runtime_error("b200-ico: a neighbour loop may only write its own element (" +
                                 lhs.name + ")");
      std::string val = rhs;
      if(s->expr->op != "=")
        val = "(" + read(lhs, top) + " " + s->expr->op.substr(0, 1) + " " + rhs + ")";
      std::string var = fname(lhs.accessID) + "_v" + std::to_string(uid++);
      out.push_back(pad + "const ::dawn::float_type " + var + " = " + val + ";");
      if(!registerOnly.count(lhs.accessID)) {
        if(storeStream.count(lhs.accessID))
          out.push_back(pad + "__stcs(&" + fieldRef(lhs, top) + ", " + var + ");");
        else
          out.push_back(pad + fieldRef(lhs, top) + " = " + var + ";");
      }
      if(!conditional)
        forwarded[lhs.accessID] = var;
      else
        forwarded.erase(lhs.accessID);
      break;
    }
    case Stmt::VarDecl: {
      std::string t = s->typeId == "Integer" ? "int" : s->typeId == "Boolean" ? "bool" : "::dawn::float_type";
      std::string init = s->expr ? " = " + expr(s->expr, top, out, indent) : "";
      out.push_back(pad + t + " " + cIdent(s->name) + init + ";");
      break;
    }
    case Stmt::If: {
      std::string cond = expr(s->expr, top, out, indent);
      out.push_back(pad + "if(" + cond + ") {");
      for(auto const& c : s->body)
        stmt(c, out, indent + 1, true, top);
      if(s->hasElse) {
        out.push_back(pad + "} else {");
        for(auto const& c : s->orelse)
          stmt(c, out, indent + 1, true, top);
      }
      out.push_back(pad + "}");
      break;
    }
    case Stmt::Block:
      out.push_back(pad + "{");
      for(auto const& c : s->body)
        stmt(c, out, indent + 1, conditional, top);
      out.push_back(pad + "}");
      break;
    case Stmt::Loop: {
      // neighbour loop (reference: Cuda-ico/ASTStencilBody.cpp:52-74), unrolled over the chain size;
      // position n in the neighbour list is the sparse index of the iteration
      if(!top.nbh.empty())
        throw std::runtime_error("b200-ico: nested neighbour loops are not supported");
      const int size = icoChainSize(s->chain, s->includeCenter);
      const std::string tag = chainTag(s->chain, s->includeCenter);
      if(s->chain.front() != top.centerLoc)
        throw std::runtime_error("b200-ico: neighbour loop over " + tag + " in a stage on " +
                                 std::string(1, locChar(top.centerLoc)));
      for(int n = 0; n < size; ++n) {
        Ctx c = top;
        c.nbh = "nb_" + tag + "_" + std::to_string(n);
        c.sparseIdx = std::to_string(n);
        out.push_back(pad + "if(" + c.nbh + " >= 0) {");
        for(auto const& b : s->body)
          stmt(b, out, indent + 1, true, c);
        out.push_back(pad + "}");
      }
      for(auto const& b : s->body) // values written in the loop are not forwarded past it
        analysis::visitStmts({b}, [&](const Expr& e, bool w) {
          if(w && e.kind == Expr::Field)
            forwarded.erase(e.accessID);
        });
      break;
    }
    }
  }
};

void IcoEmitter::planCoScheduling(std::vector<Group>& groups) const {
  if(!opt_.CoSchedule)
    return;
  auto disjoint = [](const std::set<int>& a, const std::set<int>& b) {
    for(int x : a)
      if(b.count(x))
        return false;
    return true;
  };
  for(size_t i = 0; i + 1 < groups.size(); ++i) {
    Group& a = groups[i];
    Group& b = groups[i + 1];
    if(a.coMember || a.coWith >= 0 || a.ms != b.ms || a.ms->loopOrder != LoopOrder::Parallel ||
       a.hasSpace || b.hasSpace || a.usesGlobals != b.usesGlobals)
      continue;
    // independent: neither touches what the other writes
    if(!disjoint(a.fieldsWritten, b.fieldsRead) || !disjoint(a.fieldsWritten, b.fieldsWritten) ||
       !disjoint(b.fieldsWritten, a.fieldsRead))
      continue;
    // worth it: both gather a common field through a chain
    std::set<int> ga, gb;
    for(auto const* st : a.stages)
      forEachAccess(*st, [&](const Expr& e, int nest, bool) {
        if(nest > 0 && (e.hasOffset || fieldLoc(e.accessID) != a.loc))
          ga.insert(e.accessID);
      });
    for(auto const* st : b.stages)
      forEachAccess(*st, [&](const Expr& e, int nest, bool) {
        if(nest > 0 && (e.hasOffset || fieldLoc(e.accessID) != b.loc))
          gb.insert(e.accessID);
      });
    if(disjoint(ga, gb))
      continue;
    a.coWith = static_cast<int>(i + 1);
    b.coMember = true;
  }
}

// ico_hoist: the loads of read-only fields of one level (`__ldg` at the thread's own element or at a neighbour
// whose index the thread loaded once) move to the top of the level, each into a constant, so that they are all
// in flight together.  Left where they are used they are not: a double-precision division ends a basic block
// (slow-path branch) and the compiler does not move loads across it, so ICON's edge kernel has two gathers in
// flight per thread at a time.  Safe: the fields are not written by the launch, the addresses are in range for
// every level of the block (accesses with a vertical offset stay where they are) and a missing neighbour (-1)
// keeps its guard.  Identical loads share one constant.
static std::string hoistLevelLoads(const std::string& body, const std::string& ft) {
  std::string out, head;
  std::map<std::string, std::string> seen;
  size_t pos = 0;
  for(;;) {
    size_t b = body.find("__ldg(&", pos);
    if(b == std::string::npos)
      break;
    size_t lb = body.find('[', b), rb = lb == std::string::npos ? lb : body.find(']', lb);
    if(rb == std::string::npos || rb + 1 >= body.size() || body[rb + 1] != ')') {
      out += body.substr(pos, b + 7 - pos);
      pos = b + 7;
      continue;
    }
    const std::string load = body.substr(b, rb + 2 - b), idx = body.substr(lb + 1, rb - lb - 1);
    const bool plainIndex = idx.find('[') == std::string::npos && idx.find("nbi") == std::string::npos &&
                            idx.find("(k + (") == std::string::npos && idx.find("__ld") == std::string::npos &&
                            idx.find("(int)") == std::string::npos;
    if(!plainIndex) {
      out += body.substr(pos, b + 7 - pos);
      pos = b + 7;
      continue;
    }
    auto it = seen.find(load);
    if(it == seen.end()) {
      const std::string var = "hl" + std::to_string(seen.size());
      std::string guard;
      size_t nb = idx.find("nb_");
      if(nb != std::string::npos) {
        size_t e = nb;
        while(e < idx.size() && (std::isalnum((unsigned char)idx[e]) || idx[e] == '_'))
          ++e;
        guard = idx.substr(nb, e - nb) + " >= 0 ? ";
      }
      head += "    const " + ft + " " + var + " = " + guard + load + (guard.empty() ? "" : " : (" + ft + ")0") + ";\n";
      it = seen.emplace(load, var).first;
    }
    out += body.substr(pos, b - pos) + it->second;
    pos = rb + 2;
  }
  out += body.substr(pos);
  return head + out;
}

// One role of a launch: the code of a group for the element block `blockExpr` of [lo, hi).
void IcoEmitter::emitRole(const Group& g, const std::string& blockExpr, const std::string& lo,
                          const std::string& hi, const RoleOpts& ro) {
  if(!ro.stage.empty()) {
    // the staged variant is rendered first: only the body knows which rows it reads at its own element
    const std::string saved = os_.str();
    os_.str("");
    const bool staged = emitRoleStaged(g, blockExpr, lo, hi, ro);
    const std::string text = os_.str();
    os_.str("");
    os_ << saved;
    if(staged) {
      os_ << text << "  } else {\n";
      emitRolePlain(g, blockExpr, lo, hi, ro);
      os_ << "  }\n";
      return;
    }
  }
  emitRolePlain(g, blockExpr, lo, hi, ro);
}

// ico_stage: the variant of a role for FULL element blocks.  The rows the body reads at the thread's own element
// (dense fields of the role's location, sparse weights [k][nbh][element]) are fetched by 1-D bulk copies
// (cp.async.bulk -> shared memory, completion on an mbarrier) that the first lanes of the block issue for all
// levels of the block at once; the threads load their neighbour indices meanwhile and then only gather.  A row
// starts at an arbitrary element index, so a copy starts on the 16-byte boundary at or below it (`alignElems`
// elements) and the readers add the lead-in; the block must therefore lie alignElems elements inside its level
// row at the upper end (checked per block, as is the alignment of the field bases: `stage_ok` from the host).
// Returns false (nothing emitted) if the body has no such row.
bool IcoEmitter::emitRoleStaged(const Group& g, const std::string& blockExpr, const std::string& lo,
                                const std::string& hi, const RoleOpts& ro) {
  const bool sequential = g.ms->loopOrder != LoopOrder::Parallel;
  if(sequential || ro.noReturn || lpt_ > 16 || opt_.IcoFullBlocks)
    return false;
  const int ftBytes = opt_.SinglePrecision ? 4 : 8;
  const int E = 16 / ftBytes;
  const int rowElems = block_ + E;
  if(block_ % E)
    return false;
  int lo_k = Interval::End * 2, hi_k = -Interval::End;
  for(auto const* st : g.stages)
    for(auto const& dm : st->doMethods) {
      lo_k = std::min(lo_k, dm.interval.lowerBound());
      hi_k = std::max(hi_k, dm.interval.upperBound());
    }
  std::vector<std::pair<int, int>> rows;
  IcoBody body{inst_, g, registerOnly_,
               [this](int id) { return fname(id); },
               [this](int id) -> const FieldDims& { return dims(id); },
               [this](int id) { return sparseSize(id); }, {}, 0, ro.cg, ro.streamOnce, ro.storeStream, ro.plainLoad};
  body.stage = &ro.stage;
  body.rows = &rows;
  body.rowElems = rowElems;
  body.alignElems = E;
  std::ostringstream bodyText;
  for(auto const* st : g.stages) {
    bodyText << "    // stage " << st->id << "\n";
    for(auto const& dm : st->doMethods) {
      bool guard = dm.interval.lowerBound() != lo_k || dm.interval.upperBound() != hi_k;
      std::vector<std::string> lines;
      for(auto const& s : dm.stmts)
        body.stmt(s, lines, 0, guard);
      if(guard)
        bodyText << "    if(k >= " << analysis::boundExpr(dm.interval.lowerBound()) << " && k <= "
                 << analysis::boundExpr(dm.interval.upperBound()) << ") {\n";
      for(auto const& l : lines)
        bodyText << (guard ? "      " : "    ") << l << "\n";
      if(guard) {
        bodyText << "    }\n";
        body.forwarded.clear();
      }
    }
  }
  if(rows.empty())
    return false;
  const size_t smem = 128 + (size_t)lpt_ * rows.size() * rowElems * ftBytes;
  if(smem > 200 * 1024)
    return false;
  std::string text = bodyText.str();
  for(size_t p = text.find("@ROWS@"); p != std::string::npos; p = text.find("@ROWS@", p))
    text.replace(p, 6, std::to_string(rows.size()));
  if(opt_.IcoHoist)
    text = hoistLevelLoads(text, "::dawn::float_type");
  curStage_.smemBytes = std::max(curStage_.smemBytes, smem);
  for(auto const& r : rows)
    curStage_.fields.insert(r.first);
  const bool oneBarrier = opt_.IcoStage == 2;
  const std::string n = IcoBody::stride(g.loc);
  os_ << "  const int blk_lo = " << lo << " + (int)(" << blockExpr << ") * " << block_ << ";\n";
  os_ << "  if(stage_ok && blk_lo + " << block_ << " <= " << hi << " && blk_lo + " << block_ + E << " <= " << n
      << ") { // full block: " << rows.size() << " row(s) per level arrive as bulk copies\n";
  os_ << "  const int pidx = blk_lo + threadIdx.x;\n";
  os_ << "  unsigned long long* const stg_bar = reinterpret_cast<unsigned long long*>(b200_smem);\n";
  os_ << "  ::dawn::float_type* const stg = reinterpret_cast<::dawn::float_type*>(b200_smem + 128);\n";
  os_ << "  const int kmin = 0, kmax = k_size - 1;\n";
  os_ << "  int kb = " << analysis::boundExpr(lo_k) << ", ke = " << analysis::boundExpr(hi_k) << ";\n";
  os_ << "  kb = kb < 0 ? 0 : kb; ke = ke > kmax ? kmax : ke;\n";
  os_ << "  const int k0 = kb + " << ro.levelBlock << " * " << lpt_ << ";\n";
  os_ << "  const int nlev = ke - k0 + 1 < " << lpt_ << " ? ke - k0 + 1 : " << lpt_ << "; // levels of this block\n";
  if(oneBarrier)
    os_ << "  if(threadIdx.x == 0 && nlev > 0) {\n    ::dawn_b200::tma::barrier_init(&stg_bar[0], (unsigned)nlev);\n";
  else
    os_ << "  if(threadIdx.x < " << lpt_ << ") {\n    ::dawn_b200::tma::barrier_init(&stg_bar[threadIdx.x], 1);\n";
  os_ << "    ::dawn_b200::tma::fence_barrier_init();\n  }\n  __syncthreads();\n";
  os_ << "  if((int)threadIdx.x < nlev) { // lane kk requests the rows of level k0 + kk\n";
  os_ << "    const int kk = threadIdx.x;\n    const int k = k0 + kk;\n";
  os_ << "    unsigned long long* const bar = &stg_bar[" << (oneBarrier ? "0" : "kk") << "];\n";
  for(size_t r = 0; r < rows.size(); ++r) {
    os_ << "    const size_t gi" << r << " = " << body.stagedIndex(rows[r]) << ";\n";
    os_ << "    const unsigned sh" << r << " = (unsigned)(gi" << r << " % " << E << ");\n";
  }
  os_ << "    ::dawn_b200::tma::expect_bytes(bar, " << ftBytes << "u * (" << rows.size() * block_ << "u";
  for(size_t r = 0; r < rows.size(); ++r)
    os_ << " + (sh" << r << " ? " << E << "u : 0u)";
  os_ << "));\n";
  for(size_t r = 0; r < rows.size(); ++r)
    os_ << "    ::dawn_b200::tma::bulk_load_1d(stg + (kk * " << rows.size() << " + " << r << ") * " << rowElems << ", "
        << fname(rows[r].first) << "_p + (gi" << r << " - sh" << r << "), " << ftBytes << "u * (sh" << r << " ? "
        << block_ + E << "u : " << block_ << "u), bar);\n";
  os_ << "  }\n";
  for(auto const& ch : g.chains) {
    int size = icoChainSize(ch.first, ch.second);
    std::string tag = chainTag(ch.first, ch.second);
    if(!g.topChains.count(tag))
      continue;
    for(int q = 0; q < size; ++q)
      os_ << "  const int nb_" << tag << "_" << q << " = __ldg(&tbl_" << tag << "[pidx + "
          << IcoBody::stride(ch.first.front()) << " * " << q << "]);\n";
  }
  if(oneBarrier)
    os_ << "  if(nlev > 0)\n    ::dawn_b200::tma::wait(&stg_bar[0], 0u);\n";
  const int unroll = opt_.UnrollK > 0 ? opt_.UnrollK : 0;
  os_ << "#pragma unroll" << (unroll ? " " + std::to_string(unroll) : std::string()) << "\n  for(int kk = 0; kk < "
      << lpt_ << "; ++kk) {\n    const int k = k0 + kk;\n    if(k > ke)\n      break;\n";
  if(!oneBarrier)
    os_ << "    ::dawn_b200::tma::wait(&stg_bar[kk], 0u);\n";
  os_ << text << "  }\n";
  return true;
}

void IcoEmitter::emitRolePlain(const Group& g, const std::string& blockExpr, const std::string& lo,
                               const std::string& hi, const RoleOpts& ro) {
  const bool sequential = g.ms->loopOrder != LoopOrder::Parallel;
  const bool backward = g.ms->loopOrder == LoopOrder::Backward;
  // grid: element blocks fastest (x), level blocks in y.  (Level blocks fastest, so that the blocks
  // sharing an element range find their neighbour-table rows in L2, measured slower on B200: ICON
  // nabla2 71 % against 78 % of the HBM peak - too many concurrent DRAM streams.)
  os_ << "  const int pidx = " << lo << " + (int)(" << blockExpr << ") * " << block_
      << " + threadIdx.x; // elements [" << lo << ", " << hi << ")\n";
  if(ro.noReturn)
    os_ << "  if(pidx < " << hi << ") {\n";
  else
    os_ << "  if(pidx >= " << hi << ")\n    return;\n";
  // neighbour indices, once per thread
  for(auto const& ch : g.chains) {
    int size = icoChainSize(ch.first, ch.second);
    std::string tag = chainTag(ch.first, ch.second);
    if(!g.topChains.count(tag))
      continue; // only reached from neighbours: indices are fetched where they are used
    for(int n = 0; n < size; ++n)
      os_ << "  const int nb_" << tag << "_" << n << " = __ldg(&tbl_" << tag << "[pidx + "
          << IcoBody::stride(ch.first.front()) << " * " << n << "]);\n";
  }
  // k range: hull of the DoMethod intervals of the group
  int lo_k = Interval::End * 2, hi_k = -Interval::End;
  for(auto const* st : g.stages)
    for(auto const& dm : st->doMethods) {
      lo_k = std::min(lo_k, dm.interval.lowerBound());
      hi_k = std::max(hi_k, dm.interval.upperBound());
    }
  os_ << "  const int kmin = 0, kmax = k_size - 1;\n";
  os_ << "  int kb = " << analysis::boundExpr(lo_k) << ", ke = " << analysis::boundExpr(hi_k) << ";\n";
  os_ << "  kb = kb < 0 ? 0 : kb; ke = ke > kmax ? kmax : ke;\n";
  if(sequential) {
    if(backward)
      os_ << "  for(int k = ke; k >= kb; --k) {\n";
    else
      os_ << "  for(int k = kb; k <= ke; ++k) {\n";
  } else {
    os_ << "  const int k0 = kb + " << ro.levelBlock << " * " << lpt_ << ";\n";
  }
  // blocks of up to 16 levels are unrolled completely (independent loads of several levels in flight); deeper
  // ones in steps of 8, or the body would outgrow the instruction cache (unroll_k overrides)
  const int unroll = opt_.UnrollK > 0 ? opt_.UnrollK : (lpt_ <= 16 ? 0 : 8);
  const std::string loopHead = std::string("#pragma unroll") + (unroll ? " " + std::to_string(unroll) : std::string()) +
                               "\n  for(int kk = 0; kk < " + std::to_string(lpt_) +
                               "; ++kk) {\n    const int k = k0 + kk;\n";
  if(!sequential && !opt_.IcoFullBlocks)
    os_ << loopHead << "    if(k > ke)\n      break;\n";
  std::ostringstream bodyText;
  IcoBody body{inst_, g, registerOnly_,
               [this](int id) { return fname(id); },
               [this](int id) -> const FieldDims& { return dims(id); },
               [this](int id) { return sparseSize(id); }, {}, 0, ro.cg, ro.streamOnce, ro.storeStream, ro.plainLoad};
  for(auto const* st : g.stages) {
    bodyText << "    // stage " << st->id << "\n";
    for(auto const& dm : st->doMethods) {
      bool guard = dm.interval.lowerBound() != lo_k || dm.interval.upperBound() != hi_k;
      std::vector<std::string> lines;
      for(auto const& s : dm.stmts)
        body.stmt(s, lines, 0, guard);
      if(guard)
        bodyText << "    if(k >= " << analysis::boundExpr(dm.interval.lowerBound()) << " && k <= "
            << analysis::boundExpr(dm.interval.upperBound()) << ") {\n";
      for(auto const& l : lines)
        bodyText << (guard ? "      " : "    ") << l << "\n";
      if(guard) {
        bodyText << "    }\n";
        body.forwarded.clear();
      }
    }
  }
  const std::string levelText = opt_.IcoHoist ? hoistLevelLoads(bodyText.str(), "::dawn::float_type") : bodyText.str();
  if(sequential || !opt_.IcoFullBlocks) {
    os_ << levelText << "  }\n";
    if(ro.noReturn)
      os_ << "  }\n";
    return;
  }
  // ico_full_blocks: blocks of levels that lie completely inside the range run the unrolled loop
  // without a bound check per level (loads of several levels may be scheduled together); the last,
  // partial block takes a plain loop
  os_ << "  if(k0 + " << lpt_ - 1 << " <= ke) {\n" << loopHead << levelText << "  }\n  } else {\n";
  os_ << "  for(int k = k0; k <= ke; ++k) {\n" << levelText << "  }\n  }\n";
  if(ro.noReturn)
    os_ << "  }\n";
}

// Tables (top-level chains from the consumer's own element) through which `c` reads fields `p` writes.
bool IcoEmitter::chainTables(const Group& c, const Group& p,
                             std::vector<std::pair<std::vector<LocationType>, bool>>& out) const {
  bool ok = true;
  std::set<std::string> seen;
  std::function<void(const ExprPtr&, const Expr*, int)> walk = [&](const ExprPtr& e, const Expr* red, int nest) {
    if(!e)
      return;
    if(e->kind == Expr::Field && p.fieldsWritten.count(e->accessID)) {
      if(nest == 0) {
        if(fieldLoc(e->accessID) != c.loc)
          ok = false; // a dense field of another location type read outside a reduction: not a valid access
      } else if(nest == 1 && red && (e->hasOffset || fieldLoc(e->accessID) != c.loc)) {
        if(fieldLoc(e->accessID) != red->chain.back() || fieldLoc(e->accessID) != p.loc)
          ok = false;
        else if(seen.insert(chainTag(red->chain, red->includeCenter)).second)
          out.push_back({red->chain, red->includeCenter});
      } else if(nest > 1)
        ok = false; // reached through a nested reduction: the dependence range is not a single table
    }
    if(e->kind == Expr::Reduce) {
      walk(e->kids[0], e.get(), nest + 1);
      for(size_t n = 1; n < e->kids.size(); ++n)
        walk(e->kids[n], red, nest);
      return;
    }
    for(auto const& k : e->kids)
      walk(k, red, nest);
  };
  std::function<void(const StmtPtr&)> ws = [&](const StmtPtr& st) {
    if(st->kind == Stmt::Loop) { // neighbour loops touching produced fields: not covered
      analysis::visitStmts(st->body, [&](const Expr& e, bool) {
        if(e.kind == Expr::Field && p.fieldsWritten.count(e.accessID))
          ok = false;
      });
      return;
    }
    walk(st->expr, nullptr, 0);
    for(auto const& b : st->body)
      ws(b);
    for(auto const& b : st->orelse)
      ws(b);
  };
  for(auto const* stg : c.stages)
    for(auto const& dm : stg->doMethods)
      for(auto const& st : dm.stmts)
        ws(st);
  return ok;
}

void IcoEmitter::planChains(std::vector<Group>& groups) const {
  if(!opt_.IcoChain)
    return;
  auto disjoint = [](const std::set<int>& a, const std::set<int>& b) {
    for(int x : a)
      if(b.count(x))
        return false;
    return true;
  };
  for(size_t i = 0; i < groups.size(); ++i) {
    Group& a = groups[i];
    if(a.coMember || a.chainMember || a.chainNext >= 0)
      continue;
    const Group* partner = a.coWith >= 0 ? &groups[a.coWith] : nullptr;
    const size_t j = partner ? (size_t)a.coWith + 1 : i + 1;
    if(j >= groups.size())
      continue;
    Group& c = groups[j];
    if(c.coWith >= 0 || c.coMember || c.chainMember || c.ms != a.ms || a.ms->loopOrder != LoopOrder::Parallel ||
       a.hasSpace || c.hasSpace || (partner && partner->hasSpace) || c.usesGlobals != a.usesGlobals)
      continue;
    // the consumer must not write what the producers touch, and at its own element it may only read produced
    // fields of its own location type (register forwarding does not cross roles, flags cover neighbourhoods)
    std::set<int> prodTouched = a.fieldsRead, prodWritten = a.fieldsWritten;
    prodTouched.insert(a.fieldsWritten.begin(), a.fieldsWritten.end());
    if(partner) {
      prodTouched.insert(partner->fieldsRead.begin(), partner->fieldsRead.end());
      prodTouched.insert(partner->fieldsWritten.begin(), partner->fieldsWritten.end());
      prodWritten.insert(partner->fieldsWritten.begin(), partner->fieldsWritten.end());
    }
    if(!disjoint(c.fieldsWritten, prodTouched) || disjoint(c.fieldsRead, prodWritten))
      continue;
    bool temporaries = false; // produced temporaries live in scratch with the same layout: fine; register-only ones never cross
    (void)temporaries;
    std::vector<std::pair<std::vector<LocationType>, bool>> t0, t1;
    if(!chainTables(c, a, t0) || (partner && !chainTables(c, *partner, t1)))
      continue;
    if(t0.empty() && t1.empty())
      continue;
    // a produced field of the consumer's own location read at the own element would need the producer block
    // of the same index: keep such programs as two launches
    bool ownRead = false;
    for(auto const* stg : c.stages)
      forEachAccess(*stg, [&](const Expr& e, int nest, bool) {
        if(prodWritten.count(e.accessID) && (nest == 0 || (!e.hasOffset && fieldLoc(e.accessID) == c.loc)))
          ownRead = true;
      });
    if(ownRead)
      continue;
    a.chainNext = (int)j;
    c.chainMember = true;
  }
}

void IcoEmitter::emitKernel(const Group& g, const Group* partner, const Group* consumer) {
  const bool sequential = g.ms->loopOrder != LoopOrder::Parallel;
  std::vector<const Group*> roles{&g};
  if(partner)
    roles.push_back(partner);
  if(consumer)
    roles.push_back(consumer);
  std::set<int> args, written;
  for(auto const* r : roles) {
    args.insert(r->fieldsRead.begin(), r->fieldsRead.end());
    args.insert(r->fieldsWritten.begin(), r->fieldsWritten.end());
    written.insert(r->fieldsWritten.begin(), r->fieldsWritten.end());
  }
  RoleOpts plain;
  if(opt_.IcoCacheHints && !consumer) {
    std::set<int> gathered;
    for(auto const* r : roles)
      for(auto const* stg : r->stages)
        forEachAccess(*stg, [&](const Expr& e, int nest, bool) {
          if(nest > 0 && (e.hasOffset || nest > 1 || fieldLoc(e.accessID) != r->loc))
            gathered.insert(e.accessID);
        });
    for(int f : args)
      if(!written.count(f) && !gathered.count(f) && !isVertical(f) && hasK(f))
        plain.streamOnce.insert(f);
    for(int f : written)
      if(!laterReads_.count({&g, f}) && !isVertical(f))
        plain.storeStream.insert(f);
  }
  if(opt_.IcoStage && !consumer && !sequential)
    for(int f : args)
      if(!written.count(f) && !isVertical(f) && hasK(f) && !registerOnly_.count(f) && !scratch_.count(f))
        plain.stage.insert(f);
  // the roles are rendered first: whether the kernel has a staged variant (dynamic shared memory, one more
  // argument, fewer resident blocks) is known only afterwards
  curStage_ = StageMeta{};
  std::string rolesText;
  if(!consumer) {
    const std::string saved = os_.str();
    os_.str("");
    if(!partner) {
      emitRole(g, "blockIdx.x", "elem_lo", "elem_hi", plain);
    } else {
      // blocks alternate between the two roles in proportion nb0 : nb1 (role 0 owns the blocks where
      // floor((b+1) nb0 / tot) steps), so both sweep their element ranges at the same relative pace
      os_ << "  const unsigned long long tot = (unsigned long long)nb0 + nb1;\n";
      os_ << "  const unsigned r0 = (unsigned)(((unsigned long long)blockIdx.x * nb0) / tot);\n";
      os_ << "  const unsigned r0n = (unsigned)(((unsigned long long)(blockIdx.x + 1) * nb0) / tot);\n";
      os_ << "  if(r0n > r0) {\n";
      emitRole(g, "r0", "elem_lo", "elem_hi", plain);
      os_ << "  } else {\n";
      emitRole(*partner, "blockIdx.x - r0", "elem_lo1", "elem_hi1", plain);
      os_ << "  }\n";
    }
    rolesText = os_.str();
    os_.str("");
    os_ << saved;
  }
  const bool staged = curStage_.smemBytes > 0;
  if(staged)
    stageMeta_[g.name] = curStage_;
  os_ << "// kernel " << g.name << ": " << g.stages.size() << " fused stage(s) on " << locName(g.loc);
  if(partner)
    os_ << " co-scheduled with " << partner->stages.size() << " stage(s) on " << locName(partner->loc);
  if(consumer)
    os_ << ", chained with " << consumer->stages.size() << " stage(s) on " << locName(consumer->loc)
        << " that gather their results (flags + L2 hand-over)";
  os_ << ", " << block_ << " threads, "
      << (sequential ? std::string("sequential k sweep") : std::to_string(lpt_) + " level(s)/thread");
  if(staged)
    os_ << ", own-element rows staged by cp.async.bulk (" << curStage_.smemBytes << " bytes of shared memory)";
  os_ << "\n";
  // --max-blocks-sm (CodeGen/Options.inc): resident blocks per SM the register allocator must allow.
  // Default: full occupancy (2048 threads per SM, i.e. <= 32 registers per thread) - these kernels are
  // bound by the latency of their gathers, and ICON nabla2 on B200 measured 78 % of the HBM peak at 64
  // resident warps against 72 % at the 48 warps ptxas' own 40-register allocation allowed.  A staged kernel
  // cannot have more resident blocks than its shared memory allows (227 KB per SM, 1 KB reserved per block).
  int minBlocks = opt_.MaxBlocksPerSM > 0 ? opt_.MaxBlocksPerSM : std::max(1, 2048 / block_);
  if(staged && opt_.MaxBlocksPerSM <= 0)
    minBlocks = std::max(1, std::min<int>(minBlocks, (int)(227 * 1024 / (curStage_.smemBytes + 1024))));
  os_ << "__global__ void __launch_bounds__(" << block_ << ", " << minBlocks << ")\n" << g.name << "(";
  if(g.usesGlobals)
    os_ << "const globals g, ";
  os_ << "const int k_size, const int elem_lo, const int elem_hi, const int n_cells, const int n_edges, "
         "const int n_vertices";
  if(partner)
    os_ << ", const int elem_lo1, const int elem_hi1, const unsigned nb0, const unsigned nb1";
  if(consumer)
    os_ << ", const int elem_lo2, const int elem_hi2, const b200_chain_view cv, const unsigned epoch";
  if(staged)
    os_ << ", const int stage_ok";
  std::set<std::string> tags;
  for(auto const* r : roles)
    for(auto const& ch : r->chains)
      if(tags.insert(chainTag(ch.first, ch.second)).second)
        os_ << ",\n    const int* __restrict__ tbl_" << chainTag(ch.first, ch.second);
  for(int f : args) {
    if(registerOnly_.count(f))
      continue;
    bool ro = !written.count(f);
    os_ << ",\n    " << (ro ? "const " : "") << "::dawn::float_type* __restrict__ " << fname(f) << "_p";
  }
  os_ << ") {\n";
  os_ << "  (void)n_cells; (void)n_edges; (void)n_vertices;\n";
  if(staged)
    os_ << "  extern __shared__ __align__(128) unsigned char b200_smem[];\n";
  if(!consumer) {
    os_ << rolesText << "}\n\n";
    return;
  }
  // ---- chained launch: blockIdx.x = level block * sched_len + position in the host-built schedule ----------
  // cache policy: fields every role only reads at its own element stream through (evict-first) so that the
  // handed-over fields stay in L2 until their consumers arrive; final results nobody re-reads stream out
  RoleOpts prod, cons;
  prod.levelBlock = cons.levelBlock = "lb";
  prod.noReturn = cons.noReturn = true;
  std::set<int> gathered, producedSet = g.fieldsWritten;
  if(partner)
    producedSet.insert(partner->fieldsWritten.begin(), partner->fieldsWritten.end());
  for(auto const* r : roles)
    for(auto const* stg : r->stages)
      forEachAccess(*stg, [&](const Expr& e, int nest, bool) {
        if(nest > 0 && (e.hasOffset || nest > 1 || fieldLoc(e.accessID) != r->loc))
          gathered.insert(e.accessID);
      });
  if(opt_.IcoCacheHints) {
    for(int f : args)
      if(!written.count(f) && !gathered.count(f) && !isVertical(f) && hasK(f))
        prod.streamOnce.insert(f), cons.streamOnce.insert(f);
    for(int f : consumer->fieldsWritten)
      cons.storeStream.insert(f);
  }
  // hand-over reads.  ico_chain_l1=0: ld.global.cg (L2 is the point of coherence).  Default: ordinary loads
  // that may stay in L1 - safe because the consumer also waits for the producer block on either side of its
  // range, so every 128-byte line it touches holds final values only (a line spans at most two neighbouring
  // producer blocks; the handed-over fields are written once per launch and L1 does not outlive a launch)
  for(int f : consumer->fieldsRead)
    if(producedSet.count(f)) {
      if(opt_.IcoChainL1)
        cons.plainLoad.insert(f);
      else
        cons.cg.insert(f);
    }
  os_ << "  const unsigned lb = blockIdx.x / cv.sched_len;   // block of levels\n";
  os_ << "  const unsigned item = __ldg(&cv.sched[blockIdx.x - lb * cv.sched_len]);\n";
  os_ << "  const unsigned role = item >> 30, bidx = item & 0x3fffffffu;\n";
  os_ << "  unsigned* const flags = cv.flags + (size_t)lb * (cv.nb0 + cv.nb1);\n";
  os_ << "  if(role == 0u) {\n";
  emitRole(g, "bidx", "elem_lo", "elem_hi", prod);
  os_ << "    __syncthreads();\n    if(threadIdx.x == 0)\n      ::dawn_b200::chain::publish(&flags[bidx], epoch);\n";
  if(partner) {
    os_ << "  } else if(role == 1u) {\n";
    emitRole(*partner, "bidx", "elem_lo1", "elem_hi1", prod);
    os_ << "    __syncthreads();\n    if(threadIdx.x == 0)\n      ::dawn_b200::chain::publish(&flags[cv.nb0 + bidx], epoch);\n";
  }
  os_ << "  } else {\n";
  os_ << "    // wait for the producer blocks this block's elements reach through the neighbour tables\n";
  os_ << "    int lo0 = __ldg(&cv.need[4 * bidx]), hi0 = __ldg(&cv.need[4 * bidx + 1]);\n";
  os_ << "    int lo1 = __ldg(&cv.need[4 * bidx + 2]), hi1 = __ldg(&cv.need[4 * bidx + 3]);\n";
  if(opt_.IcoChainL1) {
    os_ << "    // one more producer block on either side: whole cache lines are final (the schedule allows for it)\n";
    os_ << "    if(hi0 > lo0) { lo0 = lo0 > 0 ? lo0 - 1 : 0; hi0 = hi0 < (int)cv.nb0 ? hi0 + 1 : (int)cv.nb0; }\n";
    os_ << "    if(hi1 > lo1) { lo1 = lo1 > 0 ? lo1 - 1 : 0; hi1 = hi1 < (int)cv.nb1 ? hi1 + 1 : (int)cv.nb1; }\n";
  }
  os_ << "    for(int q = lo0 + (int)threadIdx.x; q < hi0; q += " << block_ << ")\n"
      << "      ::dawn_b200::chain::wait(&flags[q], epoch, cv.error);\n";
  os_ << "    for(int q = lo1 + (int)threadIdx.x; q < hi1; q += " << block_ << ")\n"
      << "      ::dawn_b200::chain::wait(&flags[cv.nb0 + q], epoch, cv.error);\n";
  os_ << "    __syncthreads();\n";
  emitRole(*consumer, "bidx", "elem_lo2", "elem_hi2", cons);
  os_ << "  }\n";
  os_ << "}\n\n";
}

void IcoEmitter::emitHost(const std::vector<Group>& groups) {
  const std::string cls = cIdent(inst_.name);
  const bool hasGlobals = !inst_.globals.empty();
  // distinct tables over all kernels
  std::vector<std::pair<std::vector<LocationType>, bool>> tables;
  std::set<std::string> seen;
  for(auto const& g : groups)
    for(auto const& ch : g.chains)
      if(seen.insert(chainTag(ch.first, ch.second)).second)
        tables.push_back(ch);

  os_ << "class " << cls << " {\n";
  os_ << "  b200_mesh_t m_mesh;\n  int m_k_size;\n";
  if(hasGlobals)
    os_ << "public:\n  globals m_globals;\nprivate:\n";
  for(auto const& t : tables)
    os_ << "  const int* m_tbl_" << chainTag(t.first, t.second) << " = nullptr;\n";
  for(int f : scratch_)
    os_ << "  ::dawn_b200::scratch_field m_" << fname(f) << ";\n";
  os_ << "  ::dawn_b200::timer_b200 m_timer;\n";
  for(size_t gi = 0; gi < groups.size(); ++gi)
    if(groups[gi].chainNext >= 0)
      os_ << "  b200_chain_plan* m_chain_" << gi << " = nullptr; // block schedule + flags of the chained launch\n";
  for(size_t gi = 0; gi < groups.size(); ++gi)
    if(stageMeta_.count(groups[gi].name))
      os_ << "  bool m_stage_smem_" << gi << " = false; // dynamic shared memory limit of the staged kernel raised\n";
  os_ << "  std::vector<b200_splitter_t> m_splitters; // unstructured_domain twin\n";
  os_ << "  b200_metrics_record* m_record = nullptr;  // the json* the reference hands to setup_<N>\n";
  os_ << "  int m_error = 0;\npublic:\n  long m_launches = 0;\n";
  os_ << "  " << cls << "(const " << cls << "&) = delete;\n";
  os_ << "  " << cls << "(const b200_mesh_t* mesh, int k_size) : m_mesh(*mesh), m_k_size(k_size), m_timer(\""
      << inst_.name << "\") {\n";
  os_ << "    for(int n = 0; n < mesh->num_splitters && mesh->splitters; ++n)\n"
         "      m_splitters.push_back(mesh->splitters[n]);\n"
         "    m_mesh.num_splitters = 0; m_mesh.splitters = nullptr; // owned copy above\n";
  for(auto const& t : tables) {
    os_ << "    {\n      const int chain[] = {";
    for(size_t n = 0; n < t.first.size(); ++n)
      os_ << (n ? ", " : "") << static_cast<int>(t.first[n]);
    os_ << "};\n";
    os_ << "      m_tbl_" << chainTag(t.first, t.second) << " = ::dawn_b200::find_table(m_mesh, chain, "
        << t.first.size() << ", " << (t.second ? 1 : 0) << ", " << icoChainSize(t.first, t.second)
        << ");\n";
    os_ << "      if(!m_tbl_" << chainTag(t.first, t.second)
        << ") m_error = ::dawn_b200::set_error(B200_ERR_ARG, \"mesh lacks the neighbour table for chain "
        << chainTag(t.first, t.second) << " (size " << icoChainSize(t.first, t.second) << ")\");\n";
    os_ << "    }\n";
  }
  os_ << "  }\n";
  os_ << "  ~" << cls << "() {";
  for(size_t gi = 0; gi < groups.size(); ++gi)
    if(groups[gi].chainNext >= 0)
      os_ << " b200rt_chain_plan_destroy(m_chain_" << gi << ");";
  os_ << " }\n";
  os_ << "  int error() const { return m_error; }\n";
  os_ << "  // chained launches: did every consumer block see its producers?  (synchronises the stream)\n";
  os_ << "  int chain_status(void* stream) {\n";
  for(size_t gi = 0; gi < groups.size(); ++gi)
    if(groups[gi].chainNext >= 0)
      os_ << "    if(int err = b200rt_chain_plan_status(m_chain_" << gi << ", stream)) return err;\n";
  os_ << "    (void)stream;\n    return 0;\n  }\n";
  os_ << "  void set_metrics_record(b200_metrics_record* r) { m_record = r; }\n";
  os_ << "  // dawn::unstructured_domain::set_splitter_index (driver-includes/unstructured_domain.hpp:31-33)\n";
  os_ << "  void set_splitter_index(int loc, int space, int offset, int index) {\n"
         "    for(auto& s : m_splitters)\n"
         "      if(s.location == loc && s.subdomain == space && s.offset == offset) { s.index = index; return; }\n"
         "    m_splitters.push_back(b200_splitter_t{loc, space, offset, index});\n  }\n";
  os_ << "  bool splitter(int loc, int space, int offset, int* index) const {\n"
         "    for(auto const& s : m_splitters)\n"
         "      if(s.location == loc && s.subdomain == space && s.offset == offset) { *index = s.index; return true; }\n"
         "    return false;\n  }\n";
  os_ << "  void set_stream(void* s) { m_timer.set_stream(s); }\n";
  for(auto const& gl : inst_.globals) {
    std::string t = gl.second.type == "Integer" ? "int" : gl.second.type == "Boolean" ? "bool" : "double";
    os_ << "  " << t << " get_" << cIdent(gl.first) << "() { return m_globals." << cIdent(gl.first) << "; }\n";
    os_ << "  void set_" << cIdent(gl.first) << "(" << t << " v) { m_globals." << cIdent(gl.first) << " = v; }\n";
  }
  // launch
  os_ << "\n  // fields: device pointers in API order; dense [k][element] (stride_k = element count),\n"
         "  // sparse [k][nbh][element] (stride_j = element count), vertical [k]\n";
  os_ << "  int launch(const b200_field_t* f, void* stream_) {\n";
  os_ << "    if(m_error) return m_error;\n";
  os_ << "    cudaStream_t stream = static_cast<cudaStream_t>(stream_);\n";
  os_ << "    const int n_cells = m_mesh.num_cells, n_edges = m_mesh.num_edges, n_vertices = m_mesh.num_vertices;\n";
  os_ << "    const int k_size = m_k_size;\n    if(k_size <= 0) return 0;\n";
  for(size_t a = 0; a < inst_.apiFieldIDs.size(); ++a) {
    int f = inst_.apiFieldIDs[a];
    os_ << "    ::dawn::float_type* " << fname(f) << "_p = static_cast<::dawn::float_type*>(f[" << a
        << "].data);\n";
    if(!isVertical(f)) {
      std::string n = IcoBody::stride(fieldLoc(f));
      if(isSparse(f))
        os_ << "    if(f[" << a << "].stride_j != " << n
            << (hasK(f) ? " || f[" + std::to_string(a) + "].stride_k != " + n + " * " +
                              std::to_string(sparseSize(f))
                        : std::string())
            << ") return ::dawn_b200::set_error(B200_ERR_LAYOUT, \"sparse field "
            << fname(f) << " must be laid out [k][" << sparseSize(f) << "][" << n << "]\");\n";
      else if(hasK(f))
        os_ << "    if(f[" << a << "].stride_k != " << n
            << ") return ::dawn_b200::set_error(B200_ERR_LAYOUT, \"field " << fname(f)
            << " must be laid out [k][" << n << "]\");\n";
    }
  }
  for(int f : scratch_) {
    std::string n = IcoBody::stride(fieldLoc(f));
    // a sparse temporary holds sparseSize values per element and level: [k][nbh][element]
    const std::string levelStride = isSparse(f) ? n + " * " + std::to_string(sparseSize(f)) : n;
    os_ << "    if(int err = m_" << fname(f) << ".ensure(" << (isSparse(f) ? n : std::string("0")) << ", "
        << levelStride << ", k_size)) return err;\n";
    os_ << "    ::dawn::float_type* " << fname(f) << "_p = m_" << fname(f) << ".data;\n";
  }
  for(size_t gi = 0; gi < groups.size(); ++gi) {
    const Group& g = groups[gi];
    if(g.coMember || g.chainMember)
      continue;
    const Group* partner = g.coWith >= 0 ? &groups[g.coWith] : nullptr;
    const Group* consumer = g.chainNext >= 0 ? &groups[g.chainNext] : nullptr;
    bool sequential = g.ms->loopOrder != LoopOrder::Parallel;
    std::string n = IcoBody::stride(g.loc);
    os_ << "    {\n";
    if(g.hasSpace) {
      const int loc = static_cast<int>(g.loc);
      os_ << "      int elem_lo = 0, elem_hi = 0;\n";
      os_ << "      if(!splitter(" << loc << ", " << g.space.lowerLevel << ", " << g.space.lowerOffset
          << ", &elem_lo) || !splitter(" << loc << ", " << g.space.upperLevel << ", "
          << g.space.upperOffset << ", &elem_hi))\n"
          << "        return ::dawn_b200::set_error(B200_ERR_ARG, \"" << inst_.name
          << ": splitter index of an unstructured iteration space is not set (set_splitter_index)\");\n";
      os_ << "      if(elem_lo < 0 || elem_hi > " << n << ")\n"
          << "        return ::dawn_b200::set_error(B200_ERR_ARG, \"" << inst_.name
          << ": splitter index outside the mesh\");\n";
    } else
      os_ << "      const int elem_lo = 0, elem_hi = " << n << ";\n";
    const std::string levelBlocks =
        sequential ? std::string("1")
                   : "(k_size + " + std::to_string(lpt_ - 1) + ") / " + std::to_string(lpt_);
    if(partner) {
      os_ << "      const int elem_lo1 = 0, elem_hi1 = " << IcoBody::stride(partner->loc) << ";\n";
      os_ << "      const unsigned nb0 = (unsigned)((elem_hi - elem_lo + " << block_ - 1 << ") / " << block_
          << "), nb1 = (unsigned)((elem_hi1 - elem_lo1 + " << block_ - 1 << ") / " << block_ << ");\n";
    }
    if(consumer) {
      // the schedule and the flags of the chained launch: built on first use from the device tables
      os_ << "      const int elem_lo2 = 0, elem_hi2 = " << IcoBody::stride(consumer->loc) << ";\n";
      os_ << "      if(!m_chain_" << gi << ") {\n";
      std::vector<std::pair<std::vector<LocationType>, bool>> t0, t1;
      chainTables(*consumer, g, t0);
      if(partner)
        chainTables(*consumer, *partner, t1);
      os_ << "        const int n_prod[2] = {elem_hi, " << (partner ? "elem_hi1" : "0") << "};\n";
      os_ << "        const int* tabs[] = {";
      for(auto const& t : t0)
        os_ << "m_tbl_" << chainTag(t.first, t.second) << ", ";
      for(auto const& t : t1)
        os_ << "m_tbl_" << chainTag(t.first, t.second) << ", ";
      os_ << "};\n        const int tab_role[] = {";
      for(size_t q = 0; q < t0.size(); ++q)
        os_ << "0, ";
      for(size_t q = 0; q < t1.size(); ++q)
        os_ << "1, ";
      os_ << "};\n        const int tab_nbh[] = {";
      for(auto const& t : t0)
        os_ << icoChainSize(t.first, t.second) << ", ";
      for(auto const& t : t1)
        os_ << icoChainSize(t.first, t.second) << ", ";
      os_ << "};\n";
      os_ << "        if(int err = b200rt_chain_plan_create(&m_chain_" << gi << ", elem_hi2, " << block_ << ", "
          << (partner ? 2 : 1) << ", n_prod, tabs, tab_role, tab_nbh, " << t0.size() + t1.size() << ", "
          << levelBlocks << ", " << opt_.IcoChainLag << ", stream)) return err;\n";
      os_ << "      }\n";
      os_ << "      b200_chain_view cv;\n      b200rt_chain_plan_view(m_chain_" << gi << ", &cv);\n";
      os_ << "      const unsigned epoch = b200rt_chain_plan_next_epoch(m_chain_" << gi << ");\n";
      os_ << "      if(cv.sched_len > 0) {\n";
      os_ << "      dim3 grid(cv.sched_len * cv.level_blocks);\n";
    } else if(partner) {
      os_ << "      if(nb0 + nb1 > 0) {\n";
      os_ << "      dim3 grid(nb0 + nb1, " << levelBlocks << ");\n";
    } else {
      os_ << "      if(elem_hi > elem_lo) {\n";
      os_ << "      dim3 grid((elem_hi - elem_lo + " << block_ - 1 << ") / " << block_ << ", " << levelBlocks
          << ");\n";
    }
    auto sm = stageMeta_.find(g.name);
    if(sm != stageMeta_.end()) {
      // the staged variant needs 16-byte aligned field bases (cp.async.bulk); full blocks take it then
      os_ << "      const bool stage_ok = true";
      for(int f : sm->second.fields)
        os_ << " && reinterpret_cast<uintptr_t>(" << fname(f) << "_p) % 16 == 0";
      os_ << ";\n";
      os_ << "      if(stage_ok && !m_stage_smem_" << gi << ") {\n"
          << "        cudaError_t ae = cudaFuncSetAttribute(reinterpret_cast<const void*>(&" << g.name
          << "), cudaFuncAttributeMaxDynamicSharedMemorySize, " << sm->second.smemBytes << ");\n"
          << "        if(ae != cudaSuccess) return ::dawn_b200::set_error(B200_ERR_CUDA, cudaGetErrorString(ae));\n"
          << "        m_stage_smem_" << gi << " = true;\n      }\n";
      os_ << "      " << g.name << "<<<grid, " << block_ << ", stage_ok ? " << sm->second.smemBytes
          << " : 0, stream>>>(";
    } else
      os_ << "      " << g.name << "<<<grid, " << block_ << ", 0, stream>>>(";
    if(g.usesGlobals)
      os_ << "m_globals, ";
    os_ << "k_size, elem_lo, elem_hi, n_cells, n_edges, n_vertices";
    if(partner)
      os_ << ", elem_lo1, elem_hi1, nb0, nb1";
    if(consumer)
      os_ << ", elem_lo2, elem_hi2, cv, epoch";
    if(sm != stageMeta_.end())
      os_ << ", stage_ok ? 1 : 0";
    std::set<std::string> tags;
    std::vector<const Group*> roles{&g};
    if(partner)
      roles.push_back(partner);
    if(consumer)
      roles.push_back(consumer);
    for(auto const* r : roles)
      for(auto const& ch : r->chains)
        if(tags.insert(chainTag(ch.first, ch.second)).second)
          os_ << ", m_tbl_" << chainTag(ch.first, ch.second);
    std::set<int> args;
    for(auto const* r : roles) {
      args.insert(r->fieldsRead.begin(), r->fieldsRead.end());
      args.insert(r->fieldsWritten.begin(), r->fieldsWritten.end());
    }
    for(int f : args)
      if(!registerOnly_.count(f))
        os_ << ", " << fname(f) << "_p";
    os_ << ");\n      ++m_launches;\n      }\n    }\n";
  }
  os_ << "    return ::dawn_b200::check_launch();\n  }\n";
  // ---- host-buffer entry points (reference: run_<N>_from_{c,fort}_host, CudaIcoCodeGen.cpp:896-927;
  // cuda_utils.hpp:81-125,181-195 do the transposition on one host thread, here it is a GPU kernel) --
  // the emitted host code names ::dawn::float_type; the runtime's typed entry points carry a suffix
  const std::string sfx = opt_.SinglePrecision ? "_f32" : "";
  os_ << "  // number of elements of API field `a`\n";
  os_ << "  size_t field_elems(int a) const {\n    switch(a) {\n";
  for(size_t a = 0; a < inst_.apiFieldIDs.size(); ++a) {
    int f = inst_.apiFieldIDs[a];
    os_ << "    case " << a << ": return (size_t)" << (hasK(f) ? "m_k_size" : "1");
    if(!isVertical(f))
      os_ << " * m_mesh.num_" << locName(fieldLoc(f)) << (isSparse(f) ? " * " + std::to_string(sparseSize(f)) : std::string());
    os_ << ";\n";
  }
  os_ << "    default: return 0;\n    }\n  }\n";
  os_ << "  int run_from_host(::dawn::float_type* const* host, bool element_major, void* stream) {\n";
  os_ << "    const int nf = " << inst_.apiFieldIDs.size() << ";\n";
  os_ << "    ::dawn::float_type* dev[" << std::max<size_t>(1, inst_.apiFieldIDs.size()) << "] = {};\n";
  os_ << "    ::dawn::float_type* tmp = nullptr;\n    size_t tmp_elems = 0;\n    int err = 0;\n";
  os_ << "    for(int a = 0; a < nf; ++a) tmp_elems = field_elems(a) > tmp_elems ? field_elems(a) : tmp_elems;\n";
  os_ << "    if(element_major) err = b200rt_malloc((void**)&tmp, tmp_elems * sizeof(::dawn::float_type));\n";
  os_ << "    b200_field_t f[" << std::max<size_t>(1, inst_.apiFieldIDs.size()) << "];\n";
  os_ << "    static const int sparse_of[] = {";
  for(size_t a = 0; a < inst_.apiFieldIDs.size(); ++a)
    os_ << (a ? ", " : "") << (isSparse(inst_.apiFieldIDs[a]) ? sparseSize(inst_.apiFieldIDs[a]) : 1);
  os_ << "};\n    const int elems_of[] = {";
  for(size_t a = 0; a < inst_.apiFieldIDs.size(); ++a) {
    int f = inst_.apiFieldIDs[a];
    os_ << (a ? ", " : "") << (isVertical(f) ? std::string("1") : "m_mesh.num_" + locName(fieldLoc(f)));
  }
  os_ << "};\n    const int levels_of[] = {";
  for(size_t a = 0; a < inst_.apiFieldIDs.size(); ++a)
    os_ << (a ? ", " : "") << (hasK(inst_.apiFieldIDs[a]) ? "m_k_size" : "1");
  os_ << "};\n    static const bool written_of[] = {";
  {
    std::set<int> written;
    for(auto const& g : groups)
      written.insert(g.fieldsWritten.begin(), g.fieldsWritten.end());
    for(size_t a = 0; a < inst_.apiFieldIDs.size(); ++a)
      os_ << (a ? ", " : "") << (written.count(inst_.apiFieldIDs[a]) ? "true" : "false");
  }
  os_ << "};\n";
  os_ << "    for(int a = 0; a < nf && !err; ++a) {\n";
  os_ << "      const size_t bytes = field_elems(a) * sizeof(::dawn::float_type);\n";
  os_ << "      err = b200rt_malloc((void**)&dev[a], bytes);\n";
  os_ << "      if(err) break;\n";
  os_ << "      if(element_major && elems_of[a] > 1) {\n";
  os_ << "        err = b200rt_memcpy_h2d(tmp, host[a], bytes, stream);\n";
  os_ << "        if(!err) err = b200rt_reshape" << sfx << "(tmp, dev[a], levels_of[a], elems_of[a], sparse_of[a], stream);\n";
  os_ << "      } else\n        err = b200rt_memcpy_h2d(dev[a], host[a], bytes, stream);\n";
  os_ << "      f[a].data = dev[a];\n";
  os_ << "      f[a].stride_j = sparse_of[a] > 1 ? elems_of[a] : 0;\n";
  os_ << "      f[a].stride_k = elems_of[a] * sparse_of[a];\n";
  os_ << "    }\n";
  os_ << "    if(!err) err = launch(f, stream);\n";
  os_ << "    for(int a = 0; a < nf && !err; ++a) {\n      if(!written_of[a]) continue;\n";
  os_ << "      const size_t bytes = field_elems(a) * sizeof(::dawn::float_type);\n";
  os_ << "      if(element_major && elems_of[a] > 1) {\n";
  os_ << "        err = b200rt_reshape_back" << sfx << "(dev[a], tmp, levels_of[a], elems_of[a], sparse_of[a], stream);\n";
  os_ << "        if(!err) err = b200rt_memcpy_d2h(host[a], tmp, bytes, stream);\n";
  os_ << "        if(!err) err = b200rt_stream_synchronize(stream);\n";
  os_ << "      } else\n        err = b200rt_memcpy_d2h(host[a], dev[a], bytes, stream);\n";
  os_ << "    }\n";
  os_ << "    if(!err) err = b200rt_stream_synchronize(stream);\n";
  os_ << "    for(int a = 0; a < nf; ++a) b200rt_free(dev[a]);\n    b200rt_free(tmp);\n    return err;\n  }\n";
  // verification of the written fields against reference copies on the device (reference:
  // verify_<N>, CudaIcoCodeGen.cpp:949-1028 + cuda_verify.hpp:84-196; defaults rel 1e-12, abs 0)
  os_ << "  int verify(const b200_field_t* f, const ::dawn::float_type* const* ref, const double* rel_tol, "
         "const double* abs_tol, int iteration, b200_verification_metrics* metrics, void* stream) {\n";
  os_ << "    int w = 0;\n";
  {
    std::set<int> written;
    for(auto const& g : groups)
      written.insert(g.fieldsWritten.begin(), g.fieldsWritten.end());
    for(size_t a = 0; a < inst_.apiFieldIDs.size(); ++a)
      if(written.count(inst_.apiFieldIDs[a])) {
        os_ << "    if(int err = b200rt_verify_field" << sfx << "(static_cast<const ::dawn::float_type*>(f[" << a << "].data), ref[w], field_elems("
            << a << "), rel_tol ? rel_tol[w] : " << (opt_.SinglePrecision ? "1e-5" : "1e-12") << ", abs_tol ? abs_tol[w] : 0.0, iteration, &metrics[w], stream)) return err;\n";
        // reference: MetricsSerialiser(jsonRecord, metrics, "<stencil>", "<field>").writeJson(iteration)
        // (CudaIcoCodeGen.cpp:1009-1016)
        os_ << "    if(m_record) b200rt_metrics_record_add(m_record, \"" << inst_.name << "\", \""
            << inst_.nameOf(inst_.apiFieldIDs[a]) << "\", &metrics[w]);\n";
        os_ << "    ++w;\n";
      }
  }
  os_ << "    return 0;\n  }\n";
  os_ << "  void run(const b200_field_t* f) {\n    m_timer.start();\n"
         "    if(launch(f, m_timer.stream())) ::dawn_b200::throw_last_error();\n    m_timer.pause();\n  }\n";
  os_ << "  std::string get_name() const { return \"" << inst_.name << "\"; }\n";
  os_ << "  void reset_meters() { m_timer.reset(); }\n";
  os_ << "  double get_total_time() { return m_timer.total_time(); }\n";
  os_ << "};\n";
}

std::string IcoEmitter::generate() {
  if(inst_.stencils.size() != 1)
    throw std::runtime_error("b200-ico: exactly one stencil per instantiation is supported "
                             "(as in the reference, CudaIcoCodeGen.cpp:811)");
  os_ << "namespace dawn_generated {\nnamespace b200ico {\n\n";
  if(!inst_.globals.empty()) {
    os_ << "struct globals {\n";
    std::string init;
    for(auto const& g : inst_.globals) {
      std::string t = g.second.type == "Integer" ? "int" : g.second.type == "Boolean" ? "bool" : "double";
      os_ << "  " << t << " " << cIdent(g.first) << ";\n";
      if(g.second.valueIsSet) {
        std::ostringstream v;
        v.precision(17);
        v << g.second.value;
        init += (init.empty() ? " : " : ", ") + cIdent(g.first) + "(" + v.str() + ")";
      }
    }
    os_ << "  globals()" << init << " {}\n};\n\n";
  }
  auto groups = plan(inst_.stencils[0]);
  planCoScheduling(groups);
  planChains(groups);
  for(size_t a = 0; a < groups.size(); ++a)   // which results of a launch does a later launch read again?
    for(size_t b = a + 1; b < groups.size(); ++b) {
      if(b == (size_t)groups[a].coWith || (groups[a].coMember && b == a))
        continue;
      for(int f : groups[b].fieldsRead) {
        laterReads_.insert({&groups[a], f});
        if(a > 0 && groups[a].coMember)
          laterReads_.insert({&groups[a - 1], f}); // the launch of a co-scheduled pair belongs to its first group
      }
    }
  for(auto const& g : groups) {
    if(g.coMember || g.chainMember)
      continue;
    emitKernel(g, g.coWith >= 0 ? &groups[g.coWith] : nullptr, g.chainNext >= 0 ? &groups[g.chainNext] : nullptr);
  }
  emitHost(groups);
  os_ << "} // namespace b200ico\n} // namespace dawn_generated\n\n";

  const std::string n = cIdent(inst_.name);
  const std::string cls = "dawn_generated::b200ico::" + n;
  os_ << "#define B200_EXPORT __attribute__((visibility(\"default\")))\n";
  os_ << "extern \"C\" {\n";
  // twin of the reference's setup_<N>(GlobalGpuTriMesh*, k_size, stream, ...) (CudaIcoCodeGen.cpp:1198-1219)
  os_ << "B200_EXPORT int setup_" << n << "(void** handle, const b200_mesh_t* mesh, int k_size, void* stream) {\n"
      << "  try {\n    auto* s = new " << cls << "(mesh, k_size);\n    s->set_stream(stream);\n"
      << "    if(int err = s->error()) { delete s; return err; }\n    *handle = s;\n    return 0;\n"
      << "  } catch(std::exception& e) { return ::dawn_b200::set_error(B200_ERR_RUNTIME, e.what()); }\n}\n";
  os_ << "B200_EXPORT int run_" << n << "(void* handle, const b200_field_t* fields, int nfields, void* stream) {\n"
      << "  if(!handle || nfields != " << inst_.apiFieldIDs.size()
      << ") return ::dawn_b200::set_error(B200_ERR_ARG, \"run_" << n << ": expected "
      << inst_.apiFieldIDs.size() << " fields\");\n"
      << "  return static_cast<" << cls << "*>(handle)->launch(fields, stream);\n}\n";
  os_ << "B200_EXPORT int run_" << n << "_from_c_host(void* handle, ::dawn::float_type* const* host_fields, int nfields, void* stream) {\n"
      << "  if(!handle || nfields != " << inst_.apiFieldIDs.size()
      << ") return ::dawn_b200::set_error(B200_ERR_ARG, \"run_" << n << "_from_c_host: field count\");\n"
      << "  return static_cast<" << cls << "*>(handle)->run_from_host(host_fields, true, stream);\n}\n";
  os_ << "B200_EXPORT int run_" << n << "_from_fort_host(void* handle, ::dawn::float_type* const* host_fields, int nfields, void* stream) {\n"
      << "  if(!handle || nfields != " << inst_.apiFieldIDs.size()
      << ") return ::dawn_b200::set_error(B200_ERR_ARG, \"run_" << n << "_from_fort_host: field count\");\n"
      << "  return static_cast<" << cls << "*>(handle)->run_from_host(host_fields, false, stream);\n}\n";
  os_ << "B200_EXPORT int verify_" << n << "(void* handle, const b200_field_t* fields, const ::dawn::float_type* const* ref_outputs, "
         "const double* rel_tol, const double* abs_tol, int iteration, b200_verification_metrics* metrics, void* stream) {\n"
      << "  return static_cast<" << cls << "*>(handle)->verify(fields, ref_outputs, rel_tol, abs_tol, iteration, metrics, stream);\n}\n";
  os_ << "B200_EXPORT int run_and_verify_" << n << "(void* handle, const b200_field_t* fields, int nfields, "
         "const ::dawn::float_type* const* ref_outputs, const double* rel_tol, const double* abs_tol, int iteration, "
         "b200_verification_metrics* metrics, void* stream) {\n"
      << "  if(int err = run_" << n << "(handle, fields, nfields, stream)) return err;\n"
      << "  return verify_" << n << "(handle, fields, ref_outputs, rel_tol, abs_tol, iteration, metrics, stream);\n}\n";
  os_ << "B200_EXPORT void free_" << n << "(void* handle) { delete static_cast<" << cls << "*>(handle); }\n";
  os_ << "B200_EXPORT long launches_" << n << "(void* handle) { return static_cast<" << cls
      << "*>(handle)->m_launches; }\n";
  os_ << "B200_EXPORT int chain_status_" << n << "(void* handle, void* stream) { return static_cast<" << cls
      << "*>(handle)->chain_status(stream); }\n";
  os_ << "B200_EXPORT void set_metrics_record_" << n
      << "(void* handle, b200_metrics_record* record) { static_cast<" << cls
      << "*>(handle)->set_metrics_record(record); }\n";
  // reference: set_splitter_index(GlobalGpuTriMesh*, loc, space, offset, index) (cuda_utils.cpp:3-8)
  os_ << "B200_EXPORT void set_splitter_index_" << n
      << "(void* handle, int loc, int space, int offset, int index) { static_cast<" << cls
      << "*>(handle)->set_splitter_index(loc, space, offset, index); }\n";
  for(auto const& g : inst_.globals) {
    os_ << "B200_EXPORT void set_" << n << "_" << cIdent(g.first) << "(void* handle, double v) { static_cast<"
        << cls << "*>(handle)->set_" << cIdent(g.first) << "(v); }\n";
    os_ << "B200_EXPORT double get_" << n << "_" << cIdent(g.first) << "(void* handle) { return static_cast<"
        << cls << "*>(handle)->get_" << cIdent(g.first) << "(); }\n";
  }
  os_ << "B200_EXPORT const char* b200_module_info_" << n << "() {\n  return \"{\\\"name\\\": \\\"" << inst_.name
      << "\\\", \\\"grid\\\": \\\"Unstructured\\\", \\\"float_type\\\": \\\""
      << (opt_.SinglePrecision ? "float" : "double") << "\\\", \\\"fields\\\": [";
  std::set<int> written;
  for(auto const& g : groups)
    written.insert(g.fieldsWritten.begin(), g.fieldsWritten.end());
  for(size_t a = 0; a < inst_.apiFieldIDs.size(); ++a) {
    int f = inst_.apiFieldIDs[a];
    os_ << (a ? ", " : "") << "{\\\"name\\\": \\\"" << inst_.nameOf(f) << "\\\", \\\"location\\\": \\\""
        << (isVertical(f) ? "none" : locName(fieldLoc(f))) << "\\\", \\\"sparse\\\": "
        << (isSparse(f) ? sparseSize(f) : 0) << ", \\\"k\\\": " << (hasK(f) ? 1 : 0) << "}";
  }
  os_ << "], \\\"written\\\": [";
  bool firstW = true;
  for(int f : inst_.apiFieldIDs)
    if(written.count(f)) {
      os_ << (firstW ? "" : ", ") << "\\\"" << inst_.nameOf(f) << "\\\"";
      firstW = false;
    }
  os_ << "], \\\"tables\\\": [";
  {
    std::set<std::string> seen;
    bool firstT = true;
    for(auto const& g : groups)
      for(auto const& ch : g.chains)
        if(seen.insert(chainTag(ch.first, ch.second)).second) {
          os_ << (firstT ? "" : ", ") << "{\\\"chain\\\": [";
          for(size_t q = 0; q < ch.first.size(); ++q)
            os_ << (q ? ", " : "") << static_cast<int>(ch.first[q]);
          os_ << "], \\\"include_center\\\": " << (ch.second ? 1 : 0) << ", \\\"size\\\": "
              << icoChainSize(ch.first, ch.second) << "}";
          firstT = false;
        }
  }
  {
    size_t launches = 0;
    for(auto const& g : groups)
      launches += (g.coMember || g.chainMember) ? 0 : 1;
    os_ << "], \\\"kernels\\\": " << launches << ", \\\"globals\\\": [";
  }
  for(size_t g = 0; g < inst_.globals.size(); ++g)
    os_ << (g ? ", " : "") << "\\\"" << inst_.globals[g].first << "\\\"";
  os_ << "]}\";\n}\n} // extern \"C\"\n";
  return os_.str();
}
} // namespace

// C header and Fortran interface of the per-stencil C ABI (reference: --output-c-header /
// --output-f90-interface, CudaIcoCodeGen.cpp:1525-1853,1869-1892).
std::string icoCHeader(const iir::Instantiation& inst) {
  std::string n;
  for(char c : inst.name)
    n += (std::isalnum(static_cast<unsigned char>(c)) || c == '_') ? c : '_';
  std::ostringstream os;
  os << "/* generated by dawn-b200 (b200-ico): C interface of stencil " << inst.name << " */\n";
  os << "#pragma once\n#include \"dawn_b200_rt.h\"\n";
  os << "/* element type of the fields: double, or float for a module generated with single_precision=1 */\n";
  os << "#ifndef DAWN_B200_FLOAT_TYPE\n#define DAWN_B200_FLOAT_TYPE double\n#endif\n";
  os << "#ifdef __cplusplus\nextern \"C\" {\n#endif\n";
  os << "/* fields in this order:";
  for(int f : inst.apiFieldIDs)
    os << " " << inst.nameOf(f);
  os << " */\n";
  os << "int setup_" << n << "(void** handle, const b200_mesh_t* mesh, int k_size, void* stream);\n";
  os << "void set_splitter_index_" << n << "(void* handle, int loc, int space, int offset, int index);\n";
  os << "/* verification metrics of later verify_/run_and_verify_ calls are also merged into `record` */\n";
  os << "void set_metrics_record_" << n << "(void* handle, b200_metrics_record* record);\n";
  os << "/* chained launches (several stages in one launch with an in-launch hand-over): 0 if every block saw its\n"
        " * producers so far; synchronises `stream` */\n";
  os << "int chain_status_" << n << "(void* handle, void* stream);\n";
  os << "int run_" << n << "(void* handle, const b200_field_t* fields, int nfields, void* stream);\n";
  os << "int run_" << n << "_from_c_host(void* handle, DAWN_B200_FLOAT_TYPE* const* host_fields, int nfields, void* stream);\n";
  os << "int run_" << n << "_from_fort_host(void* handle, DAWN_B200_FLOAT_TYPE* const* host_fields, int nfields, void* stream);\n";
  os << "int verify_" << n << "(void* handle, const b200_field_t* fields, const DAWN_B200_FLOAT_TYPE* const* ref_outputs, "
        "const double* rel_tol, const double* abs_tol, int iteration, b200_verification_metrics* metrics, void* stream);\n";
  os << "int run_and_verify_" << n << "(void* handle, const b200_field_t* fields, int nfields, "
        "const DAWN_B200_FLOAT_TYPE* const* ref_outputs, const double* rel_tol, const double* abs_tol, int iteration, "
        "b200_verification_metrics* metrics, void* stream);\n";
  os << "void free_" << n << "(void* handle);\n";
  for(auto const& g : inst.globals) {
    os << "void set_" << n << "_" << g.first << "(void* handle, double value);\n";
    os << "double get_" << n << "_" << g.first << "(void* handle);\n";
  }
  os << "#ifdef __cplusplus\n}\n#endif\n";
  return os.str();
}

std::string icoF90Interface(const iir::Instantiation& inst) {
  std::string n;
  for(char c : inst.name)
    n += (std::isalnum(static_cast<unsigned char>(c)) || c == '_') ? c : '_';
  std::ostringstream os;
  os << "! generated by dawn-b200 (b200-ico): Fortran interface of stencil " << inst.name << "\n";
  os << "module " << n << "_b200\n  use, intrinsic :: iso_c_binding\n  implicit none\n";
  os << "  type, bind(c) :: b200_field_t\n    type(c_ptr) :: data\n    integer(c_int) :: stride_j\n"
        "    integer(c_int) :: stride_k\n  end type b200_field_t\n";
  os << "  interface\n";
  os << "    integer(c_int) function setup_" << n << "(handle, mesh, k_size, stream) bind(c)\n"
        "      import :: c_int, c_ptr\n      type(c_ptr), intent(out) :: handle\n"
        "      type(c_ptr), value :: mesh, stream\n      integer(c_int), value :: k_size\n"
        "    end function\n";
  for(const char* fn : {"run_", "run_from_fort_host_"}) {
    bool host = std::string(fn) != "run_";
    std::string name = host ? "run_" + n + "_from_fort_host" : "run_" + n;
    os << "    integer(c_int) function " << name << "(handle, fields, nfields, stream) bind(c)\n"
       << "      import :: c_int, c_ptr" << (host ? "" : ", b200_field_t") << "\n"
       << "      type(c_ptr), value :: handle, stream\n"
       << (host ? "      type(c_ptr), intent(in) :: fields(*)\n" : "      type(b200_field_t), intent(in) :: fields(*)\n")
       << "      integer(c_int), value :: nfields\n    end function\n";
  }
  os << "    subroutine free_" << n << "(handle) bind(c)\n      import :: c_ptr\n"
        "      type(c_ptr), value :: handle\n    end subroutine\n";
  os << "    subroutine set_splitter_index_" << n << "(handle, loc, space, offset, index) bind(c)\n"
        "      import :: c_ptr, c_int\n      type(c_ptr), value :: handle\n"
        "      integer(c_int), value :: loc, space, offset, index\n    end subroutine\n";
  os << "  end interface\n";
  os << "  ! field order of fields(*):";
  for(int f : inst.apiFieldIDs)
    os << " " << inst.nameOf(f);
  os << "\nend module " << n << "_b200\n";
  return os.str();
}

TranslationUnit runIco(const std::map<std::string, iir::Instantiation>& ctx, const Options& options) {
  TranslationUnit tu;
  for(auto const& kv : ctx) {
    if(!kv.second.unstructured)
      throw std::runtime_error("b200-ico: instantiation '" + kv.first +
                               "' is Cartesian; use the b200 backend");
    IcoEmitter em(kv.second, options);
    tu.stencils[kv.second.name] = em.generate();
    if(tu.filename.empty())
      tu.filename = kv.second.filename;
  }
  tu.ppDefines = {
      "#define DAWN_GENERATED 1",
      "#undef DAWN_BACKEND_T",
      "#define DAWN_BACKEND_T B200ICO",
      // precision: chosen at generation time (option single_precision), checked at compile time; cf. the
      // Cartesian emitter
      options.SinglePrecision ? "#ifndef DAWN_PRECISION\n#define DAWN_PRECISION 0 /* DAWN_SINGLE_PRECISION */\n#endif"
                              : "#ifndef DAWN_PRECISION\n#define DAWN_PRECISION 1 /* DAWN_DOUBLE_PRECISION */\n#endif",
      "#include <driver-includes/unstructured_includes.hpp>",
      std::string("static_assert(sizeof(::dawn::float_type) == ") + (options.SinglePrecision ? "4" : "8") +
          ", \"this translation unit was generated for " + (options.SinglePrecision ? "single" : "double") +
          " precision (b200-ico codegen option single_precision)\");",
  };
  return tu;
}

} // namespace codegen
} // namespace b200
