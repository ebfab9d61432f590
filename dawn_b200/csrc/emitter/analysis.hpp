// Derived information the IIR wire format does not carry; recomputed from the statement trees.
//   * per-stage field accesses (read extents, written set)     ~ iir::DoMethod::update / Stage::update
//   * stage extents: backward propagation of read extents over all stages of a stencil
//                                      dawn/src/dawn/IIR/StencilInstantiation.cpp:304-373
//   * accumulated field extents exported as `static constexpr cartesian_extent <f>_extent`
//                                      dawn/src/dawn/CodeGen/CodeGen.cpp:509-522
//   * interval partition of a multistage (disjoint k ranges)   dawn/src/dawn/IIR/Interval.cpp:224-330
#pragma once
#include "iir.hpp"

#include <algorithm>
#include <functional>
#include <map>
#include <set>
#include <vector>

namespace b200 {
namespace analysis {

using iir::Expr;
using iir::ExprPtr;
using iir::Extent;
using iir::Stmt;
using iir::StmtPtr;

struct FieldUse {
  bool read = false, written = false;
  bool hasReadExtent = false;
  Extent readExtent; // merged over all reads (single offset => [o,o], like iir::Extents)
  bool readNeighbour = false; // unstructured: read through a reduction (has_offset)
};

struct StageInfo {
  std::map<int, FieldUse> fields; // by accessID, fields only
  std::set<int> globalsRead;
};

inline void visitExpr(const ExprPtr& e, const std::function<void(const Expr&, bool isWrite)>& fn) {
  if(!e)
    return;
  if(e->kind == Expr::Assign) {
    if(e->kids[0]->kind == Expr::Field || e->kids[0]->kind == Expr::Var) {
      fn(*e->kids[0], true);
      if(e->op != "=")
        fn(*e->kids[0], false);
    }
    visitExpr(e->kids[1], fn);
    return;
  }
  if(e->kind == Expr::Field || e->kind == Expr::Var) {
    fn(*e, false);
    if(e->kind == Expr::Field && !e->kFrom.empty()) { // vertical indirection also reads the index field
      Expr idx;
      idx.kind = Expr::Field;
      idx.name = e->kFrom;
      idx.accessID = e->kFromID;
      fn(idx, false);
    }
    return;
  }
  for(auto const& k : e->kids)
    visitExpr(k, fn);
}

inline void visitStmts(const std::vector<StmtPtr>& stmts,
                       const std::function<void(const Expr&, bool isWrite)>& fn) {
  for(auto const& s : stmts) {
    switch(s->kind) {
    case Stmt::ExprStmt: visitExpr(s->expr, fn); break;
    case Stmt::VarDecl: visitExpr(s->expr, fn); break;
    case Stmt::If:
      visitExpr(s->expr, fn);
      visitStmts(s->body, fn);
      visitStmts(s->orelse, fn);
      break;
    case Stmt::Block:
    case Stmt::Loop: visitStmts(s->body, fn); break;
    }
  }
}

inline void accumulate(const iir::Instantiation& inst, const std::vector<StmtPtr>& stmts,
                       StageInfo& info) {
  visitStmts(stmts, [&](const Expr& e, bool isWrite) {
    if(e.kind == Expr::Var) {
      if(e.isExternal && !isWrite)
        info.globalsRead.insert(e.accessID);
      return;
    }
    FieldUse& u = info.fields[e.accessID];
    if(isWrite) {
      u.written = true;
      return;
    }
    u.read = true;
    Extent ext{e.di, e.di, e.dj, e.dj, e.dk, e.dk};
    if(u.hasReadExtent)
      u.readExtent.merge(ext);
    else
      u.readExtent = ext, u.hasReadExtent = true;
    u.readNeighbour = u.readNeighbour || e.hasOffset;
  });
  (void)inst;
}

// A Parallel multistage executes its levels in any order (level blocks of different thread blocks run
// concurrently), so a field it writes must not be read at another level inside it - Dawn's
// PassSetLoopOrder would have made such a multistage Forward or Backward.  Diagnosed instead of racing.
inline void checkParallelLevels(const iir::Instantiation& inst, const iir::MultiStage& ms) {
  if(ms.loopOrder != iir::LoopOrder::Parallel)
    return;
  std::set<int> written;
  for(auto const& st : ms.stages)
    for(auto const& dm : st.doMethods)
      visitStmts(dm.stmts, [&](const Expr& e, bool w) {
        if(w && e.kind == Expr::Field)
          written.insert(e.accessID);
      });
  for(auto const& st : ms.stages)
    for(auto const& dm : st.doMethods)
      visitStmts(dm.stmts, [&](const Expr& e, bool w) {
        if(!w && e.kind == Expr::Field && e.dk != 0 && written.count(e.accessID))
          throw std::runtime_error("IIR: field '" + inst.nameOf(e.accessID) +
                                   "' is written by Parallel multistage " + std::to_string(ms.id) +
                                   " and read at a vertical offset inside it (the loop order must "
                                   "be Forward or Backward)");
      });
}

// The invariant every Dawn backend relies on (PassStageSplitter / PassFieldVersioning establish it): a
// stage never reads, at a horizontal offset, a field it writes itself - the points of a stage run in
// parallel.  IIR that has not been through those passes (e.g. the inputs of the optimizer's own unit
// tests) violates it and would race; it is rejected.
inline void checkStageIndependence(const iir::Instantiation& inst, const iir::Stage& st) {
  std::set<int> written;
  for(auto const& dm : st.doMethods)
    visitStmts(dm.stmts, [&](const Expr& e, bool w) {
      if(w && e.kind == Expr::Field)
        written.insert(e.accessID);
    });
  for(auto const& dm : st.doMethods)
    visitStmts(dm.stmts, [&](const Expr& e, bool w) {
      if(!w && e.kind == Expr::Field && (e.di != 0 || e.dj != 0) && written.count(e.accessID))
        throw std::runtime_error("IIR: stage " + std::to_string(st.id) + " reads field '" +
                                 inst.nameOf(e.accessID) + "' at a horizontal offset and writes it "
                                 "(the stage has to be split / the field versioned first)");
    });
}

inline StageInfo stageInfo(const iir::Instantiation& inst, const iir::Stage& st) {
  StageInfo info;
  for(auto const& dm : st.doMethods)
    accumulate(inst, dm.stmts, info);
  return info;
}

// Sets Stage::extent for every stage of every stencil (horizontal part only, as in the reference).
inline void computeStageExtents(iir::Instantiation& inst) {
  for(auto& stencil : inst.stencils) {
    std::vector<iir::Stage*> stages;
    for(auto& ms : stencil.multiStages)
      for(auto& st : ms.stages)
        stages.push_back(&st);
    std::vector<StageInfo> infos;
    for(auto* st : stages) {
      infos.push_back(stageInfo(inst, *st));
      st->extent = Extent{};
    }
    for(int i = static_cast<int>(stages.size()) - 1; i >= 0; --i) {
      iir::Stage& from = *stages[i];
      if(from.hasIRange || from.hasJRange) { // user-defined global iteration space: never extended
        from.extent = Extent{};
        continue;
      }
      for(auto const& fp : infos[i].fields) {
        Extent fieldExtent = fp.second.hasReadExtent ? fp.second.readExtent.add(from.extent)
                                                     : from.extent;
        if(!fp.second.hasReadExtent)
          fieldExtent = Extent{}.add(from.extent);
        for(int j = i - 1; j >= 0; --j) {
          auto it = infos[j].fields.find(fp.first);
          if(it == infos[j].fields.end() || !it->second.written)
            continue;
          Extent ext = stages[j]->extent;
          ext.merge(fieldExtent);
          ext.kMinus = ext.kPlus = 0; // redundant computation is horizontal only
          stages[j]->extent = ext;
        }
      }
    }
  }
}

// Accumulated access extent per field over a whole stencil (read extent + extent of the reading
// stage; vertical part from the accesses).
inline std::map<int, Extent> fieldExtents(const iir::Instantiation& inst, const iir::Stencil& st) {
  std::map<int, Extent> out;
  for(auto const& ms : st.multiStages)
    for(auto const& stage : ms.stages) {
      StageInfo info = stageInfo(inst, stage);
      for(auto const& fp : info.fields) {
        Extent e{};
        if(fp.second.hasReadExtent) {
          e = fp.second.readExtent;
          e.iMinus += stage.extent.iMinus, e.iPlus += stage.extent.iPlus;
          e.jMinus += stage.extent.jMinus, e.jPlus += stage.extent.jPlus;
        }
        auto it = out.find(fp.first);
        if(it == out.end())
          out[fp.first] = Extent{}, it = out.find(fp.first);
        it->second.merge(e);
      }
    }
  return out;
}

struct KRange {
  int lo, hi; // symbolic bounds (level + offset with End = 1<<20)
};

inline std::vector<KRange> partition(const iir::MultiStage& ms) {
  std::vector<KRange> ivs;
  for(auto const& st : ms.stages)
    for(auto const& dm : st.doMethods)
      ivs.push_back({dm.interval.lowerBound(), dm.interval.upperBound()});
  std::set<int> cuts;
  for(auto const& iv : ivs) {
    cuts.insert(iv.lo);
    cuts.insert(iv.hi + 1);
  }
  std::vector<int> c(cuts.begin(), cuts.end());
  std::vector<KRange> out;
  for(size_t n = 0; n + 1 < c.size(); ++n) {
    int a = c[n], b = c[n + 1] - 1;
    bool covered = false;
    for(auto const& iv : ivs)
      covered = covered || (iv.lo <= a && b <= iv.hi);
    if(covered)
      out.push_back({a, b});
  }
  if(ms.loopOrder == iir::LoopOrder::Backward)
    std::reverse(out.begin(), out.end());
  return out;
}

// C expression for a symbolic level bound given variables kmin/kmax in scope.
inline std::string boundExpr(int bound) {
  if(bound >= iir::Interval::End - (1 << 19))
    return "kmax + " + std::to_string(bound - iir::Interval::End);
  return "kmin + " + std::to_string(bound);
}

} // namespace analysis
} // namespace b200
