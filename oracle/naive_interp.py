"""ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported by the product (`dawn_b200/`).

CPU restatement of what the reference's *naive* backends execute for a given IIR, written as a
NumPy interpreter over the same IIR JSON the B200 emitter consumes.

Cartesian (`cxxnaive`) semantics followed, file:line in /root/reference:
  * run(): per MultiStage, per partitioned k-interval (reversed for Backward loop order), per k,
    per Stage: `for i in [iMin+ext.iMinus, iMax+ext.iPlus]: for j in [...]`: all statements of every
    DoMethod whose interval overlaps                 dawn/src/dawn/CodeGen/CXXNaive/CXXNaiveCodeGen.cpp:380-495
  * loop bounds iMin=iminus, iMax=isize-iplus-1 ...   CXXNaiveCodeGen.cpp:381-387
  * interval bounds: Start -> kMin+level+offset, End -> kMax+offset   CXXNaiveCodeGen.cpp:56-67
  * interval partition                               dawn/src/dawn/IIR/Interval.cpp:224-330
  * stage extents (backward propagation of read extents over all stages of a stencil)
                                                     dawn/src/dawn/IIR/StencilInstantiation.cpp:304-373
  * expressions are evaluated exactly as the fully parenthesised tree, one IEEE operation per
    node (no contraction)                            dawn/src/dawn/CodeGen/ASTCodeGenCXX.cpp:109-153
  * field(i+a, j+b, k+c) view indexing, masked dimensions ignored
                                                     dawn/src/dawn/CodeGen/CXXNaive/ASTStencilBody.cpp:121-178
  * temporaries are full 3-D storages                CXXNaiveCodeGen.cpp:318-345
  * global iteration spaces (i_range/j_range): `checkOffset(lo, hi, globalOffset + idx)`
                                                     CXXNaiveCodeGen.cpp:455-486, CodeGen.cpp:447-470

Unstructured (`naive-ico`) semantics:
  * per k, per stage `for loc in getCells/Edges/Vertices(mesh)`; reductions iterate the neighbour
    list of `chain` in mesh-library order, accumulating from `init`, weights indexed by position
                                                     dawn/src/dawn/CodeGen/CXXNaive-ico/CXXNaiveCodeGen.cpp:366-617
                                                     dawn/src/interface/toylib_interface.hpp:280-323

Because a Dawn stage never reads at a horizontal offset something it writes itself (that is what
stage splitting guarantees, `dawn/src/dawn/Optimizer/PassStageSplitter.cpp`), executing a stage
statement-by-statement over the whole (i,j) plane is equivalent to the reference's point-by-point
loop; NumPy float64 elementwise kernels perform one correctly rounded IEEE operation per node.

Parity status: PINNED — checked against the reference's own pre-generated cxxnaive code compiled
from /root/reference (oracle/_ref) on `conditional_stencil.iir`, `update_dz_c.iir`
(tests/test_oracle_vs_ref.py) and against committed golden vectors produced by that build
(tests/golden/).
"""
from __future__ import annotations

import json
import math
from typing import Dict, List, Optional, Tuple

import numpy as np

END = 1 << 20
_HALF = 1 << 19


# ------------------------------------------------------------------------------------------------
# IIR helpers
# ------------------------------------------------------------------------------------------------
# proto field names that contain an underscore (IIR.proto / statements.proto): proto3 JSON writers other
# than Dawn's IIRSerializer emit them in lowerCamelCase (block_stmt -> blockStmt), readers accept both
# (VerticalRegion.loop_order is left out on purpose: MultiStage has a field spelled loopOrder)
_SNAKE_FIELDS = """
mask_cart_i mask_cart_j iter_space cartesian_horizontal_dimension unstructured_horizontal_dimension mask_k
is_temporary field_dimensions field_value direction_value offset_value special_lower_level lower_level
special_upper_level upper_level lower_offset upper_offset type_id builtin_type is_const is_volatile
i_range j_range i_extent j_extent has_extent cartesian_extent unstructured_extent zero_extent vertical_extent
include_center block_stmt expr_stmt return_stmt var_decl_stmt stencil_call_decl_stmt
vertical_region_decl_stmt boundary_condition_decl_stmt if_stmt loop_stmt unary_operator binary_operator
assignment_expr ternary_operator fun_call_expr stencil_fun_call_expr stencil_fun_arg_expr var_access_expr
field_access_expr literal_access_expr reduction_over_neighbor_expr loop_descriptor_chain
loop_descriptor_general loop_descriptor init_list var_decl_stmt_data vertical_region stencil_call cond_part
then_part else_part argument_index is_external i_offset j_offset has_offset vertical_shift
vertical_indirection vertical_indirection_data cartesian_offset unstructured_offset zero_offset argument_map
argument_offset negate_offset boolean_value integer_value double_value string_value float_value is_constexpr
dense_location_type sparse_part
""".split()
_CAMEL = {}
for _n in _SNAKE_FIELDS:
    _p = _n.split("_")
    _CAMEL[_p[0] + "".join(w[:1].upper() + w[1:] for w in _p[1:])] = _n


def _normalize_keys(x):
    if isinstance(x, dict):
        return {_CAMEL.get(k, k): _normalize_keys(v) for k, v in x.items()}
    if isinstance(x, list):
        return [_normalize_keys(v) for v in x]
    return x


def load_iir(path_or_dict):
    if isinstance(path_or_dict, dict):
        return path_or_dict
    with open(path_or_dict) as f:
        return _normalize_keys(json.load(f))


def _kind(node: dict) -> Tuple[str, dict]:
    (k, v), = node.items()
    return k, v


def interval_bounds(iv: dict) -> Tuple[int, int]:
    lo_level = END if iv.get("special_lower_level") == "End" else iv.get("lower_level", 0)
    hi_level = END if iv.get("special_upper_level") == "End" else iv.get("upper_level", 0)
    if "special_lower_level" in iv and iv["special_lower_level"] == "Start":
        lo_level = 0
    if "special_upper_level" in iv and iv["special_upper_level"] == "Start":
        hi_level = 0
    return lo_level + iv.get("lower_offset", 0), hi_level + iv.get("upper_offset", 0)


def partition(bounds: List[Tuple[int, int]]) -> List[Tuple[int, int]]:
    """Disjoint cover of the union of closed integer intervals, split at every bound."""
    cuts = set()
    for lo, hi in bounds:
        cuts.add(lo)
        cuts.add(hi + 1)
    cuts = sorted(cuts)
    out = []
    for a, b in zip(cuts[:-1], cuts[1:]):
        if any(lo <= a and b - 1 <= hi for lo, hi in bounds):
            out.append((a, b - 1))
    return out


def resolve(bound: int, k_min: int, k_max: int) -> int:
    if bound >= END - _HALF:
        return k_max + (bound - END)
    return k_min + bound


class Ext:
    __slots__ = ("im", "ip", "jm", "jp")

    def __init__(self, im=0, ip=0, jm=0, jp=0):
        self.im, self.ip, self.jm, self.jp = im, ip, jm, jp

    def merge(self, o):
        self.im, self.ip = min(self.im, o.im), max(self.ip, o.ip)
        self.jm, self.jp = min(self.jm, o.jm), max(self.jp, o.jp)

    def add(self, o):
        return Ext(self.im + o.im, self.ip + o.ip, self.jm + o.jm, self.jp + o.jp)

    def tup(self):
        return (self.im, self.ip, self.jm, self.jp)


def _walk_exprs(node, fn):
    """Call fn(kind, payload, is_write) on every field_access_expr below an expression node."""
    k, v = _kind(node)
    if k == "field_access_expr":
        fn(v)
    elif k == "binary_operator":
        _walk_exprs(v["left"], fn)
        _walk_exprs(v["right"], fn)
    elif k == "assignment_expr":
        _walk_exprs(v["left"], fn)
        _walk_exprs(v["right"], fn)
    elif k == "unary_operator":
        _walk_exprs(v["operand"], fn)
    elif k == "ternary_operator":
        _walk_exprs(v["cond"], fn)
        _walk_exprs(v["left"], fn)
        _walk_exprs(v["right"], fn)
    elif k == "fun_call_expr":
        for a in v.get("arguments", []):
            _walk_exprs(a, fn)
    elif k == "reduction_over_neighbor_expr":
        _walk_exprs(v["rhs"], fn)
        _walk_exprs(v["init"], fn)
        for w in v.get("weights", []):
            _walk_exprs(w, fn)


def _stmt_accesses(stmt, reads: Dict[int, Ext], writes: set):
    k, v = _kind(stmt)

    def rd(fa):
        off = fa.get("cartesian_offset", {})
        e = Ext(off.get("i_offset", 0), off.get("i_offset", 0), off.get("j_offset", 0),
                off.get("j_offset", 0))
        aid = fa["data"]["accessID"]
        if aid in reads:
            reads[aid].merge(e)
        else:
            reads[aid] = e

    if k == "expr_stmt":
        ek, ev = _kind(v["expr"])
        if ek == "assignment_expr":
            lk, lv = _kind(ev["left"])
            if lk == "field_access_expr":
                writes.add(lv["data"]["accessID"])
                if ev.get("op", "=") != "=":
                    rd(lv)
            _walk_exprs(ev["right"], rd)
        else:
            _walk_exprs(v["expr"], rd)
    elif k == "var_decl_stmt":
        for e in v.get("init_list", []):
            _walk_exprs(e, rd)
    elif k == "if_stmt":
        _stmt_accesses(v["cond_part"], reads, writes)
        _stmt_accesses(v["then_part"], reads, writes)
        if "else_part" in v and v["else_part"]:
            _stmt_accesses(v["else_part"], reads, writes)
    elif k == "block_stmt":
        for s in v.get("statements", []):
            _stmt_accesses(s, reads, writes)


def stage_extents(stencil: dict) -> List[Ext]:
    """StencilInstantiation.cpp:304-373 restated; returns one Ext per stage in stencil order."""
    stages = [s for ms in stencil["multiStages"] for s in ms["stages"]]
    info = []
    for s in stages:
        reads: Dict[int, Ext] = {}
        writes: set = set()
        for dm in s["doMethods"]:
            _stmt_accesses(dm["ast"], reads, writes)
        info.append((reads, writes))
    exts = [Ext() for _ in stages]
    for i in range(len(stages) - 1, -1, -1):
        if "i_range" in stages[i] or "j_range" in stages[i]:
            exts[i] = Ext()
            continue
        reads, writes_i = info[i]
        # every field of the stage participates (written-only fields contribute zero extents)
        fields = dict(reads)
        for w in writes_i:
            fields.setdefault(w, Ext())
        for aid, rext in fields.items():
            fext = rext.add(exts[i])
            for j in range(i - 1, -1, -1):
                if aid in info[j][1]:
                    exts[j].merge(fext)
    return exts


# ------------------------------------------------------------------------------------------------
# Cartesian interpreter
# ------------------------------------------------------------------------------------------------
_MATH = {
    "sqrt": np.sqrt, "fabs": np.fabs, "abs": np.fabs, "exp": np.exp, "log": np.log,
    "pow": np.power, "min": np.minimum, "max": np.maximum, "floor": np.floor, "ceil": np.ceil,
    "trunc": np.trunc, "sin": np.sin, "cos": np.cos, "tan": np.tan, "asin": np.arcsin,
    "acos": np.arccos, "atan": np.arctan, "fmod": np.fmod, "isnan": np.isnan, "isinf": np.isinf,
}

_BIN = {
    "+": np.add, "-": np.subtract, "*": np.multiply, "/": np.divide,
    "<": np.less, ">": np.greater, "<=": np.less_equal, ">=": np.greater_equal,
    "==": np.equal, "!=": np.not_equal, "&&": np.logical_and, "||": np.logical_or,
}


class CartesianRun:
    """Executes one IIR on storages given as dict name -> ndarray.

    Storage convention: arrays are indexed [k, j, i] (i fastest).  Fields with masked dimensions
    have extent 1 along them (e.g. a `storage_j` is shape (1, nj, 1)).
    `dom` = (isize, jsize, ksize) *including* halos, `halos` = (iminus, iplus, jminus, jplus,
    kminus, kplus) as in driver-includes/domain.hpp:42-58.
    """

    def __init__(self, iir, dom, halos=(3, 3, 3, 3, 0, 0), globals_: Optional[dict] = None,
                 global_offsets=(0, 0)):
        self.iir = load_iir(iir)
        self.meta = self.iir["metadata"]
        self.ir = self.iir["internalIR"]
        self.dom = dom
        self.halos = halos
        self.names = {int(k): v for k, v in self.meta["accessIDToName"].items()}
        self.globals = {}
        for name, gv in self.ir.get("globalVariableToValue", {}).items():
            self.globals[name] = gv.get("value", 0)
        if globals_:
            self.globals.update(globals_)
        self.global_offsets = global_offsets
        self.tmp_ids = set(self.meta.get("temporaryFieldIDs", []))

    def run(self, fields: Dict[str, np.ndarray]):
        ni, nj, nk = self.dom
        im, ip, jm, jp, km, kp = self.halos
        self.iMin, self.iMax = im, ni - ip - 1
        self.jMin, self.jMax = jm, nj - jp - 1
        self.kMin, self.kMax = km, nk - kp - 1
        self.f = dict(fields)
        # ::dawn::float_type of this run: float if the caller's storages are (driver-includes/defs.hpp:47-81);
        # temporaries and float_type locals follow it, literals are weak scalars (`(float_type)2.0 * x`)
        self.ft = _float_type(fields)
        for st in self._stencil_sequence():
            exts = stage_extents(st)
            # temporaries: full 3-D storages (CXXNaiveCodeGen.cpp:318-345), +1 safety level
            for aid in self.tmp_ids:
                nm = self.names[aid]
                if nm not in self.f:
                    self.f[nm] = np.zeros((nk + 1, nj + 1, ni + 1), dtype=self.ft)
            sidx = 0
            for ms in st["multiStages"]:
                stages = ms["stages"]
                s_exts = exts[sidx:sidx + len(stages)]
                sidx += len(stages)
                self._run_ms(ms, stages, s_exts)
        return self.f

    def _stencil_sequence(self):
        """Stencils in the order the wrapper's run() executes them: the controlFlowStatements of the
        IIR (stencil calls, possibly inside if-statements over globals and local variables of the
        stencil description; reference: CodeGen::generateStencilWrapperRun + ASTStencilDesc,
        CXXNaiveCodeGen.cpp:232-259), or all stencils in order when there are none."""
        by_id = {st["stencilID"]: st for st in self.ir["stencils"]}
        flow = self.ir.get("controlFlowStatements", [])
        if not flow:
            yield from self.ir["stencils"]
            return
        local = {}

        def ev(node):
            kd, v = _kind(node)
            if kd == "literal_access_expr":
                t = v.get("type", {}).get("type_id", "Float")
                return int(v["value"]) if t == "Integer" else (v["value"] in ("true", "1") if t == "Boolean"
                                                               else float(v["value"]))
            if kd == "var_access_expr":
                return self.globals[v["name"]] if v.get("is_external", False) else local[v["name"]]
            if kd == "binary_operator":
                return _BIN[v["op"]](ev(v["left"]), ev(v["right"]))
            if kd == "unary_operator":
                a = ev(v["operand"])
                return -a if v["op"] == "-" else ((not a) if v["op"] == "!" else a)
            if kd == "ternary_operator":
                return ev(v["left"]) if ev(v["cond"]) else ev(v["right"])
            raise NotImplementedError("control flow expression " + kd)

        def ex(node):
            # a generator: the caller runs each stencil as it is reached, so assignments to globals
            # between two calls (globals_dependencies_issue401) take effect in order
            kd, v = _kind(node)
            if kd == "stencil_call_decl_stmt":
                yield by_id[int(v["stencil_call"]["callee"].rsplit("_", 1)[1])]
            elif kd == "block_stmt":
                for s in v.get("statements", []):
                    yield from ex(s)
            elif kd == "if_stmt":
                _, cv = _kind(v["cond_part"])
                if ev(cv["expr"]):
                    yield from ex(v["then_part"])
                elif v.get("else_part"):
                    yield from ex(v["else_part"])
            elif kd == "var_decl_stmt":
                init = v.get("init_list", [])
                local[v["name"]] = ev(init[0]) if init else 0.0
            elif kd == "expr_stmt":
                ek, e2 = _kind(v["expr"])
                assert ek == "assignment_expr", ek
                _, lv = _kind(e2["left"])
                rhs = ev(e2["right"])
                op = e2.get("op", "=")
                store = self.globals if lv.get("is_external", False) else local
                store[lv["name"]] = rhs if op == "=" else _BIN[op[0]](store[lv["name"]], rhs)
            else:
                raise NotImplementedError("control flow statement " + kd)

        for s in flow:
            yield from ex(s)

    def _run_ms(self, ms, stages, s_exts):
        backward = ms.get("loopOrder", "Forward") == "Backward"
        bounds = [interval_bounds(dm["interval"]) for s in stages for dm in s["doMethods"]]
        parts = partition(bounds)
        if backward:
            parts = parts[::-1]
        for lo, hi in parts:
            k0, k1 = resolve(lo, self.kMin, self.kMax), resolve(hi, self.kMin, self.kMax)
            ks = range(k1, k0 - 1, -1) if backward else range(k0, k1 + 1)
            for k in ks:
                for stage, ext in zip(stages, s_exts):
                    dms = [dm for dm in stage["doMethods"]
                           if _overlap(interval_bounds(dm["interval"]), (lo, hi))]
                    if not dms:
                        continue
                    self._run_stage(stage, dms, ext, k)

    def _run_stage(self, stage, dms, ext: Ext, k):
        self.k = k
        self.i0, self.i1 = self.iMin + ext.im, self.iMax + ext.ip
        self.j0, self.j1 = self.jMin + ext.jm, self.jMax + ext.jp
        if self.i1 < self.i0 or self.j1 < self.j0:
            return
        self.locals: Dict[int, np.ndarray] = {}
        mask = None
        if "i_range" in stage or "j_range" in stage:
            # CXXNaiveCodeGen.cpp:455-486 + generated ctor (`stageNGlobalIIndices({dom.iminus()+lo,
            # dom.iminus()+hi})`, reference/global_indexing.cpp:72-73): lo <= idx < hi with the
            # interval's End level mapped to the global size.
            shape = (self.j1 - self.j0 + 1, self.i1 - self.i0 + 1)
            mask = np.ones(shape, dtype=bool)
            ii = np.arange(self.i0, self.i1 + 1)[None, :] + self.global_offsets[0]
            jj = np.arange(self.j0, self.j1 + 1)[:, None] + self.global_offsets[1]
            if "i_range" in stage:
                lo, hi = self._global_range(stage["i_range"], self.dom[0], self.halos[0], self.halos[1])
                mask &= (ii >= lo) & (ii < hi)
            if "j_range" in stage:
                lo, hi = self._global_range(stage["j_range"], self.dom[1], self.halos[2], self.halos[3])
                mask &= (jj >= lo) & (jj < hi)
        for dm in dms:
            self._block(dm["ast"], mask)

    @staticmethod
    def _global_range(iv, size, minus, plus):
        def one(special, level, off):
            if special == "End":
                return size - plus + off
            base = 0 if special == "Start" else level
            return minus + base + off
        lo = one(iv.get("special_lower_level"), iv.get("lower_level", 0), iv.get("lower_offset", 0))
        hi = one(iv.get("special_upper_level"), iv.get("upper_level", 0), iv.get("upper_offset", 0))
        return lo, hi

    # -- statements -------------------------------------------------------------------------------
    def _block(self, node, mask):
        k, v = _kind(node)
        if k != "block_stmt":
            return self._stmt(node, mask)
        for s in v.get("statements", []):
            self._stmt(s, mask)

    def _stmt(self, node, mask):
        k, v = _kind(node)
        if k == "block_stmt":
            self._block(node, mask)
        elif k == "expr_stmt":
            ek, ev = _kind(v["expr"])
            if ek == "assignment_expr":
                self._assign(ev, mask)
            else:
                self._eval(v["expr"])
        elif k == "var_decl_stmt":
            aid = v["var_decl_stmt_data"]["accessID"]
            init = v.get("init_list", [])
            val = self._eval(init[0]) if init else 0.0
            tid = v.get("type", {}).get("builtin_type", {}).get("type_id", "Float")
            val = self._cast(val, tid)
            if tid in ("Float", "Double") and not isinstance(val, (bool, np.bool_)):
                val = np.asarray(val, dtype=self.ft)   # both spell ::dawn::float_type (ASTCodeGenCXX.cpp:163-181)
            shape = (self.j1 - self.j0 + 1, self.i1 - self.i0 + 1)
            self.locals[aid] = np.broadcast_to(np.asarray(val), shape).copy()
        elif k == "if_stmt":
            ck, cv = _kind(v["cond_part"])
            cond = self._eval(cv["expr"]) if ck == "expr_stmt" else self._eval(v["cond_part"])
            cond = np.asarray(cond).astype(bool)
            m_then = cond if mask is None else (mask & cond)
            self._block(v["then_part"], m_then)
            if "else_part" in v and v["else_part"]:
                m_else = ~cond if mask is None else (mask & ~cond)
                self._block(v["else_part"], m_else)
        else:
            raise NotImplementedError(k)

    @staticmethod
    def _cast(val, tid):
        if tid == "Integer":
            return np.trunc(val)
        if tid == "Boolean":
            return np.asarray(val).astype(bool)
        return val

    def _assign(self, ev, mask):
        op = ev.get("op", "=")
        rhs = self._eval(ev["right"])
        lk, lv = _kind(ev["left"])
        if lk == "var_access_expr":
            aid = lv["data"]["accessID"]
            if lv.get("is_external", False):
                raise NotImplementedError("assignment to global")
            cur = self.locals[aid]
            new = rhs if op == "=" else _BIN[op[0]](cur, rhs)
            self.locals[aid] = np.where(mask, new, cur) if mask is not None else \
                np.broadcast_to(np.asarray(new, dtype=float), cur.shape).copy()
            return
        assert lk == "field_access_expr"
        view = self._field_view(lv, write=True)
        new = rhs if op == "=" else _BIN[op[0]](view, rhs)
        if mask is not None:
            m = np.broadcast_to(mask, (self.j1 - self.j0 + 1, self.i1 - self.i0 + 1))
            if view.shape != m.shape:  # masked-dimension storage
                m = m[:view.shape[0], :view.shape[1]]
            view[...] = np.where(m, new, view)
        else:
            view[...] = new

    # -- expressions ------------------------------------------------------------------------------
    def _field_view(self, fa, write=False):
        name = fa["name"]
        arr = self.f[name]
        off = fa.get("cartesian_offset", {})
        di, dj = off.get("i_offset", 0), off.get("j_offset", 0)
        dk = fa.get("vertical_shift", 0)
        nk_a, nj_a, ni_a = arr.shape
        kk = self.k + dk if nk_a > 1 else 0
        js = slice(self.j0 + dj, self.j1 + dj + 1) if nj_a > 1 else slice(0, 1)
        is_ = slice(self.i0 + di, self.i1 + di + 1) if ni_a > 1 else slice(0, 1)
        return arr[kk, js, is_]

    def _eval(self, node):
        k, v = _kind(node)
        if k == "field_access_expr":
            return self._field_view(v)
        if k == "literal_access_expr":
            t = v.get("type", {}).get("type_id", "Float")
            s = v["value"]
            if t == "Boolean":
                return s in ("true", "1")
            if t == "Integer":
                return int(s)
            return float(s)
        if k == "var_access_expr":
            if v.get("is_external", False):
                return _global_value(self.globals[v["name"]])
            return self.locals[v["data"]["accessID"]]
        if k == "binary_operator":
            a, b = self._eval(v["left"]), self._eval(v["right"])
            op = v["op"]
            if op == "/" and isinstance(a, int) and isinstance(b, int):
                return int(a / b)  # C integer division truncates
            with np.errstate(all="ignore"):
                return _BIN[op](a, b)
        if k == "unary_operator":
            a = self._eval(v["operand"])
            op = v["op"]
            if op == "-":
                return np.negative(a) if not isinstance(a, (int, float)) else -a
            if op == "+":
                return a
            if op == "!":
                return np.logical_not(a)
            raise NotImplementedError(op)
        if k == "ternary_operator":
            c = self._eval(v["cond"])
            a, b = self._eval(v["left"]), self._eval(v["right"])
            if isinstance(c, (bool, np.bool_)):
                return a if c else b
            return np.where(c, a, b)
        if k == "fun_call_expr":
            callee = v["callee"].split("::")[-1]
            args = [self._eval(a) for a in v.get("arguments", [])]
            with np.errstate(all="ignore"):
                return _MATH[callee](*args)
        raise NotImplementedError(k)


def _global_value(g):
    """A global of type double stays a C `double` whatever ::dawn::float_type is (the globals struct keeps
    the declared types, CodeGen.cpp:110-136), so `field * global` is evaluated in double even in a single-
    precision build: a NumPy scalar (strong type), not a weak Python float."""
    return np.float64(g) if isinstance(g, float) else g


def _float_type(fields):
    """float32 if any floating-point storage of the run is single precision, else float64."""
    return np.float32 if any(isinstance(a, np.ndarray) and a.dtype == np.float32 for a in fields.values()) \
        else np.float64


def _overlap(a, b):
    return a[0] <= b[1] and b[0] <= a[1]


def run_cartesian(iir, dom, fields, halos=(3, 3, 3, 3, 0, 0), globals_=None, global_offsets=(0, 0)):
    return CartesianRun(iir, dom, halos, globals_, global_offsets).run(fields)


def field_extents(iir) -> Dict[str, Tuple[int, int, int, int, int, int]]:
    """Accumulated read extents per API field (`static constexpr cartesian_extent <f>_extent`,
    dawn/src/dawn/CodeGen/CodeGen.cpp:509-522): field read extent + extent of the reading stage."""
    iir = load_iir(iir)
    names = {int(k): v for k, v in iir["metadata"]["accessIDToName"].items()}
    out: Dict[str, List[int]] = {}
    for st in iir["internalIR"]["stencils"]:
        exts = stage_extents(st)
        stages = [s for ms in st["multiStages"] for s in ms["stages"]]
        for s, ext in zip(stages, exts):
            reads: Dict[int, Ext] = {}
            writes: set = set()
            kext: Dict[int, List[int]] = {}
            for dm in s["doMethods"]:
                _stmt_accesses(dm["ast"], reads, writes)
                _collect_k(dm["ast"], kext)
            for aid in set(reads) | writes:
                r = reads.get(aid, Ext()).add(ext) if aid in reads else Ext()
                kk = kext.get(aid, [0, 0])
                cur = out.setdefault(names[aid], [0, 0, 0, 0, 0, 0])
                new = [r.im, r.ip, r.jm, r.jp, kk[0], kk[1]]
                for n in (0, 2, 4):
                    cur[n] = min(cur[n], new[n])
                for n in (1, 3, 5):
                    cur[n] = max(cur[n], new[n])
    return {k: tuple(v) for k, v in out.items()}


def _collect_k(node, kext):
    def visit(n):
        if isinstance(n, dict):
            if "field_access_expr" in n:
                fa = n["field_access_expr"]
                aid = fa["data"]["accessID"]
                dk = fa.get("vertical_shift", 0)
                cur = kext.setdefault(aid, [0, 0])
                cur[0], cur[1] = min(cur[0], dk), max(cur[1], dk)
            for key, val in n.items():
                if key not in ("data", "loc"):
                    visit(val)
        elif isinstance(n, list):
            for x in n:
                visit(x)
    visit(node)


# ------------------------------------------------------------------------------------------------
# Unstructured (naive-ico) interpreter
# ------------------------------------------------------------------------------------------------
_LOC = {"Cell": 0, "Edge": 1, "Vertex": 2}


class _Ctx:
    """Where an expression is evaluated: `center` = element ids the enclosing neighbour loop /
    reduction runs around (the stage's own elements at top level), `nb` = current neighbour ids
    (None outside loops / reductions), `nb_idx` = position in the neighbour list, `center_loc` =
    location type of `center`."""
    __slots__ = ("center", "center_loc", "nb", "nb_idx")

    def __init__(self, center, center_loc, nb=None, nb_idx=None):
        self.center, self.center_loc, self.nb, self.nb_idx = center, center_loc, nb, nb_idx


class UnstructuredRun:
    """Executes an unstructured IIR with the semantics of the reference's naive-ico backend
    (dawn/src/dawn/CodeGen/CXXNaive-ico/CXXNaiveCodeGen.cpp:366-617): per multistage, per k of the
    interval, per stage `for loc in get<Cells|Edges|Vertices>(mesh)` (restricted to the stage's
    iteration space when it has one, :546-583); a reduction iterates the neighbour list of its
    chain in mesh order starting from `init`, `lhs op= [weight *] rhs` (toylib_interface.hpp:280-323);
    a nested reduction runs around the current neighbour (ASTStencilBody.cpp:359-368); a loop
    statement visits the neighbours in the same order with a running sparse index (:114-133).
    Dense accesses inside a loop / reduction address the neighbour when they carry an offset or when
    the field lives on another location type than the centre (:237-252), otherwise the centre;
    sparse accesses address (centre, position).  A vertical indirection reads the level from the
    index field at (element, k) (:334-349).  Vectorised over the elements of a stage (a stage never
    reads through a chain what it writes itself).

    `mesh` must offer num(loc), valid(loc) and nbh_table(chain, include_center) (row-major, -1
    padded): dawn_b200.trimesh.TriMesh is pinned bit-exactly against the reference's toylib.
    Fields: dense [k, n]; sparse [k, n, nbh]; vertical-only [k]; horizontal-only dense [n] / sparse
    [n, nbh].  `splitters`: {(location, subdomain, offset): element index} for unstructured
    iteration spaces (driver-includes/unstructured_domain.hpp:23-36)."""

    SUBDOMAIN = {0: "LateralBoundary", 1000: "Nudging", 2000: "Interior", 3000: "Halo", 4000: "End"}

    def __init__(self, iir, mesh, k_size, globals_=None, splitters=None):
        self.iir = load_iir(iir)
        self.meta = self.iir["metadata"]
        self.ir = self.iir["internalIR"]
        self.mesh = mesh
        self.k_size = k_size
        self.globals = {n: g.get("value", 0) for n, g in self.ir.get("globalVariableToValue", {}).items()}
        if globals_:
            self.globals.update(globals_)
        self.dims = {int(k): v for k, v in self.meta.get("fieldIDtoDimensions", {}).items()}
        self.splitters = dict(splitters or {})

    def _chain_of(self, aid):
        d = self.dims.get(aid, {}).get("unstructured_horizontal_dimension")
        return d.get("iter_space", {}).get("chain", []) if d else []

    def _is_sparse(self, aid):
        return len(self._chain_of(aid)) > 1

    def _has_k(self, aid):
        return bool(self.dims.get(aid, {}).get("mask_k", 1))

    def run(self, fields):
        self.f = fields
        self.ft = _float_type(fields)
        kmin, kmax = 0, self.k_size - 1
        for st in self.ir["stencils"]:
            for ms in st["multiStages"]:
                backward = ms.get("loopOrder", "Forward") == "Backward"
                stages = ms["stages"]
                bounds = [interval_bounds(dm["interval"]) for s in stages for dm in s["doMethods"]]
                parts = partition(bounds)
                if backward:
                    parts = parts[::-1]
                for lo, hi in parts:
                    k0, k1 = resolve(lo, kmin, kmax), resolve(hi, kmin, kmax)
                    ks = range(k1, k0 - 1, -1) if backward else range(k0, k1 + 1)
                    for k in ks:
                        for stage in stages:
                            dms = [dm for dm in stage["doMethods"]
                                   if _overlap(interval_bounds(dm["interval"]), (lo, hi))]
                            if dms:
                                self._stage(stage, dms, k)
        return self.f

    def _space(self, stage):
        """Element index range of a stage with an unstructured iteration space (stored in i_range:
        levels are the subdomain magic numbers 0/1000/2000/3000/4000)."""
        iv = stage.get("i_range")
        if not iv:
            return None

        def bound(level_key, special_key, off_key):
            level = iv.get(level_key, 0)
            if special_key in iv and level_key not in iv:
                level = 0
            return self.splitters[(self.loc, level, iv.get(off_key, 0))]
        return (bound("lower_level", "special_lower_level", "lower_offset"),
                bound("upper_level", "special_upper_level", "upper_offset"))

    def _stage(self, stage, dms, k):
        self.k = k
        self.loc = _LOC[stage["locationType"]]
        valid = self.mesh.valid(self.loc)
        space = self._space(stage)
        if space is not None:
            sel = np.zeros_like(valid)
            sel[space[0]:space[1]] = True
            valid = valid & sel
        self.elems = np.nonzero(valid)[0]
        self.locals = {}
        top = _Ctx(self.elems, self.loc)
        for dm in dms:
            for s in dm["ast"]["block_stmt"].get("statements", []):
                self._stmt(s, None, top)

    # ---- statements ---------------------------------------------------------------------------
    def _stmt(self, node, mask, ctx):
        kd, v = _kind(node)
        if kd == "block_stmt":
            for s in v.get("statements", []):
                self._stmt(s, mask, ctx)
        elif kd == "expr_stmt":
            ek, ev = _kind(v["expr"])
            assert ek == "assignment_expr", ek
            rhs = self._eval(ev["right"], ctx)
            op = ev.get("op", "=")
            lk, lv = _kind(ev["left"])
            if lk == "var_access_expr":
                aid = lv["data"]["accessID"]
                cur = self.locals[aid]
                new = rhs if op == "=" else _BIN[op[0]](cur, rhs)
                self.locals[aid] = np.where(mask, new, cur) if mask is not None else \
                    np.broadcast_to(np.asarray(new, dtype=float), cur.shape).copy()
            else:
                idx = self._index(lv, ctx, write=True)
                arr = self.f[lv["name"]]
                cur = arr[idx]
                new = rhs if op == "=" else _BIN[op[0]](cur, rhs)
                new = np.broadcast_to(np.asarray(new, dtype=float), np.shape(cur))
                if mask is not None:
                    if np.ndim(cur) == 0:      # vertical-only field: every element stores the same slot
                        if np.any(mask):
                            arr[idx] = np.asarray(new).flat[0] if np.ndim(new) else new
                        return
                    sel = tuple(np.asarray(i)[mask] if np.ndim(i) else i for i in idx)
                    arr[sel] = new[mask]
                else:
                    arr[idx] = new
        elif kd == "var_decl_stmt":
            aid = v["var_decl_stmt_data"]["accessID"]
            init = v.get("init_list", [])
            val = self._eval(init[0], ctx) if init else 0.0
            tid = v.get("type", {}).get("builtin_type", {}).get("type_id", "Float")
            if tid == "Integer":
                val = np.trunc(val)
            self.locals[aid] = np.broadcast_to(np.asarray(val, dtype=self.ft), self.elems.shape).copy()
        elif kd == "if_stmt":
            ck, cv = _kind(v["cond_part"])
            cond = np.broadcast_to(np.asarray(self._eval(cv["expr"], ctx)).astype(bool), self.elems.shape)
            self._stmt(v["then_part"], cond if mask is None else mask & cond, ctx)
            if v.get("else_part"):
                self._stmt(v["else_part"], ~cond if mask is None else mask & ~cond, ctx)
        elif kd == "loop_stmt":
            if ctx.nb is not None:
                raise NotImplementedError("nested neighbour loops")
            space = v["loop_descriptor"]["loop_descriptor_chain"]["iter_space"]
            table = self.mesh.nbh_table(space["chain"], space.get("include_center", False))[ctx.center]
            for idx in range(table.shape[1]):
                nb = table[:, idx]
                valid = nb >= 0
                if not valid.any():
                    continue
                inner = _Ctx(ctx.center, ctx.center_loc, nb, idx)
                self._stmt(v["statements"], valid if mask is None else mask & valid, inner)
        else:
            raise NotImplementedError(kd)

    # ---- field addressing ------------------------------------------------------------------------
    def _index(self, v, ctx, write=False):
        aid = v["data"]["accessID"]
        name = v["name"]
        kk = self.k + v.get("vertical_shift", 0)
        if v.get("vertical_indirection"):
            iname = v["vertical_indirection"]
            iid = v.get("vertical_indirection_data", {}).get("accessID")
            if iid is None:
                iid = next(int(a) for a, n in self.meta["accessIDToName"].items() if n == iname)
            ind = {"name": iname, "data": {"accessID": iid}, "unstructured_offset": {"has_offset": False}}
            kk = self.f[iname][self._index(ind, ctx)].astype(np.int64) + v.get("vertical_shift", 0)
        chain = self._chain_of(aid)
        has_k = self._has_k(aid)
        if not chain:                          # vertical-only
            return (kk,)
        if len(chain) > 1:                     # sparse: (centre, position in the neighbour list)
            if ctx.nb_idx is None:
                raise ValueError("sparse field %s used outside a reduction or neighbour loop" % name)
            return (kk, ctx.center, ctx.nb_idx) if has_k else (ctx.center, ctx.nb_idx)
        has_offset = v.get("unstructured_offset", {}).get("has_offset", False)
        if ctx.nb is None:
            if has_offset:
                raise ValueError("neighbour access to %s outside a reduction or neighbour loop" % name)
            elem = ctx.center
        elif has_offset or _LOC[chain[0]] != ctx.center_loc:
            if write:
                raise ValueError("a neighbour loop may only write its own element (%s)" % name)
            elem = np.maximum(ctx.nb, 0)
        else:
            elem = ctx.center
        return (kk, elem) if has_k else (elem,)

    # ---- expressions -----------------------------------------------------------------------------
    def _eval(self, node, ctx):
        kd, v = _kind(node)
        if kd == "literal_access_expr":
            t = v.get("type", {}).get("type_id", "Float")
            return int(v["value"]) if t == "Integer" else (v["value"] in ("true", "1") if t == "Boolean"
                                                           else float(v["value"]))
        if kd == "var_access_expr":
            if v.get("is_external", False):
                return _global_value(self.globals[v["name"]])
            return self.locals[v["data"]["accessID"]]
        if kd == "field_access_expr":
            return self.f[v["name"]][self._index(v, ctx)]
        if kd == "binary_operator":
            a, b = self._eval(v["left"], ctx), self._eval(v["right"], ctx)
            with np.errstate(all="ignore"):
                return _BIN[v["op"]](a, b)
        if kd == "unary_operator":
            a = self._eval(v["operand"], ctx)
            return np.negative(a) if v["op"] == "-" else (np.logical_not(a) if v["op"] == "!" else a)
        if kd == "ternary_operator":
            return np.where(self._eval(v["cond"], ctx), self._eval(v["left"], ctx),
                            self._eval(v["right"], ctx))
        if kd == "fun_call_expr":
            with np.errstate(all="ignore"):
                return _MATH[v["callee"].split("::")[-1]](*[self._eval(a, ctx)
                                                            for a in v.get("arguments", [])])
        if kd == "reduction_over_neighbor_expr":
            space = v["iter_space"]
            chain = space["chain"]
            # inside a loop / reduction the nested reduction runs around the current neighbour
            center = ctx.center if ctx.nb is None else np.maximum(ctx.nb, 0)
            around = _Ctx(center, _LOC[chain[0]])
            table = self.mesh.nbh_table(chain, space.get("include_center", False))[center]
            acc = np.broadcast_to(np.asarray(self._eval(v["init"], around), dtype=self.ft),
                                  center.shape).copy()   # `::dawn::float_type lhs = init` in the reference text
            weights = v.get("weights", [])
            op = v.get("op", "+")
            for idx in range(table.shape[1]):
                nb = table[:, idx]
                valid = nb >= 0
                if not valid.any():
                    continue
                inner = _Ctx(center, around.center_loc, nb, idx)
                rhs = self._eval(v["rhs"], inner)
                if weights:
                    rhs = np.multiply(self._eval(weights[idx], around), rhs)
                if op in ("+", "-", "*", "/"):
                    new = _BIN[op](acc, rhs)
                elif op == "min":
                    new = np.minimum(acc, rhs)
                elif op == "max":
                    new = np.maximum(acc, rhs)
                else:
                    raise NotImplementedError(op)
                acc = np.where(valid, new, acc)
            return acc
        raise NotImplementedError(kd)


def run_unstructured(iir, mesh, k_size, fields, globals_=None, splitters=None):
    return UnstructuredRun(iir, mesh, k_size, globals_, splitters).run(fields)
