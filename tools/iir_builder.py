"""Programmatic construction of Dawn IIR (proto3-JSON) fixtures.

Test/bench infrastructure: the role `dawn/src/dawn/Unittest/IIRBuilder.h:61-365` plays in the
reference (build an optimized IIR in code, hand it to a code generator).  The JSON written here
follows `dawn/src/dawn/IIR/proto/IIR/IIR.proto` and `dawn/src/dawn/AST/proto/AST/statements.proto`
exactly as `IIRSerializer` emits it (compare `dawn/test/unit-test/dawn/CodeGen/input/*.iir`), so the
same files could be piped into the reference's `dawn-codegen` and are what the B200 emitter reads.

Stencil functions of the gtclang DSL are expressed as plain Python functions returning expression
trees, i.e. the IIR produced is the *inlined* form a CUDA-class backend requires
(`dawn/src/dawn/CodeGen/Cuda/ASTStencilBody.cpp:75-81`).
"""
from __future__ import annotations

import json
from dataclasses import dataclass, field as dc_field
from typing import Dict, List, Optional, Sequence, Tuple, Union

LOC = {"Line": -1, "Column": -1}

# iir::FieldAccessType (dawn/src/dawn/IIR/FieldAccessMetadata.h:130-138)
T_GLOBAL, T_LITERAL, T_LOCAL, T_TMP, T_INTERSTENCIL_TMP, T_FIELD, T_APIFIELD = range(7)

LOC_TYPES = {"Cell": "Cell", "Edge": "Edge", "Vertex": "Vertex"}


class _Ids:
    def __init__(self):
        self.stmt = 0

    def next(self):
        self.stmt += 1
        return self.stmt


# ------------------------------------------------------------------------------------------------
# expression tree
# ------------------------------------------------------------------------------------------------
class Expr:
    def _bin(self, op, other, swap=False):
        other = lit(other)
        return BinOp(op, other, self) if swap else BinOp(op, self, other)

    def __add__(self, o): return self._bin("+", o)
    def __radd__(self, o): return self._bin("+", o, True)
    def __sub__(self, o): return self._bin("-", o)
    def __rsub__(self, o): return self._bin("-", o, True)
    def __mul__(self, o): return self._bin("*", o)
    def __rmul__(self, o): return self._bin("*", o, True)
    def __truediv__(self, o): return self._bin("/", o)
    def __rtruediv__(self, o): return self._bin("/", o, True)
    def __gt__(self, o): return self._bin(">", o)
    def __lt__(self, o): return self._bin("<", o)
    def __ge__(self, o): return self._bin(">=", o)
    def __le__(self, o): return self._bin("<=", o)
    def eq(self, o): return self._bin("==", o)
    def ne(self, o): return self._bin("!=", o)
    def logical_and(self, o): return self._bin("&&", o)
    def logical_or(self, o): return self._bin("||", o)
    def __neg__(self): return UnOp("-", self)


@dataclass(eq=False)
class Lit(Expr):
    value: str
    type_id: str  # "Float" | "Integer" | "Boolean" | "Double"


@dataclass(eq=False)
class BinOp(Expr):
    op: str
    left: Expr
    right: Expr


@dataclass(eq=False)
class UnOp(Expr):
    op: str
    operand: Expr


@dataclass(eq=False)
class Ternary(Expr):
    cond: Expr
    left: Expr
    right: Expr


@dataclass(eq=False)
class FunCall(Expr):
    callee: str
    args: List[Expr]


@dataclass(eq=False)
class VarAccess(Expr):
    name: str
    access_id: int
    is_external: bool = False


@dataclass(eq=False)
class FieldAccess(Expr):
    name: str
    access_id: int
    i: int = 0
    j: int = 0
    k: int = 0
    unstructured: bool = False
    has_offset: bool = False  # unstructured: value read at the neighbour (inside a reduction)
    vertical_indirection: Optional["FieldAccess"] = None  # level taken from this field's value


@dataclass(eq=False)
class Reduce(Expr):
    """reduction_over_neighbor_expr (statements.proto:711-724)."""
    op: str
    rhs: Expr
    init: Expr
    chain: List[str]
    weights: Optional[List[Expr]] = None
    include_center: bool = False


def lit(v) -> Expr:
    if isinstance(v, Expr):
        return v
    if isinstance(v, bool):
        return Lit("true" if v else "false", "Boolean")
    if isinstance(v, int):
        return Lit(str(v), "Integer")
    if isinstance(v, float):
        s = repr(v)
        return Lit(s, "Double")
    raise TypeError(v)


def reduce_over(chain, rhs, init=0.0, op="+", weights=None, include_center=False) -> Expr:
    """reduction_over_neighbor_expr over `chain` (e.g. ["Edge", "Cell"])."""
    return Reduce(op, _lit_or_field(rhs), _lit_or_field(init), list(chain),
                  None if weights is None else [_lit_or_field(w) for w in weights], include_center)


def _lit_or_field(x):
    if hasattr(x, "_e"):
        return x._e()
    return lit(x)


def where(cond, a, b) -> Expr:
    return Ternary(lit(cond), lit(a), lit(b))


def call(callee: str, *args) -> Expr:
    return FunCall(callee, [lit(a) for a in args])


class FieldHandle:
    def __init__(self, name, access_id, unstructured=False):
        self.name, self.access_id, self.unstructured = name, access_id, unstructured

    def __call__(self, i=0, j=0, k=0) -> FieldAccess:
        return FieldAccess(self.name, self.access_id, i, j, k, self.unstructured)

    def at(self, k=0, nbh=False, k_from=None) -> FieldAccess:
        """k_from: field handle whose value at (element, k) is the level to read (vertical indirection,
        statements.proto FieldAccessExpr.vertical_indirection); k is then added to it."""
        ind = None
        if k_from is not None:
            ind = k_from.at() if isinstance(k_from, FieldHandle) else k_from
        return FieldAccess(self.name, self.access_id, 0, 0, k, True, nbh, ind)

    @property
    def nbh(self) -> FieldAccess:
        """value at the neighbour inside a reduction"""
        return self.at(0, True)

    # allow using the bare handle as a centre access
    def _e(self):
        return self.at() if self.unstructured else self()

    def __add__(self, o): return self._e() + o
    def __radd__(self, o): return o + self._e()
    def __sub__(self, o): return self._e() - o
    def __rsub__(self, o): return lit(o) - self._e()
    def __mul__(self, o): return self._e() * o
    def __rmul__(self, o): return lit(o) * self._e()
    def __truediv__(self, o): return self._e() / o
    def __rtruediv__(self, o): return lit(o) / self._e()
    def __neg__(self): return -self._e()


def _as_expr(x):
    if isinstance(x, FieldHandle):
        return x._e()
    return lit(x)


# ------------------------------------------------------------------------------------------------
# statements
# ------------------------------------------------------------------------------------------------
@dataclass(eq=False)
class Assign:
    left: Expr
    right: Expr
    op: str = "="


@dataclass(eq=False)
class VarDecl:
    name: str
    access_id: int
    init: Optional[Expr]
    type_id: str = "Float"
    is_const: bool = False


@dataclass(eq=False)
class If:
    cond: Expr
    then: List
    otherwise: Optional[List] = None


@dataclass(eq=False)
class Call:
    """stencil_call_decl_stmt of the n-th stencil() of the builder (control flow of the description)."""
    index: int


@dataclass(eq=False)
class Loop:
    """loop_stmt over a neighbour chain (statements.proto LoopStmt + LoopDescriptorChain)."""
    chain: List[str]
    body: List
    include_center: bool = False


def loop_over(chain, *stmts, include_center=False) -> Loop:
    return Loop(list(chain), list(stmts), include_center)


@dataclass
class Interval:
    """statements.proto:120-147.  Levels: 'Start' | 'End' | int."""
    lower: Union[str, int] = "Start"
    upper: Union[str, int] = "End"
    lower_offset: int = 0
    upper_offset: int = 0

    def to_json(self):
        d = {"lower_offset": self.lower_offset, "upper_offset": self.upper_offset}
        if isinstance(self.lower, str):
            d["special_lower_level"] = self.lower
        else:
            d["lower_level"] = self.lower
        if isinstance(self.upper, str):
            d["special_upper_level"] = self.upper
        else:
            d["upper_level"] = self.upper
        return d


FULL = Interval()


# ------------------------------------------------------------------------------------------------
# extents bookkeeping (per-statement accesses, IIR.proto / statements.proto `Accesses`)
# ------------------------------------------------------------------------------------------------
class _Ext:
    def __init__(self, i, j, k, unstructured=False, has_extent=False):
        self.i = [i, i]
        self.j = [j, j]
        self.k = [k, k]
        self.unstructured = unstructured
        self.has_extent = has_extent

    def merge(self, o):
        for a, b in ((self.i, o.i), (self.j, o.j), (self.k, o.k)):
            a[0] = min(a[0], b[0])
            a[1] = max(a[1], b[1])
        self.has_extent = self.has_extent or o.has_extent

    def to_json(self):
        vert = {"minus": self.k[0], "plus": self.k[1], "undefined": False}
        if self.unstructured:
            return {"unstructured_extent": {"has_extent": self.has_extent}, "vertical_extent": vert}
        if self.i == [0, 0] and self.j == [0, 0]:
            return {"zero_extent": {}, "vertical_extent": vert}
        return {
            "cartesian_extent": {
                "i_extent": {"minus": self.i[0], "plus": self.i[1], "undefined": False},
                "j_extent": {"minus": self.j[0], "plus": self.j[1], "undefined": False},
            },
            "vertical_extent": vert,
        }


def _add(acc: Dict[int, _Ext], aid: int, e: _Ext):
    if aid in acc:
        acc[aid].merge(e)
    else:
        acc[aid] = e


class IIRBuilder:
    def __init__(self, name: str, grid: str = "Cartesian"):
        self.name = name
        self.grid = grid
        self._next_id = 10
        self._next_lit = -10
        self._ids = _Ids()
        self.names: Dict[int, str] = {}
        self.types: Dict[int, int] = {}
        self.dims: Dict[int, dict] = {}
        self.api: List[int] = []
        self.tmps: List[int] = []
        self.fields: List[int] = []
        self.globals: List[int] = []
        self.global_values: Dict[str, dict] = {}
        self.literals: Dict[int, str] = {}
        self.stencils: List[dict] = []
        self._nid = 100

    # -- ids --------------------------------------------------------------------------------------
    def _new(self):
        self._next_id += 1
        return self._next_id

    def _node(self):
        self._nid += 1
        return self._nid

    # -- declarations -----------------------------------------------------------------------------
    def _dims_json(self, dims, location=None, sparse_chain=None, include_center=False):
        if self.grid == "Cartesian":
            return {
                "cartesian_horizontal_dimension": {
                    "mask_cart_i": 1 if "i" in dims else 0,
                    "mask_cart_j": 1 if "j" in dims else 0,
                },
                "mask_k": 1 if "k" in dims else 0,
            }
        if location is None and not sparse_chain:  # vertical-only field
            return {"mask_k": 1}
        chain = sparse_chain if sparse_chain else [location]
        return {
            "unstructured_horizontal_dimension": {
                "iter_space": {"chain": list(chain), "include_center": bool(include_center)}},
            "mask_k": 1 if "k" in dims else 0,
        }

    def field(self, name, dims="ijk", location=None, sparse_chain=None, include_center=False) -> FieldHandle:
        aid = self._new()
        self.names[aid], self.types[aid] = name, T_APIFIELD
        self.dims[aid] = self._dims_json(dims, location, sparse_chain, include_center)
        self.api.append(aid)
        self.fields.append(aid)
        return FieldHandle(name, aid, self.grid != "Cartesian")

    def tmp(self, name, dims="ijk", location=None, sparse_chain=None, include_center=False) -> FieldHandle:
        aid = self._new()
        self.names[aid], self.types[aid] = name, T_TMP
        self.dims[aid] = self._dims_json(dims, location, sparse_chain, include_center)
        self.tmps.append(aid)
        self.fields.append(aid)
        return FieldHandle(name, aid, self.grid != "Cartesian")

    def global_var(self, name, type_id="Double", value=None) -> VarAccess:
        aid = self._new()
        self.names[aid], self.types[aid] = name, T_GLOBAL
        self.globals.append(aid)
        self.global_values[name] = {
            "type": type_id, "value": 0 if value is None else value, "valueIsSet": value is not None}
        return VarAccess(name, aid, True)

    def local(self, name, init=None, type_id="Float", is_const=False) -> Tuple[VarDecl, VarAccess]:
        aid = self._new()
        full = f"__local_{name}_{aid}"
        self.names[aid], self.types[aid] = full, T_LOCAL
        decl = VarDecl(full, aid, None if init is None else _as_expr(init), type_id, is_const)
        return decl, VarAccess(full, aid, False)

    # -- structure --------------------------------------------------------------------------------
    @staticmethod
    def assign(left, right, op="=") -> Assign:
        return Assign(_as_expr(left), _as_expr(right), op)

    @staticmethod
    def if_(cond, then, otherwise=None) -> If:
        return If(_as_expr(cond), list(then), None if otherwise is None else list(otherwise))

    def do(self, interval: Interval, *stmts) -> dict:
        return {"interval": interval, "stmts": list(stmts)}

    def stage(self, *do_methods, location=None, i_range=None, j_range=None) -> dict:
        return {"dos": list(do_methods), "location": location, "i_range": i_range, "j_range": j_range}

    def multistage(self, loop_order: str, *stages, caches: Optional[dict] = None) -> dict:
        return {"order": loop_order, "stages": list(stages), "caches": caches or {}}

    def stencil(self, *multistages):
        self.stencils.append({"ms": list(multistages)})
        return Call(len(self.stencils) - 1)

    def control_flow(self, *stmts):
        """Statements of the stencil description around the stencil calls (if / local variables over
        globals), as gtclang records them in controlFlowStatements.  Without it the stencils run in
        order, unconditionally."""
        self.flow = list(stmts)

    # -- serialization ----------------------------------------------------------------------------
    def _lit_id(self, value):
        self._next_lit -= 1
        self.literals[self._next_lit] = value
        return self._next_lit

    def _expr(self, e: Expr, reads: Dict[int, _Ext]) -> dict:
        nid = self._node()
        if isinstance(e, Lit):
            aid = self._lit_id(e.value)
            _add(reads, aid, _Ext(0, 0, 0))
            return {"literal_access_expr": {"value": e.value, "type": {"type_id": e.type_id},
                                            "loc": LOC, "data": {"accessID": aid}, "ID": nid}}
        if isinstance(e, BinOp):
            return {"binary_operator": {"left": self._expr(e.left, reads), "op": e.op,
                                        "right": self._expr(e.right, reads), "loc": LOC, "ID": nid}}
        if isinstance(e, UnOp):
            return {"unary_operator": {"op": e.op, "operand": self._expr(e.operand, reads),
                                       "loc": LOC, "ID": nid}}
        if isinstance(e, Ternary):
            return {"ternary_operator": {"cond": self._expr(e.cond, reads),
                                         "left": self._expr(e.left, reads),
                                         "right": self._expr(e.right, reads), "loc": LOC, "ID": nid}}
        if isinstance(e, FunCall):
            return {"fun_call_expr": {"callee": e.callee,
                                      "arguments": [self._expr(a, reads) for a in e.args],
                                      "loc": LOC, "ID": nid}}
        if isinstance(e, VarAccess):
            _add(reads, e.access_id, _Ext(0, 0, 0))
            return {"var_access_expr": {"name": e.name, "is_external": e.is_external, "loc": LOC,
                                        "data": {"accessID": e.access_id}, "ID": nid}}
        if isinstance(e, FieldAccess):
            _add(reads, e.access_id, _Ext(e.i, e.j, e.k, e.unstructured, e.has_offset))
            if e.vertical_indirection is not None:
                _add(reads, e.vertical_indirection.access_id, _Ext(0, 0, 0, True, False))
            return self._field_json(e, nid)
        if isinstance(e, Reduce):
            d = {"op": e.op, "rhs": self._expr(e.rhs, reads), "init": self._expr(e.init, reads),
                 "iter_space": {"chain": list(e.chain), "include_center": e.include_center},
                 "weights": [self._expr(w, reads) for w in (e.weights or [])],
                 "loc": LOC}  # no ID: statements.proto gives ReductionOverNeighborExpr none
            return {"reduction_over_neighbor_expr": d}
        raise TypeError(type(e))

    def _field_json(self, e: FieldAccess, nid: int) -> dict:
        d = {"name": e.name, "vertical_shift": e.k, "vertical_indirection": ""}
        if e.vertical_indirection is not None:
            d["vertical_indirection"] = e.vertical_indirection.name
            d["vertical_indirection_data"] = {"accessID": e.vertical_indirection.access_id}
        if e.unstructured:
            d["unstructured_offset"] = {"has_offset": e.has_offset}
        elif e.i == 0 and e.j == 0:
            d["zero_offset"] = {}
        else:
            d["cartesian_offset"] = {"i_offset": e.i, "j_offset": e.j}
        d.update({"argument_map": [-1, -1, -1], "argument_offset": [0, 0, 0], "negate_offset": False,
                  "loc": LOC, "data": {"accessID": e.access_id}, "ID": nid})
        return {"field_access_expr": d}

    @staticmethod
    def _acc_json(reads, writes):
        return {"accesses": {
            "writeAccess": {str(k): v.to_json() for k, v in writes.items()},
            "readAccess": {str(k): v.to_json() for k, v in reads.items()}}}

    def _stmt(self, s, up_reads, up_writes) -> dict:
        reads: Dict[int, _Ext] = {}
        writes: Dict[int, _Ext] = {}
        nid = self._node()
        if isinstance(s, Assign):
            right = self._expr(s.right, reads)
            if s.op != "=":  # compound assignment also reads the lhs
                dummy: Dict[int, _Ext] = {}
                left = self._expr(s.left, dummy)
                for k, v in dummy.items():
                    _add(reads, k, v)
            else:
                dummy = {}
                left = self._expr(s.left, dummy)
            lhs = s.left
            unstr = isinstance(lhs, FieldAccess) and lhs.unstructured
            _add(writes, lhs.access_id, _Ext(0, 0, 0, unstr))
            out = {"expr_stmt": {"expr": {"assignment_expr": {
                "left": left, "op": s.op, "right": right, "loc": LOC, "ID": self._node()}},
                "loc": LOC, "data": self._acc_json(reads, writes), "ID": nid}}
        elif isinstance(s, VarDecl):
            init = [] if s.init is None else [self._expr(s.init, reads)]
            _add(writes, s.access_id, _Ext(0, 0, 0))
            out = {"var_decl_stmt": {
                "type": {"builtin_type": {"type_id": s.type_id}, "is_const": s.is_const,
                         "is_volatile": False},
                "name": s.name, "dimension": 0, "op": "=", "init_list": init, "loc": LOC,
                "data": self._acc_json(reads, writes),
                "var_decl_stmt_data": {"accessID": s.access_id}, "ID": nid}}
        elif isinstance(s, If):
            cond_reads: Dict[int, _Ext] = {}
            cond = self._expr(s.cond, cond_reads)
            cond_stmt = {"expr_stmt": {"expr": cond, "loc": LOC,
                                       "data": self._acc_json(cond_reads, {}), "ID": self._node()}}
            for k, v in cond_reads.items():
                _add(reads, k, v)
            then = self._block(s.then, reads, writes)
            d = {"cond_part": cond_stmt, "then_part": then}
            if s.otherwise is not None:
                d["else_part"] = self._block(s.otherwise, reads, writes)
            d.update({"loc": LOC, "data": self._acc_json(reads, writes), "ID": nid})
            out = {"if_stmt": d}
        elif isinstance(s, Loop):
            body = self._block(s.body, reads, writes)
            out = {"loop_stmt": {
                "statements": body,
                "loop_descriptor": {"loop_descriptor_chain": {
                    "iter_space": {"chain": list(s.chain), "include_center": s.include_center}}},
                "loc": LOC, "data": self._acc_json(reads, writes), "ID": nid}}
        else:
            raise TypeError(type(s))
        for k, v in reads.items():
            _add(up_reads, k, _copy_ext(v))
        for k, v in writes.items():
            _add(up_writes, k, _copy_ext(v))
        return out

    def _flow_stmt(self, s) -> dict:
        if isinstance(s, Call):
            return self.stencils[s.index]["call_json"]
        if isinstance(s, If):
            d = {"cond_part": {"expr_stmt": {"expr": self._expr(s.cond, {}), "loc": LOC, "data": {},
                                             "ID": self._node()}},
                 "then_part": {"block_stmt": {"statements": [self._flow_stmt(x) for x in s.then],
                                              "loc": LOC, "data": {}, "ID": self._node()}}}
            if s.otherwise is not None:
                d["else_part"] = {"block_stmt": {"statements": [self._flow_stmt(x) for x in s.otherwise],
                                                 "loc": LOC, "data": {}, "ID": self._node()}}
            d.update({"loc": LOC, "data": {}, "ID": self._node()})
            return {"if_stmt": d}
        if isinstance(s, (VarDecl, Assign)):
            out = self._stmt(s, {}, {})
            return out
        raise TypeError(type(s))

    def _block(self, stmts, up_reads, up_writes) -> dict:
        reads: Dict[int, _Ext] = {}
        writes: Dict[int, _Ext] = {}
        js = [self._stmt(s, reads, writes) for s in stmts]
        for k, v in reads.items():
            _add(up_reads, k, _copy_ext(v))
        for k, v in writes.items():
            _add(up_writes, k, _copy_ext(v))
        return {"block_stmt": {"statements": js, "loc": LOC,
                               "data": self._acc_json(reads, writes), "ID": self._node()}}

    def build(self) -> dict:
        stencils_json = []
        id_to_call = {}
        control_flow = []
        for st in self.stencils:
            sid = self._new()
            ms_json = []
            for ms in st["ms"]:
                stages_json = []
                for stage in ms["stages"]:
                    dos_json = []
                    for do in stage["dos"]:
                        r: Dict[int, _Ext] = {}
                        w: Dict[int, _Ext] = {}
                        block = self._block(do["stmts"], r, w)
                        block["block_stmt"]["data"] = {}
                        dos_json.append({"ast": block, "doMethodID": self._new(),
                                         "interval": do["interval"].to_json()})
                    sj = {"doMethods": dos_json, "stageID": self._new(),
                          "locationType": stage["location"] or "LocationTypeUnknown"}
                    if stage["i_range"] is not None:
                        sj["i_range"] = stage["i_range"].to_json()
                    if stage["j_range"] is not None:
                        sj["j_range"] = stage["j_range"].to_json()
                    stages_json.append(sj)
                caches = {}
                for aid, c in ms["caches"].items():
                    caches[str(aid)] = dict(c, accessID=aid)
                ms_json.append({"stages": stages_json, "loopOrder": ms["order"],
                                "multiStageID": self._new(), "Caches": caches})
            stencils_json.append({"multiStages": ms_json, "stencilID": sid,
                                  "attr": {"attributes": []}})
            call = {"stencil_call_decl_stmt": {
                "stencil_call": {"loc": LOC, "callee": f"__code_gen_{sid}", "arguments": []},
                "loc": LOC, "data": {}, "ID": self._node()}}
            id_to_call[str(sid)] = call
            control_flow.append(call)
            st["call_json"] = call
        if getattr(self, "flow", None):
            control_flow = [self._flow_stmt(s) for s in self.flow]

        meta = {
            "accessIDToName": {str(k): v for k, v in self.names.items()},
            "accessIDToType": {str(k): v for k, v in self.types.items()},
            "literalIDToName": {str(k): v for k, v in self.literals.items()},
            "fieldAccessIDs": list(self.fields),
            "APIFieldIDs": list(self.api),
            "temporaryFieldIDs": list(self.tmps),
            "globalVariableIDs": list(self.globals),
            "versionedFields": {"variableVersionMap": {}},
            "fieldnameToBoundaryCondition": {},
            "fieldIDtoDimensions": {str(k): v for k, v in self.dims.items()},
            "idToStencilCall": id_to_call,
            "boundaryCallToExtent": {},
            "allocatedFieldIDs": [],
            "stencilLocation": LOC,
            "stencilName": self.name,
        }
        ir = {
            "gridType": self.grid,
            "globalVariableToValue": self.global_values,
            "stencils": stencils_json,
            "controlFlowStatements": control_flow,
            "boundaryConditions": [],
        }
        return {"metadata": meta, "internalIR": ir, "filename": self.name + ".cpp"}

    def dumps(self) -> str:
        return json.dumps(self.build(), indent=1)


def _copy_ext(e: _Ext) -> _Ext:
    c = _Ext(0, 0, 0, e.unstructured, e.has_extent)
    c.i, c.j, c.k = list(e.i), list(e.j), list(e.k)
    return c


def k_cache(policy: str, window: Optional[Tuple[int, int]] = None, ctype="CT_K") -> dict:
    """Cache entry for MultiStage.Caches (IIR.proto:34-58)."""
    d = {"type": ctype, "policy": policy}
    if window is not None:
        d["cacheWindow"] = {"minus": window[0], "plus": window[1]}
    return d
