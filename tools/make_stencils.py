"""Writes the optimized-IIR fixtures under stencils/ (the emitter's input for tests and bench).

Dawn's optimizer cannot be built in this image, and the reference ships no optimized `.iir` for its
headline stencils (only `dawn/test/unit-test/dawn/CodeGen/input/{conditional_stencil,update_dz_c}.iir`),
so the IIRs are authored here with tools/iir_builder.py in the shape the reference's pass pipeline
produces for a CUDA-class backend (stencil functions inlined, stages merged per multistage, K-caches
set by PassSetCaches):

  hori_diff_type2_stencil   gtclang/test/integration-test/CodeGen/hori_diff_type2_stencil.cpp:5-55
  hori_diff_stencil_01      gtclang/test/integration-test/CodeGen/hori_diff_stencil_01.cpp:21-31
  hori_diff_stencil         dawn/test/integration-test/dawn4py-tests/hori_diff_stencil.py
                            (expression trees pinned by data/hori_diff_stencil_ref.cpp:105-119)
  tridiagonal_solve         gtclang/test/integration-test/CodeGen/tridiagonal_solve.cpp:21-38
                            (caches as in data/tridiagonal_solve_stencil_ref.cpp: c,d K-cached)
  laplacian_stencil         dawn/examples/tutorial/laplacian_stencil.cpp:1-13
  copy_stencil              gtclang/test/integration-test/CodeGen/copy_stencil.cpp
  global_indexing           dawn/test/unit-test/dawn/CodeGen/Stencils.cpp (getGlobalIndexStencil)
  nonoverlapping_stencil    dawn/test/unit-test/dawn/CodeGen/Stencils.cpp (getNonOverlappingInterval)

Run:  python tools/make_stencils.py   (idempotent; output is committed)
"""
import os
import sys

sys.path.insert(0, os.path.dirname(__file__))
from iir_builder import (FULL, IIRBuilder, Interval, k_cache, lit, loop_over, reduce_over,  # noqa: E402
                         where)

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "stencils")


def hori_diff_type2():
    b = IIRBuilder("hori_diff_type2_stencil")
    out, inn = b.field("out"), b.field("in")
    crlato, crlatu = b.field("crlato", "j"), b.field("crlatu", "j")
    hdmask = b.field("hdmask")
    lap = b.tmp("lap")

    def delta(data, di, dj, ci=0, cj=0):  # data(off) - data, evaluated at centre shift (ci,cj)
        return data(ci + di, cj + dj) - data(ci, cj)

    lap_expr = (inn(1, 0) + inn(-1, 0) - 2.0 * inn() + crlato() * delta(inn, 0, 1)
                + crlatu() * delta(inn, 0, -1))
    s1 = b.stage(b.do(FULL, b.assign(lap, lap_expr)))

    stmts = []

    def flux_x(ci):
        decl, flx = b.local("flx", delta(lap, 1, 0, ci, 0), "Double", True)
        stmts.append(decl)
        return where((flx * delta(inn, 1, 0, ci, 0)) > 0.0, 0.0, flx)

    def flux_y(cj):
        decl, fly = b.local("fly", crlato(0, cj) * delta(lap, 0, 1, 0, cj), "Double", True)
        stmts.append(decl)
        return where((fly * delta(inn, 0, 1, 0, cj)) > 0.0, 0.0, fly)

    fx = flux_x(0) - flux_x(-1)
    d1, dfx = b.local("delta_flux_x", fx, "Double", True)
    stmts.append(d1)
    fy = flux_y(0) - flux_y(-1)
    d2, dfy = b.local("delta_flux_y", fy, "Double", True)
    stmts.append(d2)
    stmts.append(b.assign(out, inn() - hdmask() * (dfx + dfy)))
    s2 = b.stage(b.do(FULL, *stmts))
    b.stencil(b.multistage("Parallel", s1, s2,
                           caches={lap.access_id: k_cache("CP_Local", ctype="CT_IJ")}))
    return b


def hori_diff_01():
    b = IIRBuilder("hori_diff_stencil_01")
    u, out = b.field("u"), b.field("out")
    lap = b.tmp("lap")
    s1 = b.stage(b.do(FULL, b.assign(lap, u(1, 0) + u(-1, 0) + u(0, 1) + u(0, -1) - 4.0 * u())))
    s2 = b.stage(b.do(FULL, b.assign(
        out, lap(1, 0) + lap(-1, 0) + lap(0, 1) + lap(0, -1) - 4.0 * lap())))
    b.stencil(b.multistage("Parallel", s1, s2,
                           caches={lap.access_id: k_cache("CP_Local", ctype="CT_IJ")}))
    return b


def hori_diff_dawn4py():
    b = IIRBuilder("hori_diff_stencil")
    inn, out, coeff = b.field("in"), b.field("out"), b.field("coeff")
    lap = b.tmp("lap")

    def five(f):
        return (-4.0 * f()) + (coeff() * (f(1, 0) + (f(-1, 0) + (f(0, 1) + f(0, -1)))))

    s1 = b.stage(b.do(FULL, b.assign(lap, five(inn))))
    s2 = b.stage(b.do(FULL, b.assign(out, five(lap))))
    b.stencil(b.multistage("Parallel", s1, s2,
                           caches={lap.access_id: k_cache("CP_Local", ctype="CT_IJ")}))
    return b


def tridiagonal_solve():
    b = IIRBuilder("tridiagonal_solve")
    d, a, bb, c = b.field("d"), b.field("a"), b.field("b"), b.field("c")
    first = Interval("Start", "Start", 0, 0)
    rest = Interval("Start", "End", 1, 0)
    back = Interval("Start", "End", 0, -1)
    mdecl, m = b.local("m", 1.0 / (bb() - a() * c(0, 0, -1)), "Double")
    fwd = b.stage(
        b.do(first, b.assign(c, c() / bb()), b.assign(d, d() / bb())),
        b.do(rest, mdecl, b.assign(c, c() * m), b.assign(d, (d() - a() * d(0, 0, -1)) * m)),
    )
    bwd = b.stage(b.do(back, b.assign(d, c() * d(0, 0, 1), op="-=")))
    b.stencil(
        b.multistage("Forward", fwd, caches={
            c.access_id: k_cache("CP_FillFlush", (-1, 0)),
            d.access_id: k_cache("CP_FillFlush", (-1, 0))}),
        b.multistage("Backward", bwd, caches={d.access_id: k_cache("CP_FillFlush", (0, 1))}),
    )
    return b


def tutorial_laplacian():
    b = IIRBuilder("laplacian_stencil")
    dx = b.global_var("dx", "Double", 1.0)
    out, inn = b.field("out_field", "ij"), b.field("in_field", "ij")
    e = (-4 * inn() + inn(1, 0) + inn(-1, 0) + inn(0, -1) + inn(0, 1)) / (dx * dx)
    b.stencil(b.multistage("Parallel", b.stage(b.do(FULL, b.assign(out, e)))))
    return b


def copy_stencil():
    b = IIRBuilder("copy_stencil")
    inn, out = b.field("in"), b.field("out")
    b.stencil(b.multistage("Parallel", b.stage(b.do(FULL, b.assign(out, inn())))))
    return b


def global_indexing():
    # reference/global_indexing.cpp:84-106: out=in everywhere; then out=10 where the global j
    # index lies in [jminus+0, jminus+2)
    b = IIRBuilder("global_indexing")
    inn, out = b.field("in_field"), b.field("out_field")
    s1 = b.stage(b.do(FULL, b.assign(out, inn())))
    s2 = b.stage(b.do(FULL, b.assign(out, 10)), j_range=Interval("Start", "Start", 0, 2))
    b.stencil(b.multistage("Parallel", s1, s2))
    return b


def nonoverlapping():
    # reference/nonoverlapping_stencil.cpp:75-98: [0,10] laplacian / (dx*dx) with a LOCAL dx that the
    # reference leaves uninitialised (not a numeric oracle there); [15,end] out = 10.
    b = IIRBuilder("nonoverlapping_stencil")
    inn, out = b.field("in"), b.field("out")
    ddecl, dx = b.local("dx", 2.0, "Double")
    e = (-4 * (inn() + (inn(1, 0) + (inn(-1, 0) + (inn(0, -1) + inn(0, 1)))))) / (dx * dx)
    s1 = b.stage(b.do(Interval("Start", 10, 0, 0), ddecl, b.assign(out, e)),
                 b.do(Interval(15, "End", 0, 0), b.assign(out, 10)))
    b.stencil(b.multistage("Parallel", s1))
    return b


def icon_laplacian():
    # dawn/test/integration-test/dawn4py-tests/ICON_laplacian_stencil.py:36-216; statement trees as in
    # the reference's generated naive-ico text data/ICON_laplacian_stencil_ref.cpp:62-139
    b = IIRBuilder("ICON_laplacian_stencil", grid="Unstructured")
    vec = b.field("vec", "k", "Edge")
    div_vec = b.field("div_vec", "k", "Cell")
    rot_vec = b.field("rot_vec", "k", "Vertex")
    nabla2 = b.field("nabla2_vec", "k", "Edge")
    pel = b.field("primal_edge_length", "k", "Edge")
    del_ = b.field("dual_edge_length", "k", "Edge")
    tang = b.field("tangent_orientation", "k", "Edge")
    geofac_rot = b.field("geofac_rot", "k", sparse_chain=["Vertex", "Edge"])
    geofac_div = b.field("geofac_div", "k", sparse_chain=["Cell", "Edge"])
    t60 = b.tmp("__tmp_nab_60", "k", "Edge")
    t61 = b.tmp("__tmp_nab_61", "k", "Edge")
    s = [
        b.stage(b.do(FULL, b.assign(rot_vec, reduce_over(["Vertex", "Edge"], vec.nbh * geofac_rot.at()))),
                location="Vertex"),
        b.stage(b.do(FULL, b.assign(div_vec, reduce_over(["Cell", "Edge"], vec.nbh * geofac_div.at()))),
                location="Cell"),
        b.stage(b.do(FULL, b.assign(t60, reduce_over(["Edge", "Vertex"], rot_vec.nbh,
                                                     weights=[-1.0, 1.0]))), location="Edge"),
        b.stage(b.do(FULL, b.assign(t60, (tang.at() * t60.at()) / pel.at())), location="Edge"),
        b.stage(b.do(FULL, b.assign(t61, reduce_over(["Edge", "Cell"], div_vec.nbh,
                                                     weights=[-1.0, 1.0]))), location="Edge"),
        b.stage(b.do(FULL, b.assign(t61, t61.at() / del_.at())), location="Edge"),
        b.stage(b.do(FULL, b.assign(nabla2, t61.at() - t60.at())), location="Edge"),
    ]
    b.stencil(b.multistage("Parallel", *s))
    return b


def icon_nabla4():
    # BASELINE.json configs[3] names "nabla2/nabla4" reductions; the reference holds no nabla4 stencil
    # ("parity unpinned"): it is defined here as ICON does, nabla4 = nabla2(nabla2(vec)),
    # i.e. the seven stages of ICON_laplacian_stencil.py:36-131 applied a second time to nabla2_vec
    b = IIRBuilder("ICON_nabla4_stencil", grid="Unstructured")
    vec = b.field("vec", "k", "Edge")
    nabla2 = b.field("nabla2_vec", "k", "Edge")
    nabla4 = b.field("nabla4_vec", "k", "Edge")
    pel = b.field("primal_edge_length", "k", "Edge")
    del_ = b.field("dual_edge_length", "k", "Edge")
    tang = b.field("tangent_orientation", "k", "Edge")
    geofac_rot = b.field("geofac_rot", "k", sparse_chain=["Vertex", "Edge"])
    geofac_div = b.field("geofac_div", "k", sparse_chain=["Cell", "Edge"])
    s = []
    for n, (src, dst) in enumerate(((vec, nabla2), (nabla2, nabla4))):
        rot = b.tmp("__tmp_rot_%d" % n, "k", "Vertex")
        div = b.tmp("__tmp_div_%d" % n, "k", "Cell")
        t60 = b.tmp("__tmp_nab_6%d" % (2 * n), "k", "Edge")
        t61 = b.tmp("__tmp_nab_6%d" % (2 * n + 1), "k", "Edge")
        s += [
            b.stage(b.do(FULL, b.assign(rot, reduce_over(["Vertex", "Edge"], src.nbh * geofac_rot.at()))),
                    location="Vertex"),
            b.stage(b.do(FULL, b.assign(div, reduce_over(["Cell", "Edge"], src.nbh * geofac_div.at()))),
                    location="Cell"),
            b.stage(b.do(FULL, b.assign(t60, reduce_over(["Edge", "Vertex"], rot.nbh, weights=[-1.0, 1.0]))),
                    location="Edge"),
            b.stage(b.do(FULL, b.assign(t60, (tang.at() * t60.at()) / pel.at())), location="Edge"),
            b.stage(b.do(FULL, b.assign(t61, reduce_over(["Edge", "Cell"], div.nbh, weights=[-1.0, 1.0]))),
                    location="Edge"),
            b.stage(b.do(FULL, b.assign(t61, t61.at() / del_.at())), location="Edge"),
            b.stage(b.do(FULL, b.assign(dst, t61.at() - t60.at())), location="Edge"),
        ]
    b.stencil(b.multistage("Parallel", *s))
    return b


def icon_laplacian_diamond():
    # dawn/test/integration-test/dawn4py-tests/ICON_laplacian_diamond_stencil.py:37-245 (ICON's diamond
    # nabla2 with the Smagorinsky coefficient; one statement per stage like the optimizer leaves them,
    # all on edges; field order of the script :252-376)
    from iir_builder import call
    b = IIRBuilder("ICON_laplacian_diamond_stencil", grid="Unstructured")
    ecv = ["Edge", "Cell", "Vertex"]
    smag = b.field("diff_multfac_smag", "k", "Edge")
    tang = b.field("tangent_orientation", "k", "Edge")
    ipel = b.field("inv_primal_edge_length", "k", "Edge")
    ivvl = b.field("inv_vert_vert_length", "k", "Edge")
    u, v = b.field("u_vert", "k", "Vertex"), b.field("v_vert", "k", "Vertex")
    pnx, pny = b.field("primal_normal_x", "k", sparse_chain=ecv), b.field("primal_normal_y", "k", sparse_chain=ecv)
    dnx, dny = b.field("dual_normal_x", "k", sparse_chain=ecv), b.field("dual_normal_y", "k", sparse_chain=ecv)
    vn_vert = b.field("vn_vert", "k", sparse_chain=ecv)
    vn = b.field("vn", "k", "Edge")
    dvt_tang, dvt_norm = b.field("dvt_tang", "k", "Edge"), b.field("dvt_norm", "k", "Edge")
    k1, k2, kh = b.field("kh_smag_1", "k", "Edge"), b.field("kh_smag_2", "k", "Edge"), b.field("kh_smag", "k", "Edge")
    nabla2 = b.field("nabla2", "k", "Edge")
    dual = lambda: u.nbh * dnx.nbh + v.nbh * dny.nbh  # noqa: E731
    first, second = [-1.0, 1.0, 0.0, 0.0], [0.0, 0.0, -1.0, 1.0]
    stmts = [
        loop_over(ecv, b.assign(vn_vert, u.nbh * pnx.nbh + v.nbh * pny.nbh)),
        b.assign(dvt_tang, reduce_over(ecv, dual(), weights=first)),
        b.assign(dvt_tang, dvt_tang.at() * tang.at()),
        b.assign(dvt_norm, reduce_over(ecv, dual(), weights=second)),
        b.assign(k1, reduce_over(ecv, vn_vert.at(), weights=first)),
        b.assign(k1, ((k1.at() * tang.at()) * ipel.at()) + (dvt_norm.at() * ivvl.at())),
        b.assign(k1, k1.at() * k1.at()),
        b.assign(k2, reduce_over(ecv, vn_vert.at(), weights=second)),
        b.assign(k2, (k2.at() * ivvl.at()) + (dvt_tang.at() * ipel.at())),
        b.assign(k2, k2.at() * k2.at()),
        b.assign(kh, smag.at() * call("gridtools::dawn::math::sqrt", k1.at() + k2.at())),
        b.assign(nabla2, reduce_over(ecv, lit(4.0) * vn_vert.at(), weights=[
            ipel.at() * ipel.at(), ipel.at() * ipel.at(), ivvl.at() * ivvl.at(), ivvl.at() * ivvl.at()])),
        b.assign(nabla2, nabla2.at() - (((lit(8.0) * vn.at()) * (ipel.at() * ipel.at())) +
                                        ((lit(8.0) * vn.at()) * (ivvl.at() * ivvl.at())))),
    ]
    b.stencil(b.multistage("Parallel", *[b.stage(b.do(FULL, st), location="Edge") for st in stmts]))
    return b


def icon_gradient():
    # dawn/test/integration-test/dawn4py-tests/ICON_gradient.py:38-63 (the script compiles it with the
    # cuda-ico backend): a loop with centre fills the sparse geofac_grg, a reduction with centre uses it
    b = IIRBuilder("ICON_gradient", grid="Unstructured")
    cec = ["Cell", "Edge", "Cell"]
    p_grad, p_ccpr = b.field("p_grad", "k", "Cell"), b.field("p_ccpr", "k", "Cell")
    grg = b.field("geofac_grg", "k", sparse_chain=cec, include_center=True)
    b.stencil(b.multistage("Parallel",
                           b.stage(b.do(FULL, loop_over(cec, b.assign(grg, lit(2.0)), include_center=True)),
                                   location="Cell"),
                           b.stage(b.do(FULL, b.assign(p_grad, reduce_over(cec, grg.at() * p_ccpr.nbh,
                                                                           include_center=True))),
                                   location="Cell")))
    return b


def generate_versioned_field():
    # dawn/test/integration-test/dawn4py-tests/generate_versioned_field.py after PassFieldVersioning, as
    # the reference's own generated text shows it (data/generate_versioned_field_ref.cpp:42-66): c is
    # copied to its version c_0 in a first multistage, the second reads c_0 and conditionally writes c
    b = IIRBuilder("generate_versioned_field", grid="Unstructured")
    a, bb, c, d, e = (b.field(n, "k", "Edge") for n in "abcde")
    c0 = b.tmp("c_0", "k", "Edge")
    s1 = b.stage(b.do(FULL, b.assign(c0, c.at())), location="Edge")
    s2 = b.stage(b.do(FULL, b.assign(a, (bb.at() / c0.at()) + lit(5)),
                      b.if_(d.at(), [b.assign(a, bb.at())],
                            [b.if_(e.at(), [b.assign(c, a.at() + lit(1))])])), location="Edge")
    b.stencil(b.multistage("Parallel", s1), b.multistage("Parallel", s2))
    return b


def unstructured_diffusion():
    # dawn/test/integration-test/unstructured/reference/reference_diffusion.hpp:31-69
    b = IIRBuilder("diffusion", grid="Unstructured")
    inn = b.field("in_field", "k", "Cell")
    out = b.field("out_field", "k", "Cell")
    chain = ["Cell", "Edge", "Cell"]
    cdecl, cnt = b.local("cnt", reduce_over(chain, 1, init=0), "Integer")
    st = b.stage(b.do(FULL, cdecl,
                      b.assign(out, reduce_over(chain, inn.nbh, init=(-cnt) * inn.at())),
                      b.assign(out, inn.at() + 0.1 * out.at())), location="Cell")
    b.stencil(b.multistage("Parallel", st))
    return b


def unstructured_gradient():
    # reference/reference_gradient.hpp:28-60
    b = IIRBuilder("gradient", grid="Unstructured")
    cell = b.field("cell_field", "k", "Cell")
    edge = b.field("edge_field", "k", "Edge")
    s1 = b.stage(b.do(FULL, b.assign(edge, reduce_over(["Edge", "Cell"], cell.nbh, weights=[1.0, -1.0]))),
                 location="Edge")
    s2 = b.stage(b.do(FULL, b.assign(cell, reduce_over(["Cell", "Edge"], edge.nbh,
                                                       weights=[0.5, 0.0, 0.0, 0.5]))), location="Cell")
    b.stencil(b.multistage("Parallel", s1, s2))
    return b


def unstructured_diamond():
    # reference/reference_diamond.hpp:28-50 (chain E -> C -> V, 4 vertices)
    b = IIRBuilder("diamond", grid="Unstructured")
    edge = b.field("edge_field", "k", "Edge")
    vert = b.field("vertex_field", "k", "Vertex")
    b.stencil(b.multistage("Parallel", b.stage(
        b.do(FULL, b.assign(edge, reduce_over(["Edge", "Cell", "Vertex"], vert.nbh))), location="Edge")))
    return b


def unstructured_diamond_weights():
    # reference/reference_diamondWeights.hpp:35-62 (weights are expressions of edge fields)
    b = IIRBuilder("diamondWeights", grid="Unstructured")
    out = b.field("out", "k", "Edge")
    iel = b.field("inv_edge_length", "k", "Edge")
    ivl = b.field("inv_vert_length", "k", "Edge")
    inn = b.field("in", "k", "Vertex")
    w = [iel.at() * iel.at(), iel.at() * iel.at(), ivl.at() * ivl.at(), ivl.at() * ivl.at()]
    b.stencil(b.multistage("Parallel", b.stage(
        b.do(FULL, b.assign(out, reduce_over(["Edge", "Cell", "Vertex"], inn.nbh, weights=w))),
        location="Edge")))
    return b


def unstructured_intp():
    # reference/reference_intp.hpp:28-50 (chain C-E-C-E-C, 9 cells)
    b = IIRBuilder("intp", grid="Unstructured")
    inn = b.field("in", "k", "Cell")
    out = b.field("out", "k", "Cell")
    b.stencil(b.multistage("Parallel", b.stage(
        b.do(FULL, b.assign(out, reduce_over(["Cell", "Edge", "Cell", "Edge", "Cell"], inn.nbh))),
        location="Cell")))
    return b


# ---- reference toylib integration stencils (dawn/test/integration-test/unstructured/
# GenerateUnstructuredStencils.cpp; expected values: ToylibIntegrationTestCompareOutput.cpp) ---------
def _ico(name):
    return IIRBuilder(name, grid="Unstructured")


def ico_copy_cell():  # :37-60
    b = _ico("copyCell")
    i, o = b.field("in_field", "k", "Cell"), b.field("out_field", "k", "Cell")
    b.stencil(b.multistage("Parallel", b.stage(b.do(FULL, b.assign(o, i.at())), location="Cell")))
    return b


def ico_copy_edge():  # :62-82
    b = _ico("copyEdge")
    i, o = b.field("in_field", "k", "Edge"), b.field("out_field", "k", "Edge")
    b.stencil(b.multistage("Parallel", b.stage(b.do(FULL, b.assign(o, i.at())), location="Edge")))
    return b


def ico_accumulate_edge_to_cell():  # :84-110
    b = _ico("accumulateEdgeToCell")
    i, o = b.field("in_field", "k", "Edge"), b.field("out_field", "k", "Cell")
    b.stencil(b.multistage("Parallel", b.stage(
        b.do(FULL, b.assign(o, reduce_over(["Cell", "Edge"], i.nbh))), location="Cell")))
    return b


def ico_vertical_sum():  # :112-137
    b = _ico("verticalSum")
    i, o = b.field("in_field", "k", "Cell"), b.field("out_field", "k", "Cell")
    b.stencil(b.multistage("Parallel", b.stage(
        b.do(Interval("Start", "End", 1, -1), b.assign(o, i.at(1) + i.at(-1))), location="Cell")))
    return b


def ico_tridiagonal_solve():  # :139-205 (Thomas algorithm on cells, test "verticalSolver")
    b = _ico("tridiagonalSolve")
    a, bb, c, d = (b.field(n, "k", "Cell") for n in "abcd")
    md, m = b.local("m", 1.0 / (bb.at() - a.at() * c.at(-1)))
    f0 = b.stage(b.do(Interval("Start", "Start", 0, 0), b.assign(c, c.at() / bb.at()),
                      b.assign(d, d.at() / bb.at())), location="Cell")
    f1 = b.stage(b.do(Interval("Start", "End", 1, 0), md, b.assign(c, c.at() * m),
                      b.assign(d, (d.at() - a.at() * d.at(-1)) * m)), location="Cell")
    f2 = b.stage(b.do(Interval("Start", "End", 0, -1), b.assign(d, d.at() - c.at() * d.at(1))),
                 location="Cell")
    b.stencil(b.multistage("Forward", f0, f1), b.multistage("Backward", f2))
    return b


def ico_sparse_dimension(twice=False):  # :382-457 (un-flagged weights longer than the chain)
    b = _ico("sparseDimensionTwice" if twice else "sparseDimension")
    cell = b.field("cell_field", "k", "Cell")
    edge = b.field("edge_field", "k", "Edge")
    sp = b.field("sparse_dim", "k", sparse_chain=["Cell", "Edge"])
    red = lambda: reduce_over(["Cell", "Edge"], edge.nbh * sp.nbh, weights=[1.0, 1.0, 1.0, 1.0])
    stmts = [b.assign(cell, red())] * (2 if twice else 1)
    if twice:
        stmts = [b.assign(cell, red()), b.assign(cell, red())]
    b.stencil(b.multistage("Parallel", b.stage(b.do(FULL, *stmts), location="Cell")))
    return b


def ico_nested_simple():  # :459-490 (accesses carry no offset flag, as in the reference)
    b = _ico("nestedSimple")
    cell, edge, vert = (b.field("cell_field", "k", "Cell"), b.field("edge_field", "k", "Edge"),
                        b.field("vertex_field", "k", "Vertex"))
    b.stencil(b.multistage("Parallel", b.stage(b.do(FULL, b.assign(
        cell, reduce_over(["Cell", "Edge"], reduce_over(["Edge", "Vertex"], vert.at())))),
        location="Cell")))
    return b


def ico_nested_with_field():  # :492-525
    b = _ico("nestedWithField")
    cell, edge, vert = (b.field("cell_field", "k", "Cell"), b.field("edge_field", "k", "Edge"),
                        b.field("vertex_field", "k", "Vertex"))
    b.stencil(b.multistage("Parallel", b.stage(b.do(FULL, b.assign(
        cell, reduce_over(["Cell", "Edge"], edge.at() + reduce_over(["Edge", "Vertex"], vert.at())))),
        location="Cell")))
    return b


def ico_nested_with_sparse():  # :527-569
    b = _ico("nestedWithSparse")
    cell, edge, vert = (b.field("cell_field", "k", "Cell"), b.field("edge_field", "k", "Edge"),
                        b.field("vertex_field", "k", "Vertex"))
    ce = b.field("ce_sparse", "k", sparse_chain=["Cell", "Edge"])
    ev = b.field("ev_sparse", "k", sparse_chain=["Edge", "Vertex"])
    b.stencil(b.multistage("Parallel", b.stage(b.do(FULL, b.assign(
        cell, reduce_over(["Cell", "Edge"], edge.at() * ce.at() +
                          reduce_over(["Edge", "Vertex"], vert.at() * ev.at())))), location="Cell")))
    return b


def ico_sparse_assignment(n):  # :571-779
    b = _ico("sparseAssignment%d" % n)
    ecv = ["Edge", "Cell", "Vertex"]
    if n in (0, 1):
        vn = b.field("vn", "k", sparse_chain=ecv)
        u, v = b.field("uVert", "k", "Vertex"), b.field("vVert", "k", "Vertex")
        if n == 0:
            nx, ny = b.field("nx", "k", "Vertex"), b.field("ny", "k", "Vertex")
        else:
            nx, ny = b.field("nx", "k", sparse_chain=ecv), b.field("ny", "k", sparse_chain=ecv)
        body = loop_over(ecv, b.assign(vn, u.nbh * nx.nbh + v.nbh * ny.nbh))
        loc = "Edge"
    elif n == 2:
        sp = b.field("sparse", "k", sparse_chain=ecv)
        e, v = b.field("e", "k", "Edge"), b.field("v", "k", "Vertex")
        body = loop_over(ecv, b.assign(sp, lit(-4.0) * e.at() + v.nbh))
        loc = "Edge"
    elif n == 3:
        chain = ["Cell", "Edge", "Cell", "Edge", "Cell"]
        sp = b.field("sparse", "k", sparse_chain=chain)
        A, B = b.field("A", "k", "Cell"), b.field("B", "k", "Cell")
        body = loop_over(chain, b.assign(sp, A.nbh - B.at()))
        loc = "Cell"
    elif n == 4:
        sp = b.field("sparse", "k", sparse_chain=["Cell", "Edge"])
        v = b.field("v", "k", "Vertex")
        body = loop_over(["Cell", "Edge"], b.assign(sp, reduce_over(["Edge", "Vertex"], v.at())))
        loc = "Cell"
    else:
        sp = b.field("sparse", "k", sparse_chain=["Cell", "Edge"])
        v, c = b.field("v", "k", "Vertex"), b.field("c", "k", "Cell")
        body = loop_over(["Cell", "Edge"], b.assign(sp, reduce_over(
            ["Edge", "Vertex"], v.at() * reduce_over(["Vertex", "Cell"], c.at()))))
        loc = "Cell"
    b.stencil(b.multistage("Parallel", b.stage(b.do(FULL, body), location=loc)))
    return b


def ico_horizontal_vertical():  # :781-817 (horizontal-only, sparse horizontal-only and vertical fields)
    b = _ico("horizontalVertical")
    hor = b.field("horizontal", "", "Edge")
    sph = b.field("horizontal_sparse", "", sparse_chain=["Edge", "Cell"])
    ver = b.field("vertical", "k")
    full, o1, o2 = (b.field(n, "k", "Edge") for n in ("full", "out1", "out2"))
    b.stencil(b.multistage("Parallel", b.stage(b.do(
        FULL, b.assign(o1, full.at() + hor.at() + ver.at()),
        b.assign(o2, reduce_over(["Edge", "Cell"], sph.at()))), location="Edge")))
    return b


def ico_vertical_indirection():  # :819-845 (the reference spells the name "verticalIndirecion")
    b = _ico("verticalIndirecion")
    i, o, kidx = (b.field(n, "k", "Cell") for n in ("in", "out", "kidx"))
    b.stencil(b.multistage("Parallel", b.stage(
        b.do(Interval("Start", "End", 0, -1), b.assign(o, i.at(k_from=kidx))), location="Cell")))
    return b


def ico_iteration_space():  # :847-872 (Interval levels are the subdomain magic numbers)
    b = _ico("iterationSpaceUnstructured")
    o, i1, i2 = (b.field(n, "k", "Cell") for n in ("out", "in_1", "in_2"))
    s1 = b.stage(b.do(FULL, b.assign(o, i1.at())), location="Cell", i_range=Interval(3000, 4000, 0, 0))
    s2 = b.stage(b.do(FULL, b.assign(o, i2.at())), location="Cell", i_range=Interval(2000, 3000, 0, 0))
    b.stencil(b.multistage("Parallel", s1, s2))
    return b


def ico_global_var():  # :874-899 (test globalVar: dt set to 2 at run time)
    b = _ico("globalVar")
    i, o = b.field("in_field", "k", "Cell"), b.field("out_field", "k", "Cell")
    dt = b.global_var("dt", "Double", 0.5)
    b.stencil(b.multistage("Parallel", b.stage(b.do(FULL, b.assign(o, dt * i.at())), location="Cell")))
    return b


def ico_temp_field_allocation():  # :941-966
    b = _ico("tempFieldAllocation")
    i = b.field("in_field", "k", "Cell")
    t = b.tmp("tmp_field", "k", "Cell")
    o = b.field("out_field", "k", "Cell")
    b.stencil(b.multistage("Parallel", b.stage(
        b.do(FULL, b.assign(t, lit(1.0)), b.assign(o, t.at() + i.at())), location="Cell")))
    return b


def ico_sparse_temp_field_allocation():  # :968-998 (a temporary with a sparse dimension)
    b = _ico("sparseTempFieldAllocation")
    i = b.field("in_field", "k", "Cell")
    t = b.tmp("tmp_field", "k", sparse_chain=["Cell", "Edge"])
    o = b.field("out_field", "k", "Cell")
    b.stencil(b.multistage("Parallel", b.stage(b.do(
        FULL, loop_over(["Cell", "Edge"], b.assign(t, lit(1.0))),
        b.assign(o, reduce_over(["Cell", "Edge"], t.at() * i.at()))), location="Cell")))
    return b


def ico_temp_field(split=True):
    # GenerateUnstructuredStencils.cpp:900-937 `tempField`: dense = sum over E>C>V>E of sparse; then a neighbour
    # loop sparse = dense(neighbour).  The generator runs PassFieldVersioning (the cycle dense <- sparse <- dense
    # with a horizontal access makes it rename `sparse` on the right-hand side of the first statement to a new
    # version, PassFieldVersioning.cpp:268-286, CreateVersionAndRename.cpp:52-60) and PassFixVersionedInputFields
    # (a Parallel multistage in front copies the original into the version; a sparse field is copied inside a
    # neighbour loop over its chain, PassFixVersionedInputFields.cpp:36-63,176-186).  It does NOT run
    # PassStageSplitter, so the second stage as generated still gathers `dense` right after writing it
    # (split=False: every backend with one thread per element races there, the naive loops depend on the element
    # order, and no test of the reference runs the generated file; the b200-ico emitter rejects it).  split=True
    # is the same program after the stage split Dawn's default pipeline applies: the fixture.
    b = _ico("tempField")
    chain = ["Edge", "Cell", "Vertex", "Edge"]
    dense = b.field("dense", "k", "Edge")
    sparse = b.field("sparse", "k", sparse_chain=chain)
    sparse0 = b.tmp("sparse_0", "k", sparse_chain=chain)
    copy = b.stage(b.do(FULL, loop_over(chain, b.assign(sparse0, sparse.at()))), location="Edge")
    red = b.assign(dense, reduce_over(chain, sparse0.at(), init=0.0))
    fill = loop_over(chain, b.assign(sparse, dense.nbh))
    if split:
        stages = [b.stage(b.do(FULL, red), location="Edge"), b.stage(b.do(FULL, fill), location="Edge")]
    else:
        stages = [b.stage(b.do(FULL, red, fill), location="Edge")]
    b.stencil(b.multistage("Parallel", copy), b.multistage("Parallel", *stages))
    return b


def ico_reduction_in_if_conditional():  # :1000-1036 (a reduction as the condition of an if)
    b = _ico("reductionInIfConditional")
    c = b.field("c_field", "k", "Cell")
    sp = b.field("sparse_field", "k", sparse_chain=["Cell", "Edge"])
    e = b.field("e_field", "k", "Edge")
    o = b.field("out_field", "k", "Cell")
    ce = ["Cell", "Edge"]
    b.stencil(b.multistage("Parallel", b.stage(b.do(FULL, b.if_(
        reduce_over(ce, e.at()) < lit(3.0),
        [b.assign(o, reduce_over(ce, sp.at() * lit(2.0)))],
        [b.assign(o, reduce_over(ce, sp.at() * lit(4.0)))])), location="Cell")))
    return b


def ico_reduction_with_center(kind):  # :1038-1134 (chains that include the centre element)
    cec = ["Cell", "Edge", "Cell"]
    name = {0: "reductionWithCenter", 1: "reductionWithCenterSparse",
            2: "reductionAndFillWithCenterSparse"}[kind]
    b = _ico(name)
    cin, cout = b.field("cin_field", "k", "Cell"), b.field("cout_field", "k", "Cell")
    if kind == 0:
        body = [b.assign(cout, reduce_over(cec, cin.nbh, include_center=True))]
    elif kind == 1:
        sp = b.field("sparse", "k", sparse_chain=cec, include_center=True)
        body = [b.assign(cout, reduce_over(cec, cin.nbh * sp.at(), include_center=True))]
    else:
        sp = b.tmp("sparse", "k", sparse_chain=cec, include_center=True)
        body = [loop_over(cec, b.assign(sp, lit(2.0)), include_center=True),
                b.assign(cout, reduce_over(cec, cin.nbh * sp.at(), include_center=True))]
    b.stencil(b.multistage("Parallel", b.stage(b.do(FULL, *body), location="Cell")))
    return b


def ico_overwrite_gathered():
    """Not from the reference: a later stage overwrites a field an earlier stage of the same location type
    gathered through a chain (write-after-read).  cuda-ico runs one kernel per stage
    (CudaIcoCodeGen.cpp:1316-1464), so this is valid and deterministic there; a backend that fuses stages
    must keep the overwrite out of the launch that still gathers the old values."""
    b = _ico("overwriteGathered")
    g, h, o, o2 = (b.field(n, "k", "Cell") for n in ("g", "h", "o", "o2"))
    s1 = b.stage(b.do(FULL, b.assign(o, reduce_over(["Cell", "Edge", "Cell"], g.nbh))), location="Cell")
    s2 = b.stage(b.do(FULL, b.assign(g, h.at() * 2.0)), location="Cell")
    s3 = b.stage(b.do(FULL, b.assign(o2, o.at() + g.at())), location="Cell")
    b.stencil(b.multistage("Parallel", s1, s2, s3))
    return b


# ---- reference integration-test stencils (gtclang/test/integration-test/CodeGen/<name>.cpp) --------
def kcache_fill():
    b = IIRBuilder("kcache_fill")
    i, o = b.field("in"), b.field("out")
    st = b.stage(
        b.do(Interval("Start", "Start", 2, 2),
             b.assign(o, i() + i(0, 0, 1) + i(0, 0, 2) + i(0, 0, -1) + i(0, 0, -2))),
        b.do(Interval("Start", "End", 3, -2),
             b.assign(o, i() + i(0, 0, -1) + i(0, 0, 1) + i(0, 0, 2) + o(0, 0, -1))),
        b.do(Interval("End", "End", -1, -1), b.assign(o, i() + i(0, 0, -1) + i(0, 0, 1) + o(0, 0, -1))),
        b.do(Interval("End", "End", 0, 0), b.assign(o, i() + i(0, 0, -1) + o(0, 0, -1))))
    b.stencil(b.multistage("Forward", st))
    return b


def kcache_flush():
    b = IIRBuilder("kcache_flush")
    i, o = b.field("in"), b.field("out")
    st = b.stage(
        b.do(Interval("Start", "Start", 1, 2), b.assign(o, i() + i(0, 0, 1) + i(0, 0, 2) + i(0, 0, -1))),
        b.do(Interval("Start", "End", 3, -2),
             b.assign(o, i() + i(0, 0, -1) + i(0, 0, 1) + i(0, 0, 2) + o(0, 0, -1) + o(0, 0, -2))),
        b.do(Interval("End", "End", -1, -1),
             b.assign(o, i() + i(0, 0, -1) + i(0, 0, 1) + o(0, 0, -1) + o(0, 0, -2) + o(0, 0, -3))),
        b.do(Interval("End", "End", 0, 0), b.assign(o, i() + i(0, 0, -1) + o(0, 0, -1))))
    b.stencil(b.multistage("Forward", st))
    return b


def kcache_fill_backward():
    b = IIRBuilder("kcache_fill_backward")
    i, o = b.field("in"), b.field("out")
    st = b.stage(
        b.do(Interval("End", "End", -2, -2),
             b.assign(o, i() + i(0, 0, -1) + i(0, 0, -2) + i(0, 0, 1) + i(0, 0, 2))),
        b.do(Interval("Start", "End", 2, -3),
             b.assign(o, i() + i(0, 0, 1) + i(0, 0, -1) + i(0, 0, -2) + o(0, 0, 1))),
        b.do(Interval("Start", "Start", 1, 1), b.assign(o, i() + i(0, 0, 1) + i(0, 0, -1) + o(0, 0, 1))),
        b.do(Interval("Start", "Start", 0, 0), b.assign(o, i() + i(0, 0, 1) + o(0, 0, 1))))
    b.stencil(b.multistage("Backward", st))
    return b


def local_kcache():
    b = IIRBuilder("local_kcache")
    oa, ob, oc, od = b.field("out_a"), b.field("out_b"), b.field("out_c"), b.field("out_d")
    a, bb, c, d = b.tmp("a"), b.tmp("b"), b.tmp("c"), b.tmp("d")
    fwd = b.stage(
        b.do(Interval("Start", "Start", 0, 0), b.assign(c, 2.1), b.assign(d, 3.1), b.assign(oc, c()),
             b.assign(od, d())),
        b.do(Interval("Start", "Start", 1, 1), b.assign(d, 4.1), b.assign(od, d()), b.assign(oc, c())),
        b.do(Interval("Start", "End", 2, 0), b.assign(c, c(0, 0, -1) * 1.1),
             b.assign(d, d(0, 0, -1) * 1.1 + d(0, 0, -2) * 1.2), b.assign(oc, c()), b.assign(od, d())))
    bwd = b.stage(
        b.do(Interval("End", "End", 0, 0), b.assign(a, 2.1), b.assign(bb, 3.1), b.assign(oa, a()),
             b.assign(ob, bb())),
        b.do(Interval("End", "End", -1, -1), b.assign(bb, 4.1), b.assign(ob, bb()), b.assign(oa, a())),
        b.do(Interval("Start", "End", 0, -2), b.assign(a, a(0, 0, 1) * 1.1),
             b.assign(bb, bb(0, 0, 1) * 1.1 + bb(0, 0, 2) * 1.2), b.assign(oa, a()), b.assign(ob, bb())))
    b.stencil(b.multistage("Forward", fwd), b.multistage("Backward", bwd))
    return b


def intervals02():
    # k_flat = 4 (an interval level of the DSL; any fixed level below k_end)
    b = IIRBuilder("intervals02")
    i, o = b.field("in"), b.field("out")
    st = b.stage(
        b.do(Interval("Start", "Start", 1, 1), b.assign(o, 0)),
        b.do(Interval("Start", 4, 2, 0), b.assign(o, o(0, 0, -1) + i() + 1)),
        b.do(Interval(4, 4, 2, 2), b.assign(o, 0)),
        b.do(Interval(4, "End", 3, -1), b.assign(o, o(0, 0, -1) + i() + 3)))
    b.stencil(b.multistage("Forward", st))
    return b


def intervals03():
    b = IIRBuilder("intervals03")
    i, o = b.field("in"), b.field("out")
    st = b.stage(
        b.do(Interval("End", "End", -1, -1), b.assign(o, 0)),
        b.do(Interval(6, "End", 0, -2), b.assign(o, o(0, 0, 1) + i() + 3)),
        b.do(Interval(6, 6, -2, -2), b.assign(o, 1)),
        b.do(Interval("Start", 6, 0, -3), b.assign(o, o(0, 0, 1) + i() + 3)))
    b.stencil(b.multistage("Backward", st))
    return b


def asymmetric():
    b = IIRBuilder("asymmetric_stencil")
    i, o = b.field("in"), b.field("out")
    tmp = b.tmp("tmp")
    s1 = b.stage(b.do(Interval("Start", "Start", 0, 0), b.assign(tmp, i(1, 0) + i(0, -2))),
                 b.do(Interval("Start", "End", 1, 0), b.assign(tmp, i(-3, 0) + i(0, 1))))
    s2 = b.stage(b.do(Interval("Start", "Start", 0, 0), b.assign(o, 1)),
                 b.do(Interval("Start", "End", 1, 0), b.assign(o, tmp(0, 0, -1))))
    b.stencil(b.multistage("Forward", s1, s2))
    return b


def coriolis():
    b = IIRBuilder("coriolis_stencil")
    ut, un, vt, vn, fc = (b.field(n) for n in ("u_tens", "u_nnow", "v_tens", "v_nnow", "fc"))
    d1, zn = b.local("z_fv_north", fc() * (vn() + vn(1, 0)), "Double")
    d2, zs = b.local("z_fv_south", fc(0, -1) * (vn(0, -1) + vn(1, -1)), "Double")
    d3, ze = b.local("z_fu_east", fc() * (un() + un(0, 1)), "Double")
    d4, zw = b.local("z_fu_west", fc(-1, 0) * (un(-1, 0) + un(-1, 1)), "Double")
    st = b.stage(b.do(FULL, d1, d2, b.assign(ut, 0.25 * (zn + zs), op="+="), d3, d4,
                      b.assign(vt, 0.25 * (ze + zw), op="-=")))
    b.stencil(b.multistage("Parallel", st))
    return b


def globals_stencil():
    b = IIRBuilder("globals_stencil")
    vr = b.global_var("var_runtime", "Integer", 1)
    vd = b.global_var("var_default", "Double", 2.0)
    vc = b.global_var("var_compiletime", "Boolean", 1)
    i, o = b.field("in"), b.field("out")
    st = b.stage(b.do(FULL, b.if_(vc, [b.assign(o, i() + vr + vd)])))
    b.stencil(b.multistage("Parallel", st))
    return b


def kparallel_solver():
    b = IIRBuilder("kparallel_solver")
    d, a, bb, c = b.field("d"), b.field("a"), b.field("b"), b.field("c")
    s1 = b.stage(b.do(Interval("Start", "Start", 0, 0), b.assign(c, a(0, 0, 1)), b.assign(d, bb() * 2.1)),
                 b.do(Interval("Start", "End", 1, 0), b.assign(c, a(0, 0, -1)), b.assign(d, bb() * 2.0 - 1)))
    s2 = b.stage(b.do(Interval("Start", "End", 0, -1), b.assign(d, d() + a(), op="-=")))
    b.stencil(b.multistage("Parallel", s1), b.multistage("Parallel", s2))
    return b


# ---- the rest of gtclang/test/integration-test/CodeGen/CMakeLists.txt:169-198 ------------------------
def intervals_stencil():  # intervals_stencil.cpp:17-42 (k_flat = k_start + 4)
    b = IIRBuilder("intervals_stencil")
    i, o = b.field("in"), b.field("out")
    st = b.stage(b.do(Interval("Start", 4, 0, 0), b.assign(o, i() + 1)),
                 b.do(Interval(4, 4, 1, 1), b.assign(o, i() + 2)),
                 b.do(Interval(4, "End", 2, 0), b.assign(o, i() + 3)))
    b.stencil(b.multistage("Parallel", st))
    return b


def intervals01():  # intervals01.cpp: levels k_start and k_flat+1 and k_end are never written
    b = IIRBuilder("intervals01")
    i, o = b.field("in"), b.field("out")
    st = b.stage(b.do(Interval("Start", 4, 1, 0), b.assign(o, i() + 1)),
                 b.do(Interval(4, "End", 2, -1), b.assign(o, i() + 3)))
    b.stencil(b.multistage("Parallel", st))
    return b


def kcache_fill_kparallel():  # kcache_fill_kparallel.cpp: overlapping intervals, reads k-1 and k+1
    b = IIRBuilder("kcache_fill_kparallel")
    i, o = b.field("in"), b.field("out")
    st = b.stage(b.do(Interval("Start", "Start", 0, 1), b.assign(o, i() + i(0, 0, 1))),
                 b.do(Interval("Start", "End", 1, -1), b.assign(o, i() + i(0, 0, -1) + i(0, 0, 1))),
                 b.do(Interval("End", "End", -1, 0), b.assign(o, i() + i(0, 0, -1))))
    b.stencil(b.multistage("Parallel", st))
    return b


def kcache_epflush():  # kcache_epflush.cpp: temporaries with vertical offsets across two multistages
    b = IIRBuilder("kcache_epflush")
    i, o = b.field("in"), b.field("out")
    bb, tmp = b.tmp("b"), b.tmp("tmp")
    fwd = b.stage(b.do(Interval("Start", "Start", 0, 0), b.assign(tmp, i())),
                  b.do(Interval("Start", "End", 1, 0), b.assign(tmp, i() * 2), b.assign(bb, tmp(0, 0, -1))))
    bwd = b.stage(b.do(Interval("End", "End", -2, 0), b.assign(o, 2.2)),
                  b.do(Interval("End", "End", -3, -3),
                       b.assign(o, tmp(0, 0, 3) + tmp(0, 0, 2) + tmp(0, 0, 1))),
                  b.do(Interval("Start", "End", 0, -4), b.assign(tmp, bb()), b.assign(o, tmp(0, 0, 1))))
    b.stencil(b.multistage("Forward", fwd, caches={tmp.access_id: k_cache("CP_EPFlush")}),
              b.multistage("Backward", bwd))
    return b


def lap_stencil():  # lap.cpp: temporary read at +-1 of a stage reading +-2
    b = IIRBuilder("lap")
    i, o = b.field("in"), b.field("out")
    tmp = b.tmp("tmp")
    s1 = b.stage(b.do(FULL, b.assign(tmp, i(0, -2) + i(0, 2) + i(-2, 0) + i(2, 0))))
    s2 = b.stage(b.do(FULL, b.assign(o, tmp(0, -1) + tmp(0, 1) + tmp(-1, 0) + tmp(1, 0))))
    b.stencil(b.multistage("Parallel", s1, s2))
    return b


def hori_diff_02():  # hori_diff_stencil_02.cpp: laplacian of a laplacian
    b = IIRBuilder("hori_diff_stencil_02")
    u, o = b.field("u"), b.field("out")
    lap = b.tmp("lap")

    def laplacian(f):
        return f(1, 0) + f(-1, 0) + f(0, 1) + f(0, -1) - 4.0 * f()
    s1 = b.stage(b.do(FULL, b.assign(lap, laplacian(u))))
    s2 = b.stage(b.do(FULL, b.assign(o, laplacian(lap))))
    b.stencil(b.multistage("Parallel", s1, s2))
    return b


def hd_smagorinsky():  # hd_smagorinsky.cpp:47-100 (stencil functions inlined, two stages)
    from iir_builder import call
    b = IIRBuilder("hd_smagorinsky_stencil")
    u_out, v_out, u_in, v_in, hdmaskvel = (b.field(n) for n in ("u_out", "v_out", "u_in", "v_in", "hdmaskvel"))
    crlavo, crlavu, crlato, crlatu, acrlat0 = (b.field(n, "j") for n in
                                               ("crlavo", "crlavu", "crlato", "crlatu", "acrlat0"))
    eddlon, eddlat, tau_smag, weight_smag = (b.field(n) for n in ("eddlon", "eddlat", "tau_smag", "weight_smag"))
    T_sqr_s, S_sqr_uv = b.tmp("T_sqr_s"), b.tmp("S_sqr_uv")

    def delta(di, dj, f):
        return f(di, dj) - f()

    def avg(di, dj, f):
        return 0.5 * (f(di, dj) + f())

    def laplacian(f, lo, lu):
        return f(1, 0) + f(-1, 0) - 2.0 * f() + lo() * delta(0, 1, f) + lu() * delta(0, -1, f)
    d1, fdx = b.local("frac_1_dx", acrlat0() * eddlon(), "Double", True)
    d2, fdy = b.local("frac_1_dy", eddlat() / lit(6371.229e3), "Double", True)
    d3, T_s = b.local("T_s", delta(0, -1, v_in) * fdy - delta(-1, 0, u_in) * fdx, "Double", True)
    d4, S_uv = b.local("S_uv", delta(0, 1, u_in) * fdy + delta(1, 0, v_in) * fdx, "Double", True)
    s1 = b.stage(b.do(FULL, d1, d2, d3, b.assign(T_sqr_s, T_s * T_s), d4, b.assign(S_sqr_uv, S_uv * S_uv)))
    d5, hdw = b.local("hdweight", weight_smag() * hdmaskvel(), "Double", True)
    d6, smag_u = b.local("smag_u", tau_smag() * call("gridtools::dawn::math::sqrt",
                                                     avg(1, 0, T_sqr_s) + avg(0, -1, S_sqr_uv)) - hdw, "Double")
    d7, smag_v = b.local("smag_v", tau_smag() * call("gridtools::dawn::math::sqrt",
                                                     avg(0, 1, T_sqr_s) + avg(-1, 0, S_sqr_uv)) - hdw, "Double")
    clamp = lambda x: call("gridtools::dawn::math::min", lit(0.5), call("gridtools::dawn::math::max", lit(0.0), x))
    d8, lapu = b.local("lapu", laplacian(u_in, crlato, crlatu), "Double", True)
    d9, lapv = b.local("lapv", laplacian(v_in, crlavo, crlavu), "Double", True)
    s2 = b.stage(b.do(FULL, d5, d6, b.assign(smag_u, clamp(smag_u)), d7, b.assign(smag_v, clamp(smag_v)),
                      d8, d9, b.assign(u_out, u_in() + smag_u * lapu), b.assign(v_out, v_in() + smag_v * lapv)))
    b.stencil(b.multistage("Parallel", s1, s2))
    return b


def p_grad_c():  # p_grad_c.cpp: conditional ij temporary, mixed horizontal + vertical offsets, +=
    b = IIRBuilder("p_grad_c")
    dt2 = b.global_var("dt2", "Double", 0.5)
    hyd = b.global_var("hydrostatic", "Boolean", 1)
    delpc = b.field("delpc", "ij")
    pkc, gz, uc, vc, rdxc, rdyc = (b.field(n) for n in ("pkc", "gz", "uc", "vc", "rdxc", "rdyc"))
    wk = b.tmp("wk")
    s1 = b.stage(b.do(Interval("Start", "End", 0, -1),
                      b.if_(hyd, [b.assign(wk, pkc(0, 0, 1) - pkc())], [b.assign(wk, delpc())])))
    s2 = b.stage(b.do(
        Interval("Start", "End", 0, -1),
        b.assign(uc, dt2 * rdxc() / (wk(-1, 0) + wk()) *
                 ((gz(-1, 0, 1) - gz()) * (pkc(0, 0, 1) - pkc(-1, 0)) +
                  (gz(-1, 0) - gz(0, 0, 1)) * (pkc(-1, 0, 1) - pkc())), op="+="),
        b.assign(vc, dt2 * rdyc() / (wk(0, -1) + wk()) *
                 ((gz(0, -1, 1) - gz()) * (pkc(0, 0, 1) - pkc(0, -1)) +
                  (gz(0, -1) - gz(0, 0, 1)) * (pkc(0, -1, 1) - pkc())), op="+=")))
    b.stencil(b.multistage("Parallel", s1, s2))
    return b


def stencil_desc_ast(n):  # stencil_desc_ast.cpp: control flow of the stencil description
    from iir_builder import If, VarDecl
    b = IIRBuilder("stencil_desc_ast_%02d" % n)
    vr = b.global_var("var_runtime", "Integer", 1)
    vc = b.global_var("var_compiletime", "Integer", 2)
    i, o = b.field("in"), b.field("out")
    if n == 3:    # nested ifs around one stencil
        c0 = b.stencil(b.multistage("Parallel", b.stage(b.do(FULL, b.assign(o, i() + vr + vc)))))
        b.control_flow(If(vr.eq(1), [If(vc.eq(2), [c0])]))
    elif n == 4:  # two stencils, the second one level deeper
        c0 = b.stencil(b.multistage("Parallel", b.stage(b.do(FULL, b.assign(o, 0.0)))))
        c1 = b.stencil(b.multistage("Parallel", b.stage(b.do(FULL, b.assign(o, 2 + i(), op="+=")))))
        b.control_flow(If(vc.eq(2), [If(vc.ne(1), [c0, If(vc.ne(1), [c1])])]))
    else:         # 7: local variables of the description feed the condition
        c0 = b.stencil(b.multistage("Parallel", b.stage(b.do(FULL, b.assign(o, 2 * i())))))
        d1, some = b.local("some_var", 5.0, "Double")
        d2, other = b.local("some_other_var", vc, "Double")
        b.control_flow(If(vc.eq(2), [d1, d2, b.assign(some, 1.0, op="+="),
                                     If((vc + some + other).eq(10), [c0])]))
    return b


ALL = {
    "hori_diff_type2_stencil": hori_diff_type2,
    "hori_diff_stencil_01": hori_diff_01,
    "hori_diff_stencil": hori_diff_dawn4py,
    "tridiagonal_solve": tridiagonal_solve,
    "laplacian_stencil": tutorial_laplacian,
    "copy_stencil": copy_stencil,
    "global_indexing": global_indexing,
    "nonoverlapping_stencil": nonoverlapping,
    "kcache_fill": kcache_fill,
    "kcache_flush": kcache_flush,
    "kcache_fill_backward": kcache_fill_backward,
    "local_kcache": local_kcache,
    "intervals02": intervals02,
    "intervals03": intervals03,
    "asymmetric_stencil": asymmetric,
    "coriolis_stencil": coriolis,
    "globals_stencil": globals_stencil,
    "kparallel_solver": kparallel_solver,
    "ICON_laplacian_stencil": icon_laplacian,
    "ICON_nabla4_stencil": icon_nabla4,
    "ICON_laplacian_diamond_stencil": icon_laplacian_diamond,
    "ICON_gradient": icon_gradient,
    "generate_versioned_field": generate_versioned_field,
    "diffusion": unstructured_diffusion,
    "gradient": unstructured_gradient,
    "diamond": unstructured_diamond,
    "diamondWeights": unstructured_diamond_weights,
    "intp": unstructured_intp,
    "intervals_stencil": intervals_stencil,
    "intervals01": intervals01,
    "kcache_fill_kparallel": kcache_fill_kparallel,
    "kcache_epflush": kcache_epflush,
    "lap": lap_stencil,
    "hori_diff_stencil_02": hori_diff_02,
    "hd_smagorinsky_stencil": hd_smagorinsky,
    "p_grad_c": p_grad_c,
    "stencil_desc_ast_03": lambda: stencil_desc_ast(3),
    "stencil_desc_ast_04": lambda: stencil_desc_ast(4),
    "stencil_desc_ast_07": lambda: stencil_desc_ast(7),
    "overwriteGathered": ico_overwrite_gathered,
    "copyCell": ico_copy_cell,
    "copyEdge": ico_copy_edge,
    "accumulateEdgeToCell": ico_accumulate_edge_to_cell,
    "verticalSum": ico_vertical_sum,
    "tridiagonalSolve": ico_tridiagonal_solve,
    "sparseDimension": ico_sparse_dimension,
    "sparseDimensionTwice": lambda: ico_sparse_dimension(True),
    "nestedSimple": ico_nested_simple,
    "nestedWithField": ico_nested_with_field,
    "nestedWithSparse": ico_nested_with_sparse,
    "sparseAssignment0": lambda: ico_sparse_assignment(0),
    "sparseAssignment1": lambda: ico_sparse_assignment(1),
    "sparseAssignment2": lambda: ico_sparse_assignment(2),
    "sparseAssignment3": lambda: ico_sparse_assignment(3),
    "sparseAssignment4": lambda: ico_sparse_assignment(4),
    "sparseAssignment5": lambda: ico_sparse_assignment(5),
    "horizontalVertical": ico_horizontal_vertical,
    "verticalIndirecion": ico_vertical_indirection,
    "iterationSpaceUnstructured": ico_iteration_space,
    "globalVar": ico_global_var,
    "tempFieldAllocation": ico_temp_field_allocation,
    "sparseTempFieldAllocation": ico_sparse_temp_field_allocation,
    "tempField": ico_temp_field,
    "reductionInIfConditional": ico_reduction_in_if_conditional,
    "reductionWithCenter": lambda: ico_reduction_with_center(0),
    "reductionWithCenterSparse": lambda: ico_reduction_with_center(1),
    "reductionAndFillWithCenterSparse": lambda: ico_reduction_with_center(2),
}


def main():
    os.makedirs(OUT, exist_ok=True)
    for name, fn in ALL.items():
        path = os.path.join(OUT, name + ".iir")
        with open(path, "w") as f:
            f.write(fn().dumps())
        print("wrote", path)


if __name__ == "__main__":
    main()
