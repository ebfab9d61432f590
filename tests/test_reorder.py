"""Mesh renumbering (dawn_b200/reorder.py) is index work and must be exact: permutations are
bijections, renumbered neighbour tables are the relabelled originals with the neighbour order kept,
chain composition commutes with relabelling, and every oracle result maps back bit for bit."""
import numpy as np
import pytest

from util import bit_equal, iir_path
from oracle import naive_interp
from dawn_b200 import reorder as ro
from dawn_b200.trimesh import CELL, EDGE, VERTEX, TriMesh

CHAINS = [[CELL, EDGE], [CELL, VERTEX], [EDGE, CELL], [EDGE, VERTEX], [VERTEX, CELL], [VERTEX, EDGE],
          [EDGE, CELL, VERTEX], [CELL, EDGE, CELL], [EDGE, CELL, EDGE], [VERTEX, EDGE, CELL],
          [CELL, EDGE, CELL, EDGE, CELL], [EDGE, VERTEX, EDGE], [CELL, VERTEX, CELL],
          [VERTEX, EDGE, VERTEX], [EDGE, CELL, VERTEX, CELL]]


def test_renumbering_must_be_a_bijection():
    m = TriMesh(3, 2)
    r = ro.Renumbering.random(m, 1)
    for loc in ro.LOCS:
        assert np.array_equal(r.old_to_new[loc][r.new_to_old[loc]], np.arange(m.num(loc)))
    bad = {loc: np.arange(m.num(loc)) for loc in ro.LOCS}
    bad[CELL] = bad[CELL].copy()
    bad[CELL][0] = 1
    with pytest.raises(ValueError, match="bijection"):
        ro.Renumbering(bad)
    a = np.arange(m.num_edges, dtype=float)
    assert np.array_equal(r.to_old(EDGE, r.to_new(EDGE, a)), a)


@pytest.mark.parametrize("center", [False, True])
def test_table_mesh_composes_chains_like_the_structured_adapter(center):
    m = TriMesh(5, 4)
    t = ro.TableMesh.from_mesh(m)
    for chain in CHAINS:
        if center and chain[0] != chain[-1]:
            continue
        assert np.array_equal(t.nbh_table(chain, center), m.nbh_table(chain, center)), chain
    assert np.array_equal(t.nbh_table([CELL, EDGE], nbh_size=5)[:, :3], m.nbh_table([CELL, EDGE]))
    with pytest.raises(ValueError, match="wider"):
        t.nbh_table([VERTEX, EDGE], nbh_size=2)


@pytest.mark.parametrize("seed", [3, 4])
def test_renumbered_tables_are_the_relabelled_originals(seed):
    m = TriMesh(6, 5)
    r = ro.Renumbering.random(m, seed)
    rm = ro.ReorderedMesh(m, r)
    for chain in CHAINS:
        old, new = m.nbh_table(chain), rm.nbh_table(chain)
        assert new.dtype == np.int32 and new.shape == old.shape
        for n in (0, 1, new.shape[0] // 2, new.shape[0] - 1):
            src = old[r.new_to_old[chain[0]][n]]
            want = [r.old_to_new[chain[-1]][x] if x >= 0 else -1 for x in src]
            assert list(new[n]) == want, (chain, n)          # same neighbours, same order, new names
    # composing the chains from the RENUMBERED two-element tables gives the same thing: relabelling
    # commutes with the reference's chain walk (first occurrences kept, origin dropped)
    composed = ro.TableMesh.from_mesh(rm)
    for chain in CHAINS:
        assert np.array_equal(composed.nbh_table(chain), rm.nbh_table(chain)), chain
    assert np.array_equal(composed.nbh_table([CELL, EDGE, CELL], True), rm.nbh_table([CELL, EDGE, CELL], True))
    assert np.array_equal(rm.valid(EDGE), m.edge_valid[r.new_to_old[EDGE]])


def _orderings(m):
    scr = ro.ReorderedMesh(m, ro.Renumbering.random(m, 11))
    return {"random": scr, "rcm": ro.reorder(m, "rcm"), "morton": ro.reorder(m, "morton"),
            "rcm of scrambled": ro.reorder(scr, "rcm")}


def _total(rm):
    """renumbering from the TriMesh at the bottom of a stack of ReorderedMesh objects"""
    r, m = rm.renumbering, rm.base
    while isinstance(m, ro.ReorderedMesh):
        r, m = m.renumbering.then(r), m.base
    return r


@pytest.mark.parametrize("which", ["random", "rcm", "morton", "rcm of scrambled"])
def test_icon_nabla2_maps_back_bit_for_bit(which):
    m, k = TriMesh(9, 7), 4
    rng = np.random.default_rng(5)
    f = {n: rng.uniform(0.5, 1.5, (k, m.num(loc))) for n, loc in (
        ("vec", EDGE), ("div_vec", CELL), ("rot_vec", VERTEX), ("nabla2_vec", EDGE),
        ("primal_edge_length", EDGE), ("dual_edge_length", EDGE), ("tangent_orientation", EDGE),
        ("__tmp_nab_60", EDGE), ("__tmp_nab_61", EDGE))}
    f["geofac_rot"] = rng.uniform(-1, 1, (k, m.num_vertices, 6))
    f["geofac_div"] = rng.uniform(-1, 1, (k, m.num_cells, 3))
    loc_of = {"vec": EDGE, "div_vec": CELL, "rot_vec": VERTEX, "nabla2_vec": EDGE, "primal_edge_length": EDGE,
              "dual_edge_length": EDGE, "tangent_orientation": EDGE, "__tmp_nab_60": EDGE,
              "__tmp_nab_61": EDGE, "geofac_rot": VERTEX, "geofac_div": CELL}
    want = naive_interp.run_unstructured(iir_path("ICON_laplacian_stencil"), m, k, {n: v.copy() for n, v in f.items()})
    rm = _orderings(m)[which]
    r = _total(rm)
    g = {n: r.to_new(loc_of[n], v, axis=1).copy() for n, v in f.items()}   # dense [k, n], sparse [k, n, nbh]
    got = naive_interp.run_unstructured(iir_path("ICON_laplacian_stencil"), rm, k, g)
    for n, loc in (("rot_vec", VERTEX), ("div_vec", CELL), ("nabla2_vec", EDGE)):
        back = r.to_old(loc, got[n], axis=1)
        v = m.valid(loc)
        assert bit_equal(back[:, v], want[n][:, v]), n


@pytest.mark.parametrize("name,spec,out", [
    ("intp", {"in": CELL, "out": CELL}, "out"),
    ("diamond", {"edge_field": EDGE, "vertex_field": VERTEX}, "edge_field"),
    ("reductionWithCenter", {"cin_field": CELL, "cout_field": CELL}, "cout_field")])
def test_long_chains_map_back_bit_for_bit(name, spec, out):
    m, k = TriMesh(8, 6), 3
    rng = np.random.default_rng(6)
    f = {n: rng.uniform(0.5, 1.5, (k, m.num(loc))) for n, loc in spec.items()}
    want = naive_interp.run_unstructured(iir_path(name), m, k, {n: v.copy() for n, v in f.items()})
    rm = ro.reorder(ro.ReorderedMesh(m, ro.Renumbering.random(m, 12)), "rcm")
    r = _total(rm)
    got = naive_interp.run_unstructured(iir_path(name), rm, k, {n: r.to_new(spec[n], v).copy() for n, v in f.items()})
    v = m.valid(spec[out])
    assert bit_equal(r.to_old(spec[out], got[out])[:, v], want[out][:, v])


def test_follow_cells_keeps_invalid_edge_slots_last_and_is_deterministic():
    m = TriMesh(7, 5)
    a, b = ro.reorder(m, "morton"), ro.reorder(m, "morton")
    for loc in ro.LOCS:
        assert np.array_equal(a.renumbering.new_to_old[loc], b.renumbering.new_to_old[loc])
    valid = a.valid(EDGE)
    n_valid = int(m.edge_valid.sum())
    assert valid[:n_valid].all() and not valid[n_valid:].any()


def test_reordering_restores_gather_locality():
    # what a warp pays for a gather: distinct 32-byte sectors per neighbour slot (8 = perfectly
    # consecutive doubles, 32 = scattered).  A scrambled numbering is near the worst case; both
    # orderings undo that, and reverse Cuthill-McKee with edges and vertices following the cells even
    # beats the structured numbering where that one strides (3 edge slots per vertex: 24 sectors)
    m = TriMesh(48, 40)
    scr = ro.ReorderedMesh(m, ro.Renumbering.random(m, 13))
    for chain in ([CELL, EDGE], [EDGE, CELL], [VERTEX, EDGE], [EDGE, VERTEX], [CELL, EDGE, CELL]):
        natural = ro.sectors_per_warp(m, chain)
        worst = ro.sectors_per_warp(scr, chain)
        assert worst > 29 and natural < 25, (chain, natural, worst)
        rcm = ro.sectors_per_warp(ro.reorder(scr, "rcm"), chain)
        morton = ro.sectors_per_warp(ro.reorder(scr, "morton"), chain)
        assert rcm < 0.6 * worst and morton < 0.8 * worst, (chain, natural, rcm, morton, worst)
        if chain != [EDGE, VERTEX]:
            assert rcm < natural, (chain, natural, rcm)
