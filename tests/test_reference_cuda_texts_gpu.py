"""The reference's own pre-generated CUDA texts, recompiled UNMODIFIED for sm_100a against a stand-in for the
GridTools device-storage names they mention (oracle/ref_build/shim_cuda -> oracle/_ref/libdawn_refcuda*.so),
compared with ours, the oracle and the goldens of the reference's generated NAIVE code.  This is test
infrastructure around the reference (a second baseline), not the product path: a failure here says
something about the stand-in or the old backend's text, never about a b200 kernel - which is why these
cases live in their own file."""
import numpy as np
import pytest

from util import bit_equal

pytestmark = [pytest.mark.gpu]


def _diff(got, want):
    """where two arrays differ, for the assertion message"""
    bad = np.argwhere(got.view(np.int64) != want.view(np.int64))
    if not len(bad):
        return "equal"
    return "%d of %d differ, first %s (got %r want %r), index box %s..%s" % (
        len(bad), got.size, tuple(bad[0]), got[tuple(bad[0])], want[tuple(bad[0])],
        tuple(bad.min(0)), tuple(bad.max(0)))


@pytest.mark.parametrize("I,J,K", [(12, 12, 10), (70, 41, 9), (128, 64, 20)])
def test_reference_cuda_kernel_and_ours_agree_bit_for_bit(I, J, K):
    """hori_diff (dawn4py flavour: in, out, coeff): the reference's own pre-generated CUDA code, recompiled
    unmodified for sm_100a (oracle/_ref/libdawn_refcuda.so), against the b200 kernel and the oracle."""
    import ctypes
    import torch
    import dawn_b200
    from oracle import capi
    from util import HALO, run_oracle
    lib = capi.refcuda()
    if lib is None:
        pytest.skip("oracle/_ref/libdawn_refcuda.so not built")
    ni, nj = I + 2 * HALO, J + 2 * HALO
    rng = np.random.default_rng(56)
    f = {n: rng.uniform(0.5, 1.5, (K + 1, nj, ni)) for n in ("in", "out", "coeff")}
    f["out"] = np.full((K + 1, nj, ni), -1.0)
    want = run_oracle("hori_diff_stencil", (ni, nj, K), f)["out"]
    mod = dawn_b200.load_stencil("hori_diff_stencil")
    assert [fd["name"] for fd in mod.fields] == ["in", "out", "coeff"]
    dom = dawn_b200.Domain(ni, nj, K)
    dom.set_halos(HALO, HALO, HALO, HALO, 0, 0)
    ours = [dawn_b200.Storage(ni, nj, K + 1, "ijk").from_numpy(f[n]) for n in ("in", "out", "coeff")]
    mod(dom).run(*ours)
    theirs = [dawn_b200.Storage(ni, nj, K + 1, "ijk").from_numpy(f[n]) for n in ("in", "out", "coeff")]
    sec = ctypes.c_double()
    rc = lib.refcuda_hori_diff(ni, nj, K, HALO, theirs[0].stride_j, theirs[0].stride_k,
                               *[ctypes.c_void_p(s.buf.data_ptr() + 8 * s.lead) for s in theirs], 1,
                               ctypes.byref(sec))
    torch.cuda.synchronize()
    assert rc == 0
    a, b = ours[1].to_numpy(), theirs[1].to_numpy()
    assert bit_equal(a, want) and bit_equal(b, want)


# global_indexing.cu is left out: the old backend's text tests `globalOffsets_[1] + jblock` - the thread's row
# INSIDE its block - against the global index range (global_indexing.cu:110-111), so on any domain higher
# than one 4-row block it differs from the reference's own naive code (first run on a B200, round 2: 1200 of
# 7560 values differ; the reference only ever string-compares this file).  Ours matches the naive golden
# (tests/test_cartesian_gpu.py).  update_dz_c.cu likewise differs from the naive golden at 213 of 7560
# points of gz behind the stand-in (block-private temporaries; not investigated further - the b200 kernels
# match that golden bit for bit, tests/test_reference_tree_gpu.py).
@pytest.mark.parametrize("case", ["conditional_var1_1", "conditional_var1_0", "nonoverlapping", "copy"])
def test_reference_cuda_texts_reproduce_the_reference_naive_goldens(case):
    """The reference's pre-generated CUDA texts (CodeGen/reference/*.cu, copy_stencil_ref.cpp), recompiled
    unmodified for sm_100a, against the goldens its own generated NAIVE code produced (tests/golden):
    reference naive == reference CUDA on a B200 == ours (tests/test_cartesian_gpu.py)."""
    import ctypes
    import os
    import torch
    import dawn_b200
    from oracle import capi
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    name = {"conditional_var1_1": "conditional_stencil", "conditional_var1_0": "conditional_stencil",
            "nonoverlapping": "nonoverlapping_stencil", "copy": "copy_stencil"}.get(case, case)
    lib = capi.refcuda_small(name)
    if lib is None:
        pytest.skip("oracle/_ref/libdawn_refcuda_%s.so not built" % name)
    g = np.load(os.path.join(gold, "conditional_stencil.npz" if name == "conditional_stencil" else "indexing.npz"))
    ni, nj, nk = (int(x) for x in g["dims"])
    inn = np.random.default_rng(int(g["seed"])).uniform(-1, 1, (nk + 1, nj, ni))
    s_in = dawn_b200.Storage(ni, nj, nk + 1).from_numpy(inn)
    s_out = dawn_b200.Storage(ni, nj, nk + 1).fill(-1.0)
    var1 = 0 if case.endswith("_0") else 1
    rc = lib.refcuda_run(ni, nj, nk, 3, s_in.stride_j, s_in.stride_k,
                         ctypes.c_void_p(s_in.buf.data_ptr() + 8 * s_in.lead),
                         ctypes.c_void_p(s_out.buf.data_ptr() + 8 * s_out.lead), var1)
    torch.cuda.synchronize()
    assert rc == 0
    out = s_out.to_numpy()
    if name == "conditional_stencil":
        assert bit_equal(out, g["out_var1_%d" % var1]), _diff(out, g["out_var1_%d" % var1])
    elif name == "nonoverlapping_stencil":
        # below level 11: an uninitialised local in the text
        assert bit_equal(out[11:], g["nonoverlapping_from_k11"]), _diff(out[11:], g["nonoverlapping_from_k11"])
    else:
        h = 3   # dawn4py-tests/copy_stencil.py:47 builds `out = in[i+1]`
        want = np.full_like(inn, -1.0)
        want[:nk, h:-h, h:-h] = inn[:nk, h:-h, h + 1:ni - h + 1]
        assert bit_equal(out, want), _diff(out, want)


@pytest.mark.parametrize("case", ["equal", "one_off", "noise", "zero_reference", "both_zero", "large", "empty_tol",
                                  "nan_in_both", "inf_equal"])
def test_verify_kernel_against_the_reference_verifier(case):
    """b200rt_verify_field (one pass) against the reference's verify_field (cuda_verify.hpp:84-196: compare
    kernels + thrust reductions), compiled unmodified into oracle/_ref/libdawn_refcuda_verify.so: the four
    error extrema and is_valid on mixed inputs, bit for bit.  NaN errors: thrust's maximum / minimum leave
    the extrema open (order dependent), so for those only is_valid is compared - a NaN always fails."""
    import ctypes
    import os
    import torch
    from dawn_b200 import runtime as rt
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref",
                     "libdawn_refcuda_verify.so")
    if not os.path.exists(p):
        pytest.skip("oracle/_ref/libdawn_refcuda_verify.so not built")
    lib = ctypes.CDLL(p)
    lib.refcuda_verify_field.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_double,
                                         ctypes.c_double, ctypes.POINTER(ctypes.c_double),
                                         ctypes.POINTER(ctypes.c_int)]
    rng = np.random.default_rng(71)
    n = 100003
    ref = rng.standard_normal(n)
    rel_tol, abs_tol = 1e-9, 1e-12
    dsl = ref.copy()
    if case == "one_off":
        dsl[777] += 1e-3
    elif case == "noise":
        dsl = ref * (1.0 + 1e-11 * rng.standard_normal(n))
    elif case == "zero_reference":          # relative error inf where the reference is 0 and the value is not
        ref[5] = 0.0
        dsl = ref + 1e-14
    elif case == "both_zero":
        ref[:100] = 0.0
        dsl = ref.copy()
        dsl[50:] += 1e-13
    elif case == "large":
        ref *= 1e200
        dsl = ref * (1 + 1e-10)
    elif case == "empty_tol":
        rel_tol = abs_tol = 0.0
        dsl[3] = np.nextafter(dsl[3], 10.0)
    elif case == "nan_in_both":
        ref[11] = dsl[11] = np.nan
    elif case == "inf_equal":
        ref[11] = dsl[11] = np.inf
    a, b = torch.from_numpy(dsl).cuda(), torch.from_numpy(ref).cuda()
    out4, valid = (ctypes.c_double * 4)(), ctypes.c_int(-1)
    assert lib.refcuda_verify_field(n, a.data_ptr(), b.data_ptr(), rel_tol, abs_tol, out4, ctypes.byref(valid)) == 0
    m = rt.verify_field(a.data_ptr(), b.data_ptr(), n, rel_tol=rel_tol, abs_tol=abs_tol)
    assert m.is_valid == valid.value, (m.is_valid, valid.value, m.n_failed)
    if case == "nan_in_both":
        assert m.is_valid == 0 and m.n_failed == 1
        return
    got = np.array([m.max_rel_err, m.min_rel_err, m.max_abs_err, m.min_abs_err])
    want = np.array(list(out4))
    assert bit_equal(got, want), (got, want)
