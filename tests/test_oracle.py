"""Pins the CPU restatement (oracle/restate.py) to the reference: against the committed golden vectors
(outputs of the unmodified reference, tests/golden/make_golden.py), the known-answer anchors of SURVEY.md 8c,
and -- when oracle/_ref/libbasix_ref.so is present -- the compiled reference itself on fresh inputs.

No GPU needed.  Bars: bit-exact for permutations, the simplex polysets and the Piola maps (same operation
order); 1e-13 * max|ref| for the contraction (BLAS summation order) and the matrix DOF transformations.
"""

import os

import numpy as np
import pytest

from conftest import ROOT, random_points, slab_close
from oracle import restate as R

GOLDEN = os.path.join(ROOT, "tests", "golden")
VEC = np.load(os.path.join(GOLDEN, "vectors.npz"))
ELEMENT_NAMES = sorted(f[:-4] for f in os.listdir(os.path.join(GOLDEN, "elements")) if f.endswith(".npz"))
T_OPS = ["T_apply", "Tt_apply", "Tt_inv_apply", "Tinv_apply", "Tt_apply_right", "Tinv_apply_right", "T_apply_right",
         "Tt_inv_apply_right"]


def element(name):
    return R.RestatedElement(np.load(os.path.join(GOLDEN, "elements", name + ".npz")))


# ---- golden vectors ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("key", sorted(k[:-2] for k in VEC.files if k.startswith("polyset/") and k.endswith("/x")))
def test_polyset_golden(key):
    cell, deg, nd = (int(v) for v in key.split("/")[1].split("_"))
    got = R.tabulate_polyset(cell, deg, nd, VEC[key + "/x"])
    assert np.array_equal(got, VEC[key + "/P"])


@pytest.mark.parametrize("key", sorted(k[:-2] for k in VEC.files if k.startswith("tabulate/") and k.endswith("/x")))
def test_tabulate_golden(key):
    _, name, nd = key.split("/")
    got = element(name).tabulate(int(nd), VEC[key + "/x"])
    slab_close(got, VEC[key + "/basis"], 1e-13)


@pytest.mark.parametrize("name", sorted({k.split("/")[1] for k in VEC.files if k.startswith("map/")}))
def test_maps_golden(name):
    m = {k: VEC[f"map/{name}/{k}"] for k in ("U", "J", "detJ", "K", "push", "pull")}
    e = element(name)
    u = e.push_forward(m["U"], m["J"], m["detJ"], m["K"])
    assert np.array_equal(u, m["push"])
    assert np.array_equal(e.pull_back(m["push"], m["J"], m["detJ"], m["K"]), m["pull"])


@pytest.mark.parametrize("name", sorted({k.split("/")[1] for k in VEC.files
                                         if k.startswith("transform/") and k.count("/") == 3}))
@pytest.mark.parametrize("bs", [1, 2])
def test_transform_golden(name, bs):
    e = element(name)
    ci = VEC["transform/cell_info"]
    data = VEC[f"transform/{name}/{bs}/in"]
    for op in T_OPS:
        want = VEC[f"transform/{name}/{bs}/{op}"]
        got = e.T_apply_batch(op, data.copy(), bs, ci)
        if e.perms:
            assert np.array_equal(got, want), op
        else:
            assert np.max(np.abs(got - want)) <= 1e-13 * np.max(np.abs(want)), op


@pytest.mark.parametrize("name", sorted({k.split("/")[1] for k in VEC.files if k.startswith("permute/")}))
def test_permute_golden(name):
    e = element(name)
    ci = VEC["transform/cell_info"]
    for key, inverse in (("fwd", False), ("inv", True)):
        want = VEC[f"permute/{name}/{key}"]
        got = np.stack([e.permute(np.arange(e.dim, dtype=np.int32), c, inverse) for c in ci])
        assert np.array_equal(got, want)


# ---- SURVEY 8c anchors (reference outputs quoted in the survey) -------------------------------------------
def test_anchors():
    got = R.tabulate_polyset(R.TRIANGLE, 2, 1, np.array([[0.3, 0.2]]))[:, :, 0]
    want = np.array([[1.4142135623731, -0.8, -0.692820323027551, -0.489897948556636, 0, -1.42407864951343],
                     [0, 0, 6.92820323027551, 0, 0, -6.57267069006199],
                     [0, 6, 3.46410161513775, -9.79795897113271, -4.24264068711928, 1.09544511501033]])
    assert np.allclose(got, want, rtol=0, atol=5e-14)
    got = R.tabulate_polyset(R.HEXAHEDRON, 1, 0, np.array([[.25, .5, .75]]))[0, :, 0]
    assert np.allclose(got, [1, 0.866025403784439, 0, 0, -0.866025403784439, -0.75, 0, 0], rtol=0, atol=5e-15)
    got = element("P1_triangle_gll").tabulate(1, np.array([[0.3, 0.2]]))[:, 0, :, 0]
    assert np.allclose(got, [[.5, .3, .2], [-1, 1, 0], [-1, 0, 1]], rtol=0, atol=1e-14)
    J = np.array([[[2, .5, .25], [.125, 1.5, .75], [.0625, .3125, 1]]])
    detJ, K, U = np.linalg.det(J), np.linalg.inv(J), np.array([[[1.0, 2.0, 3.0]]])
    assert abs(detJ[0] - 2.478515625) < 1e-15
    assert np.allclose(R.push_forward(R.MAP_COV, U, J, detJ, K)[0, 0],
                       [0.38140267927502, 0.7123719464145, 2.37037037037037], rtol=0, atol=1e-13)
    assert np.allclose(R.pull_back(R.MAP_COV, 3, U, J, detJ, K)[0, 0], [2.4375, 4.4375, 4.75], rtol=0, atol=1e-13)
    assert np.allclose(R.push_forward(R.MAP_CONTRA, U, J, detJ, K)[0, 0],
                       [1.51300236406619, 2.16863672182821, 1.48778565799842], rtol=0, atol=1e-13)
    e = element("P5_tetrahedron_gll")
    d = e.permute(np.arange(56, dtype=np.int32), 0x01000)
    assert list(d[4:8]) == [7, 6, 5, 4]
    d = e.permute(np.arange(56, dtype=np.int32), 0x00001)
    assert list(d[28:34]) == [28, 31, 33, 29, 32, 30]
    assert np.array_equal(e.permute(np.arange(56, dtype=np.int32), 0x00006), np.arange(56))
    want = [0, 1, 2, 3, 4, 5, 6, 7, 11, 10, 9, 8, 12, 13, 14, 15, 19, 18, 17, 16, 20, 21, 22, 23, 27, 26, 25, 24, 33,
            32, 30, 31, 29, 28, 34, 37, 39, 35, 38, 36, 40, 43, 45, 41, 44, 42, 51, 50, 48, 49, 47, 46, 52, 53, 54, 55]
    assert list(e.permute(np.arange(56, dtype=np.int32), 0x2ABCD)) == want
    e = element("RT2_tetrahedron_unset")
    for ci, head in ((0x001, [-1, -3, -2, 4]), (0x002, [3, 1, 2, 4]), (0x004, [2, 3, 1, 4])):
        u = np.arange(1.0, 16.0)
        e.T_op("T_apply", u, 1, ci)
        assert np.allclose(u[:4], head, rtol=0, atol=1e-13) and np.allclose(u[4:], np.arange(5.0, 16.0))
    u = np.arange(1.0, 16.0)
    e.T_op("T_apply", u, 1, 0x5A7)
    assert np.allclose(u, [-1, -3, -2, 5, 6, 4, 7, 8, 9, 12, 10, 11, 13, 14, 15], rtol=0, atol=1e-13)


# ---- live against the compiled reference (skipped when oracle/_ref was not built) ---------------------------
@pytest.mark.parametrize("cell,degree,nd", [(1, 6, 2), (2, 5, 2), (3, 4, 2), (3, 6, 1), (4, 4, 2), (5, 3, 2), (6, 4, 2),
                                            (6, 1, 3), (7, 4, 2), (7, 2, 3)])
def test_polyset_vs_compiled_reference(ref, cell, degree, nd):
    x = random_points(cell, 211, seed=degree)
    assert np.array_equal(R.tabulate_polyset(cell, degree, nd, x), ref.tabulate_polynomial_set(cell, 0, degree, nd, x))


@pytest.mark.parametrize("name,args", [("P5_tetrahedron_gll", (1, 3, 5, 2)), ("Q4_hexahedron_gll", (1, 5, 4, 2)),
                                       ("N1curl3_tetrahedron_legendre", (3, 3, 3, 11)),
                                       ("RT4_tetrahedron_legendre", (2, 3, 4, 11))])
def test_element_vs_compiled_reference(ref, name, args):
    e, r = element(name), ref.RefElement(*args)
    rng = np.random.default_rng(5)
    x = random_points(r.cell_type, 53, seed=9)
    slab_close(e.tabulate(2, x), r.tabulate(2, x), 1e-13)
    nc = 17
    J = np.eye(3)[None] + 0.3 * rng.standard_normal((nc, 3, 3))
    detJ, K = np.linalg.det(J), np.linalg.inv(J)
    U = rng.standard_normal((nc, 3, r.value_size))
    u = r.push_forward(U, J, detJ, K)
    assert np.array_equal(e.push_forward(U, J, detJ, K), u)
    assert np.array_equal(e.pull_back(u, J, detJ, K), r.pull_back(u, J, detJ, K))
    ci = rng.integers(0, 2 ** 30, nc).astype(np.uint32)
    for bs in (1, 3):
        data = rng.standard_normal((nc, r.dim * bs))
        for op in T_OPS:
            want = r.T_apply_batch(op, data.copy(), bs, ci)
            got = e.T_apply_batch(op, data.copy(), bs, ci)
            if e.perms:
                assert np.array_equal(got, want), op
            else:
                assert np.max(np.abs(got - want)) <= 1e-13 * np.max(np.abs(want)), op
