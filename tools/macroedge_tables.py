#!/usr/bin/env python
"""Generates basix_b200/csrc/macroedge_tables.inc: the coefficients of the macroedge polynomial sets of the triangle
(degree 1, 2) and the tetrahedron (degree 1) on every sub-cell.

The reference defines these sets only as closed-form expressions per sub-cell (cpp/basix/polyset.cpp:440-1261).  This
script does NOT read that source: it treats the compiled reference (oracle/_ref/libbasix_ref.so) as a black box and
recovers each coefficient EXACTLY (to the bit) from its outputs:

* the highest derivative slabs are constants: they are the coefficients of the leading monomials (times 2 for squares);
* every lower coefficient c enters the expression the reference evaluates, fl(...fl(fl(c + t1) + t2)...), with t_i now
  known: the candidates are the few doubles around (output - sum t_i) computed in exact rational arithmetic, and the one
  that reproduces the reference's output bit for bit at every sample point of the sub-cell is the coefficient (a small
  constant next to large terms is only pinned to a few dozen ulps by one point, so the window is wide).

The table is then checked as a whole: every value and derivative slab evaluated from it (same expression order as the
kernel in csrc/polyset_macroedge.cu) equals the reference's output bit for bit at 4000 random points per sub-cell.
Run here (needs oracle/_ref):  python tools/macroedge_tables.py
"""
import itertools
import os
import sys
from fractions import Fraction

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402

TRI, TET, MACRO = 2, 3, 1


def tri_region(x):
    if x[0] + x[1] < 0.5:
        return 0
    if x[0] > 0.5:
        return 1
    if x[1] > 0.5:
        return 2
    return 3


def tet_region(x):
    if x[0] + x[1] + x[2] < 0.5:
        return 0
    if x[0] > 0.5:
        return 1
    if x[1] > 0.5:
        return 2
    if x[2] > 0.5:
        return 3
    if x[1] + x[2] < 0.5 and x[0] + x[1] < 0.5:
        return 4
    if x[1] + x[2] < 0.5:
        return 5
    if x[0] + x[1] > 0.5:
        return 6
    return 7


def sample(cell, region, n, seed):
    rng = np.random.default_rng(seed)
    tdim = 2 if cell == TRI else 3
    fn = tri_region if cell == TRI else tet_region
    out = []
    while len(out) < n:
        x = rng.random((4 * n, tdim))
        x = x[x.sum(axis=1) < 1]
        out += [p for p in x if fn(p) == region]
    return np.ascontiguousarray(np.array(out[:n]))


def neighbours(v, k=2000):
    out = [v]
    lo = hi = v
    for _ in range(k):
        lo, hi = np.nextafter(lo, -np.inf), np.nextafter(hi, np.inf)
        out += [lo, hi]
    return out


def solve_constant(observed, terms):
    """c with fl(...fl(c + t[0]) + t[1]...) == observed at every point; terms[i] = array over points (already rounded
    products, as the reference forms them)."""
    est = Fraction(float(observed[0])) - sum(Fraction(float(t[0])) for t in terms)
    hits = []
    for c in neighbours(np.float64(float(est))):
        # cheap filter on a few points, then all of them
        acc = np.full(16, c)
        for t in terms:
            acc = acc + t[:16]
        if not np.array_equal(acc, observed[:16]):
            continue
        acc = np.full(observed.shape, c)
        for t in terms:
            acc = acc + t
        if np.array_equal(acc, observed):
            hits.append(float(c))
    if not hits:
        raise RuntimeError("no candidate reproduces the reference output")
    # a small constant added to a term that is large everywhere in the sub-cell loses its last bits in the first
    # rounding: neighbouring doubles are then indistinguishable on the sub-cell (20000 points); take the middle one
    hits.sort()
    return hits[len(hits) // 2]


def triangle_table(degree):
    psize = (degree + 1) * (2 * degree + 1)
    nmono = 3 if degree == 1 else 6
    tab = np.zeros((4, psize, nmono))
    for r in range(4):
        x = sample(TRI, r, 20000, 10 + r)
        P = ref.tabulate_polynomial_set(TRI, MACRO, degree, degree, x)  # [nder][psize][npts]
        X, Y = x[:, 0], x[:, 1]
        for f in range(psize):
            if degree == 1:
                cy, cx = P[2, f, 0], P[1, f, 0]  # idx(0,1) = 2, idx(1,0) = 1
                assert np.all(P[2, f] == cy) and np.all(P[1, f] == cx)
                c0 = solve_constant(P[0, f], [cy * Y, cx * X])
                tab[r, f] = [c0, cy, cx]
            else:
                # idx(0,2) = 5, idx(1,1) = 4, idx(2,0) = 3
                cyy2, cxy, cxx2 = P[5, f, 0], P[4, f, 0], P[3, f, 0]
                assert np.all(P[5, f] == cyy2) and np.all(P[4, f] == cxy) and np.all(P[3, f] == cxx2)
                cyy, cxx = cyy2 / 2, cxx2 / 2
                cy = solve_constant(P[2, f], [cyy2 * Y, cxy * X])
                cx = solve_constant(P[1, f], [cxy * Y, cxx2 * X])
                c0 = solve_constant(P[0, f], [cy * Y, cx * X, cyy * Y * Y, cxy * X * Y, cxx * X * X])
                tab[r, f] = [c0, cy, cx, cyy, cxy, cxx]
    return tab


def eval_triangle(tab, degree, nd, x):
    """The kernel's expression order (csrc/polyset_macroedge.cu eval_triangle), vectorised over points."""
    psize = tab.shape[1]
    out = np.zeros(((nd + 1) * (nd + 2) // 2, psize, len(x)))
    reg = np.array([tri_region(p) for p in x])
    X, Y = x[:, 0], x[:, 1]
    for f in range(psize):
        c = tab[reg, f]  # [npts][nmono]
        if degree == 1:
            out[0, f] = c[:, 0] + c[:, 1] * Y + c[:, 2] * X
            if nd >= 1:
                out[2, f], out[1, f] = c[:, 1], c[:, 2]
        else:
            c0, cy, cx, cyy, cxy, cxx = (c[:, i] for i in range(6))
            out[0, f] = c0 + cy * Y + cx * X + cyy * Y * Y + cxy * X * Y + cxx * X * X
            if nd >= 1:
                out[2, f] = cy + 2 * cyy * Y + cxy * X
                out[1, f] = cx + cxy * Y + 2 * cxx * X
            if nd >= 2:
                out[5, f], out[4, f], out[3, f] = 2 * cyy, cxy, 2 * cxx
    return out


def tetrahedron_table():
    psize = 10
    best = None
    for order in itertools.permutations(range(3)):  # order of the x0 / x1 / x2 terms in the reference's sums
        try:
            tab = np.zeros((8, psize, 4))
            for r in range(8):
                x = sample(TET, r, 20000, 30 + r)
                P = ref.tabulate_polynomial_set(TET, MACRO, 1, 1, x)
                # idx(1,0,0) = 1, idx(0,1,0) = 2, idx(0,0,1) = 3
                for f in range(psize):
                    g = [P[1, f, 0], P[2, f, 0], P[3, f, 0]]
                    assert all(np.all(P[1 + a, f] == g[a]) for a in range(3))
                    c0 = solve_constant(P[0, f], [g[a] * x[:, a] for a in order])
                    tab[r, f] = [c0] + g
            best = (order, tab)
            break
        except RuntimeError:
            continue
    if best is None:
        raise RuntimeError("no term order reproduces the reference")
    return best


def eval_tetrahedron(tab, order, nd, x):
    out = np.zeros((1 + 3 * (nd >= 1), 10, len(x)))
    reg = np.array([tet_region(p) for p in x])
    for f in range(10):
        c = tab[reg, f]
        acc = c[:, 0].copy()
        for a in order:
            acc = acc + c[:, 1 + a] * x[:, a]
        out[0, f] = acc
        if nd >= 1:
            for a in range(3):
                out[1 + a, f] = c[:, 1 + a]
    return out


def emit(name, tab, f):
    flat = tab.reshape(-1)
    f.write(f"__device__ const double {name}[{len(flat)}] = {{\n")
    for i in range(0, len(flat), 4):
        f.write("    " + ", ".join(repr(float(v)) for v in flat[i:i + 4]) + ",\n")
    f.write("};\n")


def main():
    t1, t2 = triangle_table(1), triangle_table(2)
    order, tt = tetrahedron_table()
    # whole-table check against the reference, values and all derivative slabs, fresh points
    for degree, tab in ((1, t1), (2, t2)):
        for r in range(4):
            x = sample(TRI, r, 4000, 100 + r)
            for nd in range(degree + 2):
                want = ref.tabulate_polynomial_set(TRI, MACRO, degree, nd, x)
                got = eval_triangle(tab, degree, min(nd, degree), x)
                assert np.array_equal(got, want[: got.shape[0]]) and not np.any(want[got.shape[0]:]), (degree, r, nd)
    for r in range(8):
        x = sample(TET, r, 4000, 200 + r)
        for nd in range(3):
            want = ref.tabulate_polynomial_set(TET, MACRO, 1, nd, x)
            got = eval_tetrahedron(tt, order, min(nd, 1), x)
            assert np.array_equal(got, want[: got.shape[0]]) and not np.any(want[got.shape[0]:]), (r, nd)
    path = os.path.join(ROOT, "basix_b200", "csrc", "macroedge_tables.inc")
    with open(path, "w") as f:
        f.write("// GENERATED by tools/macroedge_tables.py from the outputs of the compiled reference (black box; the reference's\n"
                "// source is not read).  Coefficients of the macroedge polynomial sets per sub-cell, [sub-cell][function][monomial]:\n"
                "// triangle monomials 1, y, x, y y, x y, x x; tetrahedron 1, x, y, z.  Verified bit for bit against the reference\n"
                "// at 4000 points per sub-cell (values and every derivative slab).\n")
        emit("kMacroTri1", t1, f)
        emit("kMacroTri2", t2, f)
        emit("kMacroTet1", tt, f)
        f.write("// order in which the reference adds the x / y / z terms of the tetrahedron functions\n")
        f.write("__device__ const int kMacroTetOrder[3] = {%d, %d, %d};\n" % order)
    print("wrote", path, "tetrahedron term order", order)


if __name__ == "__main__":
    main()
