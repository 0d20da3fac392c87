"""Generate the committed golden fixtures from the UNMODIFIED reference (oracle/_ref).

Run in the build container (where /root/reference exists and oracle/build_ref.sh has been run):

    python tests/golden/make_golden.py

Writes
  tests/golden/elements/<name>.npz   defining arrays of the BASELINE.json config elements
                                     (coefficient matrix, entity dofs, entity transformations ...)
  tests/golden/vectors.npz           small input/output vectors of the reference's hot path:
                                     polyset::tabulate, FiniteElement::tabulate, push_forward,
                                     pull_back, T_apply (all 8 flavours), permute / permute_inv.
Everything is seeded; re-running reproduces the files bit for bit on the same toolchain.
"""

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from conftest import random_points  # noqa: E402
from oracle import ref  # noqa: E402

P, RT, N1E = 1, 2, 3
INTERVAL, TRI, TET, QUAD, HEX = 1, 2, 3, 4, 5
GLL, EQ, LEG = 2, 1, 11

# name -> (family, cell, degree, lagrange variant)
CONFIG_ELEMENTS = {
    "P1_triangle_gll": (P, TRI, 1, GLL),
    "P5_tetrahedron_gll": (P, TET, 5, GLL),
    "Q4_hexahedron_gll": (P, HEX, 4, GLL),
    "N1curl3_tetrahedron_legendre": (N1E, TET, 3, LEG),
    "RT4_tetrahedron_legendre": (RT, TET, 4, LEG),
    "P3_triangle_gll": (P, TRI, 3, GLL),
    "P2_tetrahedron_eq": (P, TET, 2, EQ),
    "Q2_quadrilateral_gll": (P, QUAD, 2, GLL),
    "RT2_tetrahedron_legendre": (RT, TET, 2, LEG),
    "RT2_tetrahedron_unset": (RT, TET, 2, 0),  # the element of the SURVEY 8c T_apply anchors
}


def element_arrays(r, family, lvariant):
    flat, offs = [], [0]
    for row in r.entity_dofs:
        for e in row:
            flat.extend(e)
            offs.append(len(flat))
    out = {
        "meta": np.asarray([r.cell_type, r.polyset_type, r.degree, r.embedded_superdegree, r.map_type, family,
                            lvariant, 0, 0], dtype=np.int64),
        "value_shape": np.asarray(r.value_shape, dtype=np.int64),
        "coefficient_matrix": r.coefficient_matrix,
        "dof_ordering": r.dof_ordering,
        "entity_dof_offsets": np.asarray(offs, dtype=np.int32),
        "entity_dofs_flat": np.asarray(flat, dtype=np.int32),
        "points": r.points,
        "interpolation_matrix": r.interpolation_matrix,
    }
    for ct, m in r.entity_transformations.items():
        out[f"etrans_{ct}"] = m
    return out


def jacobians(nc, gdim, tdim, seed=7):
    rng = np.random.default_rng(seed)
    J = np.zeros((nc, gdim, tdim))
    J[:, :tdim, :tdim] = np.eye(tdim)
    J += 0.3 * rng.standard_normal((nc, gdim, tdim))
    detJ = np.linalg.det(J)
    K = np.linalg.inv(J)
    return np.ascontiguousarray(J), np.ascontiguousarray(detJ), np.ascontiguousarray(K)


def main():
    os.makedirs(os.path.join(HERE, "elements"), exist_ok=True)
    refs = {}
    for name, (fam, cell, deg, lv) in CONFIG_ELEMENTS.items():
        r = ref.RefElement(fam, cell, deg, lv)
        refs[name] = r
        np.savez_compressed(os.path.join(HERE, "elements", name + ".npz"), **element_arrays(r, fam, lv))

    vec = {}
    # polyset::tabulate
    for cell, deg, nd in [(INTERVAL, 5, 2), (TRI, 4, 2), (TET, 5, 2), (QUAD, 3, 1), (HEX, 4, 1), (TET, 0, 1),
                          (6, 3, 2), (7, 3, 2), (7, 0, 1)]:  # 6 = prism, 7 = pyramid
        x = random_points(cell, 37, seed=100 + cell)
        if cell == 7:
            x[0] = [0.0, 0.0, 1.0]  # the apex takes special values (polyset.cpp:2070-2072)
        vec[f"polyset/{cell}_{deg}_{nd}/x"] = x
        vec[f"polyset/{cell}_{deg}_{nd}/P"] = ref.tabulate_polynomial_set(cell, 0, deg, nd, x)
    # FiniteElement::tabulate on the config elements
    for name, nd in [("P1_triangle_gll", 1), ("P5_tetrahedron_gll", 2), ("Q4_hexahedron_gll", 1),
                     ("N1curl3_tetrahedron_legendre", 1), ("RT4_tetrahedron_legendre", 0), ("P3_triangle_gll", 2),
                     ("Q2_quadrilateral_gll", 2)]:
        r = refs[name]
        x = random_points(r.cell_type, 29, seed=200 + r.dim)
        vec[f"tabulate/{name}/{nd}/x"] = x
        vec[f"tabulate/{name}/{nd}/basis"] = r.tabulate(nd, x)
    # push_forward / pull_back
    for name in ("N1curl3_tetrahedron_legendre", "RT4_tetrahedron_legendre", "P2_tetrahedron_eq"):
        r = refs[name]
        nc, rows = 23, 4
        J, detJ, K = jacobians(nc, 3, 3)
        U = np.random.default_rng(300).standard_normal((nc, rows, r.value_size))
        u = r.push_forward(U, J, detJ, K)
        for k, v in (("U", U), ("J", J), ("detJ", detJ), ("K", K), ("push", u), ("pull", r.pull_back(u, J, detJ, K))):
            vec[f"map/{name}/{k}"] = v
    # DOF transformations
    rng = np.random.default_rng(400)
    ci = rng.integers(0, 2 ** 30, 19).astype(np.uint32)
    ci[:3] = [0, 0x2ABCD, 0x3FFFFFFF]
    vec["transform/cell_info"] = ci
    for name in ("P5_tetrahedron_gll", "RT4_tetrahedron_legendre", "N1curl3_tetrahedron_legendre", "P3_triangle_gll",
                 "Q4_hexahedron_gll", "RT2_tetrahedron_legendre"):
        r = refs[name]
        for bs in (1, 2):
            data = rng.standard_normal((len(ci), r.dim * bs))
            vec[f"transform/{name}/{bs}/in"] = data
            for op in ref.T_OPS:
                vec[f"transform/{name}/{bs}/{op}"] = r.T_apply_batch(op, data.copy(), bs, ci)
        if r.dof_transformations_are_permutations:
            d = np.tile(np.arange(r.dim, dtype=np.int32), (len(ci), 1))
            vec[f"permute/{name}/fwd"] = r.permute_batch(d.copy(), ci)
            vec[f"permute/{name}/inv"] = r.permute_batch(d.copy(), ci, inverse=True)
    np.savez_compressed(os.path.join(HERE, "vectors.npz"), **vec)
    total = sum(os.path.getsize(os.path.join(dp, f)) for dp, _, fs in os.walk(HERE) for f in fs)
    print(f"golden fixtures written: {len(CONFIG_ELEMENTS)} elements, {len(vec)} arrays, {total / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
