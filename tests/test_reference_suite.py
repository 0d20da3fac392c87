"""The reference's OWN pytest files, unmodified, against the ``basix`` shim package (basix_b200/compat/basix).

``oracle/stage_reference_tests.sh`` stages the files from /root/reference/test into oracle/_ref/reftests/ (git-ignored,
travels to the GPU box with the snapshot; nothing of the reference is committed).  Tests whose elements / lattices /
set-up data are outside what the native factory builds raise ``basix.NotBuiltNatively`` and are counted as skipped by
the staged conftest; everything else must pass.  The counts of the last GPU run are listed in INTEGRATION.md."""

import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGED = os.path.join(ROOT, "oracle", "_ref", "reftests")


def test_compat_package_imports_and_host_helpers():
    """No GPU: the shim resolves as ``basix``, keeps the reference's names and answers the host-only queries."""
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "import basix_b200.compat as c; c.install()\n"
        "import basix, basix._basixcpp as cpp, numpy as np\n"
        "assert basix.CellType.tetrahedron == 3 and basix.ElementFamily.N1E == 3 and basix.LagrangeVariant.legendre == 11\n"
        "assert basix.topology(basix.CellType.triangle)[1] == [[1, 2], [0, 2], [0, 1]]\n"
        "assert basix.geometry(basix.CellType.hexahedron).shape == (8, 3)\n"
        "assert basix.index(1, 2) == 8 and cpp.index(1, 1, 1) == basix.index(1, 1, 1)\n"
        "assert basix.create_lattice(basix.CellType.tetrahedron, 4, basix.LatticeType.equispaced, True).shape == (35, 3)\n"
        "e = basix.create_element(basix.ElementFamily.P, basix.CellType.tetrahedron, 4, basix.LagrangeVariant.equispaced)\n"
        "assert e.dim == 35 and len(e.base_transformations()) == 14 and e.dof_transformations_are_permutations\n"
        "assert hasattr(cpp, 'FiniteElement_float64') and hasattr(cpp, 'tabulate_polynomial_set_float64')\n"
        "t = basix.create_tp_element(basix.ElementFamily.P, basix.CellType.quadrilateral, 2, basix.LagrangeVariant.gll_warped)\n"
        "assert t.has_tensor_product_factorisation and t.dof_ordering == [0, 3, 1, 4, 6, 2, 5, 7, 8]\n"
        "assert basix.finite_element.lex_dof_ordering(basix.ElementFamily.P, basix.CellType.triangle, 3) == [0, 3, 9, 6, 8, 4, 7, 1, 2, 5]\n"
        "pts, wts = basix.quadrature.gauss_jacobi_rule(1.5, 4)\n"
        "assert abs(sum(w * (1 - x) ** 3 for x, w in zip(pts, wts)) - 1 / 5.5) < 1e-14\n"
        "assert basix.cell.volume(basix.CellType.pyramid) == 1 / 3 and basix.cell.facet_orientations(basix.CellType.triangle) == [False, True, False]\n"
        "assert np.allclose(basix.cell.facet_outward_normals(basix.CellType.tetrahedron)[0], np.ones(3) / np.sqrt(3))\n"
        "try:\n"
        "    basix.create_element(basix.ElementFamily.Regge, basix.CellType.triangle, 1)\n"
        "    raise SystemExit('Regge should not be built natively')\n"
        "except basix.NotBuiltNatively:\n"
        "    pass\n" % ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr


@pytest.mark.gpu
def test_reference_pytest_files_pass_against_the_shim():
    if not os.path.isdir(os.path.join(STAGED, "test")):
        pytest.skip("reference tests not staged (oracle/stage_reference_tests.sh needs /root/reference)")
    r = subprocess.run([sys.executable, "-m", "pytest", "test", "-q", "-p", "no:cacheprovider", "-rf", "-x"], cwd=STAGED,
                       capture_output=True, text=True, timeout=1500)
    tail = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else ""
    out_dir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out_dir):
        with open(os.path.join(out_dir, "reference_suite.txt"), "w") as f:
            f.write(tail + "\n")
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-2000:]
    counts = {k: int(v) for v, k in re.findall(r"(\d+) (passed|skipped|failed|xfailed|error)", tail)}
    assert counts.get("failed", 0) == 0 and counts.get("error", 0) == 0
    assert counts.get("passed", 0) >= 6000, tail
