#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY -- never part of the product path.
#
# Stages the reference's OWN pytest files (unmodified, where they lie under /root/reference/test) into
# oracle/_ref/reftests/test/ so that tests/test_reference_suite.py can run them against the `basix` shim package
# (basix_b200/compat/basix) on the GPU box, where /root/reference does not exist.  oracle/_ref/ is git-ignored (nothing
# of the reference enters the history) but travels with the gpurun snapshot, like oracle/_ref/libbasix_ref.so.
# The only files written here that are ours: conftest.py (puts the shim on sys.path and turns the shim's
# NotBuiltNatively -- element families / lattice variants / set-up data outside the accelerated path -- into skips).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${BASIX_REFERENCE:-/root/reference}"
OUT="$HERE/_ref/reftests"
if [ ! -d "$REF/test" ]; then
  echo "stage_reference_tests: $REF/test not present; keeping any staged copy under $OUT" >&2
  exit 0
fi
FILES="__init__.py utils.py test_mappings.py test_dof_transformations.py test_lagrange.py test_tensor_products.py \
test_dof_ordering.py test_nedelec.py test_rt.py test_orthonormal_basis.py test_interpolation_between_elements.py \
test_lattice.py test_entity_dofs.py test_equality.py test_polynomials.py test_create.py test_cell.py test_continuity.py \
test_interpolation.py test_lexicographic.py test_quadrature.py test_bernstein.py test_iso.py test_cr.py test_hermite.py \
test_regge.py test_nedelec_second_kind.py test_wcoeffs.py test_orthonormal_basis_symbolic.py ${EXTRA_REFERENCE_TESTS:-}"
# not staged: test_custom_element.py (asserts the texts of create_custom_element's validation errors), test_elements.py /
# test_ufl_wrapper.py (basix.ufl), test_numba.py, test_version.py (package metadata), test_cmake / test_pkgconfig
rm -rf "$OUT"
mkdir -p "$OUT/test"
for f in $FILES; do
  cp "$REF/test/$f" "$OUT/test/$f"
done
cat > "$OUT/conftest.py" <<'PY'
"""Ours (not a reference file): run the staged reference tests against the basix shim of basix_b200."""
import os
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..", ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
import basix_b200.compat  # noqa: E402

basix_b200.compat.install()
import basix  # noqa: E402


def _skip_if_unbuilt(outcome):
    exc = outcome.excinfo
    if exc is not None and issubclass(exc[0], basix.NotBuiltNatively):
        try:
            pytest.skip(f"outside the accelerated path: {exc[1]}")
        except pytest.skip.Exception as s:
            outcome.force_exception(s)


@pytest.hookimpl(hookwrapper=True)
def pytest_runtest_call(item):
    outcome = yield
    _skip_if_unbuilt(outcome)


@pytest.hookimpl(hookwrapper=True)
def pytest_runtest_setup(item):
    outcome = yield
    _skip_if_unbuilt(outcome)
PY
echo "stage_reference_tests: staged $(ls "$OUT/test" | wc -l) files into $OUT/test"
