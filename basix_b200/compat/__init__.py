"""``basix`` under its own name, served by the B200 library.

``basix_b200.compat.install()`` puts the directory holding the ``basix`` shim package at the front of ``sys.path``:
after that ``import basix`` / ``import basix._basixcpp`` resolve to thin modules that keep the reference's names,
positional orders and error behaviour (python/basix/*.py, python/basix/wrappers/basix_wrappers.cpp:224-331) and forward
every hot-path call to ``libbasix_b200.so``.  What the native factory does not build (most element families, the
non-warped lattice variants, quadrature tables, UFL / numba glue) raises ``basix.NotBuiltNatively`` (a
``RuntimeError``), never a silent fallback.
"""

import os
import sys


def path() -> str:
    return os.path.dirname(os.path.abspath(__file__))


def install() -> None:
    p = path()
    if p not in sys.path:
        sys.path.insert(0, p)
