"""Builds and runs tests/cpp/host_mirror.cpp: the C++ host mirror (include/basix_b200.hpp) of the reference
class over the C ABI.  Host mode (no GPU) checks construction, shapes, the reference's error text and that compute
calls fail loudly; GPU mode checks SURVEY 8c anchors through the C++ API."""

import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT


def dump_element(npz_path, out_path):
    a = np.load(npz_path)
    meta = [int(v) for v in a["meta"]]
    counts = {0: (1,), 1: (2, 1), 2: (3, 3, 1), 3: (4, 6, 4, 1), 4: (4, 4, 1), 5: (8, 12, 6, 1)}[meta[0]]
    offs, flat = a["entity_dof_offsets"], a["entity_dofs_flat"]
    w = []
    w.append(" ".join(str(v) for v in meta[:5]))
    vs = [int(v) for v in a["value_shape"]]
    w.append(" ".join(str(v) for v in [len(vs)] + vs))
    c = a["coefficient_matrix"].ravel()
    w.append(str(len(c)) + " " + " ".join(repr(float(v)) for v in c))
    o = [int(v) for v in a["dof_ordering"]]
    w.append(" ".join(str(v) for v in [len(o)] + o))
    w.append(str(len(counts)))
    k = 0
    for cnt in counts:
        row = [str(cnt)]
        for _ in range(cnt):
            dofs = [int(v) for v in flat[offs[k]:offs[k + 1]]]
            row += [str(len(dofs))] + [str(v) for v in dofs]
            k += 1
        w.append(" ".join(row))
    for ct in (1, 2, 4):
        key = f"etrans_{ct}"
        if key in a.files:
            m = a[key]
            w.append(f"{m.shape[1]} {m.size} " + " ".join(repr(float(v)) for v in m.ravel()))
        else:
            w.append("-1 0")
    with open(out_path, "w") as f:
        f.write("\n".join(w) + "\n")


def build(tmp_path):
    exe = str(tmp_path / "host_mirror")
    libdir = os.path.join(ROOT, "basix_b200")
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "cpp", "host_mirror.cpp"), "-o", exe, "-L", libdir, "-lbasix_b200",
                    f"-Wl,-rpath,{libdir}"], check=True)
    elem = str(tmp_path / "p5.txt")
    dump_element(os.path.join(ROOT, "tests", "golden", "elements", "P5_tetrahedron_gll.npz"), elem)
    return exe, elem


def test_cpp_host_mirror_host_mode(tmp_path):
    exe, elem = build(tmp_path)
    out = subprocess.run([exe, elem, "host"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr


@pytest.mark.gpu
def test_cpp_host_mirror_gpu(tmp_path):
    exe, elem = build(tmp_path)
    out = subprocess.run([exe, elem, "gpu"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "ok (gpu)" in out.stdout
