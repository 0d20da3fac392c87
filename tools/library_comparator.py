#!/usr/bin/env python
"""Library comparators for the FP64 tensor roofline of BASELINE config 2 (P5 tetrahedron, tabulate(nd=2)):

* cuBLAS DGEMM 8192^3 (torch.matmul, float64): the library FP64 matrix rate on this GPU, next to the DMMA microbenchmark
  (tools/peaks.cu, 37.0 TFLOP/s) that bench.py uses as the roofline denominator;
* the "library path" for the same tabulate: this repository's polyset kernel (P[10][56][npts] to HBM) + one cuBLAS DGEMM
  per derivative slab (C[56x56] @ P[d]) + the transpose to [pt][dof] -- the reference's own structure
  (finite-element.cpp:1834-1888: polyset::tabulate, math::dot -> dgemm, transposing scatter) with library kernels --
  timed against the fused derivative-operator kernel on the same points.
Prints one JSON object."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import basix_b200 as bx  # noqa: E402


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    res = {}
    n = 8192
    A = torch.randn(n, n, device="cuda", dtype=torch.float64)
    B = torch.randn(n, n, device="cuda", dtype=torch.float64)
    C = torch.empty(n, n, device="cuda", dtype=torch.float64)
    ms = min(timed(lambda: torch.matmul(A, B, out=C), 3) for _ in range(3))
    res["cublas_dgemm_8192"] = {"ms": ms, "tflops": 2 * n ** 3 / (ms * 1e-3) / 1e12}
    del A, B, C
    e = bx.create_element(1, 3, 5, 2)
    npts = 4_000_000
    g = torch.Generator(device="cuda").manual_seed(42)
    s, _ = torch.sort(torch.rand(npts, 3, generator=g, device="cuda", dtype=torch.float64), dim=1)
    x = torch.stack([s[:, 0], s[:, 1] - s[:, 0], s[:, 2] - s[:, 1]], dim=1).contiguous()
    Cm = torch.from_numpy(e.coefficient_matrix).cuda()
    out = torch.empty((10, npts, 56, 1), dtype=torch.float64, device="cuda")

    def fused():
        e.tabulate(2, x, out=out)

    tmp = torch.empty((56, npts), dtype=torch.float64, device="cuda")
    out2 = torch.empty((10, npts, 56), dtype=torch.float64, device="cuda")

    def library():
        P = bx.tabulate_polynomial_set(3, 0, 5, 2, x)  # [10][56][npts]
        for d in range(10):
            torch.matmul(Cm, P[d], out=tmp)
            out2[d].copy_(tmp.t())

    t_f = timed(fused, 5)
    t_l = timed(library, 3)
    library()
    fused()
    err = float((out2 - out[..., 0]).abs().max() / out.abs().max())
    res["tabulate_p5_tet_nd2"] = {
        "points": npts, "fused_kernel_ms": t_f, "library_path_ms": t_l, "speedup": t_l / t_f,
        "library_path": "bx polyset kernel -> 10 x cublasDgemm(56x56 @ 56xN) -> transpose copy",
        "max_rel_difference": err}
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
