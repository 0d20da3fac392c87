#!/usr/bin/env python
"""What a 20-microsecond kernel can reach under bench.py's timing rule (untimed 512 MB L2 flush, one event pair per
launch): event-pair cost of an empty launch, a copy and a fill that move the same bytes as tabulate(P1 triangle, nd=1,
1 M points) = 16 MB read + 72 MB written, and the tabulate kernel itself with and without the flush."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import basix_b200 as bx  # noqa: E402
from basix_b200 import _capi  # noqa: E402


def timed(fn, n=50, flush=True):
    scratch = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    ts = []
    for _ in range(n):
        if flush:
            scratch.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts = np.array(ts[5:])
    return {"median_us": float(np.median(ts)), "min_us": float(ts.min()), "mean_us": float(ts.mean())}


def main():
    e = bx.create_element(1, 2, 1, 2)
    n = 1_000_000
    x = torch.rand(n, 2, device="cuda", dtype=torch.float64) / 2
    out = torch.empty(e.tabulate_shape(1, n), dtype=torch.float64, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    L = _capi.lib()
    args = (e._h, 1, x.data_ptr(), n, 2, out.data_ptr(), _capi.BX_MEM_DEVICE, stream)
    src = torch.empty(44 << 20, dtype=torch.uint8, device="cuda")
    dst = torch.empty(44 << 20, dtype=torch.uint8, device="cuda")
    w72 = torch.empty(72_000_000, dtype=torch.uint8, device="cuda")
    tiny = torch.empty(32, device="cuda")
    res = {
        "empty_launch": timed(lambda: tiny.fill_(0.0)),
        "copy_44MB_to_44MB_flushed": timed(lambda: dst.copy_(src)),
        "fill_72MB_flushed": timed(lambda: w72.fill_(0)),
        "tabulate_p1_tri_flushed": timed(lambda: L.bx_tabulate_f64(*args)),
        "tabulate_p1_tri_l2_warm": timed(lambda: L.bx_tabulate_f64(*args), flush=False),
        "copy_44MB_to_44MB_l2_warm": timed(lambda: dst.copy_(src), flush=False),
    }
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
