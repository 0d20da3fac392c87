#!/usr/bin/env python
"""Benchmark of the basix hot path on B200.

  python bench.py --gpus N --steps K --warmup W [--impl ours|reference] [--workload NAME]

A "step" is one pass of the hot path over one batch of synthetic input.  The headline workload (default) is
BASELINE.json configs[1]: P5 Lagrange on the tetrahedron, tabulate(nd=2) at 10M points per GPU.  Multi-GPU is
weak scaling: every rank processes its own slice of points / cells, no data-path collective.

JSON line keys (contract):
  value         whole-job throughput, inputs resident in HBM, CUDA events on the launch stream, max over ranks
  e2e           same metric through the C ABI with HOST (pinned) buffers: H2D of the inputs and D2H of the
                result inside the timed region
  roofline      the dominant kernel: algorithmic bytes (or flops) per launch / its measured duration, against
                the measured peak (MEASURED_PEAKS.json for HBM; FP64 tensor peak measured by tools/peaks.cu,
                profiles/r01_peaks_fp64_hbm_pcie.json, since the driver file has no FP64 figure)
  cpu_baseline  the compiled reference (oracle/_ref) on this box's host cores, bounded sample (rank 0, N=1)
  --impl reference   times only that CPU reference: same metric / config / unit
Other workloads (--workload) are the remaining BASELINE.json configs; they print the same line.
"""

from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden", "elements")
FP64_TENSOR_PEAK_TFLOPS = 37.0  # measured: DMMA m8n8k4 microbenchmark, profiles/r01_peaks_fp64_hbm_pcie.json
L2_BYTES = 126 << 20
FLUSH_BYTES = 512 << 20
METRIC_VALUES = "tabulated basis values/sec (pts x dofs x derivs)"
METRIC_CELLS = "cells/sec"


# ---- synthetic inputs ---------------------------------------------------------------------------
def random_points(cell, n, seed):
    rng = np.random.default_rng(seed)
    if cell == 3:
        s = np.sort(rng.random((n, 3)), axis=1)
        return np.ascontiguousarray(np.stack([s[:, 0], s[:, 1] - s[:, 0], s[:, 2] - s[:, 1]], axis=1))
    if cell == 2:
        u = rng.random((n, 2))
        flip = u.sum(axis=1) > 1
        u[flip] = 1 - u[flip]
        return np.ascontiguousarray(u)
    return np.ascontiguousarray(rng.random((n, {1: 1, 4: 2, 5: 3}[cell])))


def random_jacobians(nc, seed):
    """J = I + 0.3 N(0,1), rejecting |det| < 0.1; detJ and K = J^-1 in float64."""
    rng = np.random.default_rng(seed)
    J = np.eye(3)[None] + 0.3 * rng.standard_normal((nc, 3, 3))
    det = np.linalg.det(J)
    bad = np.abs(det) < 0.1
    while bad.any():
        J[bad] = np.eye(3)[None] + 0.3 * rng.standard_normal((int(bad.sum()), 3, 3))
        det = np.linalg.det(J)
        bad = np.abs(det) < 0.1
    return np.ascontiguousarray(J), np.ascontiguousarray(det), np.ascontiguousarray(np.linalg.inv(J))


def random_cell_info(nc, seed):
    """Realistic tetrahedron cell_info: per face reflect in {0,1}, rotations in {0,1,2}; 6 edge bits."""
    rng = np.random.default_rng(seed)
    ci = np.zeros(nc, dtype=np.uint32)
    for f in range(4):
        ci |= (rng.integers(0, 2, nc).astype(np.uint32) | (rng.integers(0, 3, nc).astype(np.uint32) << 1)) << (3 * f)
    ci |= rng.integers(0, 64, nc).astype(np.uint32) << 12
    return ci


# Device-side generators of the same distributions (the side configs of the default line and the strong-scaling leg
# never touch host memory: generating 28 GB of normal deviates with numpy takes minutes, on the GPU milliseconds).
def device_points(cell, n, seed):
    import torch

    g = torch.Generator(device="cuda").manual_seed(seed)
    tdim = {1: 1, 2: 2, 3: 3, 4: 2, 5: 3}[cell]
    u = torch.rand(n, tdim, generator=g, device="cuda", dtype=torch.float64)
    if cell == 3:
        s, _ = torch.sort(u, dim=1)
        return torch.stack([s[:, 0], s[:, 1] - s[:, 0], s[:, 2] - s[:, 1]], dim=1).contiguous()
    if cell == 2:
        flip = u.sum(dim=1) > 1
        u[flip] = 1 - u[flip]
    return u.contiguous()


def device_jacobians(nc, seed):
    import torch

    g = torch.Generator(device="cuda").manual_seed(seed)
    eye = torch.eye(3, dtype=torch.float64, device="cuda")
    J = eye[None] + 0.3 * torch.randn(nc, 3, 3, generator=g, device="cuda", dtype=torch.float64)

    def det_inv(J):
        a, b, c, d, e, f, g_, h, i = (J[:, r, q] for r in range(3) for q in range(3))
        co = torch.stack([e * i - f * h, c * h - b * i, b * f - c * e, f * g_ - d * i, a * i - c * g_, c * d - a * f,
                          d * h - e * g_, b * g_ - a * h, a * e - b * d], dim=1)
        det = a * co[:, 0] + b * co[:, 3] + c * co[:, 6]
        return det.contiguous(), (co / det[:, None]).reshape(-1, 3, 3).contiguous()

    det, K = det_inv(J)
    J[det.abs() < 0.1] = eye
    det, K = det_inv(J)
    return J.contiguous(), det, K


def device_cell_info(nc, seed):
    import torch

    g = torch.Generator(device="cuda").manual_seed(seed)
    ci = torch.zeros(nc, dtype=torch.int64, device="cuda")
    for f in range(4):
        refl = torch.randint(0, 2, (nc,), generator=g, device="cuda")
        rot = torch.randint(0, 3, (nc,), generator=g, device="cuda")
        ci |= (refl | (rot << 1)) << (3 * f)
    ci |= torch.randint(0, 64, (nc,), generator=g, device="cuda") << 12
    return ci.to(torch.int32)


def _multi_indices(tdim, nd):
    """Derivative multi-indices of total order <= nd (one per slab of tabulate's output)."""
    if tdim == 1:
        return [(k,) for k in range(nd + 1)]
    if tdim == 2:
        return [(kx, ky) for kx in range(nd + 1) for ky in range(nd + 1 - kx)]
    return [(kx, ky, kz) for kx in range(nd + 1) for ky in range(nd + 1 - kx) for kz in range(nd + 1 - kx - ky)]


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
            nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
            nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.02)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f).get("hbm_gbs", 6650.0)), "MEASURED_PEAKS.json hbm_gbs"
    return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


def measured_traffic(workload, units_per_launch):
    """roofline.traffic: DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) of the dominant kernel per launch,
    from the committed ncu --set full captures (latest round first).  Returns (bytes or None, source or None)."""
    for name in ("r02_traffic.json", "r01_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                ent = json.load(f).get(workload)
            if ent:
                return float(ent["bytes_per_unit"]) * units_per_launch, f"ncu --set full capture, profiles/{name}"
        except (OSError, ValueError, KeyError):
            pass
    return None, None


def load_element(name, ref_args=None):
    """Elements are built natively by the library (basix_b200.create_element: P, RT, N1curl); `name` is the
    committed array fixture of the same reference element, used only if a family is not built natively."""
    import basix_b200 as bx

    if ref_args is not None and ref_args[0] in (1, 2, 3):
        return bx.create_element(*ref_args)
    return bx.FiniteElement.from_arrays(np.load(os.path.join(GOLDEN, name + ".npz")))


# ======================================================================================================
# Workloads.  Each provides setup / step (device-resident), e2e_setup / e2e_step (host buffers through the
# C ABI) and cpu_setup / cpu_step (the compiled reference), plus algorithmic bytes and flops PER LAUNCH.
class Tabulate:
    metric, unit = METRIC_VALUES, "values/s"

    def __init__(self, element, ref_args, cell, nd, npts, chunks, desc, bound):
        self.element_name, self.ref_args, self.cell, self.nd = element, ref_args, cell, nd
        self.npts, self.chunks, self.desc, self.bound = npts, chunks, desc, bound

    def setup(self, rank, device_only=False):
        import torch

        self.e = load_element(self.element_name, self.ref_args)
        e = self.e
        self.cnp = self.npts // self.chunks  # points per launch (the table of one launch stays resident)
        self.shape = e.tabulate_shape(self.nd, self.cnp)
        self.values_per_step = int(np.prod(self.shape)) * self.chunks
        self.units_per_step = self.values_per_step
        if device_only:
            self.x = device_points(self.cell, self.npts, 42 + rank)
        else:
            self.x_host = torch.from_numpy(random_points(self.cell, self.npts, 42 + rank)).pin_memory()
            self.x = self.x_host.cuda()
        self.out = torch.empty(self.shape, dtype=torch.float64, device="cuda")
        self.path = e.tabulate_path(self.nd)
        self.launches_per_step = self.chunks
        self.units_per_launch = self.cnp
        self.alg_bytes = self.cnp * e._tdim * 8 + int(np.prod(self.shape)) * 8
        # FP64 flops per point: the reference contracts every derivative slab with the full psize-column
        # coefficient matrix; the derivative-operator kernel contracts slab d with dim P_{n-|alpha_d|} columns
        # (csrc/tabulate_dop.cu).  roofline.achieved counts the flops of the algorithm that runs.
        dense = 2.0 * e.dim * e.value_size * e._psize * self.shape[0]
        self.flops_per_point = dense
        if self.path == "derivative-operator":
            from math import comb

            tdim, n = e._tdim, e.embedded_superdegree
            orders = [sum(a) for a in _multi_indices(tdim, self.nd)]
            self.flops_per_point = 2.0 * e.dim * e.value_size * sum(comb(n - m + tdim, tdim) if m <= n else 0 for m in orders)
        self.flops = self.flops_per_point * self.cnp
        self.config = {"workload": self.desc, "points_per_gpu": self.npts, "nd": self.nd, "tabulate_path": self.path,
                       "points_per_launch": self.cnp, "fp64_flops_per_point": self.flops_per_point,
                       "fp64_flops_per_point_reference_algorithm": dense,
                       "l2": "table of one launch is %.2f GB (>> 126 MB L2)" % (int(np.prod(self.shape)) * 8 / 1e9)}
        if self.alg_bytes < 2 * L2_BYTES:
            # the arrays of one launch fit in L2: run_ours() flushes it between timed steps and times step by step
            self.flush_l2 = True
            self.config["l2"] = ("table of one launch is %.0f MB (< 126 MB L2): L2 flushed between timed steps by writing a "
                                 "%d MB buffer (untimed), steps timed one by one" % (int(np.prod(self.shape)) * 8 / 1e6,
                                                                                     FLUSH_BYTES >> 20))

        # the timed call is the C-ABI entry point with device pointers (bx_tabulate_f64, BX_MEM_DEVICE) on torch's
        # current stream; arguments are prepared once so that a 20-microsecond launch (P1 triangle) is not hidden
        # behind Python argument checking
        from basix_b200 import _capi

        self._fn, self._check = _capi.lib().bx_tabulate_f64, _capi.check
        stream = torch.cuda.current_stream().cuda_stream
        row = e._tdim * 8
        self._calls = [(e._h, self.nd, self.x.data_ptr() + c * self.cnp * row, self.cnp, e._tdim, self.out.data_ptr(),
                        _capi.BX_MEM_DEVICE, stream) for c in range(self.chunks)]

    def step(self):
        for args in self._calls:
            self._check(self._fn(*args))

    def e2e_setup(self, avail_bytes):
        import torch

        per_pt = self.values_per_step // self.npts * 8
        pts = self.npts
        while pts * per_pt > 0.4 * avail_bytes and pts > 100_000:
            pts //= 2
        self.e2e_pts = pts
        self.e2e_shape = self.e.tabulate_shape(self.nd, pts)
        del self.out
        torch.cuda.empty_cache()
        self.out_host = torch.empty(self.e2e_shape, dtype=torch.float64).pin_memory()
        self.e2e_units = int(np.prod(self.e2e_shape))
        self.h2d, self.d2h = pts * self.e._tdim * 8, self.e2e_units * 8
        return {"points_per_gpu": pts}

    def e2e_step(self):
        from basix_b200 import _capi

        _capi.check(_capi.lib().bx_tabulate_f64(self.e._h, self.nd, self.x_host.data_ptr(), self.e2e_pts,
                                                self.x_host.shape[1], self.out_host.data_ptr(), _capi.BX_MEM_HOST,
                                                None))

    def cpu_setup(self, cores):
        from oracle import ref

        r = ref.RefElement(*self.ref_args)
        per_pt = r.dim * r.value_size * ref.polyset_nderivs(r.cell_type, self.nd)
        n = int(min(self.npts, 2_000_000, max(20_000, 6e7 * cores / per_pt)))
        x = random_points(self.cell, n, 42)
        out = np.empty((ref.polyset_nderivs(r.cell_type, self.nd), n, r.dim, r.value_size))
        chunk = max(1024, min(65536, n // (cores * 4) or 1024))
        self._cpu = (r, x, out, chunk, cores, n)
        return per_pt * n, (f"{n} points per step, {chunk}-point chunks over {cores} std::threads, oracle/_ref "
                            f"(unmodified reference, g++ -O2, scipy OpenBLAS, 1 BLAS thread per chunk)")

    def cpu_step(self):
        from oracle.ref import _check, _ptr, lib

        r, x, out, chunk, cores, n = self._cpu
        _check(lib().bxref_tabulate(r._h, self.nd, _ptr(x), n, x.shape[1], _ptr(out), chunk, cores))


class PushForward:
    metric, unit = METRIC_CELLS, "cells/s"
    bound = "hbm"

    def __init__(self, element, ref_args, nc, rows, pull, desc, bcast=False):
        self.element_name, self.ref_args, self.nc, self.rows, self.pull, self.desc = element, ref_args, nc, rows, pull, desc
        self.bcast = bcast  # one slab of reference values for all cells (push_forward_broadcast)

    def setup(self, rank, device_only=False):
        import torch

        self.e = load_element(self.element_name, self.ref_args)
        nc, rows = self.nc, self.rows
        if device_only:
            J, detJ, K = device_jacobians(nc, 7 + rank)
            g = torch.Generator(device="cuda").manual_seed(8 + rank)
            U = torch.randn((rows, 3) if self.bcast else (nc, rows, 3), generator=g, device="cuda", dtype=torch.float64)
            self.dev = [U, J, detJ, K]
        else:
            J, detJ, K = random_jacobians(nc, 7 + rank)
            U = np.random.default_rng(8 + rank).standard_normal((rows, 3) if self.bcast else (nc, rows, 3))
            self.host = [torch.from_numpy(a).pin_memory() for a in (U, J, detJ, K)]
            self.dev = [t.cuda() for t in self.host]
        self.units_per_step = nc
        self.launches_per_step = 1
        self.units_per_launch = nc
        # covariant push reads U and K, pull reads u and J; contravariant also reads detJ
        contra = int(self.e.map_type) == 3
        self.alg_bytes = nc * ((1 if self.bcast else 2) * rows * 3 * 8 + 72 + (8 if contra else 0))
        self.flops = 18.0 * rows * nc
        self.config = {"workload": self.desc, "cells_per_gpu": nc, "rows": rows,
                       "l2": "arrays are %.2f GB per step (>> 126 MB L2)" % (self.alg_bytes / 1e9)}

    def step(self):
        f = self.e.push_forward_broadcast if self.bcast else self.e.pull_back if self.pull else self.e.push_forward
        self.result = f(*self.dev)

    def e2e_setup(self, avail_bytes):
        import torch

        self.out_host = torch.empty((self.nc, self.rows, 3), dtype=torch.float64).pin_memory()
        self.e2e_units = self.nc
        # the library ships only the arrays the map reads (csrc/capi.cu map_any): covariant maps the "K" argument,
        # contravariant maps the "J" argument and detJ; pull_back swaps J and K
        U, J, d, K = self.host
        contra = int(self.e.map_type) == 3
        mats = ([K if self.pull else J, d] if contra else [J if self.pull else K])
        self.h2d = sum(t.numel() * 8 for t in [U] + mats)
        self.d2h = self.out_host.numel() * 8
        return {"cells_per_gpu": self.nc}

    def e2e_step(self):
        from basix_b200 import _capi

        U, J, d, K = self.host
        L = _capi.lib()
        fn = L.bx_push_forward_broadcast_f64 if self.bcast else L.bx_pull_back_f64 if self.pull else L.bx_push_forward_f64
        _capi.check(fn(self.e._h, U.data_ptr(), J.data_ptr(), d.data_ptr(), K.data_ptr(), self.nc, self.rows, 3, 3, 3,
                       self.out_host.data_ptr(), _capi.BX_MEM_HOST, None))

    def cpu_setup(self, cores):
        from oracle import ref

        r = ref.RefElement(*self.ref_args)
        n = int(min(self.nc, max(100_000, 4e6 * cores / self.rows)))
        J, detJ, K = random_jacobians(n, 7)
        U = np.random.default_rng(8).standard_normal((n, self.rows, 3))  # the reference needs the replicated array
        self._cpu = (r, U, J, detJ, K, cores)
        return n, f"{n} cells per step over {cores} std::threads, oracle/_ref (unmodified reference, g++ -O2)"

    def cpu_step(self):
        r, U, J, detJ, K, cores = self._cpu
        (r.pull_back if self.pull else r.push_forward)(U, J, detJ, K, nthreads=cores)


class TApply:
    metric, unit = METRIC_CELLS, "cells/s"
    bound = "hbm"

    def __init__(self, element, ref_args, nc, bs, permute, desc):
        self.element_name, self.ref_args, self.nc, self.bs, self.permute, self.desc = element, ref_args, nc, bs, permute, desc

    def setup(self, rank, device_only=False):
        import torch

        self.e = load_element(self.element_name, self.ref_args)
        e, nc = self.e, self.nc
        esize = 4 if self.permute else 8
        if device_only:
            self.ci = device_cell_info(nc, 11 + rank)
            if self.permute:
                self.data = torch.arange(e.dim, dtype=torch.int32, device="cuda").repeat(nc, 1)
            else:
                self.data = torch.empty(nc, e.dim * self.bs, dtype=torch.float64, device="cuda")
                g = torch.Generator(device="cuda").manual_seed(12 + rank)
                step = max(1, nc // 8)  # randn in pieces: the generator's scratch stays small next to an 84 GB array
                for c0 in range(0, nc, step):
                    self.data[c0:c0 + step].normal_(generator=g)
        else:
            ci = random_cell_info(nc, 11 + rank)
            self.ci_host = torch.from_numpy(ci.view(np.int32)).pin_memory()
            self.ci = self.ci_host.cuda()
            if self.permute:
                d = np.tile(np.arange(e.dim, dtype=np.int32), (nc, 1))
            else:
                d = np.random.default_rng(12 + rank).standard_normal((nc, e.dim * self.bs))
            self.data_host = torch.from_numpy(d).pin_memory()
            self.data = self.data_host.cuda()
        self.units_per_step = nc
        self.launches_per_step = 1
        self.units_per_launch = nc
        # only the dofs that CAN move (those on edges, and on faces of 3-D cells) need to be read
        # and written (+ 4 B of cell_info); which of them actually move depends on each cell's cell_info
        ed = e.entity_dofs
        movable = sum(len(d) for dim in range(1, e._tdim) for d in ed[dim])
        self.alg_bytes = nc * (2 * movable * esize * self.bs + 4)
        self.flops = 0.0
        # What DRAM can deliver at best for this strided in-place pattern: the movable range of a cell is one run of
        # `run` bytes every `pitch` bytes; reads arrive in whole 128-byte lines, writes leave in 32-byte sectors
        # (ncu: dram__bytes of the kernel, profiles/*traffic.json).  Exact count over one period of the alignment.
        lo = min(min(d) for dim in range(1, e._tdim) for d in ed[dim] if d) * esize * self.bs
        run, pitch = movable * esize * self.bs, e.dim * esize * self.bs
        period = 128 // np.gcd(pitch, 128)

        def touched(g):  # distinct g-byte blocks the runs of one period of cells touch, in bytes per cell
            blocks = set()
            for c in range(period):
                blocks.update(range((c * pitch + lo) // g, (c * pitch + lo + run - 1) // g + 1))
            return len(blocks) * g / period

        rd, wr = touched(128), touched(32)
        self.line_bytes = nc * (rd + wr + 4)
        self.line_what = ("movable run of %d B every %d B: %.0f B read per cell in whole 128-byte lines + %.0f B written in "
                          "32-byte sectors + 4 B cell_info" % (run, pitch, rd, wr))
        self.config = {"workload": self.desc, "cells_per_gpu": nc, "block_size": self.bs, "movable_dofs": movable,
                       "whole_array_bytes_per_cell": 2 * e.dim * esize * self.bs + 4,
                       "l2": "data is %.2f GB per step (>> 126 MB L2)" % (nc * e.dim * esize * self.bs / 1e9)}

    def step(self):
        if self.permute:
            self.e.permute_batch(self.data, self.ci)
        else:
            self.e.T_apply_batch(self.data, self.bs, self.ci)

    def e2e_setup(self, avail_bytes):
        self.e2e_units = self.nc
        self.h2d = self.data_host.numel() * self.data_host.element_size() + self.nc * 4
        self.d2h = self.data_host.numel() * self.data_host.element_size()
        return {"cells_per_gpu": self.nc}

    def e2e_step(self):
        from basix_b200 import _capi

        L = _capi.lib()
        if self.permute:
            _capi.check(L.bx_permute_i32(self.e._h, 0, self.data_host.data_ptr(), self.nc, self.ci_host.data_ptr(),
                                         _capi.BX_MEM_HOST, None))
        else:
            _capi.check(L.bx_T_apply_f64(self.e._h, 0, self.data_host.data_ptr(), self.nc, self.e.dim * self.bs,
                                         self.bs, self.ci_host.data_ptr(), _capi.BX_MEM_HOST, None))

    def cpu_setup(self, cores):
        from oracle import ref

        r = ref.RefElement(*self.ref_args)
        n = int(min(self.nc, max(100_000, (4e6 if self.permute else 1e6) * cores)))
        ci = random_cell_info(n, 11)
        if self.permute:
            d = np.tile(np.arange(r.dim, dtype=np.int32), (n, 1))
        else:
            d = np.random.default_rng(12).standard_normal((n, r.dim * self.bs))
        self._cpu = (r, d, ci, cores)
        return n, f"{n} cells per step over {cores} std::threads, oracle/_ref (unmodified reference, g++ -O2)"

    def cpu_step(self):
        r, d, ci, cores = self._cpu
        if self.permute:
            r.permute_batch(d, ci, nthreads=cores)
        else:
            r.T_apply_batch("T_apply", d, self.bs, ci, nthreads=cores)


def device_tet_coords(nc, seed):
    """Vertex coordinates of nc affine tetrahedra: the reference tetrahedron + 0.2 N(0,1) per coordinate."""
    import torch

    g = torch.Generator(device="cuda").manual_seed(seed)
    base = torch.zeros(4, 3, dtype=torch.float64, device="cuda")
    base[1:] = torch.eye(3, dtype=torch.float64, device="cuda")
    return (base[None] + 0.2 * torch.randn(nc, 4, 3, generator=g, device="cuda", dtype=torch.float64)).contiguous()


class PhysicalBasis:
    """tabulate-once -> T_apply -> push_forward fused (geometry from vertex coordinates on chip), or
    the cell geometry alone (J, detJ, K from vertex coordinates)."""
    metric, unit = METRIC_CELLS, "cells/s"
    bound = "hbm"

    def __init__(self, element, ref_args, nc, npts, desc, geometry_only=False, pointwise=False):
        self.element_name, self.ref_args, self.nc, self.npts, self.desc = element, ref_args, nc, npts, desc
        self.geometry_only = geometry_only
        self.pointwise = pointwise  # per-(cell, point) Jacobians given (non-affine cells)

    def setup(self, rank, device_only=False):
        import torch

        import basix_b200 as bx

        self.e = load_element(self.element_name, self.ref_args)
        e, nc = self.e, self.nc
        self.coords = device_tet_coords(nc, 31 + rank)
        self.ci = device_cell_info(nc, 11 + rank)
        self.X = device_points(3, self.npts, 5)
        self.U = e.tabulate(0, self.X)[0].contiguous()
        self.units_per_step = self.units_per_launch = nc
        self.launches_per_step = 1
        self.geometry = bx.geometry
        if self.geometry_only == "hex":
            # trilinear hexahedra at the 2 x 2 x 2 Gauss points: the general (non-affine) form of bx_geometry
            q1 = bx.create_element(1, 5, 1, 2)
            gp = 0.5 + 0.5 * np.array([-1.0, 1.0]) / np.sqrt(3.0)
            Xq = np.array([[a, b, c] for a in gp for b in gp for c in gp])
            self.dphi = torch.from_numpy(np.ascontiguousarray(q1.tabulate(1, Xq)[1:4, :, :, 0])).cuda()
            g = torch.Generator(device="cuda").manual_seed(41 + rank)
            base = torch.tensor([[i, j, k] for k in (0.0, 1.0) for j in (0.0, 1.0) for i in (0.0, 1.0)], dtype=torch.float64,
                                device="cuda")
            self.coords = (base[None] + 0.1 * torch.randn(nc, 8, 3, generator=g, device="cuda", dtype=torch.float64)).contiguous()
            self.alg_bytes = nc * (192 + 8 * 152)
            self.config = {"workload": self.desc, "cells_per_gpu": nc, "points_per_cell": 8,
                           "l2": "arrays are %.2f GB per step (>> 126 MB L2)" % (self.alg_bytes / 1e9)}
        elif self.geometry_only:
            self.alg_bytes = nc * (96 + 72 + 8 + 72)
            self.config = {"workload": self.desc, "cells_per_gpu": nc,
                           "l2": "arrays are %.2f GB per step (>> 126 MB L2)" % (self.alg_bytes / 1e9)}
        else:
            rows = self.npts * e.dim
            self.out_bytes = rows * 3 * 8
            self.alg_bytes = nc * (96 + 4 + self.out_bytes)
            # the same result through the separate calls: geometry writes J / detJ / K; the replicated reference values
            # are transformed in place (read + write of the movable dofs) and pushed forward (read values and K, write)
            ed = e.entity_dofs
            movable = sum(len(d) for dim in range(1, 3) for d in ed[dim])
            unfused = 96 + 152 + self.npts * 2 * movable * 24 + 4 + (rows * 24 + 72) + rows * 24
            self.config = {"workload": self.desc, "cells_per_gpu": nc, "points_per_cell": self.npts, "rows_per_cell": rows,
                           "bytes_per_cell_fused": 100 + self.out_bytes, "bytes_per_cell_three_calls": unfused,
                           "l2": "arrays are %.2f GB per step (>> 126 MB L2)" % (self.alg_bytes / 1e9)}
        if self.pointwise:
            g = torch.Generator(device="cuda").manual_seed(51 + rank)
            self.Jp = (torch.eye(3, device="cuda", dtype=torch.float64)[None, None]
                       + 0.2 * torch.randn(nc, self.npts, 3, 3, generator=g, device="cuda", dtype=torch.float64)).contiguous()
            self.dp = torch.linalg.det(self.Jp)
            self.Kp = torch.linalg.inv(self.Jp).contiguous()
            self.alg_bytes = nc * (self.npts * 72 + 4 + self.out_bytes)
            self.config["geometry"] = "one Jacobian per (cell, point): K[nc][npts][3][3] read, nothing computed on chip"
        self.flops = 0.0
        if not device_only:
            self.coords_host = self.coords.cpu().pin_memory()
            self.ci_host = self.ci.cpu().pin_memory()

    def step(self):
        if self.geometry_only == "hex":
            self.result = self.geometry(self.coords, self.dphi)
        elif self.geometry_only:
            self.result = self.geometry(self.coords)
        elif self.pointwise:
            self.result = self.e.physical_basis(None, U=self.U, J=self.Jp, detJ=self.dp, K=self.Kp, cell_info=self.ci)
        else:
            self.result = self.e.physical_basis(None, U=self.U, coords=self.coords, cell_info=self.ci)

    def extra(self):
        """The same result through the library's separate calls (geometry, T_apply on the replicated reference values,
        push_forward), timed on a 2M-cell slice: what the fusion saves."""
        import torch

        if self.geometry_only or self.pointwise:
            return None
        import basix_b200 as bx

        e, n, npts = self.e, min(self.nc, 2_000_000), self.npts
        coords, ci = self.coords[:n], self.ci[:n]
        ci_rep = ci.repeat_interleave(npts) if npts > 1 else ci

        def three_calls():
            J, detJ, K = bx.geometry(coords)
            rep = self.U.reshape(1, npts, -1).expand(n, npts, e.dim * 3).contiguous().reshape(n * npts, e.dim * 3)
            e.T_apply_batch(rep, 3, ci_rep)
            return e.push_forward(rep.reshape(n, npts * e.dim, 3), J, detJ, K)

        for _ in range(2):
            out = three_calls()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            out = three_calls()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 3
        fused = self.e.physical_basis(None, U=self.U, coords=coords, cell_info=ci)
        same = bool(torch.equal(fused.reshape(out.shape), out))
        del out, fused
        return {"three_calls_ns_per_cell": ms * 1e6 / n, "three_calls": "bx.geometry + replicate + T_apply_batch(block 3) + "
                "push_forward on %d cells, device-resident" % n, "fused_equals_three_calls_bitwise": same}

    def e2e_setup(self, avail_bytes):
        import torch

        self.e2e_units = self.nc
        self.result = None
        torch.cuda.empty_cache()
        if self.geometry_only == "hex":
            self.coords_host = self.coords.cpu().pin_memory()
            self.dphi_host = self.dphi.cpu()
            self.out_host = [torch.empty(s, dtype=torch.float64).pin_memory() for s in
                             ((self.nc, 8, 3, 3), (self.nc, 8), (self.nc, 8, 3, 3))]
            self.h2d, self.d2h = self.nc * 192, self.nc * 8 * 152
        elif self.geometry_only:
            self.out_host = [torch.empty(s, dtype=torch.float64).pin_memory() for s in
                             ((self.nc, 3, 3), (self.nc,), (self.nc, 3, 3))]
            self.h2d, self.d2h = self.nc * 96, self.nc * 152
        else:
            self.out_host = torch.empty((self.nc, self.npts, self.e.dim, 3), dtype=torch.float64).pin_memory()
            self.X_host = self.X.cpu()
            self.h2d, self.d2h = self.nc * 100 + self.X_host.numel() * 8, self.out_host.numel() * 8
            if self.pointwise:
                self.K_host = self.Kp.cpu().pin_memory()
                self.h2d = self.nc * (self.npts * 72 + 4) + self.X_host.numel() * 8
        return {"cells_per_gpu": self.nc}

    def e2e_step(self):
        from basix_b200 import _capi

        L = _capi.lib()
        if self.geometry_only == "hex":
            J, d, K = self.out_host
            _capi.check(L.bx_geometry_f64(self.coords_host.data_ptr(), self.dphi_host.data_ptr(), self.nc, 8, 3, 3, 8,
                                          J.data_ptr(), d.data_ptr(), K.data_ptr(), _capi.BX_MEM_HOST, None))
        elif self.geometry_only:
            J, d, K = self.out_host
            _capi.check(L.bx_geometry_f64(self.coords_host.data_ptr(), None, self.nc, 4, 3, 3, 1, J.data_ptr(), d.data_ptr(),
                                          K.data_ptr(), _capi.BX_MEM_HOST, None))
        elif self.pointwise:
            _capi.check(L.bx_physical_basis_pointwise_f64(self.e._h, 0, self.X_host.data_ptr(), self.npts, 3, None, None,
                                                          self.K_host.data_ptr(), self.ci_host.data_ptr(), self.nc, 3,
                                                          self.out_host.data_ptr(), _capi.BX_MEM_HOST, None))
        else:
            # points in, physical basis functions out: the reference values never exist on the host
            _capi.check(L.bx_physical_basis_f64(self.e._h, 0, self.X_host.data_ptr(), self.npts, 3,
                                                self.coords_host.data_ptr(), None, None, None, self.ci_host.data_ptr(),
                                                self.nc, 3, self.out_host.data_ptr(), _capi.BX_MEM_HOST, None))

    cpu_kind = "reference"

    def cpu_setup(self, cores):
        from oracle import ref, restate

        r = ref.RefElement(*self.ref_args)
        n = int(min(self.nc, max(50_000, 2e6 * cores / (self.npts * r.dim))))
        rng = np.random.default_rng(31)
        base = np.zeros((4, 3))
        base[1:] = np.eye(3)
        coords = base[None] + 0.2 * rng.standard_normal((n, 4, 3))
        if self.geometry_only == "hex":
            self.cpu_kind = "port"
            n = int(min(self.nc, 200_000))
            base = np.array([[i, j, k] for k in (0.0, 1.0) for j in (0.0, 1.0) for i in (0.0, 1.0)])
            coords = base[None] + 0.1 * rng.standard_normal((n, 8, 3))
            self._cpu = (restate, coords, self.dphi.cpu().numpy())
            return n, f"{n} cells x 8 points per step, oracle/restate.py geometry (numpy, vectorised over cells)"
        if self.geometry_only:
            self.cpu_kind = "port"
            self._cpu = (restate, coords)
            return n, f"{n} cells per step, oracle/restate.py geometry (numpy, vectorised over cells; the reference has no batched geometry)"
        X = random_points(3, self.npts, 5)
        U = r.tabulate(0, X)[0]
        self._cpu = (r, restate, coords, random_cell_info(n, 11), U, cores)
        return n, (f"{n} cells per step over {cores} std::threads: the reference's T_apply (block size 3) on the replicated "
                   f"reference values and push_forward, oracle/_ref (unmodified reference, g++ -O2); geometry by numpy")

    def cpu_step(self):
        if self.geometry_only == "hex":
            restate, coords, dphi = self._cpu
            self._cpu_out = restate.geometry(coords, dphi)
            return
        if self.geometry_only:
            restate, coords = self._cpu
            self._cpu_out = restate.geometry(coords)
            return
        r, restate, coords, ci, U, cores = self._cpu
        n, npts = len(coords), self.npts
        J, detJ, K = restate.geometry(coords)
        rep = np.ascontiguousarray(np.broadcast_to(U.reshape(1, npts, -1), (n, npts, U.shape[1] * 3))).reshape(n * npts, -1)
        r.T_apply_batch("T_apply", rep, 3, np.repeat(ci, npts), nthreads=cores)
        self._cpu_out = r.push_forward(rep.reshape(n, npts * U.shape[1], 3), J, detJ, K, nthreads=cores)


class Interpolate:
    """coefficients = interpolation_matrix @ values over many cells (the step after pull_back)."""
    metric, unit = METRIC_CELLS, "cells/s"
    bound = "tensor"
    cpu_kind = "port"  # numpy dgemm with the reference's matrix: the reference has no batched call to time

    def __init__(self, element, ref_args, nc, desc, pulled=False):
        self.element_name, self.ref_args, self.nc, self.desc = element, ref_args, nc, desc
        self.pulled = pulled  # pull_back fused into the interpolation (bx_interpolate_pulled)

    def setup(self, rank, device_only=False):
        import torch

        # the reference's own element (committed arrays of the unmodified reference, tests/golden/elements): its
        # interpolation points are the reference's Xiao-Gimbutas points; the native factory's Gauss-Jacobi points
        # give the same functionals with a larger matrix
        import basix_b200 as bx

        self.e = bx.FiniteElement.from_arrays(np.load(os.path.join(GOLDEN, self.element_name + ".npz")))
        e, nc = self.e, self.nc
        self.K, self.N = e.value_size * e.points.shape[0], e.dim
        g = torch.Generator(device="cuda").manual_seed(21 + rank)
        self.values = torch.randn(nc, self.K, generator=g, device="cuda", dtype=torch.float64)
        self.units_per_step = nc
        self.launches_per_step = 1
        self.units_per_launch = nc
        self.alg_bytes = nc * (self.K + self.N) * 8
        self.flops = 2.0 * self.K * self.N * nc
        self.config = {"workload": self.desc, "cells_per_gpu": nc, "values_per_cell": self.K, "ndofs": self.N,
                       "fp64_flops_per_cell": 2 * self.K * self.N,
                       "l2": "arrays are %.2f GB per step (>> 126 MB L2)" % (self.alg_bytes / 1e9)}
        if self.pulled:
            # physical values u[nc][npts][3] and the Jacobians; the pulled-back values never exist in memory
            self.u = self.values.reshape(nc, e.points.shape[0], 3)
            self.J = torch.eye(3, device="cuda", dtype=torch.float64)[None] + 0.3 * torch.randn(
                nc, 3, 3, generator=g, device="cuda", dtype=torch.float64)
            self.detJ = torch.linalg.det(self.J)
            self.Kinv = torch.linalg.inv(self.J).contiguous()
            self.alg_bytes = nc * (self.K + 9 + self.N) * 8
            self.flops = (2.0 * self.K * self.N + 5.0 * self.K) * nc
            self.config["fp64_flops_per_cell"] = 2 * self.K * self.N + 5 * self.K
            self.config["bytes_per_cell_fused"] = (self.K + 9 + self.N) * 8
            self.config["bytes_per_cell_two_calls"] = (2 * self.K + 9) * 8 + (self.K + self.N) * 8

    def step(self):
        if self.pulled:
            self.result = self.e.interpolate_pulled_batch(self.u, self.J, self.detJ, self.Kinv)
        else:
            self.result = self.e.interpolate_batch(self.values)

    def extra(self):
        """The same coefficients through the two separate calls (pull_back, interpolate): what the fusion saves."""
        import torch

        if not self.pulled:
            return None
        n = min(self.nc, 4_000_000)
        u, J, d, K = self.u[:n], self.J[:n], self.detJ[:n], self.Kinv[:n]

        def two_calls():
            return self.e.interpolate_batch(self.e.pull_back(u, J, d, K).reshape(n, -1), "point-major")

        for _ in range(2):
            out = two_calls()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            out = two_calls()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 3
        fused = self.e.interpolate_pulled_batch(u, J, d, K)
        rel = float((fused - out).abs().max() / out.abs().max())
        del out, fused
        return {"two_calls_ns_per_cell": ms * 1e6 / n, "two_calls": "pull_back + interpolate_batch(point-major) on %d cells, "
                "device-resident" % n, "fused_vs_two_calls_max_rel_difference": rel}

    def e2e_setup(self, avail_bytes):
        import torch

        self.values_host = self.values.cpu().pin_memory()
        self.out_host = torch.empty((self.nc, self.N), dtype=torch.float64).pin_memory()
        self.e2e_units = self.nc
        self.h2d, self.d2h = self.values_host.numel() * 8, self.out_host.numel() * 8
        if self.pulled:
            self.J_host, self.detJ_host, self.K_host = (a.cpu().pin_memory() for a in (self.J, self.detJ, self.Kinv))
            self.h2d += self.nc * 9 * 8
        return {"cells_per_gpu": self.nc}

    def e2e_step(self):
        from basix_b200 import _capi

        if self.pulled:
            _capi.check(_capi.lib().bx_interpolate_pulled_f64(self.e._h, self.values_host.data_ptr(), self.J_host.data_ptr(),
                                                             self.detJ_host.data_ptr(), self.K_host.data_ptr(), self.nc, 3,
                                                             self.out_host.data_ptr(), _capi.BX_MEM_HOST, None))
            return
        _capi.check(_capi.lib().bx_interpolate_f64(self.e._h, 0, self.values_host.data_ptr(), self.nc,
                                                   self.out_host.data_ptr(), _capi.BX_MEM_HOST, None))

    def cpu_setup(self, cores):
        # the reference has no batched interpolate: its documented usage is `i_m * values` per cell
        # (finite-element.h:1183-1216), i.e. one dgemm over a batch of cells with the reference's own matrix
        from oracle import ref

        r = ref.RefElement(*self.ref_args)
        n = int(min(self.nc, 2_000_000))
        self._cpu = (np.ascontiguousarray(r.interpolation_matrix.T), np.random.default_rng(21).standard_normal((n, self.K)))
        if self.pulled:
            n = int(min(n, 500_000))
            rng = np.random.default_rng(22)
            J = np.eye(3)[None] + 0.3 * rng.standard_normal((n, 3, 3))
            self._cpu = (self._cpu[0], rng.standard_normal((n, self.K // 3, 3)), r, J, np.linalg.det(J),
                         np.ascontiguousarray(np.linalg.inv(J)), cores)
            self.cpu_kind = "reference"
            return n, (f"{n} cells per step: the reference's pull_back over {cores} std::threads (oracle/_ref, unmodified "
                       f"reference), then values @ interpolation_matrix.T with numpy (OpenBLAS dgemm)")
        return n, (f"{n} cells per step, values @ interpolation_matrix.T with numpy (OpenBLAS dgemm, all threads), the "
                   f"reference's documented usage of interpolation_matrix()")

    def cpu_step(self):
        if self.pulled:
            Mt, u, r, J, d, K, cores = self._cpu
            U = r.pull_back(u, J, d, K, nthreads=cores)
            self._cpu_out = U.transpose(0, 2, 1).reshape(len(u), -1) @ Mt
            return
        Mt, v = self._cpu
        self._cpu_out = v @ Mt


WORKLOADS = {
    "tabulate_p5_tet": lambda: Tabulate(
        "P5_tetrahedron_gll", (1, 3, 5, 2), 3, 2, 10_000_000, 1,
        "P5 Lagrange (gll_warped) tetrahedron, tabulate(nd=2), 10M points per GPU", "tensor"),
    "tabulate_p1_tri": lambda: Tabulate(
        "P1_triangle_gll", (1, 2, 1, 2), 2, 1, 1_000_000, 1,
        "P1 Lagrange triangle, tabulate(nd=1), 1M points per GPU", "hbm"),
    "tabulate_q4_hex": lambda: Tabulate(
        "Q4_hexahedron_gll", (1, 5, 4, 2), 5, 1, 50_000_000, 5,
        "Q4 Lagrange (gll_warped) hexahedron, tabulate(nd=1), 50M points per GPU streamed as 5 launches of 10M "
        "through one resident 40 GB table", "hbm"),
    "push_forward_n1curl3": lambda: PushForward(
        "N1curl3_tetrahedron_legendre", (3, 3, 3, 11), 20_000_000, 1, False,
        "N1curl degree 3 tetrahedron, covariant Piola push_forward, 20M cells x 1 row"),
    "pull_back_n1curl3": lambda: PushForward(
        "N1curl3_tetrahedron_legendre", (3, 3, 3, 11), 20_000_000, 1, True,
        "N1curl degree 3 tetrahedron, covariant Piola pull_back, 20M cells x 1 row"),
    "push_forward_n1curl3_rows45": lambda: PushForward(
        "N1curl3_tetrahedron_legendre", (3, 3, 3, 11), 4_000_000, 45, False,
        "N1curl degree 3 tetrahedron, covariant Piola push_forward, 4M cells x 45 rows"),
    "pull_back_n1curl3_rows45": lambda: PushForward(
        "N1curl3_tetrahedron_legendre", (3, 3, 3, 11), 4_000_000, 45, True,
        "N1curl degree 3 tetrahedron, covariant Piola pull_back, 4M cells x 45 rows"),
    "t_apply_rt4_block3": lambda: TApply(
        "RT4_tetrahedron_legendre", (2, 3, 4, 11), 20_000_000, 3, False,
        "RT degree 4 tetrahedron, T_apply(block 3) over 20M cells, random cell_info"),
    "geometry_tet": lambda: PhysicalBasis(
        "N1curl3_tetrahedron_legendre", (3, 3, 3, 11), 20_000_000, 1,
        "J, detJ, K of 20M affine tetrahedra from their vertex coordinates", geometry_only=True),
    "geometry_hex": lambda: PhysicalBasis(
        "N1curl3_tetrahedron_legendre", (3, 3, 3, 11), 4_000_000, 1,
        "J, detJ, K of 4M trilinear hexahedra at 8 points each, from node coordinates and the derivatives of the Q1 "
        "coordinate element (general form of bx_geometry)", geometry_only="hex"),
    "physical_basis_n1curl3": lambda: PhysicalBasis(
        "N1curl3_tetrahedron_legendre", (3, 3, 3, 11), 20_000_000, 1,
        "N1curl degree 3 tetrahedron: tabulate-once -> T_apply -> covariant Piola push_forward fused, geometry from vertex "
        "coordinates on chip, 20M cells x 1 point (45 rows)"),
    "physical_basis_rt4": lambda: PhysicalBasis(
        "RT4_tetrahedron_legendre", (2, 3, 4, 11), 10_000_000, 1,
        "RT degree 4 tetrahedron: tabulate-once -> T_apply -> contravariant Piola push_forward fused, geometry from vertex "
        "coordinates on chip, 10M cells x 1 point (70 rows)"),
    "physical_basis_n1curl3_4pts": lambda: PhysicalBasis(
        "N1curl3_tetrahedron_legendre", (3, 3, 3, 11), 5_000_000, 4,
        "N1curl degree 3 tetrahedron: fused physical basis at 4 points per cell (180 rows), 5M cells"),
    "physical_basis_n1curl3_pointwise_4pts": lambda: PhysicalBasis(
        "N1curl3_tetrahedron_legendre", (3, 3, 3, 11), 5_000_000, 4,
        "N1curl degree 3 tetrahedron: fused physical basis at 4 points per cell with one Jacobian per (cell, point) "
        "(non-affine cells), 5M cells", pointwise=True),
    "interpolate_n1curl3": lambda: Interpolate(
        "N1curl3_tetrahedron_legendre", (3, 3, 3, 11), 20_000_000,
        "N1curl degree 3 tetrahedron, interpolation_matrix (45 x 141) applied to 20M cells of point values"),
    "interpolate_pulled_n1curl3": lambda: Interpolate(
        "N1curl3_tetrahedron_legendre", (3, 3, 3, 11), 20_000_000,
        "N1curl degree 3 tetrahedron, covariant Piola pull_back fused into the interpolation (45 x 141 matrix), 20M cells of "
        "physical point values", pulled=True),
    "push_forward_bcast_n1curl3_rows45": lambda: PushForward(
        "N1curl3_tetrahedron_legendre", (3, 3, 3, 11), 8_000_000, 45, False,
        "N1curl degree 3 tetrahedron, covariant Piola push_forward of ONE 45-row slab of reference values onto 8M cells "
        "(push_forward_broadcast)", bcast=True),
    "t_apply_rt4": lambda: TApply(
        "RT4_tetrahedron_legendre", (2, 3, 4, 11), 50_000_000, 1, False,
        "RT degree 4 tetrahedron, T_apply(block 1) over 50M cells, random cell_info"),
    "t_apply_p5_tet": lambda: TApply(
        "P5_tetrahedron_gll", (1, 3, 5, 2), 50_000_000, 1, False,
        "P5 Lagrange tetrahedron, T_apply<double>(block 1) over 50M cells (permutation element), random cell_info"),
    "permute_p5_tet": lambda: TApply(
        "P5_tetrahedron_gll", (1, 3, 5, 2), 50_000_000, 1, True,
        "P5 Lagrange tetrahedron, permute (int32) over 50M cells, random cell_info"),
}
DEFAULT = "tabulate_p5_tet"
# The remaining BASELINE.json configs, reported under "configs" of the default N = 1 line (device-resident, BASELINE sizes):
# config 1; config 3; config 4 with rows = 1 and rows = 45, push and pull; config 5 (i) with block sizes 1 and 3 and (ii).
SIDE_CONFIGS = [
    ("tabulate_p1_tri", {}),
    ("tabulate_q4_hex", {}),
    ("push_forward_n1curl3", {}),
    ("pull_back_n1curl3", {}),
    ("push_forward_n1curl3_rows45", {"nc": 20_000_000}),
    ("pull_back_n1curl3_rows45", {"nc": 20_000_000}),
    ("push_forward_bcast_n1curl3_rows45", {"nc": 20_000_000}),
    ("t_apply_rt4", {}),
    ("t_apply_rt4_block3", {"nc": 50_000_000}),
    ("permute_p5_tet", {}),
    ("t_apply_p5_tet", {}),
    ("interpolate_n1curl3", {}),
    ("interpolate_pulled_n1curl3", {}),
    ("geometry_tet", {}),
    ("geometry_hex", {}),
    ("physical_basis_n1curl3", {}),
    ("physical_basis_rt4", {}),
    ("physical_basis_n1curl3_4pts", {}),
]


# ======================================================================================================
def cpu_reference(w, steps, warmup):
    from oracle import ref

    if not ref.available():
        return None
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    cores = os.cpu_count() or 1
    units, sample = w.cpu_setup(cores)
    for _ in range(warmup):
        w.cpu_step()
    t0 = time.perf_counter()
    for _ in range(steps):
        w.cpu_step()
    dt = (time.perf_counter() - t0) / steps
    res = dict(value=units / dt, unit=w.unit, cores=cores, kind=getattr(w, "cpu_kind", "reference"), sample=sample,
               ms_per_step=dt * 1e3)
    if res["kind"] == "reference" and cores > 1:
        # SURVEY 8d asks for both figures: the reference as shipped (it has no threading of its own) on ONE thread,
        # over a sample 1 / cores as large, and the all-cores figure above
        units1, sample1 = w.cpu_setup(1)
        w.cpu_step()
        t0 = time.perf_counter()
        w.cpu_step()
        res["single_thread"] = {"value": units1 / (time.perf_counter() - t0), "unit": w.unit, "sample": sample1}
    return res


def run_reference(args, w, rank):
    if rank != 0:
        return
    res = cpu_reference(w, max(1, args.steps), max(0, min(args.warmup, 1)))
    if res is None:
        emit({"impl": "reference", "unavailable": "oracle/_ref/libbasix_ref.so missing"})
        return
    emit({
        "impl": "reference", "metric": w.metric, "value": res["value"], "unit": w.unit, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": w.desc, "sample": res["sample"]},
        "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample", "single_thread") if k in res},
        "e2e": {"value": res["value"], "unit": w.unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })


def run_ours(args, w, rank, local_rank, world):
    import psutil
    import torch

    import basix_b200 as bx

    if bx.device_count() < 1:
        raise RuntimeError("bench.py: no CUDA device; the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    bx.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def time_steps(w, steps):
        """(total ms over `steps` steps, kernel launches): CUDA events on the launch stream.  Workloads whose arrays fit
        in L2 are timed step by step with an untimed L2 flush before each step."""
        n0 = bx.kernel_launch_count()
        if getattr(w, "flush_l2", False):
            scratch = torch.empty(FLUSH_BYTES, dtype=torch.uint8, device="cuda")
            evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
            for a, b in evs:
                scratch.fill_(1)
                a.record()
                w.step()
                b.record()
            barrier()
            total_ms = sum(a.elapsed_time(b) for a, b in evs)
            del scratch
        else:
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            for _ in range(steps):
                w.step()
            ev1.record()
            barrier()
            total_ms = ev0.elapsed_time(ev1)
        return total_ms, bx.kernel_launch_count() - n0

    hbm_peak, hbm_src = measured_hbm_peak()

    def roofline_of(w, name, ms_step):
        """Roofline of the dominant kernel (a step is `launches_per_step` launches of that one kernel)."""
        ms_launch = ms_step / w.launches_per_step
        hbm = {"achieved": w.alg_bytes / (ms_launch * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s", "peak_source": hbm_src}
        hbm["frac"] = hbm["achieved"] / hbm_peak
        if w.bound == "tensor":
            ach = w.flops / (ms_launch * 1e-3) / 1e12
            roofline = {"bound": "tensor", "achieved": ach, "peak": FP64_TENSOR_PEAK_TFLOPS, "unit": "TFLOP/s",
                        "frac": ach / FP64_TENSOR_PEAK_TFLOPS, "traffic": None,
                        "peak_source": "FP64 tensor (DMMA m8n8k4) peak measured by tools/peaks.cu on this pool's B200 "
                                       "(profiles/r01_peaks_fp64_hbm_pcie.json); MEASURED_PEAKS.json has no FP64 figure",
                        "hbm": hbm}
        else:
            roofline = dict(bound="hbm", traffic=None, **hbm)
        roofline["traffic"], roofline["traffic_source"] = measured_traffic(name, w.units_per_launch)
        if getattr(w, "flush_l2", False):
            # a launch this short never sees the steady-state copy rate the peak was measured at: time a plain device
            # copy moving the same bytes (half read, half written) under the same rule -- untimed 512 MB L2 flush, one
            # event pair per launch -- as the size-matched ceiling
            src = torch.empty(int(w.alg_bytes) // 2, dtype=torch.uint8, device="cuda")
            dst = torch.empty_like(src)

            class _Copy:
                flush_l2 = True

                @staticmethod
                def step():
                    dst.copy_(src)

            for _ in range(3):
                _Copy.step()
            cms, _ = time_steps(_Copy, 20)
            roofline["copy_same_bytes_ms"] = cms / 20
            roofline["frac_of_copy_same_bytes"] = (cms / 20) / ms_launch
            roofline["copy_same_bytes"] = ("torch copy_ of %d bytes to a second buffer (same bytes read + written as one "
                                           "launch), same L2 flush and event-pair timing" % (int(w.alg_bytes) // 2))
            del src, dst
            # ... and the event pair around an EMPTY launch under the same rule: what the timing itself costs a kernel
            # that lives for ~20 us (reported, never subtracted from `achieved` / `frac`)
            tiny = torch.empty(32, device="cuda")

            class _Empty:
                flush_l2 = True

                @staticmethod
                def step():
                    tiny.fill_(0.0)

            for _ in range(3):
                _Empty.step()
            ems, _ = time_steps(_Empty, 20)
            roofline["event_pair_empty_launch_ms"] = ems / 20
            roofline["frac_net_of_empty_launch"] = (w.alg_bytes / max((ms_launch - ems / 20) * 1e-3, 1e-9) / 1e9) / hbm_peak
        if hasattr(w, "line_bytes"):
            roofline["line_granular"] = {"bytes": w.line_bytes, "frac": w.line_bytes / (ms_launch * 1e-3) / 1e9 / hbm_peak,
                                         "what": w.line_what}
        roofline["ms_per_launch"] = ms_launch
        roofline["launches_per_step"] = w.launches_per_step
        return roofline

    if args.device_only:
        args.no_e2e = True
    w.setup(rank, device_only=args.device_only)
    warmup = max(args.warmup, 3)
    for _ in range(warmup):
        w.step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    total_ms, launches = time_steps(w, args.steps)
    clocks = sampler.stop()
    ms_step = max_over_ranks(total_ms) / args.steps
    value = w.units_per_step * world / (ms_step * 1e-3)
    roofline = roofline_of(w, args.workload, ms_step)
    main_extra = w.extra() if hasattr(w, "extra") else None

    # ---- end to end through the C ABI with host (pinned) buffers ----
    e2e = None
    if not args.no_e2e:
        extra = w.e2e_setup(psutil.virtual_memory().available / max(1, world))
        w.e2e_step()
        barrier()
        e2e_steps = max(1, min(args.steps, 3))
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            w.e2e_step()
        torch.cuda.synchronize()
        dt = max_over_ranks((time.perf_counter() - t0) / e2e_steps)
        e2e = {"value": w.e2e_units * world / dt, "unit": w.unit, "h2d_bytes_per_step": int(w.h2d),
               "d2h_bytes_per_step": int(w.d2h), "steps": e2e_steps, "ms_per_step": dt * 1e3}
        e2e.update(extra)
        # the box's own ceiling for that leg: plain pinned device->host copies (no kernels, no library), all ranks at
        # once, into the same pinned result buffer the e2e step fills.  e2e bytes / ceiling says how much of the e2e time
        # is the bus and how much is this repository's staging.
        host = getattr(w, "out_host", None)
        if host is None:
            host = getattr(w, "data_host", None)
        if isinstance(host, (list, tuple)):  # several result arrays (geometry): the largest one stands for the leg
            host = max(host, key=lambda h: h.numel())
        if host is not None and host.numel() * host.element_size() >= (64 << 20):
            flat = host.view(-1).view(torch.uint8) if host.is_contiguous() else None
        else:
            flat = None
        if flat is not None:
            piece = min(1 << 30, flat.numel())
            npieces = max(1, min(16, flat.numel() // piece))
            src = torch.empty(piece, dtype=torch.uint8, device="cuda")
            flat[:piece].copy_(src, non_blocking=True)
            barrier()
            t0 = time.perf_counter()
            for i in range(npieces):
                flat[i * piece:(i + 1) * piece].copy_(src, non_blocking=True)
            torch.cuda.synchronize()
            dt_c = max_over_ranks(time.perf_counter() - t0)
            ceil = npieces * piece / dt_c / 1e9
            e2e["d2h_ceiling_GBps_per_gpu"] = ceil
            e2e["d2h_ceiling"] = (f"{npieces} x {piece >> 20} MiB cudaMemcpyAsync device->pinned host per rank, {world} rank(s) "
                                  "concurrently, no kernels")
            e2e["d2h_GBps_per_gpu"] = w.d2h / dt / 1e9
            e2e["frac_of_d2h_ceiling"] = (w.d2h / dt / 1e9) / ceil
            del src

    # ---- gather bandwidth (N > 1 only; SURVEY 8e: reported separately, never part of `value`) ----
    gather = None
    if dist is not None:
        from basix_b200.sharding import gather_cells

        rows = (256 << 20) // 448  # 256 MiB per rank of 56-double rows: one derivative slab of a point slice
        part = torch.full((rows, 56), float(rank), dtype=torch.float64, device="cuda")
        full = torch.empty((rows * world, 56), dtype=torch.float64, device="cuda")
        gather_cells(part, rows * world, out=full)  # warm-up (NCCL channel set-up)
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(5):
            gather_cells(part, rows * world, out=full)
        g1.record()
        barrier()
        ms = max_over_ranks(g0.elapsed_time(g1)) / 5
        ok = bool((full[:: rows, 0] == torch.arange(world, dtype=torch.float64, device="cuda")).all())
        gather = {"collective": "all_gather_into_tensor of one 256 MiB result slab per rank straight into the gathered "
                                "table (torch.distributed, NCCL over NVLink; no staging copies)",
                  "bytes_per_rank": int(part.numel() * 8), "ms": ms, "correct": ok,
                  "algbw_GBps": part.numel() * 8 * world / (ms * 1e-3) / 1e9,
                  "busbw_GBps": part.numel() * 8 * (world - 1) / (ms * 1e-3) / 1e9}
        del part, full

    # ---- strong scaling (N > 1): the SAME 10 M points of config 2 cut into contiguous slices ----
    strong = None
    if dist is not None and args.workload == DEFAULT:
        from basix_b200.sharding import shard_bounds

        total = w.npts
        lo, hi = shard_bounds(total, world, rank)
        sw = WORKLOADS[DEFAULT]()
        sw.npts = hi - lo
        for attr in ("x", "x_host", "out", "out_host"):
            if hasattr(w, attr):
                delattr(w, attr)
        torch.cuda.empty_cache()
        sw.setup(rank, device_only=True)
        for _ in range(3):
            sw.step()
        barrier()
        tms, _ = time_steps(sw, args.steps)
        sms = max_over_ranks(tms) / args.steps
        strong = {"workload": f"config 2 with its 10M points cut into {world} contiguous slices (one per GPU)",
                  "points_total": total, "points_per_gpu": hi - lo, "ms_per_step": sms,
                  "value": total * (sw.values_per_step // sw.npts) / (sms * 1e-3), "unit": w.unit, "scaling": "strong"}
        del sw

    # ---- the other BASELINE.json configs (N = 1): device-resident passes at the BASELINE sizes, same timing rules ----
    configs = None
    if world == 1 and not args.no_configs and args.workload == DEFAULT:
        import gc

        for attr in ("x", "x_host", "out", "out_host"):
            if hasattr(w, attr):
                delattr(w, attr)
        configs = {}
        for name, override in SIDE_CONFIGS:
            gc.collect()
            torch.cuda.empty_cache()
            sw = WORKLOADS[name]()
            for k, v in override.items():
                setattr(sw, k, v)
            if override:
                sw.desc += " [%s]" % ", ".join(f"{k}={v}" for k, v in override.items())
            try:
                sw.setup(rank, device_only=True)
                for _ in range(3):
                    sw.step()
                torch.cuda.synchronize()
                ksteps = max(3, min(args.steps, 10))
                if getattr(sw, "flush_l2", False):
                    ksteps = 50  # a 20-microsecond launch timed step by step: many samples, each behind its own L2 flush
                tms, nl = time_steps(sw, ksteps)
                sms = tms / ksteps
                rf = roofline_of(sw, name, sms)
                more = sw.extra() if hasattr(sw, "extra") else None
                configs[name] = {"workload": sw.desc, "ms_per_step": sms, "value": sw.units_per_step / (sms * 1e-3),
                                 "unit": sw.unit, "steps": ksteps, "gpu_launches": int(nl),
                                 "roofline": {k: rf[k] for k in ("bound", "achieved", "peak", "unit", "frac", "traffic",
                                                                 "ms_per_launch", "launches_per_step",
                                                                 "frac_of_copy_same_bytes", "copy_same_bytes_ms",
                                                                 "event_pair_empty_launch_ms", "frac_net_of_empty_launch",
                                                                 "line_granular") if k in rf},
                                 "l2": sw.config.get("l2")}
                if rf["bound"] == "tensor":
                    configs[name]["roofline"]["hbm_frac"] = rf["hbm"]["frac"]
                if more:
                    configs[name].update(more)
                    configs[name]["fused_ns_per_cell"] = sms * 1e6 / sw.units_per_step
            except Exception as ex:  # a side config never takes the headline line down
                configs[name] = {"workload": sw.desc, "error": f"{type(ex).__name__}: {ex}"}
            del sw

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_reference(w, 2, 1)
        if cpu is not None:
            cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample", "single_thread") if k in cpu}

    if rank == 0:
        cfg = dict(w.config)
        cfg["parallelism"] = f"independent slices over {world} GPU(s), no collective"
        if main_extra:
            cfg.update(main_extra)
            cfg["fused_ns_per_cell"] = ms_step * 1e6 / w.units_per_step
        emit({
            "metric": w.metric, "value": value, "unit": w.unit, "n_gpus": world, "steps": args.steps,
            "warmup": warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int32" if getattr(w, "permute", False) else "f64", "data": "synthetic",
            "config": cfg, "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
            "cpu_baseline": cpu, "gather": gather, "strong": strong, "configs": configs,
        })
    if dist is not None:
        dist.destroy_process_group()


_JSON_OUT = None


def emit(line: dict):
    """The one JSON line goes to the process's ORIGINAL stdout; everything else any library prints on fd 1
    (e.g. NCCL's version banner) has been redirected to stderr by main()."""
    os.write(_JSON_OUT, (json.dumps(line) + "\n").encode())


def main():
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT, choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (profiling runs)")
    ap.add_argument("--device-only", action="store_true",
                    help="generate the inputs on the GPU (no host copies of them; implies --no-e2e): profiling runs")
    ap.add_argument("--no-configs", action="store_true",
                    help="skip the device-resident passes over the other BASELINE.json configs (default line, N = 1)")
    ap.add_argument("--units", type=int, default=0,
                    help="override the workload's points / cells per GPU (profiling runs only; the judged line "
                         "uses the BASELINE.json size)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    w = WORKLOADS[args.workload]()
    if args.units > 0:
        if isinstance(w, Tabulate):
            w.npts = args.units
        else:
            w.nc = args.units
        w.desc += f" [size overridden to {args.units} per GPU]"
    if args.impl == "reference":
        run_reference(args, w, rank)
    else:
        run_ours(args, w, rank, local_rank, world)


if __name__ == "__main__":
    main()
