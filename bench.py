#!/usr/bin/env python
"""Benchmark of the basix hot path on B200 (contract: see the task statement / DESIGN.md "Measurement").

  python bench.py --gpus N --steps K --warmup W [--impl ours|reference] [--workload NAME]

A "step" is one pass of the hot path over one batch of synthetic input.  The headline workload is
BASELINE.json configs[1]: P5 Lagrange on the tetrahedron, tabulate(nd=2) at 10M points per GPU
(weak scaling: every rank tabulates its own 10M-point slice, no data-path collective).

`value`   whole-job throughput, inputs resident in HBM, timed with CUDA events on the launch stream.
`e2e`     the same metric through the C ABI with HOST buffers (H2D of x and D2H of the table inside the
          timed region, pinned memory).
`roofline`      the dominant kernel against the measured peak.
`cpu_baseline`  the compiled reference (oracle/_ref) on this box's host cores, bounded sample (rank 0, N=1).
`--impl reference` times only that CPU reference, same metric/config.
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

# workload -> description
WORKLOADS = {
    # name: (element fixture, kind, params)
    "tabulate_p5_tet_nd2_10M": dict(element="P5_tetrahedron_gll", kind="tabulate", cell=3, nd=2, units=10_000_000,
                                    ref_args=(1, 3, 5, 2), unit="values/s",
                                    desc="P5 Lagrange (gll_warped) tetrahedron, tabulate(nd=2), 10M points per GPU"),
    "tabulate_p1_tri_nd1_1M": dict(element="P1_triangle_gll", kind="tabulate", cell=2, nd=1, units=1_000_000,
                                   ref_args=(1, 2, 1, 2), unit="values/s",
                                   desc="P1 Lagrange triangle, tabulate(nd=1), 1M points per GPU"),
}
METRIC = "tabulated basis values/sec (pts x dofs x derivs)"


def random_points(cell, n, seed):
    rng = np.random.default_rng(seed)
    if cell == 3:
        s = np.sort(rng.random((n, 3)), axis=1)
        return np.ascontiguousarray(np.stack([s[:, 0], s[:, 1] - s[:, 0], s[:, 2] - s[:, 1]], axis=1))
    u = rng.random((n, 2))
    flip = u.sum(axis=1) > 1
    u[flip] = 1 - u[flip]
    return np.ascontiguousarray(u)


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
            time.sleep(0.05)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d.get("hbm_gbs", 6650.0)), "MEASURED_PEAKS.json"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ======================================================================================================
def cpu_reference_run(w, steps, warmup, sample_pts=None):
    """Times the compiled reference (oracle/_ref) on all host threads over a bounded sample."""
    from oracle import ref

    if not ref.available():
        return None
    cores = os.cpu_count() or 1
    r = ref.RefElement(*w["ref_args"])
    nd = w["nd"]
    per_pt = r.dim * r.value_size * ref.polyset_nderivs(r.cell_type, nd)
    # ~1.5 s of work per core per step at ~4e7 values/s/core
    if sample_pts is None:
        sample_pts = int(min(w["units"], 2_000_000, max(20_000, 6e7 * cores / per_pt)))
    x = random_points(w["cell"], sample_pts, 42)
    chunk = max(1024, min(65536, sample_pts // (cores * 4) or 1024))
    out = np.empty((ref.polyset_nderivs(r.cell_type, nd), sample_pts, r.dim, r.value_size))
    from oracle.ref import _check, _ptr, lib

    def step():
        _check(lib().bxref_tabulate(r._h, nd, _ptr(x), sample_pts, x.shape[1], _ptr(out), chunk, cores))

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return dict(value=per_pt * sample_pts / dt, unit=w["unit"], cores=cores, kind="reference",
                sample=f"{sample_pts} points per step, {chunk}-point chunks over {cores} std::threads, "
                       f"oracle/_ref (unmodified reference, g++ -O2, scipy OpenBLAS 1 thread per chunk)",
                ms_per_step=dt * 1e3)


def run_reference(args, w, rank, world):
    if rank != 0:
        return
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    res = cpu_reference_run(w, max(1, args.steps), max(0, min(args.warmup, 1)))
    if res is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libbasix_ref.so missing"}))
        return
    line = {
        "impl": "reference", "metric": METRIC, "value": res["value"], "unit": w["unit"], "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": w["desc"], "sample": res["sample"]},
        "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": res["value"], "unit": w["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ======================================================================================================
def run_ours(args, w, rank, local_rank, world):
    import torch

    import basix_b200 as bx
    from basix_b200 import _capi

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

    e = bx.FiniteElement.from_arrays(np.load(os.path.join(GOLDEN, w["element"] + ".npz")))
    nd, npts = w["nd"], w["units"]
    shape = e.tabulate_shape(nd, npts)
    values_per_step = int(np.prod(shape))
    x_host = torch.from_numpy(random_points(w["cell"], npts, 42 + rank)).pin_memory()
    x = x_host.cuda()
    out = torch.empty(shape, dtype=torch.float64, device="cuda")

    def step():
        e.tabulate(nd, x, out=out)

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    n0 = bx.kernel_launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    ev[0].record()
    for i in range(args.steps):
        step()
        ev[i + 1].record()
    barrier()
    launches = bx.kernel_launch_count() - n0
    clocks = sampler.stop()
    ms_total = ev[0].elapsed_time(ev[-1])
    t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    value = values_per_step * world / (ms_step * 1e-3)

    # ---- end to end through the C ABI with host buffers (pinned), H2D + D2H inside the timed region ----
    import psutil

    out_bytes = values_per_step * 8
    avail = psutil.virtual_memory().available / max(1, world)
    e2e_pts = npts
    while e2e_pts * (out_bytes / npts) > 0.4 * avail and e2e_pts > 100_000:
        e2e_pts //= 2
    e2e_shape = e.tabulate_shape(nd, e2e_pts)
    out_host = None if args.no_e2e else torch.empty(e2e_shape, dtype=torch.float64).pin_memory()
    xh = x_host[:e2e_pts]
    del out
    torch.cuda.empty_cache()
    lib = _capi.lib()

    def e2e_step():
        _capi.check(lib.bx_tabulate_f64(e._h, nd, xh.data_ptr(), e2e_pts, xh.shape[1], out_host.data_ptr(),
                                        _capi.BX_MEM_HOST, None))

    e2e_steps = max(1, min(args.steps, 3))
    if args.no_e2e:
        e2e_value = None
    else:
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        torch.cuda.synchronize()
        dt = torch.tensor([(time.perf_counter() - t0) / e2e_steps], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e_value = int(np.prod(e2e_shape)) * world / float(dt.item())

    # ---- roofline of the dominant kernel ----
    hbm_peak, peak_src = measured_peaks()
    path = e.tabulate_path(nd)
    flops = 2.0 * e.dim * e.value_size * e._psize * shape[0] * npts
    bytes_alg = npts * (e._tdim * 8 + values_per_step / npts * 8)
    roofline = {
        "kernel": f"tabulate ({path} path)", "bound": "tensor", "achieved": flops / (ms_step * 1e-3) / 1e12,
        "peak": args.fp64_peak, "unit": "TFLOP/s", "frac": flops / (ms_step * 1e-3) / 1e12 / args.fp64_peak,
        "traffic": None, "peak_source": "nominal B200 FP64 (tensor = vector) 37 TFLOP/s; not in MEASURED_PEAKS.json",
        "hbm": {"achieved": bytes_alg / (ms_step * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "frac": bytes_alg / (ms_step * 1e-3) / 1e9 / hbm_peak, "peak_source": peak_src},
    }

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
        cpu = cpu_reference_run(w, 2, 1)
        if cpu is not None:
            cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": w["unit"], "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": w["desc"], "points_per_gpu": npts, "nd": nd, "tabulate_path": path,
                       "l2": "inputs+outputs larger than L2 (table is %.1f GB per step)" % (out_bytes / 1e9),
                       "parallelism": f"points sharded over {world} GPU(s), no collective"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": w["unit"], "h2d_bytes_per_step": int(e2e_pts * e._tdim * 8),
                    "d2h_bytes_per_step": int(np.prod(e2e_shape)) * 8, "points_per_gpu": e2e_pts,
                    "steps": e2e_steps},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="tabulate_p5_tet_nd2_10M", choices=sorted(WORKLOADS))
    ap.add_argument("--fp64-peak", type=float, default=37.0, help="FP64 peak in TFLOP/s used for roofline.frac")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (profiling runs)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, w, rank, world)
    else:
        run_ours(args, w, rank, local_rank, world)


if __name__ == "__main__":
    main()
