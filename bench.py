#!/usr/bin/env python
"""bench.py -- headline benchmark of the numgrids hot path on B200 (contract: see the task brief).

Headline workload (BASELINE.json configs[4], the one the metric's 1/2/4/8-GPU scaling is quoted
on): fused Cartesian FD Laplacian  Diff(g,2,0)(f)+Diff(g,2,1)(f)+Diff(g,2,2)(f)  on an
equidistant grid with 1024^3 points PER GPU, slab-sharded along the last axis with a halo
exchange (weak scaling: 1024^3, 1024x1024x2048, 1024x2048x2048, 2048^3 at N = 1, 2, 4, 8).
A "step" is one Laplacian apply over the whole (global) grid.  The other BASELINE configs
(cfg1 FD 512^2, cfg2 Chebyshev 256^3, cfg3 FFT 512x512x1024, cfg4 spherical 128x128x256) are
timed at N=1 and reported under "per_config" in the same JSON line.

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference      # CPU arm: the oracle port on the host cores
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Gpts/s per Laplacian apply (fp64), fused Cartesian FD Laplacian, 1024^3 points per GPU"
UNIT = "Gpts/s"

GLOBAL_SHAPES = {1: (1024, 1024, 1024), 2: (1024, 1024, 2048), 4: (1024, 2048, 2048),
                 8: (2048, 2048, 2048)}


def global_shape_for(n_gpus: int, scale: int):
    if n_gpus in GLOBAL_SHAPES:
        g = GLOBAL_SHAPES[n_gpus]
    else:                                           # other N: grow the sharded axis only
        g = (1024, 1024, 1024 * n_gpus)
    return tuple(max(16, s // scale) for s in g)


def workload_config(n_gpus: int, scale: int):
    """The `config` object both arms print (identical by construction, so the driver can see it is the same
    workload): the CPU arm times the same operator on `cpu_sample_shape`, a sub-block of it, and reports a rate."""
    return {"workload": "cfg5: fused Cartesian FD Laplacian Diff(g,2,0)+Diff(g,2,1)+Diff(g,2,2) "
                        "(acc=4), 1024^3 points per GPU, last-axis slabs + halo exchange",
            "global_shape": list(global_shape_for(n_gpus, scale)),
            "cpu_sample_shape": list(cpu_sample_shape(scale, True)),
            "field": "sin(2 pi X) cos(2 pi Y) exp(Z) from global coordinates",
            "l2": "inputs (8 GiB in + 8 GiB out per GPU) are far larger than the 126 MB L2; no flush needed"}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------
# clocks sampling during the timed region
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.FIELDS}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


class NvmlSampler:
    """In-process NVML polling (~1 kHz) of SM clock, power and throttle reasons: the timed region of the
    headline is a few tens of milliseconds, which nvidia-smi's 100 ms loop barely sees."""

    def __init__(self, gpu_index: int):
        self.gpu, self.rows, self.stop_flag, self.thread, self.h = gpu_index, [], False, None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        except Exception:
            self.h = None

    def _run(self):
        nv, h = self.nv, self.h
        while not self.stop_flag:
            try:
                self.rows.append((time.perf_counter(), nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM),
                                  nv.nvmlDeviceGetPowerUsage(h) / 1000.0,
                                  nv.nvmlDeviceGetCurrentClocksEventReasons(h)))
            except Exception:
                break
            time.sleep(0.0005)

    def start(self):
        if self.h is not None:
            self.thread = threading.Thread(target=self._run, daemon=True)
            self.thread.start()

    def stop(self, t0=None, t1=None):
        if self.thread is None:
            return None
        self.stop_flag = True
        self.thread.join(timeout=2)
        rows = [r for r in self.rows if (t0 is None or r[0] >= t0) and (t1 is None or r[0] <= t1)] or self.rows
        if not rows:
            return None
        nv = self.nv
        bits = 0
        for r in rows:
            bits |= int(r[3])
        names = {"sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap, "hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown,
                 "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown}
        sm = [r[1] for r in rows]
        return {"samples": len(rows), "sm_mhz_median": statistics.median(sm), "sm_mhz_min": min(sm),
                "power_w_max": max(r[2] for r in rows), "reasons": sorted(k for k, v in names.items() if bits & v)}


# ---------------------------------------------------------------------------------------------
# CPU arm: the oracle port (numpy restatement of findiff + the reference idiom)
# ---------------------------------------------------------------------------------------------
def cpu_sample_shape(scale: int, threaded: bool = True):
    base = (1024, 512, 512) if threaded else (512, 256, 256)
    return tuple(max(16, s // scale) for s in base)


def _cpu_field(shape):
    hs = [1.0 / 1023.0] * 3
    x, y, z = (np.arange(n) * h for n, h in zip(shape, hs))
    f = (np.sin(2 * np.pi * x)[:, None, None] * np.cos(2 * np.pi * y)[None, :, None]
         * np.exp(z)[None, None, :])
    return f, hs


def time_cpu_oracle(shape, repeats: int, threaded: bool = True):
    """The CPU restatement of the reference path on a bounded sub-block of the headline workload.

    threaded: oracle/pencil_c.c (C/OpenMP, every host thread) -- same arithmetic, bit-identical to
    the NumPy port; else oracle/pencil.py (NumPy, one thread, like the reference itself).
    Returns (best seconds, points, cores used).
    """
    from oracle import pencil
    f, hs = _cpu_field(shape)
    if threaded:
        from oracle import pencil_c
        fn = lambda a: pencil_c.cartesian_laplacian3(a, hs, 4)
        cores = pencil_c.use_all_cores()
    else:
        fn = lambda a: pencil.cartesian_laplacian(a, hs, 4)
        cores = 1
    small = np.ascontiguousarray(f[:32])
    assert np.array_equal(fn(small), pencil.cartesian_laplacian(small, hs, 4))     # same bits as the NumPy port
    best = float("inf")
    for _ in range(repeats):
        t0 = time.perf_counter()
        fn(f)
        best = min(best, time.perf_counter() - t0)
    return best, int(np.prod(shape)), cores


def cpu_baseline_record(scale: int, repeats: int = 3):
    shape = cpu_sample_shape(scale, True)
    t, pts, cores = time_cpu_oracle(shape, repeats, True)
    shape1 = cpu_sample_shape(scale, False)
    t1, pts1, _ = time_cpu_oracle(shape1, 1, False)
    return {"value": pts / t / 1e9, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{shape[0]}x{shape[1]}x{shape[2]} sub-block of the workload, best of {repeats}, C/OpenMP "
                      f"restatement (oracle/pencil_c.c, bit-identical to the NumPy port) on {cores} threads; "
                      f"host has {os.cpu_count()} cores",
            "numpy_1thread": {"value": pts1 / t1 / 1e9, "unit": UNIT,
                              "sample": f"{shape1[0]}x{shape1[1]}x{shape1[2]}, oracle/pencil.py (NumPy, single-threaded "
                                        "like the reference's findiff path)"}}


def run_reference_arm(args):
    """CPU arm: the reference's algorithm for the path on the host cores (every thread it can use)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    scale = args.scale
    shape = cpu_sample_shape(scale, True)
    times = []
    cores = 1
    for _ in range(args.warmup):
        time_cpu_oracle(shape, 1, True)
    for _ in range(args.steps):
        t, pts, cores = time_cpu_oracle(shape, 1, True)
        times.append(t)
    t_step = sum(times) / len(times)
    value = pts / t_step / 1e9
    gshape = global_shape_for(args.gpus, scale)
    shape1 = cpu_sample_shape(scale, False)
    t1, pts1, _ = time_cpu_oracle(shape1, 1, False)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_step * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": workload_config(args.gpus, scale),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"each step = one Laplacian of a {shape[0]}x{shape[1]}x{shape[2]} sub-block of the "
                                   f"workload; C/OpenMP restatement of the reference path (oracle/pencil_c.c, "
                                   f"bit-identical to the NumPy port) on {cores} threads; host has {os.cpu_count()} cores",
                         "numpy_1thread": {"value": pts1 / t1 / 1e9, "unit": UNIT,
                                           "sample": f"{shape1[0]}x{shape1[1]}x{shape1[2]}, oracle/pencil.py (NumPy, "
                                                     "single-threaded like the reference's findiff path)"}},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def make_local_field(torch, gshape, z0, z1, device):
    """f = sin(2 pi X) cos(2 pi Y) exp(Z) from GLOBAL coordinates (no global meshgrid)."""
    xs = [torch.linspace(0, 1, n, dtype=torch.float64, device=device) for n in gshape]
    fx = torch.sin(2 * np.pi * xs[0])[:, None, None]
    fy = torch.cos(2 * np.pi * xs[1])[None, :, None]
    fz = torch.exp(xs[2][z0:z1])[None, None, :]
    return (fx * fy * fz).contiguous()


def parity_check(torch, gshape, hs, z0, z1, out, dev):
    """The timed output against the oracle: three 24-plane x blocks (both global x boundaries and the middle)
    of this rank's slab, full y and z extent, bit for bit.  The oracle (oracle/pencil_c.c, the C restatement
    pinned to the NumPy one) runs on the same input values -- the field regenerated on the device over the
    block's planes +- 8 and the slab's z range +- 8 (the neighbour ranks' columns), copied to the host."""
    from oracle import pencil_c
    pencil_c.use_all_cores()
    nx, Z = gshape[0], gshape[2]
    ze0, ze1 = max(z0 - 8, 0), min(z1 + 8, Z)
    xs = [torch.linspace(0, 1, n, dtype=torch.float64, device=dev) for n in gshape]
    fx = torch.sin(2 * np.pi * xs[0])[:, None, None]
    fy = torch.cos(2 * np.pi * xs[1])[None, :, None]
    fz = torch.exp(xs[2][ze0:ze1])[None, None, :]
    worst, equal, pts = 0.0, True, 0
    mid = max(nx // 2 - 12, 0)
    for a, b in ((0, min(24, nx)), (mid, min(mid + 24, nx)), (max(nx - 24, 0), nx)):
        lo, hi = max(a - 8, 0), min(b + 8, nx)
        blk = (fx[lo:hi] * fy * fz).contiguous().cpu().numpy()
        ref = pencil_c.cartesian_laplacian3(blk, hs, 4)[a - lo:b - lo, :, z0 - ze0:z0 - ze0 + (z1 - z0)]
        got = out[a:b].cpu().numpy()
        equal = equal and bool(np.array_equal(got, ref))
        den = float(np.max(np.abs(ref))) or 1.0
        worst = max(worst, float(np.max(np.abs(got - ref))) / den)
        pts += got.size
    return equal, worst, pts


def bind_to_gpu_numa_node(gpu_index: int):
    """Run this process (and first-touch its pinned buffers) on the CPUs NVML reports as local to the GPU."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"{len(cpus)} CPUs local to GPU {gpu_index}"
    except Exception as e:  # pragma: no cover
        return f"not bound ({type(e).__name__})"
    return "not bound"


def time_events(torch, fn, steps, warmup):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in ev:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ts = [a.elapsed_time(b) for a, b in ev]
    return ts


def _lib_last_kernel():
    from numgrids_b200 import _lib
    return _lib.last_kernel()


def per_config_extras(torch, ng, peak_gbs, quick):
    """cfg1-cfg4 at N=1: device-resident timings (median over a few launches)."""
    A = ng.AxisType
    out = {}
    dev = "cuda"

    def med(ts):
        return statistics.median(ts)

    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)   # 256 MiB > L2

    def timed(fn, steps=10, warmup=3, flush_l2=False):
        def wrapped():
            fn()
        for _ in range(warmup):
            wrapped()
        torch.cuda.synchronize()
        ts = []
        for _ in range(steps):
            if flush_l2:
                flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            wrapped()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return med(ts)

    try:
        # cfg1: Diff(Grid(eq512, eq512), 1, 0); launch-bound -> also graph replay over rotating buffers
        ax = ng.create_axis(A.EQUIDISTANT, 512, 0.0, 1.0)
        g = ng.Grid(ax, ax)
        op = ng.Diff(g, 1, 0).operator.operator
        nbuf = 80                                                     # 80 x 2 x 2 MiB = 320 MiB > L2
        ins = torch.randn(nbuf, 512, 512, dtype=torch.float64, device=dev)
        outs = torch.empty_like(ins)
        ms_single = timed(lambda: op.apply_ptr(ins[0].data_ptr(), outs[0].data_ptr()), 20, 5, True)
        graph = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for i in range(nbuf):
                op.apply_ptr(ins[i].data_ptr(), outs[i].data_ptr(), stream=s.cuda_stream)
            torch.cuda.synchronize()
            with torch.cuda.graph(graph, stream=s):
                for rep in range(5):
                    for i in range(nbuf):
                        op.apply_ptr(ins[i].data_ptr(), outs[i].data_ptr(), stream=s.cuda_stream)
        torch.cuda.current_stream().wait_stream(s)
        ms_graph = timed(graph.replay, 5, 2) / (5 * nbuf)
        # the same 400 independent applies as 8 parallel branches of one graph (fork/join inside the
        # capture): the applies of different fields overlap, so this is the kernel's throughput when the
        # launch latency of a 4 MiB job is hidden -- what a user differentiating many fields gets
        nbranch = 8
        graph8 = torch.cuda.CUDAGraph()
        side = [torch.cuda.Stream() for _ in range(nbranch - 1)]
        with torch.cuda.stream(s):
            with torch.cuda.graph(graph8, stream=s):
                for b_ in side:
                    b_.wait_stream(s)
                lanes = [s] + side
                for rep in range(5):
                    for i in range(nbuf):
                        st_ = lanes[i % nbranch]
                        op.apply_ptr(ins[i].data_ptr(), outs[i].data_ptr(), stream=st_.cuda_stream)
                for b_ in side:
                    s.wait_stream(b_)
        torch.cuda.current_stream().wait_stream(s)
        ms_graph8 = timed(graph8.replay, 5, 2) / (5 * nbuf)
        # the 80 fields as one stack in ONE launch (ng_apply_batch): the streaming kernel at its large-array rate
        from numgrids_b200 import _lib as _L
        lib_ = _L.load()
        st_ptr = torch.cuda.current_stream().cuda_stream
        ms_batch = timed(lambda: _L.check(lib_.ng_apply_batch(op.plan.handle, ins.data_ptr(), outs.data_ptr(), nbuf, st_ptr)),
                         10, 3, True) / nbuf
        pts = 512 * 512
        out["cfg1_fd_512x512_order1_axis0"] = {
            "gpts_batched_80_fields_one_launch": pts / ms_batch / 1e6, "ms_per_apply_batched": ms_batch,
            "hbm_frac_batched": 16 * pts / ms_batch / 1e6 / peak_gbs,
            "gpts_single_launch_cold_l2": pts / ms_single / 1e6, "ms_single_launch": ms_single,
            "gpts_graph_replay_rotating_320MiB": pts / ms_graph / 1e6, "ms_per_apply_graph": ms_graph,
            "hbm_frac_graph": 16 * pts / ms_graph / 1e6 / peak_gbs,
            "gpts_graph_replay_8_branches": pts / ms_graph8 / 1e6, "ms_per_apply_graph_8_branches": ms_graph8,
            "hbm_frac_graph_8_branches": 16 * pts / ms_graph8 / 1e6 / peak_gbs}
        del ins, outs, graph, graph8
    except Exception as e:  # pragma: no cover
        out["cfg1_fd_512x512_order1_axis0"] = {"error": repr(e)}

    try:
        # the per-axis finite-difference Diff itself at cfg5's size: Diff(g, 2, axis) on equidistant 1024^3
        # (streaming kernels: register-window march for axes 0/1, shuffle-tap row kernel for axis 2), 16 B/pt
        nfd = 256 if quick else 1024
        ax = ng.create_axis(A.EQUIDISTANT, nfd, 0.0, 1.0)
        g = ng.Grid(ax, ax, ax)
        f = torch.rand(nfd, nfd, nfd, dtype=torch.float64, device=dev)
        o = torch.empty_like(f)
        res = {}
        for order, axis in ((2, 0), (2, 1), (2, 2), (1, 0), (1, 2)):
            op = ng.Diff(g, order, axis).operator.operator
            ms = timed(lambda: op.apply_ptr(f.data_ptr(), o.data_ptr()), 10, 3, nfd < 512)
            res[f"order{order}_axis{axis}"] = {"gpts": nfd ** 3 / ms / 1e6, "ms": ms,
                                               "hbm_frac": 16 * nfd ** 3 / ms / 1e6 / peak_gbs}
        out[f"fd_diff_{nfd}cubed"] = res
        del f, o
    except Exception as e:  # pragma: no cover
        out["fd_diff"] = {"error": repr(e)}

    try:
        # cfg5 in the reference's second idiom: CurvilinearGrid(..., scale_factors = ones).laplacian(f)
        # = sum_i D_i(D_i f) with first-derivative stencils: the single-pass kernel (csrc/curvlap3d.cu; `inner`
        # never leaves the SM), at 512^3 because the reference API evaluates the scale factors on the full
        # meshgrid (24 GiB of host arrays at 1024^3)
        nii = 128 if quick else 512
        ax = ng.create_axis(A.EQUIDISTANT, nii, 0.0, 1.0)
        one = lambda c: np.ones_like(c[0])
        gc = ng.CurvilinearGrid(ax, ax, ax, scale_factors=(one, one, one))
        f = torch.rand(nii, nii, nii, dtype=torch.float64, device=dev)
        ms = timed(lambda: gc.laplacian(f), 5, 2, nii < 512)
        out[f"cfg5_idiom_ii_unit_curvilinear_laplacian_{nii}cubed"] = {
            "gpts": nii ** 3 / ms / 1e6, "ms": ms, "passes": 1, "kernel": _lib_last_kernel(),
            "hbm_frac_algorithmic_16B": 16 * nii ** 3 / ms / 1e6 / peak_gbs}
        del f, gc
    except Exception as e:  # pragma: no cover
        out["cfg5_idiom_ii"] = {"error": repr(e)}

    try:
        # cfg2: Diff(g, 2, 0) on Chebyshev 256^3 -- FP64-pipe-bound ordered DGEMM (2*256 instr/pt)
        n = 128 if quick else 256
        ax = ng.create_axis(A.CHEBYSHEV, n, 0.0, 1.0)
        g = ng.Grid(ax, ax, ax)
        f = torch.randn(n, n, n, dtype=torch.float64, device=dev)
        o = torch.empty_like(f)
        res = {}
        for order, axis in ((2, 0), (2, 2), (1, 1)):
            op = ng.Diff(g, order, axis).operator.operator
            ms = timed(lambda: op.apply_ptr(f.data_ptr(), o.data_ptr()), 10, 3, True)
            instr = 2.0 * n * n ** 3
            res[f"order{order}_axis{axis}"] = {
                "gpts": n ** 3 / ms / 1e6, "ms": ms, "fp64_instr_per_s_T": instr / ms / 1e9,
                "fp64_pipe_frac_at_1965MHz": instr / ms / 1e9 / (148 * 64 * 1.965e-3)}
        out[f"cfg2_chebyshev_{n}cubed"] = res
        del f, o
    except Exception as e:  # pragma: no cover
        out["cfg2_chebyshev"] = {"error": repr(e)}

    try:
        # cfg3: Diff(g, 1, 2), periodic 1024 last axis, batch 512x512
        b = 128 if quick else 512
        e_ax = ng.create_axis(A.EQUIDISTANT, b, 0.0, 1.0)
        p_ax = ng.create_axis(A.EQUIDISTANT_PERIODIC, 1024, 0.0, 2 * np.pi)
        g = ng.Grid(e_ax, e_ax, p_ax)
        f = torch.randn(b, b, 1024, dtype=torch.float64, device=dev)
        o = torch.empty_like(f)
        op = ng.Diff(g, 1, 2).operator.operator
        ms = timed(lambda: op.apply_ptr(f.data_ptr(), o.data_ptr()), 10, 3, b < 512)
        pts = b * b * 1024
        out[f"cfg3_fft_{b}x{b}x1024_order1_axis2"] = {
            "gpts": pts / ms / 1e6, "ms": ms, "hbm_frac": 16 * pts / ms / 1e6 / peak_gbs}
        del f, o
    except Exception as e:  # pragma: no cover
        out["cfg3_fft"] = {"error": repr(e)}

    try:
        # the same operator along an axis that is NOT the contiguous one (cylindrical phi in the middle, doubly
        # periodic boxes): TMA tile kernel (csrc/fft_tile.cu)
        b = 128 if quick else 512
        res = {}
        for shape, axis in (((b, 1024, b), 1), ((2 * b, 256, 2 * b), 1)):
            axes = [ng.create_axis(A.EQUIDISTANT_PERIODIC if a == axis else A.EQUIDISTANT, m, 0.0,
                                   2 * np.pi if a == axis else 1.0) for a, m in enumerate(shape)]
            g = ng.Grid(*axes)
            f = torch.randn(*shape, dtype=torch.float64, device=dev)
            o = torch.empty_like(f)
            op = ng.Diff(g, 1, axis).operator.operator
            ms = timed(lambda: op.apply_ptr(f.data_ptr(), o.data_ptr()), 10, 3, b < 512)
            pts = int(np.prod(shape))
            res["x".join(map(str, shape)) + f"_axis{axis}"] = {"gpts": pts / ms / 1e6, "ms": ms,
                                                               "hbm_frac": 16 * pts / ms / 1e6 / peak_gbs}
            del f, o
        out["fft_strided_axis_order1"] = res
    except Exception as e:  # pragma: no cover
        out["fft_strided_axis"] = {"error": repr(e)}

    try:
        # periodic axes whose length is not a power of two: zero-padded transform on the register kernels (plain first
        # derivative, n <= 512), native mixed-radix transform (csrc/fft_mixed.cu) beyond that and for higher orders
        b = 128 if quick else 512
        res = {}
        for n, order in ((360, 1), (360, 3), (1000, 1), (1536, 1)):
            shape = (b, b if n < 1200 else b // 2, n)
            axes = [ng.create_axis(A.EQUIDISTANT_PERIODIC if a == 2 else A.EQUIDISTANT, m, 0.0,
                                   2 * np.pi if a == 2 else 1.0) for a, m in enumerate(shape)]
            g = ng.Grid(*axes)
            f = torch.randn(*shape, dtype=torch.float64, device=dev)
            o = torch.empty_like(f)
            op = ng.Diff(g, order, 2).operator.operator
            ms = timed(lambda: op.apply_ptr(f.data_ptr(), o.data_ptr()), 10, 3, b < 512)
            pts = int(np.prod(shape))
            res["x".join(map(str, shape)) + f"_order{order}"] = {"gpts": pts / ms / 1e6, "ms": ms, "kernel": _lib_last_kernel(),
                                                                  "hbm_frac": 16 * pts / ms / 1e6 / peak_gbs}
            del f, o
        out["fft_any_length_axis2"] = res
    except Exception as e:  # pragma: no cover
        out["fft_any_length"] = {"error": repr(e)}

    try:
        # cfg4: SphericalGrid(Cheb 128, Cheb 128, periodic 256): laplacian / gradient / divergence
        r = ng.create_axis(A.CHEBYSHEV, 128, 0.5, 2.0)
        th = ng.create_axis(A.CHEBYSHEV, 128, 0.3, np.pi - 0.3)
        ph = ng.create_axis(A.EQUIDISTANT_PERIODIC, 256, 0.0, 2 * np.pi)
        g = ng.SphericalGrid(r, th, ph)
        R, T, P = (torch.from_numpy(c).cuda() for c in g.meshed_coords)
        f = R ** 2 * torch.sin(T) ** 2 * torch.cos(2 * P) + torch.exp(-R) * torch.cos(T)
        v = (R.contiguous(), torch.sin(T).contiguous(), torch.cos(P).contiguous())
        pts = 128 * 128 * 256
        res = {}
        # FP64-pipe-bound: a Chebyshev apply issues 2*128 DMUL/DADD per point, the n = 256
        # FFT apply ~28 FP64 instructions per point; laplacian and curl = 4 Chebyshev + 2 FFT applies,
        # gradient and divergence = 2 + 1
        cheb, fft = 2.0 * 128, 28.0
        work = {"laplacian": 4 * cheb + 2 * fft, "gradient": 2 * cheb + fft,
                "divergence": 2 * cheb + fft, "curl": 4 * cheb + 2 * fft}
        for name, fn in (("laplacian", lambda: g.laplacian(f)), ("gradient", lambda: g.gradient(f)),
                         ("divergence", lambda: g.divergence(*v)), ("curl", lambda: g.curl(*v))):
            ms = timed(fn, 10, 3, True)
            res[name] = {"gpts": pts / ms / 1e6, "ms": ms, "fp64_instr_per_point": work[name],
                         "fp64_pipe_frac_at_1965MHz": work[name] * pts / ms / 1e9 / (148 * 64 * 1.965e-3)}
        out["cfg4_spherical_128x128x256"] = res
    except Exception as e:  # pragma: no cover
        out["cfg4_spherical"] = {"error": repr(e)}
    return out


def multi_gpu_extras(torch, dist, ng, rank, world, dev, quick):
    """N > 1: the other sharded paths of SURVEY 8(e), device-resident, max over ranks of the per-apply time.
    (a) FFT along the SHARDED axis (slab -> pencil redistribution over NVLink peer memory, apply, and back):
        cfg3 per GPU, global grid 512N x 512 x 1024, slabs 512N x 512 x 1024/N, pencils 512 x 512 x 1024;
    (b) spherical Laplacian on phi slabs: Chebyshev r, theta local; FFT phi through the redistribution."""
    from numgrids_b200.slab import SlabCurvilinear, SlabDiff
    A = ng.AxisType
    out = {}

    def timed(fn, steps=5, warmup=2):
        for _ in range(warmup):
            fn()
        dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            fn()
        b.record()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b) / steps], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    try:
        b_ = 128 if quick else 512
        nper = 256 if quick else 1024
        g = ng.Grid(ng.create_axis(A.EQUIDISTANT, b_ * world, 0.0, 1.0), ng.create_axis(A.EQUIDISTANT, b_, 0.0, 1.0),
                    ng.create_axis(A.EQUIDISTANT_PERIODIC, nper, 0.0, 2 * np.pi))
        d = SlabDiff(g, 1, 2, rank=rank, nranks=world)
        f = torch.randn(*d.local_shape, dtype=torch.float64, device=dev)
        o = torch.empty_like(f)
        ms = timed(lambda: d(f, out=o))
        pts = int(np.prod(g.shape))
        out["fft_on_sharded_axis"] = {
            "global_shape": list(g.shape), "local_slab": list(d.local_shape), "ms": ms, "gpts": pts / ms / 1e6,
            "exchange": "NVLink peer memory (symmetric)" if d._symm is not None and not d._symm_failed else "all_to_all_single"}
        del f, o, d
        torch.cuda.empty_cache()
    except Exception as e:  # pragma: no cover
        out["fft_on_sharded_axis"] = {"error": repr(e)}
    try:
        nr = 64 if quick else 128
        g = ng.SphericalGrid(ng.create_axis(A.CHEBYSHEV, nr, 0.5, 2.0), ng.create_axis(A.CHEBYSHEV, nr, 0.3, np.pi - 0.3),
                             ng.create_axis(A.EQUIDISTANT_PERIODIC, 256 * world, 0.0, 2 * np.pi))
        sc = SlabCurvilinear(g, rank=rank, nranks=world)
        f = torch.randn(*sc.local_shape, dtype=torch.float64, device=dev)
        ms = timed(lambda: sc.laplacian(f))
        pts = int(np.prod(g.shape))
        out["spherical_laplacian_phi_slabs"] = {"global_shape": list(g.shape), "local_slab": list(sc.local_shape),
                                                "ms": ms, "gpts": pts / ms / 1e6}
    except Exception as e:  # pragma: no cover
        out["spherical_laplacian_phi_slabs"] = {"error": repr(e)}
    return out


def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    import numgrids_b200 as ng
    from numgrids_b200 import _lib
    from numgrids_b200.slab import SlabLaplacian, neighbours, exchange_halos

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (numgrids_b200 has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n_gpus = world
    scale = args.scale
    gshape = global_shape_for(n_gpus, scale)
    hs = [1.0 / (n - 1) for n in gshape]
    # spacing with the reference's bits: coords[1]-coords[0] of linspace(0,1,n)
    hs = [float(np.linspace(0, 1, n)[1] - np.linspace(0, 1, n)[0]) for n in gshape]
    slab = SlabLaplacian(gshape, hs, rank=rank, nranks=world, acc=4)
    f = make_local_field(torch, gshape, slab.z0, slab.z1, dev)
    out = torch.empty_like(f)
    local_pts = f.numel()
    local_shape = list(f.shape)
    global_pts = int(np.prod(gshape))
    peak_gbs, peak_src = measured_peaks()
    lo_n, hi_n = neighbours(rank, world)

    kernel_ms = []
    stencil_ev = []      # N > 1 with overlap: one dict of CUDA events per step (slab.py)

    def step(record=False):
        if world == 1:
            if record:
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
            slab.apply_local(f, out)
            if record:
                b.record()
                kernel_ms.append((a, b))
            return
        slab.stencil_events = stencil_ev if (record == "parts" and not args.no_overlap) else None
        if record and args.no_overlap:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
        slab(f, out=out, overlap=not args.no_overlap)       # pack -> exchange (overlapped) -> stencil
        if record and args.no_overlap:
            b.record()
            kernel_ms.append((a, b))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    nvml = NvmlSampler(local_rank)
    if rank == 0:
        sampler.start()
        nvml.start()
        time.sleep(0.15)
    launches0 = _lib.launch_count()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_host0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        step(record=True)
    e1.record()
    barrier()
    t_host1 = time.perf_counter()
    launches = _lib.launch_count() - launches0
    elapsed_ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    if rank == 0:
        clocks["nvml_1khz"] = nvml.stop(t_host0, t_host1)
    if world > 1 and not args.no_overlap:
        # the timed steps above ran the product path (ng_slab_lap_apply: the whole step behind the C ABI); the per-part
        # times come from a few extra steps of the same sequence driven from Python with CUDA events between the parts
        for _ in range(min(args.steps, 5)):
            step(record="parts")
        barrier()
        slab.stencil_events = None
    k_ms = [a.elapsed_time(b) for a, b in kernel_ms]
    parts = None
    if stencil_ev:
        # the stencil of a step spans from the start of the interior launch to the join of the two streams
        k_ms += [e["int0"].elapsed_time(e["join"]) for e in stencil_ev]       # (pack and signals are not the stencil)
        mean = lambda a, b: statistics.mean(e[a].elapsed_time(e[b]) for e in stencil_ev)
        parts = {"interior_ms": mean("int0", "int1"), "boundary_ms": mean("edge0", "edge1"),
                 "stencil_span_ms": mean("int0", "join")}
        if "pack0" in stencil_ev[0]:
            parts["pack_ms"] = mean("pack0", "pack1")
            parts["signal_wait_ms"] = mean("pack1", "edge0")
            parts["step_span_ms"] = mean("pack0", "join")
    t = torch.tensor([elapsed_ms, statistics.mean(k_ms), float(launches)], dtype=torch.float64, device=dev)
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        elapsed_ms, kern_ms, launches = float(tmax[0]), float(tmax[1]), int(tsum[2])
        if parts is not None:                       # per-part times: max over ranks
            keys = sorted(parts)
            pt = torch.tensor([parts[k] for k in keys], dtype=torch.float64, device=dev)
            dist.all_reduce(pt, op=dist.ReduceOp.MAX)
            parts = {k: float(v) for k, v in zip(keys, pt)}
    else:
        kern_ms = float(t[1])
    ms_per_step = elapsed_ms / args.steps
    value = global_pts / ms_per_step / 1e6

    # ---- sustained number (N = 1): >= 2 s of back-to-back applies; the part settles at its 1000 W cap
    sustained = None
    if world == 1 and not args.no_extras:
        nrun = 640
        sampler2 = NvmlSampler(local_rank)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(nrun + 1)]
        torch.cuda.synchronize()
        sampler2.start()
        ts0 = time.perf_counter()
        ev[0].record()
        for i in range(nrun):
            slab.apply_local(f, out)
            ev[i + 1].record()
        torch.cuda.synchronize()
        ts1 = time.perf_counter()
        tail = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(nrun // 2, nrun))
        ms_s = tail[len(tail) // 2]
        sustained = {"applies": nrun, "seconds": ts1 - ts0, "ms_per_apply_median_second_half": ms_s,
                     "gpts": local_pts / ms_s / 1e6, "hbm_frac": 16.0 * local_pts / ms_s / 1e6 / peak_gbs,
                     "clocks_nvml_1khz": sampler2.stop(ts0 + 0.5 * (ts1 - ts0), ts1)}

    # ---- parity: every rank checks the array it just timed against the oracle (bit for bit)
    try:
        p_equal, p_err, p_pts = parity_check(torch, gshape, hs, slab.z0, slab.z1, out, dev)
        pv = torch.tensor([0.0 if p_equal else 1.0, p_err, float(p_pts)], dtype=torch.float64, device=dev)
        if world > 1:
            pmax = pv.clone()
            dist.all_reduce(pmax, op=dist.ReduceOp.MAX)
            psum = pv.clone()
            dist.all_reduce(psum, op=dist.ReduceOp.SUM)
            pv = torch.stack([pmax[0], pmax[1], psum[2]])
        parity = {"parity_checked": True, "bit_exact": bool(pv[0] == 0.0), "max_rel_linf": float(pv[1]),
                  "points_compared": int(pv[2]),
                  "sample": "per rank: x planes [0,24), [nx/2-12,nx/2+12), [nx-24,nx) of the timed output, full y/z "
                            "extent, vs oracle/pencil_c.c (C restatement pinned to the NumPy one) on the same input"}
    except Exception as e:  # pragma: no cover
        parity = {"parity_checked": False, "error": repr(e)}

    # ---- e2e: host buffers in, host buffers out, copies inside the timed region
    e2e = None
    try:
        e2e_steps = max(1, min(args.e2e_steps, args.steps))
        numa = bind_to_gpu_numa_node(local_rank)          # pinned buffers on the GPU's own NUMA node (first touch)
        h_in = torch.empty(f.shape, dtype=torch.float64, pin_memory=True)
        h_out = torch.empty(f.shape, dtype=torch.float64, pin_memory=True)
        h_in.copy_(f)
        h_out.zero_()
        if world == 1:
            lap = _lib.load()

            def e2e_step():     # the C-ABI host call: H2D + kernel + D2H, synchronous
                _lib.check(lap.ng_fdlap_apply_host(slab.plan.handle, h_in.data_ptr(), h_out.data_ptr()))
        else:
            slab.stencil_events = None

            def e2e_step():     # chunked H2D || halo exchange || stencil || D2H per rank (slab.py)
                slab.apply_host(h_in, h_out)
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        barrier()
        dt = (time.perf_counter() - t0) / e2e_steps
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt[0])
        ok = bool(torch.equal(h_out.to(dev), out))
        e2e = {"value": global_pts / dt / 1e9, "unit": UNIT, "h2d_bytes_per_step": 8 * local_pts * world,
               "d2h_bytes_per_step": 8 * local_pts * world, "ms_per_step": dt * 1e3, "steps": e2e_steps,
               "matches_device_result": ok,
               "path": "ng_fdlap_apply_host (C ABI, pinned host buffers, chunked 3-stream pipeline)" if world == 1 else
                       "SlabLaplacian.apply_host per rank: chunked H2D || per-chunk halo exchange || stencil || D2H",
               "host_numa_binding": numa}
        del h_in, h_out
    except Exception as e:  # pragma: no cover
        e2e = {"value": None, "unit": UNIT, "error": repr(e), "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}

    multi = None
    if world > 1 and not args.no_extras:
        out = None
        if getattr(slab, "_host_state", None):
            slab._host_state = None
        torch.cuda.empty_cache()
        multi = multi_gpu_extras(torch, dist, ng, rank, world, dev, quick=scale > 1)

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (fdlap3d_march_kernel): 16 B/point algorithmic
    alg_bytes = 16.0 * local_pts
    achieved = alg_bytes / (kern_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic_fdlap3d.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            if tj.get("local_shape") == list(f.shape):
                traffic = tj.get("dram_bytes_per_launch")
        except Exception:
            pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
                "frac": achieved / peak_gbs, "traffic": traffic, "peak_source": peak_src,
                "kernel": "fdlap3d_ring2_kernel<2,4,8> (plane ring, TMA)", "kernel_ms": kern_ms,
                "algorithmic_bytes_per_launch": alg_bytes,
                "fp64_instr_per_point": 32,
                "fp64_pipe_frac_at_1965MHz": 32.0 * local_pts / (kern_ms * 1e-3) / (148 * 64 * 1.965e9)}

    # ---- CPU baseline (rank 0, N=1 only): the oracle port on a bounded sample
    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        cpu_baseline = cpu_baseline_record(scale)

    extras = None
    if world == 1 and not args.no_extras:
        out = None
        torch.cuda.empty_cache()
        extras = per_config_extras(torch, ng, peak_gbs, quick=scale > 1)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(n_gpus, scale),
        "details": {"local_shape": local_shape,
                    "halo_exchange": ("none (1 GPU)" if world == 1 else
                                      ("ng_slab_lap_apply (C ABI): " if getattr(slab, "_cabi_plan", None) is not None else
                                       "slab.py (per-part events): ") +
                                      ("NVLink peer stores from the pack kernel (symmetric memory) + per-neighbour flags; "
                                       "interior z tiles on the main stream, boundary z tiles behind the flags on a side stream"
                                       if getattr(slab, "_symm", None) is not None or getattr(slab, "_cabi_plan", None) is not None
                                       else "NCCL send/recv: " + str(getattr(slab, "_symm_error", "disabled")) + " / "
                                            + str(getattr(slab, "_cabi_error", ""))))},
        "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": launches,
        "clocks": clocks, "parity": parity, "parity_checked": bool(parity.get("parity_checked") and parity.get("bit_exact")),
        "step_parts": parts, "sustained": sustained, "per_config": extras if extras is not None else multi,
    }
    print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scale", type=int, default=int(os.environ.get("NG_BENCH_SCALE", "1")),
                    help="divide every extent by this (debug only; 1 = the named workload)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--no-overlap", action="store_true", help="serialise halo exchange and stencil (N > 1)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
