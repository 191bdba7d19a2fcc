"""Interleaved A/B timing of the fused Laplacian kernel variants on ONE GPU, with the SM clock sampled
through NVML while each variant runs (a part at its power cap clocks down: compare like with like).

    python tools/ab_lap.py [nx ny nz] [--rounds N]

Variants: whole array without halos (ring slots 6 / 8 / 10, x chunks 32 / 64 / 128), the slab step of the
multi-GPU path on the same shape (interior launch + boundary launch with both halos, serial and on two
streams), each part alone.  Every round runs every variant once (5 applies each, after 2 warm-ups) in a
rotated order; the table reports the median and minimum per apply over the rounds.
"""
import os
import statistics
import sys
import threading
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from numgrids_b200 import _lib  # noqa: E402
from numgrids_b200.slab import SlabLaplacian  # noqa: E402

args = [a for i, a in enumerate(sys.argv[1:], 1) if not a.startswith("--") and sys.argv[i - 1] not in ("--rounds", "--only")]
rounds = int(sys.argv[sys.argv.index("--rounds") + 1]) if "--rounds" in sys.argv else 6
burst = "--burst" in sys.argv
only = sys.argv[sys.argv.index("--only") + 1].split(",") if "--only" in sys.argv else None
shape = tuple(int(s) for s in (args[:3] if len(args) >= 3 else (1024, 1024, 1024)))
hs = [1.0 / (n - 1) for n in shape]
lap = SlabLaplacian(shape, hs)
xs = [torch.linspace(0, 1, n, dtype=torch.float64, device="cuda") for n in shape]
f = (torch.sin(2 * np.pi * xs[0])[:, None, None] * torch.cos(2 * np.pi * xs[1])[None, :, None]
     * torch.exp(xs[2])[None, None, :]).contiguous()
out = torch.empty_like(f)
rows = shape[0] * shape[1]
lo = torch.randn(rows, 2, dtype=torch.float64, device="cuda")
hi = torch.randn(rows, 2, dtype=torch.float64, device="cuda")
side = torch.cuda.Stream()
pts = f.numel()


class Clock:
    def __init__(self):
        import pynvml
        pynvml.nvmlInit()
        self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(torch.cuda.current_device())
        self.rows, self.on = [], False

    def __enter__(self):
        self.rows, self.on = [], True
        self.t = threading.Thread(target=self.run, daemon=True)
        self.t.start()
        return self

    def run(self):
        while self.on:
            self.rows.append((self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM),
                              self.nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0))
            time.sleep(0.001)

    def __exit__(self, *a):
        self.on = False
        self.t.join()

    def summary(self):
        if not self.rows:
            return (0, 0)
        return (statistics.median(r[0] for r in self.rows), max(r[1] for r in self.rows))


def env(**kw):
    for k in ("NG_FDLAP_S", "NG_FDLAP_XCHUNK", "NG_FDLAP_RING", "NG_FDLAP_EDGE_TEST", "NG_MBAR_HINT_NS", "NG_FDLAP_TY"):
        os.environ.pop(k, None)
    for k, v in kw.items():
        os.environ[k] = str(v)
    _lib.tuning_reload()


def whole(**kw):
    def fn():
        lap.apply_local(f, out)
    return (kw, fn)


def parts_serial():
    lap.apply_part(f, out, 1)
    lap.apply_part(f, out, 2, lo, hi)


def parts_two_streams():
    main = torch.cuda.current_stream()
    side.wait_stream(main)
    with torch.cuda.stream(side):
        lap.apply_part(f, out, 2, lo, hi)
        ev = side.record_event()
    lap.apply_part(f, out, 1)
    main.wait_event(ev)


variants = {
    "whole S=8 xc=64 (default)": whole(),
    "whole S=10 xc=32": whole(NG_FDLAP_S=10, NG_FDLAP_XCHUNK=32),
    "whole xc=32": whole(NG_FDLAP_XCHUNK=32),
    "whole xc=256": whole(NG_FDLAP_XCHUNK=256),
    "whole S=6": whole(NG_FDLAP_S=6),
    "whole S=10": whole(NG_FDLAP_S=10),
    "whole, mbarrier hint 0 ns": whole(NG_MBAR_HINT_NS=0),
    "whole, mbarrier hint 200 ns": whole(NG_MBAR_HINT_NS=200),
    "whole, mbarrier hint 5000 ns": whole(NG_MBAR_HINT_NS=5000),
    "whole, mbarrier hint 50000 ns": whole(NG_MBAR_HINT_NS=50000),
    "whole S=8 xc=128": whole(NG_FDLAP_XCHUNK=128),
    "slab: interior + boundary, serial": ({}, parts_serial),
    "slab: interior || boundary (2 streams)": ({}, parts_two_streams),
    "slab: whole with halos (one call)": ({}, lambda: lap.apply_local(f, out, lo, hi)),
    "slab: interior part alone": ({}, lambda: lap.apply_part(f, out, 1)),
    "slab: boundary part alone (halos)": ({}, lambda: lap.apply_part(f, out, 2, lo, hi)),
    "slab: boundary part alone (one-sided z)": ({}, lambda: lap.apply_part(f, out, 2)),
    "experiment: tiles 1 and nt-2 alone": ({}, lambda: lap.apply_part(f, out, 3)),
    "experiment: boundary tiles as interior": ({"NG_FDLAP_EDGE_TEST": 1}, lambda: lap.apply_part(f, out, 2, lo, hi)),
}
names = [n for n in variants if only is None or any(o in n for o in only)]
res = {n: [] for n in names}
clk = {n: [] for n in names}
for r in range(rounds):
    order = names[r % len(names):] + names[:r % len(names)]
    for n in order:
        kw, fn = variants[n]
        env(**kw)
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        if burst:
            time.sleep(0.4)                         # let the part leave its power cap: the next 5 applies run at boost clocks
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with Clock() as c:
            a.record()
            for _ in range(5):
                fn()
            b.record()
            torch.cuda.synchronize()
        res[n].append(a.elapsed_time(b) / 5)
        clk[n].append(c.summary())
env()
print(f"shape {shape}, {rounds} rounds x 5 applies; 16 B/pt")
for n in names:
    med, mn = statistics.median(res[n]), min(res[n])
    mhz = statistics.median(c[0] for c in clk[n])
    pw = max(c[1] for c in clk[n])
    print(f"{n:42s} median {med:7.3f} ms  min {mn:7.3f} ms  {16 * pts / med / 1e6:7.0f} GB/s (median)  "
          f"SM {mhz:5.0f} MHz  {pw:5.0f} W max")
