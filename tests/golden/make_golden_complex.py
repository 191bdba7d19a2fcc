"""Complex-valued fields through the REAL reference (build container only): tests/golden/complex_cases.npz.

The reference differentiates complex fields componentwise on finite-difference, logarithmic and Chebyshev axes
(findiff's ``w * y``, scipy's sparse matvec) and keeps ``np.real(...)`` of the result on periodic axes
(numgrids/diff.py:141).  numgrids_b200 applies its real kernels to the real and imaginary parts; the GPU test
(tests/test_gpu_complex.py) compares with the vectors stored here.  findiff is supplied by oracle/findiff_port.py.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..")))

from oracle import reference_loader  # noqa: E402
from tests_support import scale_factor_fns, smooth_field  # noqa: E402

ref = reference_loader.load()
from numgrids import AxisType, CurvilinearGrid, Diff, Grid, create_axis  # noqa: E402

KIND = {"equidistant": AxisType.EQUIDISTANT, "periodic": AxisType.EQUIDISTANT_PERIODIC,
        "chebyshev": AxisType.CHEBYSHEV, "log": AxisType.LOGARITHMIC}

DIFF_GRIDS = [
    [("equidistant", 24, 0.0, 1.0)],
    [("periodic", 16, 0.0, 2 * np.pi)],
    [("chebyshev", 17, -1.0, 2.0)],
    [("log", 25, 0.1, 10.0)],
    [("equidistant", 20, 0.0, 2.0), ("chebyshev", 13, -1.0, 1.0)],
    [("log", 12, 0.5, 4.0), ("periodic", 16, 0.0, 1.0), ("equidistant", 16, 0.0, 1.0)],
]
CURV_GRIDS = [
    ("custom3", [("equidistant", 16, 0.5, 1.5), ("equidistant", 16, 0.25, 1.0), ("equidistant", 64, 0.0, 1.0)]),
    ("custom3", [("chebyshev", 9, 0.5, 1.5), ("equidistant", 12, 0.25, 1.0), ("chebyshev", 10, 0.0, 1.0)]),
    ("unit2", [("log", 14, 0.5, 4.0), ("chebyshev", 11, 0.0, 1.0)]),
]


def axes_of(specs):
    return [create_axis(KIND[k], n, lo, hi) for k, n, lo, hi in specs]


def cfield(coords):
    return smooth_field(coords, 0.1) + 1j * smooth_field(coords, 0.7)


arrays, meta = {}, []
cid = 0
for specs in DIFF_GRIDS:
    g = Grid(*axes_of(specs))
    f = cfield([a.coords for a in g.axes])
    for order in (1, 2):
        for axis in range(g.ndims):
            out = Diff(g, order, axis)(f)
            arrays[f"d{cid}"] = out
            meta.append({"kind": "diff", "id": cid, "axes": specs, "order": order, "axis": axis,
                         "complex_out": bool(np.iscomplexobj(out))})
            cid += 1
for name, specs in CURV_GRIDS:
    g = CurvilinearGrid(*axes_of(specs), scale_factors=scale_factor_fns(name))
    f = cfield([a.coords for a in g.axes])
    v = [cfield([a.coords for a in g.axes]) * (i + 1.5) for i in range(g.ndims)]
    arrays[f"c{cid}_lap"] = g.laplacian(f)
    big = g.size > 10000                       # the 16 x 16 x 64 grid (single-pass Laplacian kernel): laplacian only
    ncurl = 0
    if not big:
        for i, comp in enumerate(g.gradient(f)):
            arrays[f"c{cid}_grad{i}"] = comp
        arrays[f"c{cid}_div"] = g.divergence(*v)
        curl = g.curl(*v)
        ncurl = len(curl) if isinstance(curl, tuple) else 1
        for i, comp in enumerate(curl if isinstance(curl, tuple) else (curl,)):
            arrays[f"c{cid}_curl{i}"] = comp
    meta.append({"kind": "curv", "id": cid, "name": name, "axes": specs, "ncurl": ncurl, "lap_only": bool(big)})
    cid += 1
arrays["__meta__"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
np.savez_compressed(os.path.join(HERE, "complex_cases.npz"), **arrays)
print(f"{cid} cases -> complex_cases.npz; diff outputs complex: {sum(m.get('complex_out', False) for m in meta)} "
      f"real: {sum(m['kind'] == 'diff' and not m['complex_out'] for m in meta)}")
