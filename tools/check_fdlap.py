"""Correctness of the fused-Laplacian kernels against the thread-per-point kernel of fdlap.cu (bit-equal): the plane-ring
kernel at every ring depth / several x chunks, the register-march fallback (NG_FDLAP_RING=0), with and without slab halos,
and the interior / boundary split of the multi-GPU step."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from numgrids_b200 import _lib  # noqa: E402
from numgrids_b200.slab import SlabLaplacian  # noqa: E402

KEYS = ("NG_FDLAP_GENERIC", "NG_FDLAP_RING", "NG_FDLAP_S", "NG_FDLAP_XCHUNK")


def env(**kw):
    for k in KEYS:
        os.environ.pop(k, None)
    os.environ.update({k: str(v) for k, v in kw.items()})
    _lib.tuning_reload()


variants = [dict(NG_FDLAP_S=s, NG_FDLAP_XCHUNK=xc) for s in (6, 8, 10) for xc in (1, 7, 32)] + [dict(NG_FDLAP_RING=0)]
shapes = [(64, 64, 64), (40, 37, 128), (7, 16, 64), (33, 48, 192), (9, 21, 256), (20, 24, 320)]
bad = 0
for shape in shapes:
    lap = SlabLaplacian(shape, [1.0 / (n - 1) for n in shape])
    torch.manual_seed(1)
    f = torch.randn(*shape, dtype=torch.float64, device="cuda")
    rows = shape[0] * shape[1]
    lo = torch.randn(rows, 2, dtype=torch.float64, device="cuda")
    hi = torch.randn(rows, 2, dtype=torch.float64, device="cuda")
    for halos in ((None, None), (lo, hi), (lo, None), (None, hi)):
        env(NG_FDLAP_GENERIC=1)
        ref = lap.apply_local(f, torch.empty_like(f), *halos).clone()
        for v in variants:
            env(**v)
            out = torch.full_like(f, float("nan"))
            lap.apply_local(f, out, *halos)
            parts = torch.full_like(f, float("nan"))
            lap.apply_part(f, parts, 1, *halos)
            lap.apply_part(f, parts, 2, *halos)
            torch.cuda.synchronize()
            for what, o in (("whole", out), ("parts", parts)):
                if not torch.equal(o, ref):
                    bad += 1
                    print("MISMATCH", what, shape, [h is not None for h in halos], v, "n_bad", int((o != ref).sum()))
env()
print("bad", bad)
sys.exit(1 if bad else 0)
