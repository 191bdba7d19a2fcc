"""Correctness of every fused-Laplacian kernel variant against the generic kernel (bit-equal), incl. slab halos."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from numgrids_b200.slab import SlabLaplacian  # noqa: E402

variants = [dict(NG_FDLAP_V="3", NG_FDLAP_R=str(r), NG_FDLAP_WY=str(w), NG_FDLAP_S=str(s), NG_FDLAP_XCHUNK=str(xc))
            for (r, w, s) in ((2, 8, 6), (2, 4, 6), (4, 4, 6), (1, 8, 6), (2, 8, 8), (4, 8, 5), (3, 8, 6))
            for xc in (5, 32)]
variants += [dict(NG_FDLAP_V="4", NG_FDLAP_R=str(r), NG_FDLAP_WY=str(w), NG_FDLAP_S=str(s), NG_FDLAP_XCHUNK=str(xc))
             for (r, w, s) in ((2, 8, 6), (2, 4, 6), (4, 4, 6), (1, 8, 6), (2, 8, 8), (2, 8, 5), (3, 8, 6), (2, 6, 6))
             for xc in (1, 7, 32)]
variants += [dict(NG_FDLAP_V="2", NG_FDLAP_R="2", NG_FDLAP_WY="8", NG_FDLAP_XCHUNK="32")]
shapes = [(64, 64, 64), (40, 37, 128), (7, 16, 64), (33, 48, 192), (9, 21, 256)]
bad = 0
for shape in shapes:
    lap = SlabLaplacian(shape, [1.0 / (n - 1) for n in shape])
    torch.manual_seed(1)
    f = torch.randn(*shape, dtype=torch.float64, device="cuda")
    rows = shape[0] * shape[1]
    lo = torch.randn(rows, 2, dtype=torch.float64, device="cuda")
    hi = torch.randn(rows, 2, dtype=torch.float64, device="cuda")
    for halos in ((None, None), (lo, hi), (lo, None), (None, hi)):
        os.environ["NG_FDLAP_GENERIC"] = "1"
        ref = lap.apply_local(f, torch.empty_like(f), *halos).clone()
        os.environ["NG_FDLAP_GENERIC"] = "0"
        for v in variants:
            os.environ.update(v)
            out = torch.full_like(f, float("nan"))
            lap.apply_local(f, out, *halos)
            torch.cuda.synchronize()
            ok = torch.equal(out, ref)
            if not ok:
                bad += 1
                nbad = int((out != ref).sum())
                print("MISMATCH", shape, [h is not None for h in halos], v, "n_bad", nbad)
print("bad", bad)
