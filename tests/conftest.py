import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# test_slab_step_behind_the_c_abi runs several ranks' streams (with spinning flag waits) on one GPU: give every
# stream its own hardware queue so that no wait kernel sits in front of the kernel it waits for (read at context creation)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name), allow_pickle=False)
    meta = json.loads(bytes(z["__meta__"]).decode()) if "__meta__" in z.files else None
    return z, meta


def rel_linf(a, ref):
    a = np.asarray(a)
    ref = np.asarray(ref)
    den = np.max(np.abs(ref))
    return float(np.max(np.abs(a - ref)) / den) if den > 0 else float(np.max(np.abs(a)))


# tolerances stated by BASELINE.json's north_star
TOL_STENCIL = 1e-12      # finite-difference / log paths (we also assert bit-exactness)
TOL_CHEB = 1e-12         # Chebyshev path (we also assert bit-exactness)
TOL_FFT = 1e-11          # periodic spectral path


@pytest.fixture(scope="session")
def gpu_ready():
    import numgrids_b200._lib as L
    L.load()
    if L.device_count() < 1:
        pytest.fail("GPU test selected but no CUDA device is usable (no CPU fallback exists)")
    return True
