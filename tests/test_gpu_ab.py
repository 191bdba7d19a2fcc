"""A/B parity of the fast kernels against the simple kernels they replace (same library, other code path):
streaming finite-difference kernels vs the gather kernel (bit for bit), the general Chebyshev kernel vs the
fallback (bit for bit), the strided-pencil FFT kernel vs the radix-2 shared-memory kernel (<= 5e-13).
Each tool draws random shapes / orders / fused operands; the simple kernels are pinned to the oracle and the
golden vectors by the other GPU tests."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tool,cases,seed", [("check_fd_stream.py", 120, 21), ("check_dense_general.py", 80, 22),
                                             ("check_fft_strided.py", 50, 23)])
def test_fast_kernels_equal_simple_kernels(gpu_ready, tool, cases, seed):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", tool), str(cases), str(seed)],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "0 mismatches" in r.stdout


def test_fused_laplacian_kernels_equal_the_simple_kernel(gpu_ready):
    """Plane-ring kernel (every ring depth, several x chunks), register-march fallback, slab halos and the interior /
    boundary split, all against the thread-per-point kernel: tools/check_fdlap.py."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "check_fdlap.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "bad 0" in r.stdout
