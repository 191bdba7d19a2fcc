"""Import the REAL reference (maroba/numgrids) -- build container only, TEST INFRASTRUCTURE.

/root/reference exists only in the build container (never on the GPU box), so
this loader is used exclusively by ``tests/golden/make_golden.py`` and by the
CPU tests that are skipped when the path is absent.  It

* puts ``oracle.findiff_port`` on ``sys.modules['findiff']`` when the real
  findiff is not importable (reference numgrids/diff.py:10 needs it),
* stubs ``matplotlib`` with a MagicMock when it is missing (only plotting and
  the reference's tests/test_diff.py:5 import it),
* optionally makes ``FFTDiff._build_spectral_matrix`` lazy so that large
  periodic grids can be constructed (numgrids/diff.py:128 builds an
  ``n x grid.size`` Kronecker matrix eagerly).
"""
from __future__ import annotations

import importlib
import os
import sys
from unittest import mock

REFERENCE_ROOT = os.environ.get("NUMGRIDS_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "numgrids"))


def real_findiff_available() -> bool:
    try:
        mod = importlib.import_module("findiff")
    except ImportError:
        return False
    return not getattr(mod, "__version__", "").endswith("-port")


def load(lazy_fft_matrix: bool = False):
    """Return the reference's ``numgrids`` package (imported from /root/reference)."""
    if not available():
        raise RuntimeError(f"reference not present at {REFERENCE_ROOT}")
    if not real_findiff_available() and "findiff" not in sys.modules:
        from oracle import findiff_port
        sys.modules["findiff"] = findiff_port
    try:
        importlib.import_module("matplotlib")
    except ImportError:
        mpl = mock.MagicMock()
        sys.modules.setdefault("matplotlib", mpl)
        sys.modules.setdefault("matplotlib.pyplot", mpl.pyplot)
        sys.modules.setdefault("matplotlib.patches", mpl.patches)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    ref = importlib.import_module("numgrids")
    if not os.path.realpath(ref.__file__).startswith(os.path.realpath(REFERENCE_ROOT)):
        raise RuntimeError(f"'numgrids' resolved to {ref.__file__}, not the reference")
    if lazy_fft_matrix:
        import numgrids.diff as rdiff
        rdiff.FFTDiff._build_spectral_matrix = lambda self: None
    return ref
