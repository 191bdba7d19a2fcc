"""Run the REFERENCE's own test-suite against numgrids_b200 (build container only: needs /root/reference).

`numgrids` and its submodules are aliased to `numgrids_b200` before the reference's tests are collected
(`python tools/run_reference_tests.py [pytest args]`).  Without a GPU every test that computes on the device fails with
NumgridsB200Error (no CPU fallback); with one, the whole suite is the drop-in check."""
import importlib
import os
import sys
from unittest import mock

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF_TESTS = os.path.join(os.environ.get("NUMGRIDS_REFERENCE_ROOT", "/root/reference"), "tests")

import numgrids_b200  # noqa: E402

sys.modules["numgrids"] = numgrids_b200
for sub in ("api", "axes", "diff", "grids", "curvilinear", "boundary", "integration", "interpol", "amr", "io"):
    sys.modules[f"numgrids.{sub}"] = importlib.import_module(f"numgrids_b200.{sub}")
try:
    import matplotlib  # noqa: F401
except ImportError:
    mpl = mock.MagicMock()
    sys.modules.setdefault("matplotlib", mpl)
    sys.modules.setdefault("matplotlib.pyplot", mpl.pyplot)

import pytest  # noqa: E402

sys.exit(pytest.main(["-q", "-p", "no:cacheprovider", "--rootdir", "/tmp", REF_TESTS, *sys.argv[1:]]))
