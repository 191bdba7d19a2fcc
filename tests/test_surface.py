"""Drop-in surface: every public name of the reference package exists here with a compatible signature
(build container only: needs the reference sources to compare against)."""
import inspect

import pytest

import numgrids_b200 as ng
from oracle import reference_loader

pytestmark = pytest.mark.skipif(not reference_loader.available(), reason="reference sources only exist in the build container")

# members of reference classes that are deliberately absent here (plotting needs matplotlib)
ABSENT = {("Grid", "plot"), ("CurvilinearGrid", "plot"), ("SphericalGrid", "plot"), ("CylindricalGrid", "plot"),
          ("PolarGrid", "plot"), ("MultiGrid", "plot")}


def _public(obj):
    return {n for n in dir(obj) if not n.startswith("_")}


def _params(fn):
    return [p for p in inspect.signature(fn).parameters.values()]


def _compatible(ref_fn, our_fn, what):
    """Same parameter names, kinds and defaults in the same order; ours may add trailing optional parameters."""
    rp, op = _params(ref_fn), _params(our_fn)
    assert len(op) >= len(rp), what
    for r, o in zip(rp, op):
        assert r.name == o.name and r.kind == o.kind, (what, r, o)
        if r.default is not inspect.Parameter.empty:
            assert o.default == r.default, (what, r, o)
    for extra in op[len(rp):]:
        assert extra.default is not inspect.Parameter.empty or extra.kind in (extra.VAR_POSITIONAL, extra.VAR_KEYWORD), (what, extra)


def test_every_reference_name_is_exported_with_a_compatible_signature():
    ref = reference_loader.load()
    names = [n for n in vars(ref) if not n.startswith("_") and not inspect.ismodule(getattr(ref, n))]
    assert len(names) >= 25
    for name in names:
        assert hasattr(ng, name), f"numgrids_b200 lacks {name}"
        r, o = getattr(ref, name), getattr(ng, name)
        if inspect.isclass(r):
            assert inspect.isclass(o), name
            missing = {m for m in _public(r) - _public(o) if (name, m) not in ABSENT}
            assert not missing, (name, missing)
            if "__init__" in vars(r):
                _compatible(r.__init__, o.__init__, f"{name}.__init__")
            for m in sorted(_public(r) & _public(o)):
                rm, om = inspect.getattr_static(r, m), inspect.getattr_static(o, m)
                assert isinstance(rm, property) == isinstance(om, property), (name, m)
                if inspect.isfunction(rm) and inspect.isfunction(om):
                    _compatible(rm, om, f"{name}.{m}")
        elif inspect.isfunction(r):
            _compatible(r, o, name)


def test_submodule_classes_exist():
    ref = reference_loader.load()
    import importlib
    for mod, names in (("axes", ["Axis", "EquidistantAxis", "ChebyshevAxis", "LogAxis"]),
                       ("diff", ["GridDiff", "FiniteDifferenceDiff", "FFTDiff", "ChebyshevDiff", "LogDiff"]),
                       ("boundary", ["BoundaryCondition"]), ("amr", ["ErrorEstimator"]), ("io", ["save_grid"]),
                       ("interpol", ["Interpolator"]), ("integration", ["Integral"]), ("curvilinear", ["CurvilinearGrid"])):
        ours = importlib.import_module(f"numgrids_b200.{mod}")
        theirs = importlib.import_module(f"numgrids.{mod}")
        for n in names:
            assert hasattr(ours, n) and hasattr(theirs, n), (mod, n)
            rc, oc = getattr(theirs, n), getattr(ours, n)
            if inspect.isclass(rc):
                missing = {m for m in _public(rc) - _public(oc) if (n, m) not in ABSENT and m != "plot"}
                assert not missing, (mod, n, missing)
