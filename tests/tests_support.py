"""Shared helpers for the test-suite (grid construction from golden metadata)."""
import numpy as np


def mesh(coords):
    return np.meshgrid(*coords, indexing="ij")


def scale_factor_fns(name):
    ones = lambda c: np.ones_like(c[0])
    if name == "spherical":
        return (ones, lambda c: c[0], lambda c: c[0] * np.sin(c[1]))
    if name == "cylindrical":
        return (ones, lambda c: c[0], ones)
    if name == "polar":
        return (ones, lambda c: c[0])
    if name == "unit3":
        return (ones, ones, ones)
    if name == "unit2":
        return (ones, ones)
    if name == "custom3":       # must match tests/golden/make_golden.py CUSTOM3
        return (lambda c: 1.0 + c[0] ** 2, lambda c: c[0] * np.exp(c[1]),
                lambda c: 2.0 + np.sin(2 * np.pi * c[2]) * c[1])
    raise KeyError(name)


def scale_factor_arrays(name, coords):
    m = mesh(coords)
    return tuple(fn(m) for fn in scale_factor_fns(name))


def smooth_field(coords, shift=0.0):
    """Same deterministic field as tests/golden/make_golden.py."""
    nd = len(coords)
    f = np.ones([len(c) for c in coords])
    for i, c in enumerate(coords):
        shp = [1] * nd
        shp[i] = len(c)
        c = np.asarray(c).reshape(shp)
        f = f * np.sin((i + 1.3) * c + 0.2 * i + shift) + 0.1 * np.cos(c + shift)
    return f


def build_axes(ng, specs):
    kind = {"equidistant": ng.AxisType.EQUIDISTANT, "periodic": ng.AxisType.EQUIDISTANT_PERIODIC,
            "chebyshev": ng.AxisType.CHEBYSHEV, "log": ng.AxisType.LOGARITHMIC}
    return [ng.create_axis(kind[k], n, lo, hi) for k, n, lo, hi in specs]


def build_curv_grid(ng, name, axes):
    if name == "spherical":
        return ng.SphericalGrid(*axes)
    if name == "cylindrical":
        return ng.CylindricalGrid(*axes)
    if name == "polar":
        return ng.PolarGrid(*axes)
    return ng.CurvilinearGrid(*axes, scale_factors=scale_factor_fns(name))


def amr_field(name, mesh):
    """Analytic fields of the adaptive-refinement golden cases (tests/golden/make_golden_amr.py)."""
    if name == "poly5":
        return mesh[0] ** 5
    if name == "aniso":                       # needs far more resolution along axis 0 than along axis 1
        return np.sin(9.0 * mesh[0]) * np.exp(-mesh[0]) + mesh[1] ** 2
    if name == "wave3":
        return np.sin(4.0 * mesh[0]) * np.log(mesh[1]) + np.cos(3.0 * mesh[2]) * mesh[0]
    raise KeyError(name)
