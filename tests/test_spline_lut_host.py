"""Host check of the interval look-up table of the spline pass (hp_spline_lut_size / hp_spline_lut_fill,
pure host code in the C-ABI library): table + forward scan must land on scipy PPoly's interval,
searchsorted_right(x, r) - 1 clamped to [0, n - 2] (core/stockholder.py:271-302 evaluates the spline
with extrapolation), for every radial transform the reference uses and for adversarial inputs."""

import numpy as np
import pytest

SHIFT = 15  # hi32(r) >> 15: sign, exponent, 5 mantissa bits (kLutShift in csrc/hp_spline.cu)


def _device_index(x, lut, key0, r):
    """NumPy restatement of the kernel's index computation; also returns the scan lengths."""
    hi = (np.asarray(r, dtype=np.float64).view(np.int64) >> 32).astype(np.int64)
    b = np.clip((hi >> SHIFT) - key0, 0, len(lut) - 1)
    i = lut[b].astype(np.int64)
    steps = np.zeros_like(i)
    last = len(x) - 2
    while True:
        adv = (i < last) & (x[np.minimum(i + 1, len(x) - 1)] <= r)
        if not adv.any():
            return i, steps
        i = i + adv
        steps += adv


def _tables(x, built_lib):
    from horton_part_b200 import _lib

    x = np.ascontiguousarray(x, dtype=np.float64)
    nb = int(_lib.call("hp_spline_lut_size", len(x), x))
    lut, key0 = np.zeros(nb, dtype=np.uint16), np.zeros(1, dtype=np.int32)
    _lib.call("hp_spline_lut_fill", len(x), x, key0, lut)
    return lut, int(key0[0])


def _grids():
    from horton_part_b200 import gridlite as g

    yield "becke150", g.BeckeRTransform(1e-4, 1.5).transform_1d_grid(g.GaussChebyshev(150)).points
    yield "exp120", g.ExpRTransform(5e-4, 2e1, 119).transform_1d_grid(g.UniformInteger(120)).points
    yield "power59", g.PowerRTransform(3.1e-8, 34.9, 58).transform_1d_grid(g.UniformInteger(59)).points
    yield "two", np.array([0.5, 2.0])
    yield "three", np.array([1e-3, 1.0, 50.0])
    yield "starts_at_zero", np.concatenate([[0.0], np.geomspace(1e-6, 30.0, 80)])
    yield "uniform", np.linspace(0.1, 20.0, 400)  # many knots per octave: the scan does the work
    yield "huge_range", np.geomspace(1e-200, 1e200, 300)  # more bins than the cap: key0 is raised


@pytest.mark.parametrize("name,x", list(_grids()), ids=[n for n, _ in _grids()])
def test_lut_plus_scan_is_searchsorted(built_lib, name, x):
    x = np.asarray(x, dtype=np.float64)
    lut, key0 = _tables(x, built_lib)
    assert 1 <= len(lut) <= 4096 and lut.max() <= len(x) - 2
    rng = np.random.default_rng(5)
    r = np.concatenate([
        x, np.nextafter(x, 0.0), np.nextafter(x, np.inf), [0.0, 5e-324, 1e-300, 1e300, x[0] / 3, x[-1] * 3],
        np.exp(rng.uniform(np.log(max(x[0], 1e-250)) - 2, np.log(x[-1]) + 2, 20000)),
        rng.uniform(0.0, x[-1] * 1.1, 20000),
    ])
    got, steps = _device_index(x, lut, key0, r)
    want = np.clip(np.searchsorted(x, r, side="right") - 1, 0, len(x) - 2)
    assert np.array_equal(got, want)
    if name in ("becke150", "exp120", "power59"):  # the reference's radial transforms: at most one step
        assert steps.max() <= 1, steps.max()
