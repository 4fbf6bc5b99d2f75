"""GPU parity of the spline pro-atom path (ISA): spline construction vs SciPy, full partitioning vs
the reference's own outputs."""

import numpy as np
import pytest
import stockholder_oracle as oracle

pytestmark = pytest.mark.gpu


def _gold(gold, tag):
    return {k.split("/", 1)[1]: gold[k] for k in gold.files if k.startswith(tag + "/")}


def _check_spline_coefficients(coef, off, xs, ys, tol=1e-12):
    from scipy.interpolate import CubicSpline

    for a, (x, y) in enumerate(zip(xs, ys)):
        spl = CubicSpline(x, np.where(y < 0, 0.0, y), True)
        got = coef[4 * (off[a] - a) : 4 * (off[a + 1] - a - 1)].reshape(-1, 4)
        # same piecewise cubic: compare values (incl. extrapolation) on a dense sample
        t = np.linspace(x[0] - 0.5, x[-1] + 2.0, 4001)
        seg = np.clip(np.searchsorted(x, t, side="right") - 1, 0, len(x) - 2)
        dd = t - x[seg]
        mine = ((got[seg, 0] * dd + got[seg, 1]) * dd + got[seg, 2]) * dd + got[seg, 3]
        ref = spl(t)
        assert np.abs(mine - ref).max() <= tol * max(1.0, np.abs(ref).max()), a


def test_spline_build_matches_scipy():
    import torch

    from horton_part_b200 import _lib

    dev = torch.device("cuda:0")
    rng = np.random.default_rng(3)
    sizes = [2, 3, 4, 7, 120, 150]
    xs = [np.sort(rng.uniform(0.01, 20.0, n)) * np.linspace(1, 3, n) for n in sizes]
    ys = [np.exp(-x) * (1 + 0.1 * rng.normal(size=x.size)) for x in xs]
    ys[3][2] = -0.5  # exercises the negative clip
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    knots, vals = np.concatenate(xs), np.concatenate(ys)
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    coef = torch.zeros(4 * (len(knots) - len(sizes)), dtype=torch.float64, device=dev)
    work = torch.zeros(2 * len(knots), dtype=torch.float64, device=dev)
    # serial Thomas solve (any knot count)
    _lib.call("hp_spline_build", len(sizes), d(off), d(knots), d(vals), 1, coef, work, None, None, 0,
              torch.cuda.current_stream(dev).cuda_stream)
    _check_spline_coefficients(coef.cpu().numpy(), off, xs, ys)


def test_spline_build_parallel_matches_scipy():
    """One block per atom with the precomputed inverse of the not-a-knot system (hp_spline_system_inverse):
    the version the ISA iteration uses; radial grids of the reference's tests and adversarial knot sets."""
    import torch

    from horton_part_b200 import _lib, gridlite

    dev = torch.device("cuda:0")
    rng = np.random.default_rng(4)
    xs = [gridlite.BeckeRTransform(1e-4, 1.5).transform_1d_grid(gridlite.GaussChebyshev(150)).points,
          gridlite.ExpRTransform(5e-4, 2e1, 119).transform_1d_grid(gridlite.UniformInteger(120)).points,
          np.sort(rng.uniform(0.01, 20.0, 4)), np.sort(rng.uniform(0.01, 20.0, 37)) * np.linspace(1, 3, 37)]
    ys = [np.exp(-1.3 * x) * (1 + 0.1 * rng.normal(size=x.size)) for x in xs]
    ys[3][5] = -0.2
    sizes = [len(x) for x in xs]
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    knots, vals = np.concatenate(xs), np.concatenate(ys)
    pool, where, pos = [], [], 0
    for x in xs:
        mat = np.zeros(len(x) ** 2)
        _lib.call("hp_spline_system_inverse", len(x), np.ascontiguousarray(x), mat)
        where.append(pos)
        pool.append(mat)
        pos += mat.size
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    coef = torch.zeros(4 * (len(knots) - len(sizes)), dtype=torch.float64, device=dev)
    work = torch.zeros(2 * len(knots), dtype=torch.float64, device=dev)
    _lib.call("hp_spline_build", len(sizes), d(off), d(knots), d(vals), 1, coef, work, d(np.array(where, dtype=np.int64)),
              d(np.concatenate(pool)), max(sizes), torch.cuda.current_stream(dev).cuda_stream)
    _check_spline_coefficients(coef.cpu().numpy(), off, xs, ys, tol=1e-11)


def _isa(case, **kw):
    from horton_part_b200 import ISAWPart

    part = ISAWPart(case["coords"], case["numbers"], case["pseudo"], case["grid"], case["rho"], **kw)
    part.do_partitioning()
    return part


def _compare(part, ref, rtol=1e-8):
    assert part["niter"] == int(ref["niter"])
    np.testing.assert_allclose(part["charges"], ref["charges"], rtol=rtol, atol=1e-9)
    np.testing.assert_allclose(part["history_changes"], ref["history_changes"], rtol=1e-5)
    np.testing.assert_allclose(part["history_entropies"], ref["history_entropies"], rtol=rtol, atol=1e-11)
    np.testing.assert_allclose(part["propars"], ref["propars"], rtol=1e-9, atol=1e-12)  # measured 6e-14 (values > 1e-6)


def test_isa_h2o_against_reference_run(h2o):
    part = _isa(h2o)
    ref = _gold(h2o["gold"], "isa")
    _compare(part, ref)
    assert part["niter"] == 36
    assert abs(part["charges"] - np.array([-0.490017586929, 0.245018706885, 0.244998880045])).max() < 2e-3
    np.testing.assert_allclose(part["promoldens"][::97], ref["promoldens_sample"], rtol=1e-8)
    np.testing.assert_allclose(part["at_weights_0"][::53], ref["at_weights_0_sample"], rtol=1e-8, atol=1e-13)
    # first iteration: all-zero propars -> every weight is exactly 1/(2 natom)
    first = _isa(h2o, maxiter=1)
    assert np.allclose(first["at_weights_1"], 1.0 / 6.0, rtol=1e-15)
    # tests/common.py:127-138 check_proatom_splines analogue: host spline helper vs cached propars
    spl = part.get_proatom_spline(0)
    r = h2o["grid"].atgrids[0].rgrid.points
    assert abs(spl(r) - part["propars"][: len(r)]).max() < 1e-12


def test_isa_water6_against_reference_run(water6):
    part = _isa(water6, maxiter=60)
    _compare(part, _gold(water6["gold"], "isa"))
