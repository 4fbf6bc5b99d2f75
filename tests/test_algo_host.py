"""Host algebra (DIIS, CDIIS, BFGS) and the aLISA radial plug-in solvers against vectors produced
by the reference itself (oracle/gen_golden.py::case_algo).  CPU only."""

import contextlib
import io
import logging
import warnings

import numpy as np
import pytest
from conftest import GOLDEN

from horton_part_b200 import lisa_solvers, synthetic
from horton_part_b200.algo import bfgs, cdiis, diis, lstsq_solver_dyn
from horton_part_b200.core.basis import ExpBasisFuncHelper


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLDEN / "algo_host.npz")


def _toy():
    A, b = synthetic.contraction_map(12, seed=1)
    return (lambda x: A @ x + b + 0.05 * np.sin(x)), np.zeros(12)


@pytest.mark.parametrize("mode", ["R-CDIIS", "AD-CDIIS", "FD-CDIIS", "Roothaan"])
@pytest.mark.parametrize("qr", ["full", "economic"])
def test_cdiis_trajectory_matches_reference(gold, mode, qr):
    f, x0 = _toy()
    with contextlib.redirect_stdout(io.StringIO()):
        conv, n, rnorm, mk, cnorm, x, hist = cdiis(x0.copy(), f, 1e-10, 200, modeQR=qr, mode=mode)
    assert conv and n == int(gold[f"cdiis/{mode}/{qr}/niter"])
    assert list(mk) == list(gold[f"cdiis/{mode}/{qr}/mk"])  # restarts / depth changes at the same steps
    np.testing.assert_allclose(rnorm, gold[f"cdiis/{mode}/{qr}/rnorm"], rtol=1e-9, atol=1e-16)
    np.testing.assert_allclose(x, gold[f"cdiis/{mode}/{qr}/x"], rtol=1e-12)
    assert len(hist) == n + 1


@pytest.mark.parametrize("version", ["P", "A"])
@pytest.mark.parametrize("name", ["sp", "dyn"])
def test_diis_trajectory_matches_reference(gold, version, name):
    f, x0 = _toy()
    with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        x, n, hist = diis(x0.copy(), f, 1e-10, version=version, lstsq_solver=lstsq_solver_dyn if name == "dyn" else None)
    assert n == int(gold[f"diis/{version}/{name}/niter"])
    np.testing.assert_allclose(x, gold[f"diis/{version}/{name}/x"], rtol=1e-12)
    np.testing.assert_allclose(np.asarray(hist), gold[f"diis/{version}/{name}/history"], rtol=1e-9, atol=1e-14)


def test_diis_raises_when_not_converged():
    f, x0 = _toy()
    with pytest.raises(RuntimeError, match="not converge"):
        diis(x0, f, 1e-30, maxiter=3)


def test_cdiis_reports_non_convergence():
    f, x0 = _toy()
    conv, n, *_ = cdiis(x0, f, 1e-30, maxiter=4)
    assert not conv and n == 3


def test_bfgs_update(gold):
    H1 = bfgs(gold["bfgs/d1"], gold["bfgs/s"], gold["bfgs/d0"], gold["bfgs/H0"])
    np.testing.assert_allclose(H1, gold["bfgs/H1"], rtol=1e-13)
    # secant condition H1 y = s
    np.testing.assert_allclose(H1 @ (gold["bfgs/d1"] - gold["bfgs/d0"]), gold["bfgs/s"], rtol=1e-10)


SOLVERS = ["solver_sc", "solver_sc_1_iter", "solver_diis", "solver_cdiis", "solver_m_newton",
           "solver_quasi_newton", "solver_newton", "solver_trust_region"]  # fmt: skip


@pytest.mark.parametrize("func_type", ["gauss", "slater"])
@pytest.mark.parametrize("number,pop", [(8, 8.5), (1, 0.7), (6, 6.1)])
@pytest.mark.parametrize("name", SOLVERS)
def test_radial_plugin_solver_matches_reference(gold, func_type, number, pop, name):
    helper = ExpBasisFuncHelper.from_function_type(func_type)
    bs, rho, c0, r, w = synthetic.radial_problem(helper, number, pop)
    key = f"radial/{func_type}/{number}/{name}"
    log = logging.getLogger("test_algo_host")
    call = lambda: getattr(lisa_solvers, name)(bs, rho, c0.copy(), r, w, 1e-8, log, 1e-15, -1e-12, 1e-4)  # noqa: E731
    with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        if key + "/raised" in gold.files:
            with pytest.raises(Exception) as info:
                call()
            assert type(info.value).__name__ == str(gold[key + "/raised"])
            return
        got = call()
    ref = gold[key]
    # same LAPACK / SciPy calls on the same numbers: agreement far below the 1e-8 bar
    np.testing.assert_allclose(got, ref, rtol=1e-9, atol=1e-12)
    if name != "solver_sc_1_iter":
        assert abs(got.sum() - np.einsum("i,i", w, rho)) < 1e-3  # population is conserved


def test_builtin_solver_table_covers_the_reference_names():
    from horton_part_b200.alisa import LinearISAWPart

    assert set(LinearISAWPart.builtin_solvers) == {
        "cvxopt", "sc", "diis", "newton", "m-newton", "quasi-newton", "trust-region", "sc-1-iter",
        "sc-plus-convex", "cdiis"}  # fmt: skip


def test_cvxopt_solver_needs_the_package():
    try:
        import cvxopt  # noqa: F401

        pytest.skip("cvxopt present")
    except ImportError:
        pass
    with pytest.raises(ImportError, match="cvxopt"):
        lisa_solvers.solver_cvxopt(np.ones((2, 3)), np.ones(3), np.ones(2), None, np.ones(3), 1e-8,
                                   logging.getLogger("x"), 1e-15, -1e-12, 1e-4)  # fmt: skip


def test_active_set_qp_against_brute_force_enumeration():
    """algo/qp.py against the oracle shim's independent solver (every support set enumerated) on
    GISA-shaped problems: Gaussian overlap matrices and random SPD matrices."""
    import sys

    from conftest import ROOT

    sys.path.insert(0, str(ROOT / "oracle" / "qcgrid_shim"))
    try:
        import qpsolvers as shim
    finally:
        sys.path.pop(0)
    from horton_part_b200.algo import solve_qp_simplex

    rng = np.random.default_rng(0)
    for trial in range(60):
        n = int(rng.integers(2, 8))
        if trial % 2:
            a = np.sort(10 ** rng.uniform(-1.5, 2.5, size=n))
            P = 2 / np.pi**1.5 * (a[:, None] * a[None, :]) ** 1.5 / (a[:, None] + a[None, :]) ** 1.5
        else:
            B = rng.normal(size=(n, n))
            P = B @ B.T + 0.1 * np.eye(n)
        q = 3 * rng.normal(size=n)
        total = float(rng.uniform(0.1, 9))
        x = solve_qp_simplex(P, q, total)
        ref = shim.solve_qp(P, q, -np.identity(n), np.zeros((n, 1)), np.ones((1, n)), np.ones((1, 1)) * total)
        assert (x >= 0).all() and abs(x.sum() - total) < 1e-10
        np.testing.assert_allclose(x, ref, atol=1e-6 * max(1.0, total))
        cost = lambda y: 0.5 * y @ P @ y + q @ y  # noqa: E731
        assert cost(x) <= cost(ref) + 1e-9 * max(1.0, abs(cost(ref)))
