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


def test_cvxopt_engine_must_be_installed_when_asked_for():
    try:
        import cvxopt  # noqa: F401

        if not getattr(cvxopt, "__oracle_shim__", False):
            pytest.skip("cvxopt present")
    except ImportError:
        pass
    with pytest.raises(ImportError, match="cvxopt"):
        lisa_solvers.solver_cvxopt(np.ones((2, 3)), np.ones(3), np.ones(2), None, np.ones(3), 1e-8,
                                   logging.getLogger("x"), 1e-15, -1e-12, 1e-4, engine="cvxopt")  # fmt: skip
    with pytest.raises(ValueError, match="engine"):
        lisa_solvers.solver_cvxopt(np.ones((2, 3)), np.ones(3), np.ones(2), None, np.ones(3), 1e-8,
                                   logging.getLogger("x"), 1e-15, -1e-12, 1e-4, engine="nope")  # fmt: skip


def test_product_refuses_the_oracle_stand_ins():
    """With oracle/qcgrid_shim on sys.path (as the golden generator has it) ``import cvxopt`` finds
    the oracle's stand-in; the product must treat the package as absent."""
    import sys

    from conftest import ROOT

    from horton_part_b200.utils import optional_package

    saved = {k: sys.modules.pop(k, None) for k in ("cvxopt", "qpsolvers")}
    sys.path.insert(0, str(ROOT / "oracle" / "qcgrid_shim"))
    try:
        import cvxopt
        import qpsolvers

        assert cvxopt.__oracle_shim__ and qpsolvers.__oracle_shim__
        assert optional_package("cvxopt") is None and optional_package("qpsolvers") is None
    finally:
        sys.path.pop(0)
        for k, v in saved.items():
            sys.modules.pop(k, None)
            if v is not None:
                sys.modules[k] = v
    assert optional_package("a_package_that_does_not_exist") is None
    assert optional_package("json") is not None


CONVEX = np.load(GOLDEN / "convex_radial.npz")


@pytest.mark.parametrize("key", sorted(CONVEX.files))
def test_convex_programme_solver_matches_reference_through_shim(key):
    """lisa_solvers.solver_cvxopt (built-in interior-point engine) against the reference's
    solver_cvxopt whose cvxopt.solvers.cp call was answered by the oracle's SciPy-based stand-in:
    two independent methods, one minimiser."""
    func_type, number, mode = key.split("/")
    helper = ExpBasisFuncHelper.from_function_type(func_type)
    pop = {8: 8.5, 1: 0.7, 6: 6.1}[int(number)]
    bs, rho, c0, r, w = synthetic.radial_problem(helper, int(number), pop)
    log = logging.getLogger("test_algo_host")
    got = lisa_solvers.solver_cvxopt(bs, rho, c0.copy(), r, w, 1e-8, log, 1e-15, -1e-12, 1e-4,
                                     allow_neg_params=(mode == "free"), engine="builtin")  # fmt: skip
    ref = CONVEX[key]
    assert abs(got.sum() - np.einsum("i,i", w, rho)) < 1e-12
    if mode == "nonneg":
        assert (got >= 0).all()
    # the density the coefficients build is pinned tightly; single coefficients of the nearly
    # linearly dependent Slater sets are loose along the flat directions of the objective
    pro_got, pro_ref = got @ bs, ref @ bs
    assert np.sqrt(np.einsum("i,i,i", w, pro_got - pro_ref, pro_got - pro_ref)) < 1e-10
    np.testing.assert_allclose(got, ref, atol=1e-9 if func_type == "gauss" else 1e-7)


def test_batched_convex_programme_equals_the_per_atom_solver():
    """solver_cvxopt_batched (atoms of equal shape stacked, mixed shapes grouped) against per-atom
    solver_cvxopt calls and against the reference-through-shim vectors."""
    log = logging.getLogger("test_algo_host")
    for func_type in ("gauss", "slater"):
        helper = ExpBasisFuncHelper.from_function_type(func_type)
        problems, keys = [], []
        for number, pops in ((8, (8.5, 8.1, 8.9)), (1, (0.7, 0.55)), (6, (6.1,)), (8, (7.6,))):
            for pop in pops:
                bs, rho, c0, r, w = synthetic.radial_problem(helper, number, pop)
                problems.append((bs, rho, c0, r, w))
                keys.append((number, pop))
        got = lisa_solvers.solver_cvxopt_batched(problems, 1e-8, log, 1e-15, -1e-12, 1e-4)
        assert len(got) == len(problems)
        for (number, pop), prob, x in zip(keys, problems, got):
            bs, rho, c0, r, w = prob
            one = lisa_solvers.solver_cvxopt(bs, rho, c0.copy(), r, w, 1e-8, log, 1e-15, -1e-12, 1e-4, engine="builtin")
            assert abs(x.sum() - np.einsum("i,i", w, rho)) < 1e-12 and (x >= 0).all()
            pro_a, pro_b = x @ bs, one @ bs
            assert np.sqrt(np.einsum("i,i,i", w, pro_a - pro_b, pro_a - pro_b)) < 1e-10
            np.testing.assert_allclose(x, one, atol=1e-12 if func_type == "gauss" else 1e-6)
            key = f"{func_type}/{number}/nonneg"
            if (number, pop) in ((8, 8.5), (1, 0.7), (6, 6.1)):
                np.testing.assert_allclose(x, CONVEX[key], atol=1e-9 if func_type == "gauss" else 1e-6)


def test_batched_convex_programme_reports_non_convergence():
    helper = ExpBasisFuncHelper.from_function_type("gauss")
    bs, rho, c0, r, w = synthetic.radial_problem(helper, 8, 8.5)
    log = logging.getLogger("test_algo_host")
    log.disabled = True
    try:
        got = lisa_solvers.solver_cvxopt_batched([(bs, rho, c0, r, w)] * 2, 1e-8, log, 1e-15, -1e-12, 1e-4, maxiters=2)
    finally:
        log.disabled = False
    assert got == [None, None]
    with pytest.raises(ValueError, match="not finite"):
        lisa_solvers.solver_cvxopt_batched([(bs, rho * np.nan, c0, r, w)], 1e-8, log, 1e-15, -1e-12, 1e-4)


def test_sc_plus_convex_falls_back_to_the_programme():
    """With too few self-consistent iterations allowed, "sc-plus-convex" hands over to the convex
    programme (alisa.py:356-457; the reference's own hand-over call omits `population_cutoff` and
    raises TypeError) and must land on the same minimiser."""
    helper = ExpBasisFuncHelper.from_function_type("gauss")
    bs, rho, c0, r, w = synthetic.radial_problem(helper, 8, 8.5)
    log = logging.getLogger("test_algo_host")
    got = lisa_solvers.solver_sc_plus_cvxopt(bs, rho, c0.copy(), r, w, 1e-8, log, 1e-15, -1e-12, 1e-4, sc_iter_limit=3)
    np.testing.assert_allclose(got, CONVEX["gauss/8/nonneg"], atol=1e-9)
    # enough iterations: the fixed point itself, same minimiser at the fixed-point tolerance
    sc = lisa_solvers.solver_sc_plus_cvxopt(bs, rho, c0.copy(), r, w, 1e-8, log, 1e-15, -1e-12, 1e-4)
    np.testing.assert_allclose(sc, CONVEX["gauss/8/nonneg"], atol=1e-5)


def test_interior_point_on_quadratic_programmes():
    """algo/cp.py against the exact active-set solver on random simplex-constrained QPs."""
    from horton_part_b200.algo import cp, solve_qp_simplex

    rng = np.random.default_rng(3)
    for _ in range(25):
        n = int(rng.integers(2, 12))
        B = rng.normal(size=(n, n))
        P = B @ B.T + 0.05 * np.eye(n)
        q = 3 * rng.normal(size=n)
        total = float(rng.uniform(0.1, 9))

        def F(x=None, z=None):
            if x is None:
                return 0, np.full(n, total / n)
            g = P @ x + q
            f = 0.5 * x @ P @ x + q @ x
            return (f, g) if z is None else (f, g, z[0] * P)

        sol = cp(F, G=-np.identity(n), h=np.zeros(n), A=np.ones((1, n)), b=np.array([total]))
        assert sol["status"] == "optimal"
        np.testing.assert_allclose(sol["x"], solve_qp_simplex(P, q, total), atol=1e-9 * max(1.0, total))
        # KKT: multipliers non-negative, complementary to the slacks
        assert (sol["zl"] >= 0).all() and sol["gap"] < 1e-10


def test_interior_point_general_constraints_and_failure_modes():
    """Box + hyperplane projection with a closed-form answer (bisection on the multiplier), an
    equality-only programme with a closed form, an infeasible start, and the iteration limit."""
    from horton_part_b200.algo import cp

    rng = np.random.default_rng(11)
    n = 9
    target = rng.normal(size=n)
    lo, hi, total = -0.3 * np.ones(n), 0.8 * np.ones(n), 1.7

    def F(x=None, z=None):
        if x is None:
            return 0, np.zeros(n)  # violates sum x = total: infeasible start
        d = x - target
        return (d @ d, 2 * d) if z is None else (d @ d, 2 * d, z[0] * 2 * np.identity(n))

    G = np.vstack([np.identity(n), -np.identity(n)])
    h = np.concatenate([hi, -lo])
    sol = cp(F, G=G, h=h, A=np.ones((1, n)), b=[total])
    assert sol["status"] == "optimal"
    a, b = -10.0, 10.0  # projection: x = clip(target - nu, lo, hi) with sum x = total
    for _ in range(200):
        nu = 0.5 * (a + b)
        a, b = (nu, b) if np.clip(target - nu, lo, hi).sum() > total else (a, nu)
    np.testing.assert_allclose(sol["x"], np.clip(target - nu, lo, hi), atol=1e-9)

    # equality only: min sum a_i exp(x_i), sum x = b  =>  a_i exp(x_i) = lambda for all i
    coef = rng.uniform(0.5, 2.0, size=n)

    def F2(x=None, z=None):
        if x is None:
            return 0, np.zeros(n)
        e = coef * np.exp(x)
        return (e.sum(), e) if z is None else (e.sum(), e, z[0] * np.diag(e))

    sol = cp(F2, A=np.ones((1, n)), b=[2.5])
    assert sol["status"] == "optimal" and sol["zl"].size == 0
    lam = np.exp((2.5 + np.log(coef).sum()) / n)
    np.testing.assert_allclose(sol["x"], np.log(lam / coef), atol=1e-10)

    assert cp(F2, A=np.ones((1, n)), b=[2.5], options={"maxiters": 1})["status"] == "unknown"
    with pytest.raises(ValueError, match="not finite"):
        cp(lambda x=None, z=None: (0, np.zeros(2)) if x is None else (np.nan, np.zeros(2)))


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


def test_interior_point_against_scipy_on_random_entropy_programmes():
    """algo/cp.py on random programmes of the aLISA form  min -sum_i a_i ln((B x)_i)  s.t.  x >= 0,
    sum x = N  (B > 0, a > 0; some columns dominated so that bounds become active) against SciPy's
    SLSQP started from several points: an independent method on the same unique minimiser."""
    from scipy.optimize import minimize

    from horton_part_b200.algo import cp

    rng = np.random.default_rng(17)
    for trial in range(12):
        n, m = int(rng.integers(3, 9)), int(rng.integers(20, 60))
        B = rng.uniform(0.05, 1.0, size=(m, n)) * np.exp(-rng.uniform(0, 3, size=(m, 1)) * np.arange(n)[None, :])
        if trial % 3 == 0:
            B[:, -1] = 0.5 * B[:, 0]  # a dominated column: its coefficient must end on the bound
        a = rng.uniform(0.1, 2.0, size=m)
        total = float(rng.uniform(0.5, 8.0))

        def F(x=None, z=None):
            if x is None:
                return 0, np.full(n, total / n)
            p = B @ x
            if (p <= 0).any():
                return np.inf, np.full(n, np.nan)
            f = -(a * np.log(p)).sum()
            g = -B.T @ (a / p)
            return (f, g) if z is None else (f, g, z[0] * (B.T * (a / p**2)) @ B)

        sol = cp(F, G=-np.identity(n), h=np.zeros(n), A=np.ones((1, n)), b=[total])
        assert sol["status"] == "optimal"
        x = sol["x"]
        assert abs(x.sum() - total) < 1e-10 and (x > -1e-14).all()
        best = None
        for start in (np.full(n, total / n), total * rng.dirichlet(np.ones(n))):
            res = minimize(lambda v: F(np.maximum(v, 1e-300))[0], start, jac=lambda v: F(np.maximum(v, 1e-300))[1],
                           method="SLSQP", bounds=[(0, None)] * n, constraints={"type": "eq", "fun": lambda v: v.sum() - total,
                                                                               "jac": lambda v: np.ones(n)},
                           options={"ftol": 1e-15, "maxiter": 500})  # fmt: skip
            if best is None or res.fun < best.fun:
                best = res
        assert F(x)[0] <= best.fun + 1e-9 * max(1.0, abs(best.fun))  # at least as good as SLSQP's point
        np.testing.assert_allclose(B @ x, B @ best.x, rtol=2e-4)  # and the same fitted function
        # KKT at the interior-point answer: gradient + multiplier is zero on the support, >= 0 off it
        g = F(x)[1]
        mult = g + sol["y"][0]
        support = x > 1e-8 * total
        assert np.abs(mult[support]).max() < 1e-6 * max(1.0, np.abs(g).max())
        assert (mult[~support] > -1e-6 * max(1.0, np.abs(g).max())).all()
        if trial % 3 == 0:
            assert x[-1] < 1e-8 * total
